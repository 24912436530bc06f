"""The hot loop's own double-precision elementary functions (csrc/device_math.cuh
``log10_nb``, ``exp10_nb``, ``atan2_nb``, ``sqrt_nb``) against NumPy / libm, through the
C ABI (``ss_math_probe``).  Bar: <= 4 ulp of the result (CUDA's own libm documents 1-2
ulp for the same functions; the north star asks 1e-5 relative on magnitudes and
positions), and sqrt within 1 ulp."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def ulp_err(got, ref):
    ref = np.asarray(ref, dtype=np.float64)
    return np.abs(got - ref) / np.spacing(np.abs(ref))


@pytest.fixture(scope="module")
def probe():
    from stream_sim_b200.engine import math_probe
    return lambda fn, x, y=None: math_probe(fn, x, y).cpu().numpy()


def test_log10(probe):
    rng = np.random.default_rng(11)
    t = np.concatenate([
        rng.uniform(0.0, 4.0, 400000),                 # t = 1 + x of the flux noise
        1.0 + rng.normal(0, 1e-3, 100000),             # near 1: no cancellation
        np.exp(rng.uniform(-600, 600, 100000)),        # the whole normal range
        [1.0, 2.0, 0.5, 10.0, 1e-300, 1e300, np.sqrt(2.0), np.nextafter(1.0, 0), np.nextafter(1.0, 2)],
    ])
    t = t[t > 0]
    got = probe("log10", t)
    ref = np.log10(t)
    near_one = np.abs(t - 1.0) < 1e-15
    assert np.all(got[t == 1.0] == 0.0)
    assert ulp_err(got[~near_one], ref[~near_one]).max() <= 4.0
    assert np.abs(got - ref).max() < 3e-13             # 1e300: |log10| = 300, 4 ulp = 2e-13


def test_exp10(probe):
    rng = np.random.default_rng(12)
    x = np.concatenate([rng.uniform(-4, 2, 400000), rng.uniform(-299, 299, 200000),
                        [0.0, 1.0, -1.0, 2.0, 0.5, -0.5, 299.9, -299.9]])
    got = probe("exp10", x)
    ref = np.power(10.0, x)
    assert ulp_err(got, ref).max() <= 4.0
    assert np.isnan(probe("exp10", np.array([np.nan])))[0]          # unseen pixel: NaN depth


def test_atan2(probe):
    rng = np.random.default_rng(13)
    y = np.concatenate([rng.normal(size=500000), rng.normal(size=1000) * 1e-12,
                        [0.0, 1.0, -1.0, 0.0, 1.0, -1.0, 1e-300, 0.0]])
    x = np.concatenate([rng.normal(size=500000), rng.normal(size=1000),
                        [1.0, 0.0, 0.0, -1.0, 1.0, -1.0, 1.0, 0.0]])
    got = probe("atan2", y, x)
    ref = np.arctan2(y, x)
    ok = ref != 0
    assert ulp_err(got[ok], ref[ok]).max() <= 4.0
    assert np.abs(got - ref).max() <= 4 * np.spacing(np.pi)
    # unit vectors as the rotation produces them: absolute error in degrees
    v = rng.normal(size=(200000, 3))
    v /= np.linalg.norm(v, axis=1)[:, None]
    lon = np.degrees(probe("atan2", v[:, 1], v[:, 0]))
    lat = np.degrees(probe("atan2", v[:, 2], np.hypot(v[:, 0], v[:, 1])))
    assert np.abs(lon - np.degrees(np.arctan2(v[:, 1], v[:, 0]))).max() < 1e-13
    assert np.abs(lat - np.degrees(np.arctan2(v[:, 2], np.hypot(v[:, 0], v[:, 1])))).max() < 1e-13


def test_sqrt(probe):
    rng = np.random.default_rng(14)
    a = np.concatenate([rng.uniform(0, 200, 400000), np.exp(rng.uniform(-600, 600, 100000)),
                        [0.0, 1.0, 4.0, 2.0, 0.25, 2.5e-5, 100.00002000000099]])
    got = probe("sqrt", a)
    ref = np.sqrt(a)
    assert got[0 == a][0] == 0.0
    nz = a > 0
    assert ulp_err(got[nz], ref[nz]).max() <= 1.0
    assert np.mean(got[nz] == ref[nz]) > 0.999          # correctly rounded but for rare ties
    assert np.isnan(probe("sqrt", np.array([np.nan])))[0]
