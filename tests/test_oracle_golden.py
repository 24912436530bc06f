"""Pins the CPU oracle (oracle/streamsim_oracle.py) -- CPU only, no GPU.

The fixtures under tests/golden/ were produced by scripts/make_golden.py from the
unmodified reference at /root/reference.  If the oracle drifts from them, every
GPU parity claim built on it is void, so these are bit-exact comparisons.
"""
import numpy as np
import pytest

import streamsim_oracle as oracle
from conftest import (GENERATE_CASES, INJECT_CASES, ROOT, golden_generate_case,
                      golden_inject_case)
from helpers import oracle_survey


@pytest.mark.parametrize("name", GENERATE_CASES)
def test_generate_bit_equal_to_reference(name):
    cfg, draws, expect = golden_generate_case(name)
    got = oracle.generate(cfg, draws, None, data_root=ROOT)
    for k, v in expect.items():
        assert np.array_equal(got[k], v, equal_nan=True), (name, k)


@pytest.mark.parametrize("tag,dust", INJECT_CASES)
def test_inject_bit_equal_to_reference(tag, dust):
    cat, draws, expect, meta = golden_inject_case(tag, dust)
    sv = oracle_survey(meta["compl_band"], meta["first"], meta["nside_maglim"], meta["nside_ebv"])
    got = oracle.observe(sv, cat["ra"], cat["dec"], cat["mag_g"], cat["mag_r"], draws,
                         nside=meta["nside"], perfect_galstarsep=True, dust_correction=dust)
    for k, v in expect.items():
        assert np.array_equal(got[k], v, equal_nan=True), (tag, dust, k)
    # the fixture exercises every branch
    assert np.isnan(expect["mag_r_obs"]).any() and np.isnan(expect["magerr_r"]).any()
    assert expect["flag_observed"].any() and (~expect["flag_observed"]).any()
    assert (expect["flag_perfect_galstarsep"] != expect["flag_observed"]).any()
    if meta["first"] >= -10.4:   # photo-error table does not reach dsat-1: fill value 1.0 -> 10 mag
        assert np.isclose(expect["magerr_r"][cat["mag_r"] < 15.0], 10.000001, atol=1e-6).all()


def test_function_known_answers(known_answers):
    """The reference's own known-answer vectors (reference tests/test_functions.py:8-35)."""
    ka = known_answers
    x = np.array(ka["x"])
    nan = lambda v: np.array([np.nan if e is None else e for e in v])
    f = oracle.OFunc(dict(type="cubicsplineinterpolation", nodes=ka["spread_nodes"],
                          node_values=ka["spread_values"]))
    np.testing.assert_allclose(f(x), nan(ka["spread_expected"]))
    g = oracle.OFunc(dict(type="lineardensitycubicsplineinterpolation",
                          intensity_nodes=np.array(ka["intensity_nodes"]),
                          intensity_node_values=np.array(ka["intensity_values"]),
                          spread_nodes=np.array(ka["spread_nodes"]),
                          spread_node_values=np.array(ka["spread_values"])))
    np.testing.assert_allclose(g(x), nan(ka["linear_density_expected"]), atol=1e-7)
    h = oracle.OFunc(dict(type="filecubicsplineinterpolation",
                          filename="./data/patrick_2022_splines.csv", stream_name="atlas",
                          spline_type="spread"), data_root=ROOT)
    np.testing.assert_allclose(h(x), nan(ka["spread_expected"]))


def test_rotation_matches_notebook_rows(known_answers):
    """The only golden vector for the gala/astropy boundary: the first rows printed
    by reference notebooks/tutorial_inject_stream.ipynb (seed 42, frame accepted at
    trial 2).  Reproduces the frame from PCG64(42) draws 5-8 and the data from a
    separate PCG64(42) stream."""
    nb = known_answers["notebook_inject"]
    rng = np.random.default_rng(nb["seed"])
    phi1 = rng.uniform(-5, 5, nb["n"])
    phi2 = rng.uniform(-1, 1, nb["n"])
    np.testing.assert_allclose(np.c_[phi1, phi2][:5], nb["phi_head"], atol=5e-7)
    rng = np.random.default_rng(nb["seed"])
    for _ in range(nb["accepted_trial"]):
        e1, e2 = oracle.random_endpoints(rng), oracle.random_endpoints(rng)
    from stream_sim_b200.synthetic import NOTEBOOK_ENDPOINTS
    np.testing.assert_allclose([e1, e2], NOTEBOOK_ENDPOINTS, rtol=0, atol=1e-12)
    ra, dec = oracle.rotate(phi1, phi2, oracle.frame_from_endpoints(*e1, *e2))
    np.testing.assert_allclose(np.c_[ra, dec][:5], nb["radec_head"], atol=6e-7)


def test_ang2pix_known_values_and_identities():
    """HEALPix is unpinned against healpy itself (not installed); these are the
    healpy docstring examples (from memory) plus geometric identities."""
    th = lambda theta, phi: oracle.ang2pix_ring(16, np.degrees(np.atleast_1d(phi)),
                                                90 - np.degrees(np.atleast_1d(theta)))
    assert th(np.pi / 2, 0)[0] == 1440
    assert list(th([np.pi / 2, np.pi / 4, np.pi / 2, 0, np.pi],
                   [0., np.pi / 4, np.pi / 2, 0, 0])) == [1440, 427, 1520, 0, 3068]
    assert [int(oracle.ang2pix_ring(n, [0.], [0.])[0]) for n in (1, 2, 4, 8)] == [4, 12, 72, 336]
    # nside=1: the 12 base-pixel centres
    lat1 = np.degrees(np.arcsin(2 / 3))
    lon = [45, 135, 225, 315, 0, 90, 180, 270, 45, 135, 225, 315]
    lat = [lat1] * 4 + [0] * 4 + [-lat1] * 4
    assert list(oracle.ang2pix_ring(1, lon, lat)) == list(range(12))
    # poles, longitude wrap
    assert oracle.ang2pix_ring(4096, [0.], [90.])[0] == 0
    assert oracle.ang2pix_ring(4096, [0.], [-90.])[0] == 12 * 4096 ** 2 - 4
    assert oracle.ang2pix_ring(64, [360.], [10.])[0] == oracle.ang2pix_ring(64, [0.], [10.])[0]
    # equal-area: uniform points fill pixels uniformly; RING and NEST are consistent
    from stream_sim_b200.healpix import ring2nest
    rng = np.random.default_rng(0)
    lon = rng.uniform(0, 360, 400000)
    lat = np.degrees(np.arcsin(rng.uniform(-1, 1, 400000)))
    counts = np.bincount(oracle.ang2pix_ring(8, lon, lat), minlength=768)
    chi2 = ((counts - counts.mean()) ** 2 / counts.mean()).sum() / 767
    assert 0.8 < chi2 < 1.2
    for nside in (1, 2, 64, 4096):
        r = oracle.ang2pix_ring(nside, lon, lat)
        assert r.min() >= 0 and r.max() < 12 * nside ** 2
        assert np.array_equal(ring2nest(nside, r), oracle.ang2pix_nest(nside, lon, lat))
    assert sorted(ring2nest(8, np.arange(768)).tolist()) == list(range(768))


def test_philox_known_answers():
    """Random123 known-answer vectors for philox4x32-10."""
    h = lambda w: [int(x) for x in w]
    assert h(oracle.philox4x32_10(0, 0, 0, 0, 0, 0)) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    f = 0xffffffff
    assert h(oracle.philox4x32_10(f, f, f, f, f, f)) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert h(oracle.philox4x32_10(0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344,
                                  0xa4093822, 0x299f31d0)) == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]
    d = oracle.device_draws(1, 0, 200000)
    for k in ("u_phi1", "u_mass", "u_det"):
        assert 0 <= d[k].min() and d[k].max() < 1 and abs(d[k].mean() - 0.5) < 0.005
    for k in ("z_phi2", "z_r", "z_g"):
        assert abs(d[k].mean()) < 0.01 and abs(d[k].std() - 1) < 0.01
    # sharding invariance of the draw layout
    e = oracle.device_draws(1, 1000, 10)
    assert np.array_equal(e["z_r"], d["z_r"][1000:1010])


def test_imf_and_isochrone_sampling_sane():
    from stream_sim_b200.synthetic import isochrone_table
    iso = isochrone_table()
    mass, cdf = oracle.imf_cdf(iso["mass_init"].min(), iso["mass_init"].max())
    assert cdf[0] == 0.0 and cdf[-1] == 1.0 and np.all(np.diff(cdf) > 0)
    u = np.random.default_rng(2).uniform(size=100000)
    mg, mr = oracle.sample_mags(iso, u, np.full(len(u), 16.5))
    assert np.isfinite(mg).all() and np.isfinite(mr).all()
    assert np.median(mg - mr) > 0.3             # low-mass stars dominate: red colours
    # Chabrier (2003) log-normal branch peaks at m_c = 0.079 Msun
    m = np.array([0.05, 0.079, 0.12])
    p = oracle.chabrier2003_pdf(m)
    assert p[1] > p[0] and p[1] > p[2]
