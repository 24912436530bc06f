"""Parity of the CUDA path against the reference-pinned oracle (run with -m gpu).

Everything here goes through the public Python API and the C ABI of
libstreamsim_b200.so.  Bars: bit-exact for pixel indices, flags and the phi1
grid index; <=1e-9 relative (1e-5 required by the north star) for fp64 positions
and magnitudes; 1e-3 mag in the fp32 output mode.
"""
import numpy as np
import pytest

import streamsim_oracle as oracle
from conftest import GENERATE_CASES, INJECT_CASES, golden_generate_case, golden_inject_case
from helpers import assert_close, oracle_survey, stream_config

pytestmark = pytest.mark.gpu

RTOL = 1e-9          # what we hold the fp64 path to (north star asks 1e-5)


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available(), "these tests need a B200"
    return torch


# ------------------------------------------------------------------ generate
@pytest.mark.parametrize("name", GENERATE_CASES)
def test_generate_matches_reference_golden(torch_cuda, name):
    """StreamModel.sample with the reference's own MT19937 draws injected equals
    the reference's output (tests/golden/generate_*.npz)."""
    from stream_sim_b200.model import StreamModel
    cfg, draws, expect = golden_generate_case(name)
    df = StreamModel(cfg).sample(len(draws["u_phi1"]), draws=draws)
    # phi1 is a grid value (or u*scale+loc): bit-exact
    assert np.array_equal(df["phi1"].to_numpy(), expect["phi1"]), "phi1 not bit-exact"
    assert_close(df["phi2"].to_numpy(), expect["phi2"], RTOL, 1e-12, "phi2")
    if "dist" in expect:
        assert_close(df["dist"].to_numpy(dtype=float), expect["dist"], RTOL, 0, "dist")
    assert list(df.columns) == ["phi1", "phi2", "dist", "mu1", "mu2", "rv", "mag_g", "mag_r"]


def test_generate_isochrone_matches_oracle(torch_cuda):
    from stream_sim_b200.model import StreamModel
    from stream_sim_b200.synthetic import isochrone_table
    cfg = stream_config("pal5")
    n = 1 << 16
    d = oracle.device_draws(7, 0, n)
    df = StreamModel(cfg).sample(n, draws=dict(u_phi1=d["u_phi1"], z_phi2=d["z_phi2"],
                                                u_dist=d["u_dist"], u_mass=d["u_mass"]))
    ref = oracle.generate(cfg, d, isochrone_table())
    assert np.array_equal(df["phi1"].to_numpy(), ref["phi1"])
    for k in ("phi2", "dist", "mag_g", "mag_r"):
        assert_close(df[k].to_numpy(), ref[k], RTOL, 1e-12, k)


# ------------------------------------------------------------------ inject (public API)
@pytest.mark.parametrize("tag,dust", INJECT_CASES)
def test_inject_matches_reference_golden(torch_cuda, tag, dust):
    """StreamInjector.inject(seed=..., rng_compat=True) reproduces the reference's
    DataFrame for the same seed: flags bit-exact, magnitudes to 1e-9."""
    import pandas as pd
    from stream_sim_b200.observed import StreamInjector
    from stream_sim_b200.surveys import Survey
    cat, draws, expect, meta = golden_inject_case(tag, dust)
    sv = Survey.synthetic(meta["nside_maglim"], meta["nside_ebv"],
                          completeness_band=meta["compl_band"], first=meta["first"])
    inj = StreamInjector(sv)
    res = inj.inject(pd.DataFrame(cat), seed=meta["seed"], perfect_galstarsep=True,
                     verbose=False, dust_correction=dust, rng_compat=True)
    for flag in ("flag_observed", "flag_perfect_galstarsep"):
        assert np.array_equal(res[flag].to_numpy(dtype=bool), expect[flag]), flag
    for b in "rg":
        col = res[f"mag_{b}_obs"].to_numpy()
        bad = np.array([isinstance(v, str) and v == "BAD_MAG" for v in col])
        assert np.array_equal(bad, np.isnan(expect[f"mag_{b}_obs"])), f"BAD_MAG pattern {b}"
        got = np.array([np.nan if s else float(v) for v, s in zip(col, bad)])
        assert_close(got, expect[f"mag_{b}_obs"], RTOL, 1e-9, f"mag_{b}_obs")
        assert_close(res[f"magerr_{b}"].to_numpy(), expect[f"magerr_{b}"], RTOL, 0, f"magerr_{b}")
    assert list(res.columns) == ["ra", "dec", "mag_g", "mag_r", "mag_r_obs", "magerr_r",
                                 "mag_g_obs", "magerr_g", "flag_observed",
                                 "flag_perfect_galstarsep"]


def test_ang2pix_matches_oracle_all_branches(torch_cuda):
    from stream_sim_b200 import healpix as hp
    rng = np.random.default_rng(3)
    n = 1 << 18
    lon = rng.uniform(-20.0, 380.0, n)
    lat = np.degrees(np.arcsin(rng.uniform(-1, 1, n)))
    lat[:1000] = rng.uniform(89.0, 90.0, 1000)          # have_sth branch
    lat[1000:2000] = rng.uniform(-90.0, -89.0, 1000)
    lon[:4], lat[:4] = [0, 0, 360, 45], [90, -90, 0, np.degrees(np.arcsin(2 / 3))]
    for nside in (1, 2, 16, 128, 512, 4096, 8192):
        assert np.array_equal(hp.ang2pix(nside, lon, lat), oracle.ang2pix_ring(nside, lon, lat)), nside
        assert np.array_equal(hp.ang2pix(nside, lon, lat, nest=True),
                              oracle.ang2pix_nest(nside, lon, lat)), ("nest", nside)
    # healpy docstring examples (theta, phi)
    assert hp.ang2pix(16, np.pi / 2, 0, lonlat=False) == 1440
    assert list(hp.ang2pix(16, [np.pi / 2, np.pi / 4, np.pi / 2, 0, np.pi],
                           [0., np.pi / 4, np.pi / 2, 0, 0], lonlat=False)) == [1440, 427, 1520, 0, 3068]


# ------------------------------------------------------------------ fused path
def _fused_engine(cfg_name="pal5", compl_band="r", nsides=(128, 512), **kw):
    from stream_sim_b200.engine import StarEngine
    from stream_sim_b200.frames import GreatCircleFrame
    from stream_sim_b200.model import StreamModel
    from stream_sim_b200.surveys import Survey
    from stream_sim_b200.synthetic import NOTEBOOK_ENDPOINTS
    sv = Survey.synthetic(*nsides, completeness_band=compl_band,
                          **(dict(device="cuda") if max(nsides) >= 2048 else {}))
    frame = GreatCircleFrame.from_endpoints(*NOTEBOOK_ENDPOINTS)
    model = StreamModel(stream_config(cfg_name))
    return StarEngine(stream=model, survey=sv, frame=frame, **kw), frame


def _oracle_fused(cfg_name, draws, compl_band="r", nsides=(128, 512), **kw):
    from stream_sim_b200.synthetic import NOTEBOOK_ENDPOINTS, isochrone_table
    R = oracle.frame_from_endpoints(*NOTEBOOK_ENDPOINTS[0], *NOTEBOOK_ENDPOINTS[1])
    sv = oracle_survey(compl_band, nside_maglim=nsides[0], nside_ebv=nsides[1],
                       lazy=max(nsides) >= 2048)
    return oracle.generate_observe(stream_config(cfg_name), isochrone_table(), R, sv, draws, **kw)


def _normals_close(dev, emu, what=""):
    """Device normals (MUFU log2 / sqrt / sin / cos) against the NumPy float32 emulation: the
    two agree to ~1e-6; the absolute error of the MUFU log grows to ~1e-4 only for the few
    draws whose radius is nearly zero."""
    d = np.abs(dev - emu)
    assert np.median(d) < 1e-6 and np.quantile(d, 0.999) < 2e-5 and d.max() < 2e-3, (what, d.max())


def _check_draws_against_oracle(used, seed, offset, n):
    """The draws exported by the kernel vs the oracle's derivation: uniforms are exact,
    the fp32 Box-Muller normals agree with the NumPy float32 emulation to ~1e-6."""
    ref = oracle.device_draws(seed, offset, n)
    for k in ("u_phi1", "u_mass", "u_det", "u_det2"):
        assert np.array_equal(used[k], ref[k]), k
    for k in ("z_r", "z_g"):
        _normals_close(used[k], ref[k], k)
    return ref


GRID_CONFIGS = ("pal5", "atlas_spline", "phoenix_spline")   # inverse-CDF densities


def _assert_same_bytes(torch, a, b, keys=None):
    for k in (keys or a.keys()):
        assert torch.equal(a[k].view(torch.uint8), b[k].view(torch.uint8)), k


def _check_fast_against_f64(torch, fast, exact):
    """SS_MATH_FAST: every decision and every position / true magnitude identical to the
    double path, magerr and mag_obs within 1e-5 relative (the north star's fp64 bar)."""
    _assert_same_bytes(torch, fast, exact, [k for k in exact if k in (
        "pix", "flag_observed", "flag_perfect_galstarsep", "phi1", "phi2", "dist", "ra", "dec",
        "mag_g", "mag_r")])
    for k in ("mag_g_obs", "mag_r_obs", "magerr_g", "magerr_r"):
        assert_close(fast[k].cpu().numpy(), exact[k].cpu().numpy(), 1e-5, 0, k)


@pytest.mark.parametrize("cfg_name,grid_table", [("toy1", True), ("pal5", True), ("pal5", False),
                                                 ("atlas_spline", True), ("phoenix_spline", True)])
def test_fused_free_running_matches_oracle(torch_cuda, cfg_name, grid_table):
    """Free-running (Philox on the device) fused generate->rotate->observe equals the
    oracle fed with the draws the kernel used (exported through SSOut.draws_out, and
    themselves checked against the oracle's Philox derivation): pixel indices and
    flags bit-exact, everything else to 1e-9.  The production kernel (k_fused), which
    cannot export draws, must produce the very same bytes as the generic kernel, with and
    without the packed survey maps; its SS_MATH_FAST variant the same decisions."""
    n, seed, offset = 1 << 18, 12345, 1000
    eng, _ = _fused_engine(cfg_name, perfect_galstarsep=True, grid_table=grid_table)
    from stream_sim_b200 import _lib
    assert (eng.params.density_kind == _lib.DENSITY_GRID) == (grid_table and cfg_name in GRID_CONFIGS)
    cnt = eng.new_counters()
    out = eng.generate_observe(n, seed=seed, offset=offset, counters=cnt, pix=True,
                               export_draws=True)
    got = {k: v.cpu().numpy() for k, v in out.items()}
    draws = eng.exported_draws(out)
    odraws = _check_draws_against_oracle(draws, seed, offset, n)
    if cfg_name == "toy1":          # Gaussian track, Uniform distance modulus
        _normals_close(draws["z_phi2"], odraws["z_phi2"], "z_phi2")
        assert np.array_equal(draws["u_dist"], odraws["u_dist"])
    lean = eng.generate_observe(n, seed=seed, offset=offset, pix=True)
    _assert_same_bytes(torch_cuda, lean, out, lean.keys())
    if eng.params.density_kind == _lib.DENSITY_GRID:
        # (maps of different nside: the production kernel gathers the three maps separately)
        assert not eng.maps.packed
        generic, _ = _fused_engine(cfg_name, perfect_galstarsep=True, fused_kernel=False)
        _assert_same_bytes(torch_cuda, generic.generate_observe(n, seed=seed, offset=offset, pix=True), lean)
        fast, _ = _fused_engine(cfg_name, perfect_galstarsep=True, math="fast")
        _check_fast_against_f64(torch_cuda, fast.generate_observe(n, seed=seed, offset=offset, pix=True), lean)
    ref = _oracle_fused(cfg_name, draws, perfect_galstarsep=True)
    assert np.array_equal(got["phi1"], ref["phi1"])
    assert np.array_equal(got["pix"], ref["pix"]), "pixel indices"
    assert np.array_equal(got["flag_observed"].astype(bool), ref["flag_observed"])
    assert np.array_equal(got["flag_perfect_galstarsep"].astype(bool), ref["flag_perfect_galstarsep"])
    for k in ("phi2", "dist", "mag_g", "mag_r", "ra", "dec", "mag_g_obs", "mag_r_obs",
              "magerr_g", "magerr_r"):
        assert_close(got[k], ref[k], RTOL, 1e-9, k)
    s = eng.summarize(cnt)
    assert s["n_total"] == n
    assert s["n_observed"] == int(ref["flag_observed"].sum())
    assert s["n_perfect_galstarsep"] == int(ref["flag_perfect_galstarsep"].sum())
    assert s["n_badmag_r"] == int(np.isnan(ref["mag_r_obs"]).sum())
    assert s["n_badmag_g"] == int(np.isnan(ref["mag_g_obs"]).sum())


def test_device_normals_are_standard_normal(torch_cuda):
    """The fp32 Box-Muller draws: KS against the exact normal CDF at 2 million draws per
    slot, moments, independence of the pairs, and a populated 4-sigma tail."""
    from scipy import stats
    n = 1 << 21
    eng, _ = _fused_engine("toy1")            # Gaussian track: z_phi2 is a normal
    out = eng.generate_observe(n, seed=99, export_draws=True, columns=["flag_observed"])
    d = eng.exported_draws(out)
    for k in ("z_phi2", "z_r", "z_g"):
        z = d[k]
        assert stats.kstest(z, "norm").pvalue > 1e-3, k
        assert abs(z.mean()) < 4 / np.sqrt(n) and abs(z.var() - 1) < 6 * np.sqrt(2 / n), k
        assert abs(stats.kurtosis(z)) < 0.02, k
        tail = (np.abs(z) > 4).mean()
        assert 0.5 * 6.33e-5 < tail < 1.6 * 6.33e-5, (k, tail)
    assert abs(np.corrcoef(d["z_r"], d["z_g"])[0, 1]) < 4 / np.sqrt(n)
    assert abs(np.corrcoef(d["z_phi2"], d["z_r"])[0, 1]) < 4 / np.sqrt(n)
    for k in ("u_phi1", "u_mass", "u_det"):
        assert stats.kstest(d[k], "uniform").pvalue > 1e-3, k


def test_double_precision_normals_switch(torch_cuda):
    """StarEngine(normals="f64") -- the four normal draws of a star from a double-precision
    Box-Muller of the same Philox words (SSParams.normals_f64): the exported draws equal the
    oracle's NumPy emulation to a few ulp, production and generic kernel agree byte for byte,
    the oracle fed with those draws reproduces the run, and the uniforms are untouched."""
    from scipy import stats
    n, seed = 1 << 18, 31
    eng, _ = _fused_engine("toy1", normals="f64")        # Gaussian track: z_phi2 is a normal
    out = eng.generate_observe(n, seed=seed, pix=True, export_draws=True)   # generic kernel
    fused = eng.generate_observe(n, seed=seed, pix=True)                      # production kernel
    for k, v in fused.items():
        assert torch_cuda.equal(v.view(torch_cuda.uint8), out[k].view(torch_cuda.uint8)), k
    d = eng.exported_draws(out)
    want = oracle.device_draws(seed, 0, n, normals="f64")
    for k in ("z_phi2", "z_r", "z_g"):
        assert np.abs(d[k] - want[k]).max() < 1e-13, k
        assert stats.kstest(d[k], "norm").pvalue > 1e-3, k
    f32 = oracle.device_draws(seed, 0, n)
    assert np.abs(want["z_r"] - f32["z_r"]).max() > 1e-9          # it is a different draw ...
    assert np.abs(want["z_r"] - f32["z_r"]).max() < 1e-2          # ... of the same pair of words
    for k in ("u_phi1", "u_mass", "u_det", "u_det2"):
        assert np.array_equal(d[k], want[k]), k
    ref = _oracle_fused("toy1", d)
    got = {k: v.cpu().numpy() for k, v in out.items() if k != "draws"}
    assert np.array_equal(got["pix"], ref["pix"])
    assert np.array_equal(got["flag_observed"].astype(bool), ref["flag_observed"])
    for k in ("phi2", "dist", "mag_g_obs", "mag_r_obs", "magerr_r"):
        assert np.allclose(got[k], ref[k], rtol=1e-9, atol=1e-9, equal_nan=True), k


@pytest.mark.parametrize("cfg_name", ["gap", "tails"])
def test_word_table_edge_densities(torch_cuda, cfg_name):
    """Densities that stress the 32-bit word tables of the production kernel.  "gap": an almost
    empty stretch -- tens of thousands of index thresholds of the phi1 sampler fall into one
    bucket (count field saturated, binary search over word thresholds).  "tails": zero density
    at both ends -- a quarter of the thresholds sit at u = 0 and are folded into the first
    bucket's step, a quarter are never reached.  Production kernel = generic kernel (double
    thresholds) = oracle."""
    from stream_sim_b200.engine import WLUT_K_CAP
    n, seed = 1 << 21, 9
    eng, _ = _fused_engine(cfg_name)
    assert int(eng.params.grid_wlut_n) > 0
    lut = eng.grid_wlut.cpu().numpy().view(np.uint32).reshape(-1, 4)
    if cfg_name == "gap":
        assert (lut[:, 0] >> 20).max() == WLUT_K_CAP
    else:
        assert (lut[0, 0] & 0xFFFFF) > 20000                 # thresholds at or below zero
    out = eng.generate_observe(n, seed=seed, pix=True, export_draws=True)      # generic kernel
    fused = eng.generate_observe(n, seed=seed, pix=True)                         # production kernel
    _assert_same_bytes(torch_cuda, fused, out, fused.keys())
    got = {k_: v.cpu().numpy() for k_, v in out.items() if k_ != "draws"}
    ref = _oracle_fused(cfg_name, eng.exported_draws(out))
    assert np.array_equal(got["phi1"], ref["phi1"])
    if cfg_name == "gap":
        inside = np.abs(got["phi1"]) < 8.9
        assert 0 < inside.sum() < 1e-3 * n                 # a few stars do land in the stretch
    else:
        assert np.abs(got["phi1"]).max() <= 5.0
    assert np.array_equal(got["pix"], ref["pix"])
    assert np.array_equal(got["flag_observed"].astype(bool), ref["flag_observed"])


def test_whole_sky_matches_oracle(torch_cuda):
    """The production kernel over the whole sphere (the bench's stress workload): both polar caps
    of the RING scheme in its 32-bit pixel arithmetic, every base face, stars within 0.8 degrees
    of a pole (the reference's own angle chain), transverse angles far beyond the Taylor range
    of the small-angle sin/cos -- pix and flags bit-exact against the oracle, columns to 1e-9,
    and the same bytes from the generic kernel."""
    n, seed = 1 << 21, 5
    eng, _ = _fused_engine("wholesky")
    out = eng.generate_observe(n, seed=seed, pix=True, export_draws=True)      # generic kernel
    fused = eng.generate_observe(n, seed=seed, pix=True)                         # production kernel
    _assert_same_bytes(torch_cuda, fused, out, fused.keys())
    got = {k: v.cpu().numpy() for k, v in out.items() if k != "draws"}
    ref = _oracle_fused("wholesky", eng.exported_draws(out))
    z = np.sin(np.radians(got["dec"]))
    assert (z > 2 / 3).mean() > 0.1 and (z < -2 / 3).mean() > 0.1 and (np.abs(z) >= 0.9999).sum() > 50
    assert np.abs(got["phi2"]).max() > 90.0
    assert np.array_equal(got["pix"], ref["pix"]), "pixel indices"
    npix = 12 * int(eng.params.nside_pix) ** 2
    assert got["pix"].min() < 0.001 * npix and got["pix"].max() > 0.999 * npix   # cap to cap
    assert np.array_equal(got["flag_observed"].astype(bool), ref["flag_observed"])
    for k in ("phi2", "dist", "mag_g", "mag_r", "ra", "dec", "mag_g_obs", "mag_r_obs",
              "magerr_g", "magerr_r"):
        assert_close(got[k], ref[k], RTOL, 1e-9, k)


def test_fp32_map_option(torch_cuda):
    """StarEngine(map_dtype="f32") -- a labelled non-parity option: the survey maps rounded to
    fp32 and gathered as 16-byte packed records.  What it computes is the reference's chain on
    the ROUNDED maps: production and generic kernel agree byte for byte, the oracle fed with the
    rounded maps reproduces pix, both flags and the columns, and against the float64 maps only
    a handful of flags move (stars within ~1e-7 mag of a decision)."""
    n, seed = 1 << 19, 17
    nsides = (512, 512)                                   # one nside: the packed form applies
    eng, _ = _fused_engine("pal5", nsides=nsides, map_dtype="f32")
    assert eng.maps.packed32 and not eng.maps.packed
    out = eng.generate_observe(n, seed=seed, pix=True, export_draws=True)     # generic kernel
    fused = eng.generate_observe(n, seed=seed, pix=True)                        # production kernel
    for k, v in fused.items():
        assert torch_cuda.equal(v.view(torch_cuda.uint8), out[k].view(torch_cuda.uint8)), k
    from stream_sim_b200.synthetic import NOTEBOOK_ENDPOINTS, isochrone_table
    R = oracle.frame_from_endpoints(*NOTEBOOK_ENDPOINTS[0], *NOTEBOOK_ENDPOINTS[1])
    sv = oracle_survey("r", nside_maglim=nsides[0], nside_ebv=nsides[1])
    rnd = lambda m: np.asarray(m, dtype=np.float64).astype(np.float32).astype(np.float64)
    sv32 = dict(sv, ebv_map=rnd(sv["ebv_map"]), maglim_maps={b: rnd(m) for b, m in sv["maglim_maps"].items()})
    d = eng.exported_draws(out)
    got = {k: v.cpu().numpy() for k, v in out.items() if k != "draws"}
    ref = oracle.generate_observe(stream_config("pal5"), isochrone_table(), R, sv32, d)
    assert np.array_equal(got["pix"], ref["pix"])
    assert np.array_equal(got["flag_observed"].astype(bool), ref["flag_observed"])
    for k in ("mag_g_obs", "mag_r_obs", "magerr_g", "magerr_r"):
        assert np.allclose(got[k], ref[k], rtol=1e-9, atol=1e-9, equal_nan=True), k
    ref64 = oracle.generate_observe(stream_config("pal5"), isochrone_table(), R, sv, d)
    moved = np.count_nonzero(got["flag_observed"].astype(bool) != ref64["flag_observed"])
    assert moved <= 5, moved
    ok = np.isfinite(ref64["magerr_r"]) & np.isfinite(got["magerr_r"])
    assert np.abs(got["magerr_r"][ok] / ref64["magerr_r"][ok] - 1).max() < 1e-5


def test_shard_invariance_and_offsets(torch_cuda):
    """Any split of the global index range gives identical per-star bytes."""
    n = 100003
    eng, _ = _fused_engine("pal5")
    whole = eng.generate_observe(n, seed=99)
    parts = [eng.generate_observe(hi - lo, seed=99, offset=lo)
             for lo, hi in ((0, 1), (1, 33334), (33334, 70000), (70000, n))]
    for k, v in whole.items():
        cat = torch_cuda.cat([p[k] for p in parts])
        assert torch_cuda.equal(cat.view(torch_cuda.uint8), v.view(torch_cuda.uint8)), k


def test_fp32_output_mode(torch_cuda):
    n = 1 << 16
    eng, _ = _fused_engine("pal5")
    a = eng.generate_observe(n, seed=5, dtype="f64")
    b = eng.generate_observe(n, seed=5, dtype="f32")
    assert torch_cuda.equal(a["flag_observed"], b["flag_observed"])
    for k in ("mag_g", "mag_r", "mag_g_obs", "mag_r_obs", "magerr_g", "magerr_r", "dist"):
        x, y = a[k].cpu().numpy(), b[k].cpu().numpy().astype(float)
        assert_close(y, x, 0, 1e-3, k)           # 1e-3 mag in the fp32 mode
    for k in ("phi1", "phi2", "ra", "dec"):
        assert_close(b[k].cpu().numpy().astype(float), a[k].cpu().numpy(), 1e-6, 1e-6, k)


def test_fast_math_mode(torch_cuda):
    """SS_MATH_FAST under injected draws (the generic kernel) against the oracle: pixel
    indices and both flags bit-exact, positions and true magnitudes 1e-9, noisy magnitudes
    and errors 1e-5 relative (the north star's bar for fp64 outputs); and the production
    kernel free-running: same decisions as its double twin, star by star."""
    n, seed = 1 << 20, 4242
    exact, _ = _fused_engine("pal5", perfect_galstarsep=True)
    fast, _ = _fused_engine("pal5", perfect_galstarsep=True, math="fast")
    d = oracle.device_draws(seed, 0, n)
    inj = {k: d[k] for k in ("u_phi1", "z_phi2", "u_dist", "u_mass", "z_r", "z_g", "u_det", "u_det2")}
    a = exact.generate_observe(n, draws=inj, pix=True)
    b = fast.generate_observe(n, draws=inj, pix=True)
    _check_fast_against_f64(torch_cuda, b, a)
    ref = _oracle_fused("pal5", d, perfect_galstarsep=True)
    assert np.array_equal(b["pix"].cpu().numpy(), ref["pix"])
    for flag in ("flag_observed", "flag_perfect_galstarsep"):
        assert np.array_equal(b[flag].cpu().numpy().astype(bool), ref[flag]), flag
    for k in ("mag_g_obs", "mag_r_obs", "magerr_g", "magerr_r"):
        assert_close(b[k].cpu().numpy(), ref[k], 1e-5, 0, k)
    # free-running, production kernel, counters included
    ca, cb = exact.new_counters(), fast.new_counters()
    fa = exact.generate_observe(n, seed=seed, counters=ca, pix=True)
    fb = fast.generate_observe(n, seed=seed, counters=cb, pix=True)
    _check_fast_against_f64(torch_cuda, fb, fa)
    assert torch_cuda.equal(ca, cb)
    # the observe-only entry in fast mode
    cat = {k: fa[k] for k in ("ra", "dec", "mag_g", "mag_r")}
    oa, ob = exact.observe(cat, seed=seed), fast.observe(cat, seed=seed)
    assert torch_cuda.equal(oa["flag_observed"], ob["flag_observed"])
    assert_close(ob["mag_r_obs"].cpu().numpy(), oa["mag_r_obs"].cpu().numpy(), 1e-5, 0, "mag_r_obs")


def test_lsst_configuration_star_by_star(torch_cuda):
    """BASELINE config 4 as configured: nside-4096 depth and E(B-V) maps (built on the
    device; the oracle evaluates the same closed-form maps at the pixels it needs), 2^22
    stars, fused path against the oracle star by star -- pixels and flags bit-exact, the ten
    float columns to 1e-9 -- and the production kernel (packed maps: one 32-byte record per
    pixel) byte-equal to the generic kernel on separate maps."""
    n, seed = 1 << 22, 777
    eng, _ = _fused_engine("pal5", nsides=(4096, 4096), perfect_galstarsep=True)
    assert eng.maps.packed and eng.maps.nside_packed == 4096
    out = eng.generate_observe(n, seed=seed, pix=True, export_draws=True)      # generic kernel
    prod = eng.generate_observe(n, seed=seed, pix=True)                        # k_fused, packed
    _assert_same_bytes(torch_cuda, prod, out, prod.keys())
    plain, _ = _fused_engine("pal5", nsides=(4096, 4096), perfect_galstarsep=True, packed_maps=False)
    assert not plain.maps.packed
    _assert_same_bytes(torch_cuda, plain.generate_observe(n, seed=seed, pix=True), prod)
    del plain
    got = {k: v.cpu().numpy() for k, v in out.items()}
    draws = eng.exported_draws(out)
    _check_draws_against_oracle(draws, seed, 0, n)
    ref = _oracle_fused("pal5", draws, nsides=(4096, 4096), perfect_galstarsep=True)
    assert np.array_equal(got["pix"], ref["pix"]), "pixel indices"
    assert np.array_equal(got["flag_observed"].astype(bool), ref["flag_observed"])
    assert np.array_equal(got["flag_perfect_galstarsep"].astype(bool), ref["flag_perfect_galstarsep"])
    for k in ("phi2", "dist", "mag_g", "mag_r", "ra", "dec", "mag_g_obs", "mag_r_obs",
              "magerr_g", "magerr_r"):
        assert_close(got[k], ref[k], RTOL, 1e-9, k)
    assert 0.15 < ref["flag_observed"].mean() < 0.3 and np.isnan(ref["magerr_r"]).mean() > 0.01


def test_host_buffer_entry_equals_device_entry(torch_cuda):
    n, chunk = 300000, 1 << 16
    eng, _ = _fused_engine("pal5")
    dev = eng.generate_observe(n, seed=11)
    host = eng.allocate_host(n, list(dev.keys()))
    eng.generate_observe_host(n, host, eng.host_workspace(chunk), chunk, seed=11)
    for k, v in dev.items():
        assert torch_cuda.equal(v.cpu().view(torch_cuda.uint8), host[k].view(torch_cuda.uint8)), k
    assert int(host["counters"][0]) == n


def test_observe_host_entry(torch_cuda):
    n, chunk = 200000, 1 << 15
    eng, _ = _fused_engine("pal5")
    full = eng.generate_observe(n, seed=21)
    cat = {k: full[k] for k in ("ra", "dec", "mag_g", "mag_r")}
    dev = eng.observe(cat, seed=21)
    cat_h = {k: v.cpu().pin_memory() for k, v in cat.items()}
    host = eng.allocate_host(n, list(dev.keys()))
    eng.observe_host(cat_h, host, eng.host_workspace(chunk), chunk, seed=21)
    for k, v in dev.items():
        assert torch_cuda.equal(v.cpu().view(torch_cuda.uint8), host[k].view(torch_cuda.uint8)), k
    # and observe-only on the fused path's own (ra, dec, mags) reproduces its flags
    assert torch_cuda.equal(dev["flag_observed"], full["flag_observed"])


def test_summary_only_mode_equals_full_run(torch_cuda):
    """Counters/histograms-only streaming (no per-star output) gives the same
    summary as a full run, whatever the chunking."""
    n = (1 << 21) + 12345
    eng, _ = _fused_engine("pal5", hist=True)
    full = eng.new_counters()
    out = eng.generate_observe(n, seed=8, counters=full)
    chunked = eng.run_summary(n, seed=8, chunk=300000)
    assert torch_cuda.equal(full, chunked)
    s = eng.summarize(full)
    assert s["n_total"] == n and s["n_observed"] == int(out["flag_observed"].sum())
    for name in ("phi1", "phi2", "dist", "mag_g", "g_minus_r"):
        assert 0 < s["hist_" + name].sum() <= n
    assert s["hist_phi1"].sum() == n          # pal5 phi1 lies inside the default range
    # the five histograms against their definition (fp32 bins, SSParams.hist_*) on the columns
    cols = {k: out[k].cpu().numpy() for k in ("phi1", "phi2", "dist", "mag_g", "mag_r")}
    cols["g_minus_r"] = cols["mag_g"] - cols["mag_r"]
    for name in ("phi1", "phi2", "dist", "mag_g", "g_minus_r"):
        bins = eng.hist_bin(name, cols[name])
        assert np.array_equal(s["hist_" + name], np.bincount(bins[bins >= 0], minlength=256)), name
    # ... which is the plain histogram up to fp32 rounding at the bin edges
    lo, hi = eng.hist_ranges["phi1"]
    ref = np.histogram(cols["phi1"], bins=256, range=(lo, hi))[0]
    assert np.abs(s["hist_phi1"] - ref).sum() <= 1e-4 * n
    # and the generic kernel fills the same histograms
    generic, _ = _fused_engine("pal5", hist=True, fused_kernel=False)
    assert torch_cuda.equal(generic.run_summary(n, seed=8, chunk=500000), full)


def test_full_size_properties(torch_cuda):
    """BASELINE config 4 size (1e8 stars, nside-4096 maps generated on the device):
    size-independent properties instead of a star-by-star oracle run --
    idempotence (same seed, same counters), additivity over any sharding of the index
    range (the multi-GPU contract), conservation (every star lands in the phi1
    histogram; observed <= total), and agreement of the observed fraction with a
    2^20-star launch that IS checked star by star elsewhere."""
    from stream_sim_b200.engine import StarEngine
    from stream_sim_b200.frames import GreatCircleFrame
    from stream_sim_b200.model import StreamModel
    from stream_sim_b200.surveys import Survey
    from stream_sim_b200.synthetic import NOTEBOOK_ENDPOINTS
    n, seed = 100_000_000, 2026
    sv = Survey.synthetic(4096, 4096, device="cuda")
    eng = StarEngine(stream=StreamModel(stream_config("pal5")), survey=sv,
                     frame=GreatCircleFrame.from_endpoints(*NOTEBOOK_ENDPOINTS), hist=True)
    whole = eng.run_summary(n, seed=seed, chunk=n)
    again = eng.run_summary(n, seed=seed, chunk=n)
    assert torch_cuda.equal(whole, again), "idempotence"
    shards = eng.new_counters()
    cuts = [0, 1, 12_500_000, 37_500_001, 50_000_000, 99_999_999, n]
    for lo, hi in zip(cuts, cuts[1:]):
        eng.generate_observe(hi - lo, seed=seed, offset=lo, counters=shards, out={})
    assert torch_cuda.equal(whole, shards), "additivity over shards"
    s = eng.summarize(whole)
    assert s["n_total"] == n and s["hist_phi1"].sum() == n
    assert 0 < s["n_observed"] < n and s["n_domain_error"] == 0 and s["n_bad_coord"] == 0
    small = eng.summarize(eng.run_summary(1 << 20, seed=seed))
    f_big, f_small = s["n_observed"] / n, small["n_observed"] / (1 << 20)
    assert abs(f_big - f_small) < 5 * np.sqrt(f_small * (1 - f_small) / (1 << 20))
    # unseen pixels: 2 % of the map, so ~2 % of the stars
    assert abs(s["n_unseen"] / n - 0.02) < 0.004

    # BASELINE config 5 (synthetic scaling sweep, 1e9-1e11 stars): 1e10 stars streamed as
    # counters + histograms only -- in ONE launch (64-bit star indices, 32-bit per-CTA
    # accumulators), in 2^28-star launches, and as the 8 index ranges 8 GPUs would take
    # (offset + i runs past 2^32 on 7 of them): all three byte-identical.
    from stream_sim_b200.distributed import shard_range
    big = 10_000_000_000
    one = eng.run_summary(big, seed=seed, chunk=big)
    chunks = eng.run_summary(big, seed=seed)
    ranks = eng.new_counters()
    for r in range(8):
        lo, hi = shard_range(big, r, 8)
        eng.run_summary(hi - lo, seed=seed, offset=lo, chunk=hi - lo, counters=ranks)
    assert torch_cuda.equal(one, chunks) and torch_cuda.equal(one, ranks)
    sb = eng.summarize(one)
    assert sb["n_total"] == big and sb["hist_phi1"].sum() == big
    f_huge = sb["n_observed"] / big
    assert abs(f_huge - f_big) < 5 * np.sqrt(f_big * (1 - f_big) / n)


# ------------------------------------------------------------------ distributions
def _ks(sample, ref_sorted):
    sample = np.sort(sample)
    both = np.concatenate([sample, ref_sorted])
    c1 = np.searchsorted(sample, both, side="right") / len(sample)
    c2 = np.searchsorted(ref_sorted, both, side="right") / len(ref_sorted)
    d = np.abs(c1 - c2).max()
    ne = len(sample) * len(ref_sorted) / (len(sample) + len(ref_sorted))
    return d, 1.63 / np.sqrt(ne)        # 1 % critical value


@pytest.mark.parametrize("name", ["toy2", "pal5"])
def test_free_running_marginals_ks(torch_cuda, name, known_answers):
    """Free-running output matches the reference's distributions: KS test of
    phi1/phi2/dist against samples drawn by the reference itself
    (tests/golden/marginals.npz), and the notebook's describe() for Pal 5."""
    from conftest import load_golden
    from stream_sim_b200.model import StreamModel
    from stream_sim_b200.utils import parse_config
    marg = load_golden("marginals.npz")
    cfg = parse_config(f"config/{name}_config.yaml")["stream"]
    cfg.pop("isochrone", None)
    df = StreamModel(cfg).sample(200000, seed=2024)
    for k in ("phi1", "phi2", "dist"):
        key = f"{name}_{k}"
        if key not in marg:
            continue
        d, crit = _ks(df[k].to_numpy(dtype=float), marg[key].astype(float))
        assert d < crit, f"KS {key}: D={d:.4f} crit={crit:.4f}"
    if name == "pal5":
        ref = known_answers["pal5_describe"]
        assert abs(df["phi1"].mean() - ref["phi1"]["mean"]) < 0.05
        assert abs(df["phi1"].std() - ref["phi1"]["std"]) < 0.05
        assert abs(df["phi2"].mean() - ref["phi2"]["mean"]) < 0.01
        assert abs(df["dist"].mean() - ref["dist"]["mean"]) < 0.01


def test_free_running_magnitude_colour_ks(torch_cuda):
    """Magnitude and colour marginals of the free-running device path against the
    oracle driven by an independent NumPy RNG."""
    from stream_sim_b200.synthetic import isochrone_table
    n = 200000
    eng, _ = _fused_engine("pal5")
    out = eng.generate_observe(n, seed=31337)
    rng = np.random.default_rng(1)
    draws = dict(u_phi1=rng.uniform(size=n), z_phi2=rng.standard_normal(n),
                 u_dist=rng.uniform(size=n), u_mass=rng.uniform(size=n))
    ref = oracle.generate(stream_config("pal5"), draws, isochrone_table())
    g, r = out["mag_g"].cpu().numpy(), out["mag_r"].cpu().numpy()
    for a, b, what in ((g, ref["mag_g"], "mag_g"), (g - r, ref["mag_g"] - ref["mag_r"], "g-r")):
        d, crit = _ks(a, np.sort(b))
        assert d < crit, f"KS {what}: D={d:.4f} crit={crit:.4f}"


# ------------------------------------------------------------------ more configurations
def test_large_global_offsets(torch_cuda):
    """Star indices beyond 2^32 (the high word of the Philox counter) -- the range a
    1e11-star sharded run reaches."""
    n, seed, offset = 50000, 5, (1 << 35) + 12345
    eng, _ = _fused_engine("pal5")
    out = eng.generate_observe(n, seed=seed, offset=offset, pix=True, export_draws=True)
    draws = eng.exported_draws(out)
    _check_draws_against_oracle(draws, seed, offset, n)
    ref = _oracle_fused("pal5", draws)
    assert np.array_equal(out["pix"].cpu().numpy(), ref["pix"])
    assert np.array_equal(out["flag_observed"].cpu().numpy().astype(bool), ref["flag_observed"])
    assert_close(out["mag_r_obs"].cpu().numpy(), ref["mag_r_obs"], RTOL, 1e-9, "mag_r_obs")
    lo = eng.generate_observe(n, seed=seed, offset=12345)
    assert not torch_cuda.equal(lo["phi1"], out["phi1"])


def test_single_band_and_nested_maps(torch_cuda):
    """bands=['r'] only (reference observed.py:154-258: g is then ignored, including its
    SNR cut) and NESTED-ordered maps give the same answers as the oracle / the RING run."""
    from stream_sim_b200 import healpix as hp
    from stream_sim_b200.engine import StarEngine
    from stream_sim_b200.surveys import Survey
    n, seed = 100000, 17
    full, _ = _fused_engine("pal5")
    base = full.generate_observe(n, seed=seed)
    cat = {k: base[k] for k in ("ra", "dec", "mag_g", "mag_r")}
    sv = Survey.synthetic(128, 512)
    only_r = StarEngine(survey=sv, bands=("r",))
    got = only_r.observe(dict(ra=cat["ra"], dec=cat["dec"], mag_r=cat["mag_r"]), seed=seed,
                         export_draws=True)
    assert set(got) == {"mag_r_obs", "magerr_r", "flag_observed", "draws"}
    d = only_r.exported_draws(got)
    ref = oracle.observe(oracle_survey(), cat["ra"].cpu().numpy(), cat["dec"].cpu().numpy(),
                         cat["mag_g"].cpu().numpy(), cat["mag_r"].cpu().numpy(), d, bands=("r",))
    assert np.array_equal(got["flag_observed"].cpu().numpy().astype(bool), ref["flag_observed"])
    assert_close(got["mag_r_obs"].cpu().numpy(), ref["mag_r_obs"], RTOL, 1e-9, "mag_r_obs")
    assert int(got["flag_observed"].sum()) > int(base["flag_observed"].sum())   # no g cuts
    # NESTED maps
    nest_sv = Survey.synthetic(128, 512)
    nest_sv.ebv_map = hp.reorder_ring_to_nest(sv.ebv_map)
    nest_sv.maglim_maps = {b: hp.reorder_ring_to_nest(m) for b, m in sv.maglim_maps.items()}
    ring_out = StarEngine(survey=sv).observe(cat, seed=seed, pix=True)
    nest_out = StarEngine(survey=nest_sv, nest=True).observe(cat, seed=seed, pix=True)
    for k in ("mag_g_obs", "mag_r_obs", "magerr_g", "magerr_r", "flag_observed"):
        assert torch_cuda.equal(ring_out[k].view(torch_cuda.uint8), nest_out[k].view(torch_cuda.uint8)), k
    assert np.array_equal(hp.ring2nest(4096, ring_out["pix"].cpu().numpy()),
                          nest_out["pix"].cpu().numpy())


def test_bad_inputs_raise_like_the_reference(torch_cuda):
    import pandas as pd
    from stream_sim_b200.observed import StreamInjector
    from stream_sim_b200.surveys import Survey
    inj = StreamInjector(Survey.synthetic(16, 16))
    good = pd.DataFrame(dict(ra=[10.0, 20.0], dec=[-5.0, 5.0], mag_g=[22.0, 23.0], mag_r=[21.5, 22.5]))
    assert len(inj.inject(good, seed=1, verbose=False)) == 2
    empty = inj.inject(good.iloc[:0], seed=1, verbose=False)
    assert len(empty) == 0 and "flag_observed" in empty.columns
    bad = good.copy()
    bad.loc[0, "dec"] = 95.0                               # healpy: THETA out of range
    with pytest.raises(ValueError, match="THETA is out of range"):
        inj.inject(bad, seed=1, verbose=False)
    bad.loc[0, "dec"] = np.nan
    with pytest.raises(ValueError, match="THETA is out of range"):
        inj.inject(bad, seed=1, verbose=False)
    nanmag = good.copy()
    nanmag.loc[1, "mag_r"] = np.nan                        # propagates: BAD_MAG, not observed
    res = inj.inject(nanmag, seed=1, verbose=False)
    assert res["mag_r_obs"][1] == "BAD_MAG" and not res["flag_observed"][1]
    with pytest.raises(ValueError, match="survey must be a string or Survey instance"):
        StreamInjector(3)


# ------------------------------------------------------------------ edge cases
def test_empty_and_tiny_inputs(torch_cuda):
    eng, _ = _fused_engine("toy1")
    out = eng.generate_observe(0, seed=1)
    assert all(v.numel() == 0 for v in out.values())
    out = eng.generate_observe(1, seed=1)
    assert all(v.numel() == 1 for v in out.values())
    out = eng.generate_observe(513, seed=1)       # one full block + 1 ragged star
    ref = eng.generate_observe(513, seed=1)
    for k in out:
        assert torch_cuda.equal(out[k].view(torch_cuda.uint8), ref[k].view(torch_cuda.uint8))


def test_known_magerr_for_saturated_star(torch_cuda, known_answers):
    """Saturated stars get magerr = sqrt(10^2 + 0.005^2) = 10.000001 (the constant
    visible in the reference notebook outputs)."""
    from stream_sim_b200.surveys import Survey
    sv = Survey.synthetic(16, 16)
    e = sv.get_photo_error("r", np.array([14.0, 15.9]), np.array([26.0, 25.0]))
    assert np.allclose(e, known_answers["notebook_inject"]["saturated_magerr"], atol=1e-6)
    assert np.array_equal(e, np.sqrt(10.0 ** 2 + 0.005 ** 2) * np.ones(2))
    c = sv.get_completeness("r", np.array([14.0, 20.0, 40.0]), np.array([26.0, np.nan, 26.0]))
    assert np.array_equal(c, [0.0, 0.0, 0.0])
