"""GPU tests of the rows next to the hot path: great-circle
frame search, partial catalog completion, the standalone sampler / getter entry
points, and the CLI driver."""
import numpy as np
import pandas as pd
import pytest

import streamsim_oracle as oracle
from helpers import assert_close, oracle_survey, stream_config

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def survey():
    from stream_sim_b200.surveys import Survey
    return Survey.synthetic(128, 512)


def _cap_mask(nside=64):
    """Southern cap dec < -35 deg as a boolean RING mask."""
    rng = np.random.default_rng(0)
    lon = rng.uniform(0, 360, 3_000_000)
    lat = np.degrees(np.arcsin(rng.uniform(-1, np.sin(np.radians(-35.0)), 3_000_000)))
    mask = np.zeros(12 * nside * nside, dtype=bool)
    mask[oracle.ang2pix_ring(nside, lon, lat)] = True
    return mask


def test_frame_search_matches_oracle(survey):
    """Same PCG64 stream -> same accepted trial, same frame, same sky positions."""
    from stream_sim_b200.observed import StreamInjector
    rng = np.random.default_rng(5)
    phi1, phi2 = rng.uniform(-8, 8, 20000), rng.normal(0, 0.3, 20000)
    mask = _cap_mask()
    ref = oracle.find_frame(np.random.default_rng(42), phi1, phi2, mask)
    assert ref is not None and ref[2] > 1, "the search should need several trials"
    inj = StreamInjector(survey)
    sky = inj.phi_to_radec(phi1, phi2, seed=42, mask=mask, verbose=False)
    np.testing.assert_allclose(inj._last_gc_frame.R, ref[0], atol=1e-15)
    ra, dec = oracle.rotate(phi1, phi2, ref[0])
    assert_close(sky.icrs.ra.deg, ra, 1e-12, 1e-10, "ra")
    assert_close(sky.icrs.dec.deg, dec, 1e-12, 1e-10, "dec")
    # 'last' reuses the frame; an explicit 3x3 matrix is accepted too
    again = inj.phi_to_radec(phi1[:10], phi2[:10], gc_frame="last")
    assert np.array_equal(again.icrs.ra.deg, sky.icrs.ra.deg[:10])
    # impossible threshold -> no frame
    assert inj._find_gc_frame(phi1=phi1, phi2=phi2, mask=np.zeros_like(mask), seed=1,
                              max_iter=5, verbose=False) is None
    with pytest.raises(ValueError, match="cannot be empty"):
        inj.phi_to_radec([], [])


def test_mask_building_and_cache(survey):
    from stream_sim_b200 import healpix as hp
    from stream_sim_b200.observed import StreamInjector
    from stream_sim_b200.surveys import SurveyFactory
    if survey.coverage is None:
        SurveyFactory._build_coverage_map(survey, verbose=False)
    inj = StreamInjector(survey)
    StreamInjector.clear_mask_cache()
    m = inj._create_mask(["footprint", "ebv"], verbose=False, ebv_threshold=0.08)
    assert m.dtype == bool and len(m) == hp.nside2npix(128)
    ebv128 = hp.ud_grade(np.asarray(survey.ebv_map), 128)
    expect = (np.asarray(survey.coverage) > 0.5) & (ebv128 < 0.08)
    assert np.array_equal(m, expect)
    assert inj._create_mask(["ebv", "footprint"], verbose=False, ebv_threshold=0.08) is m
    assert StreamInjector.list_cached_masks() == [("synthetic", ("ebv", "footprint"), 0.08)]
    assert inj._create_mask(None, verbose=False) is None
    with pytest.raises(ValueError, match="Unknown mask type"):
        inj._create_mask(["nope"], verbose=False)
    frac = inj._fraction_inside_mask(np.array([10.0, 200.0]), np.array([-30.0, 45.0]), m)
    assert 0.0 <= frac <= 1.0


def test_complete_catalog_partial_fill_matches_oracle():
    """Only missing (NaN) phi1/phi2/dist entries are sampled, with the draws of their
    own row index; magnitudes are sampled for every row (reference model.py:159-351)."""
    from stream_sim_b200.model import StreamModel
    from stream_sim_b200.synthetic import isochrone_table
    cfg = stream_config("pal5")
    n, seed = 5000, 77
    rng = np.random.default_rng(1)
    given1 = rng.uniform(-5, 10, n)
    cat = pd.DataFrame(dict(phi1=np.where(rng.uniform(size=n) < 0.5, given1, np.nan),
                            phi2=np.where(rng.uniform(size=n) < 0.3, 0.123, np.nan)))
    out = StreamModel(cfg).complete_catalog(cat, seed=seed, verbose=False)
    assert list(out.columns) == ["phi1", "phi2", "dist", "mag_g", "mag_r"]
    keep1, keep2 = cat["phi1"].notna().to_numpy(), cat["phi2"].notna().to_numpy()
    assert np.array_equal(out["phi1"].to_numpy()[keep1], cat["phi1"].to_numpy()[keep1])
    assert np.all(out["phi2"].to_numpy()[keep2] == 0.123)
    assert not out.isna().any().any()
    d = oracle.device_draws(seed, 0, n)
    ref = oracle.generate(cfg, d, isochrone_table())
    phi1 = np.where(keep1, cat["phi1"].to_numpy(), ref["phi1"])
    assert np.array_equal(out["phi1"].to_numpy(), phi1)
    phi2_ref = oracle.sample_track(cfg["track"], phi1, z=d["z_phi2"])
    # z_phi2 is the fp32 Box-Muller draw: the oracle's float32 emulation agrees to ~1e-6
    assert_close(out["phi2"].to_numpy()[~keep2], phi2_ref[~keep2], 0, 1e-5, "phi2")
    dist_ref = oracle.sample_track(cfg["distance_modulus"], phi1, u=d["u_dist"])
    assert_close(out["dist"].to_numpy(), dist_ref, 1e-9, 0, "dist")
    mg, mr = oracle.sample_mags(isochrone_table(), d["u_mass"], dist_ref)
    assert_close(out["mag_g"].to_numpy(), mg, 1e-9, 0, "mag_g")
    assert_close(out["mag_r"].to_numpy(), mr, 1e-9, 0, "mag_r")
    assert cat["phi1"].isna().any(), "input is not modified"
    # error behaviour of the reference
    m = StreamModel(cfg)
    with pytest.raises(ValueError, match="phi1 required to sample phi2"):
        m.complete_catalog(pd.DataFrame(dict(x=[1.0, 2.0])), columns_to_add=["phi2"], verbose=False)
    with pytest.raises(ValueError, match="size must be provided"):
        m.complete_catalog(None)
    with pytest.raises(TypeError):
        m.complete_catalog(3.0)
    fresh = m.complete_catalog(None, size=100, verbose=False, seed=1)
    assert len(fresh) == 100 and not fresh.isna().any().any()
    ren = m.complete_catalog(pd.DataFrame(dict(phi_1=[0.0, 1.0], distance=[16.0, 17.0])),
                             columns_to_add=["mag_g", "mag_r"], verbose=False, seed=2)
    assert {"phi1", "dist", "mag_g", "mag_r"} <= set(ren.columns)
    assert np.array_equal(ren["dist"].to_numpy(), [16.0, 17.0])


def test_inject_from_stream_coordinates_only(survey):
    """(phi1, phi2) in, everything else sampled: frame, sky position, distance,
    magnitudes, observation -- reference notebook cell 14 flow."""
    from stream_sim_b200.frames import GreatCircleFrame
    from stream_sim_b200.observed import StreamInjector
    from stream_sim_b200.synthetic import NOTEBOOK_ENDPOINTS
    rng = np.random.default_rng(42)
    data = pd.DataFrame(dict(phi1=rng.uniform(-5, 5, 4000), phi2=rng.uniform(-1, 1, 4000)))
    scfg = dict(isochrone=dict(name="synthetic"),
                distance_modulus=dict(center=dict(type="Constant", value=16.5),
                                      spread=dict(type="Constant", value=0.0)))
    inj = StreamInjector(survey)
    frame = GreatCircleFrame.from_endpoints(*NOTEBOOK_ENDPOINTS)
    res = inj.inject(data, seed=42, gc_frame=frame, stream_config=scfg, verbose=False,
                     perfect_galstarsep=True)
    assert list(res.columns) == ["phi1", "phi2", "ra", "dec", "dist", "mag_g", "mag_r",
                                 "mag_r_obs", "magerr_r", "mag_g_obs", "magerr_g",
                                 "flag_observed", "flag_perfect_galstarsep"]
    R = oracle.frame_from_endpoints(*NOTEBOOK_ENDPOINTS[0], *NOTEBOOK_ENDPOINTS[1])
    ra, dec = oracle.rotate(data["phi1"].to_numpy(), data["phi2"].to_numpy(), R)
    assert_close(res["ra"].to_numpy(), ra, 1e-12, 1e-10, "ra")
    assert_close(res["dec"].to_numpy(), dec, 1e-12, 1e-10, "dec")
    assert np.all(res["dist"] == 16.5)
    d = oracle.device_draws(42, 0, len(data))
    ref = oracle.observe(oracle_survey(), ra, dec, res["mag_g"].to_numpy(),
                         res["mag_r"].to_numpy(), d, perfect_galstarsep=True)
    # the flux normals of the oracle are a float32 emulation (~1e-6) of the device's:
    # a star whose noisy flux or SNR sits within that of a threshold may differ
    assert (res["flag_observed"].to_numpy() != ref["flag_observed"]).sum() <= 1
    assert (res["flag_perfect_galstarsep"].to_numpy() != ref["flag_perfect_galstarsep"]).sum() <= 1
    assert 0 < res["flag_observed"].sum() < len(res)
    # reference error messages
    with pytest.raises(ValueError, match="either \\(ra, dec\\) or \\(phi1, phi2\\)"):
        inj.inject(pd.DataFrame(dict(x=[1.0])), verbose=False)
    with pytest.raises(ValueError, match="`stream_config` must be provided"):
        inj.inject(data, gc_frame=frame, verbose=False)
    with pytest.raises(ValueError, match="only 'r' and 'g'"):
        inj.inject(res, bands=["r", "i"], verbose=False)
    with pytest.raises(ValueError, match="Detection flag requires 'r'"):
        inj.inject(res, bands=["g"], verbose=False)
    nan_bad = inj.inject(res[["ra", "dec", "mag_g", "mag_r"]], seed=3, verbose=False, bad_mag="nan")
    assert nan_bad["mag_r_obs"].dtype == np.float64


def test_standalone_samplers_and_models():
    from stream_sim_b200.model import DensityModel, IsochroneModel, TrackModel
    from stream_sim_b200.samplers import (GaussianSampler, UniformSampler,
                                          inverse_transform_sample, sampler_factory)
    n = 50000
    d = oracle.device_draws(9, 0, n)
    u = UniformSampler(-2.0, 3.0).sample(n, seed=9)
    assert np.array_equal(u, d["u_phi2"] * 5.0 + -2.0)           # U32 of Philox block 1, word 0
    mu, sig = np.linspace(0, 1, n), np.full(n, 0.5)
    g = GaussianSampler(mu, sig).sample(n, seed=9)
    dg = np.abs(g - (d["z_phi2"] * sig + mu))        # MUFU normals vs their NumPy emulation
    assert np.quantile(dg, 0.999) < 1e-5 and dg.max() < 1e-3, "gaussian loc/scale arrays"
    with pytest.raises(ValueError, match="Domain error"):
        GaussianSampler(0.0, -1.0).sample(10)
    vals = np.linspace(-3, 3, 1001)
    pdf = np.exp(-0.5 * vals ** 2)
    s = inverse_transform_sample(vals, pdf, n, seed=9)
    cdf = np.cumsum(pdf)
    cdf /= cdf[-1]
    idx = np.rint(np.where((d["u_phi1"] < cdf[0]), 0.0, np.interp(d["u_phi1"], cdf, np.arange(1001.0))))
    assert np.array_equal(s, vals[idx.astype(int)])
    dens = DensityModel(dict(type="Uniform", xmin=-9.0, xmax=9.0))
    assert np.array_equal(dens.sample(n, seed=9), d["u_phi1"] * 18.0 + -9.0)
    sin = sampler_factory("sinusoid", period=3.0, offset=1.0, xmin=-9.0, xmax=9.0)
    x = sin.sample(n, seed=9)
    assert np.array_equal(x, oracle.sample_phi1(dict(type="Sinusoid", period=3.0, offset=1.0,
                                                     xmin=-9.0, xmax=9.0), d["u_phi1"]))
    tcfg = dict(center=dict(type="Sinusoid", period=9.0, amplitude=0.75),
                spread=dict(type="Sinusoid", period=6.0, amplitude=0.2, offset=0.25))
    phi2 = TrackModel(tcfg).sample(x, seed=9)
    dt = np.abs(phi2 - oracle.sample_track(tcfg, x, z=d["z_phi2"]))
    assert np.quantile(dt, 0.999) < 1e-5 and dt.max() < 1e-3, "track"
    with pytest.raises(ValueError, match="Domain error"):
        TrackModel(dict(center=dict(type="Constant", value=0.0),
                        spread=dict(type="Constant", value=-1.0))).sample(x[:100])
    iso = IsochroneModel(dict(name="synthetic"))
    iso.create_isochrone(dict(name="synthetic"))
    from stream_sim_b200.synthetic import isochrone_table
    mg, mr = iso.sample(n, 16.5, seed=9)
    rg, rr = oracle.sample_mags(isochrone_table(), d["u_mass"], np.full(n, 16.5))
    assert_close(mg, rg, 1e-9, 0, "mag_g")
    assert_close(mr, rr, 1e-9, 0, "mag_r")


def test_tables_too_large_for_shared_memory():
    """A 40000-knot track/density table (1.9 MB of tables) does not fit the 227 KB of shared
    memory: the kernel then reads the blob from global memory.  Same answers."""
    from stream_sim_b200.model import StreamModel
    rng = np.random.default_rng(11)
    xk = np.sort(rng.uniform(-10, 10, 40000))
    xk[0], xk[-1] = -10.0, 10.0
    cfg = dict(density=dict(type="Interpolation", xvals=list(xk), yvals=list(1.0 + 0.5 * np.sin(xk))),
               track=dict(center=dict(type="Interpolation", xvals=list(xk), yvals=list(0.3 * np.cos(xk))),
                          spread=dict(type="Interpolation", xvals=list(xk),
                                      yvals=list(0.2 + 0.05 * np.sin(2 * xk))),
                          sampler="Gaussian"),
               distance_modulus=dict(center=dict(type="Line", slope=0.1, intercept=16.5),
                                     spread=dict(type="Constant", value=0.05), sampler="Gaussian"))
    n = 20000
    d = oracle.device_draws(3, 0, n)
    model = StreamModel(cfg)
    eng = model.engine()
    import ctypes as C
    smem, grid, block = C.c_int32(), C.c_int32(), C.c_int32()
    eng.lib.ss_launch_info(eng.blob.numel(), n, C.byref(smem), C.byref(grid), C.byref(block))
    assert eng.blob.numel() > 227 * 1024 and smem.value == 0 and block.value == 1024
    df = model.sample(n, draws=dict(u_phi1=d["u_phi1"], z_phi2=d["z_phi2"], z_dist=d["z_dist"]))
    ref = oracle.generate(cfg, d)
    assert np.array_equal(df["phi1"].to_numpy(), ref["phi1"])
    assert_close(df["phi2"].to_numpy(), ref["phi2"], 1e-9, 1e-12, "phi2")
    assert_close(df["dist"].to_numpy(dtype=float), ref["dist"], 1e-9, 0, "dist")


def test_survey_getters_on_device(survey):
    sv = oracle_survey()
    rng = np.random.default_rng(4)
    pix = rng.integers(0, 12 * 128 * 128, 5000)
    mag = rng.uniform(14, 30, 5000)
    maglim = survey.get_maglim("r", pix)
    assert np.array_equal(maglim, sv["maglim_maps"]["r"][pix], equal_nan=True)
    ext = survey.get_extinction("g", rng.integers(0, 12 * 512 * 512, 100))
    assert ext.shape == (100,) and np.all(ext > 0)
    err = survey.get_photo_error("r", mag, maglim)
    assert_close(err, oracle.get_photo_error(sv, "r", mag, maglim), 1e-12, 0, "photo error")
    for kind, key in (("completeness", "completeness"), ("detection_efficiency", "efficiency_detection")):
        c = survey.get_efficiency("r", mag, maglim, type=kind)
        assert np.array_equal(c, oracle.get_efficiency(sv, "r", mag, maglim, key), equal_nan=True)
    assert isinstance(survey.get_completeness("r", 24.0, 26.0), float)
    with pytest.raises(ValueError, match="not recognized"):
        survey.get_efficiency("r", mag, maglim, type="nope")


def test_device_function_evaluator_matches_host():
    """ss_eval_function: the device face of every function kind equals its host face."""
    import ctypes as C
    import torch
    from stream_sim_b200 import _lib
    from stream_sim_b200.engine import BlobPacker
    from stream_sim_b200.functions import function_factory
    lib = _lib.load()
    x = np.linspace(-15, 20, 20001)
    xt = torch.from_numpy(x).cuda()
    cases = [("constant", dict(value=2.5, xmin=-1.0, xmax=1.0)), ("line", dict(slope=0.1, intercept=16.5)),
             ("sinusoid", dict(amplitude=0.75, period=9.0, phase=0.3, offset=0.1)),
             ("fileinterpolation", dict(filename="./data/erkal_2016_pal_5_input.csv", columns=["phi1", "width"])),
             ("filecubicsplineinterpolation", dict(filename="./data/patrick_2022_splines.csv",
                                                   stream_name="atlas", spline_type="center")),
             ("filelineardensitycubicsplineinterpolation",
              dict(filename="./data/patrick_2022_splines.csv", stream_name="phoenix"))]
    for name, kw in cases:
        f = function_factory(name, **kw)
        pk = BlobPacker()
        desc = f.describe(pk)
        blob = torch.from_numpy(pk.finalize()).cuda()
        t = _lib.SSTables()
        t.blob, t.blob_bytes = blob.data_ptr(), blob.numel()
        out = torch.empty_like(xt)
        rc = lib.ss_eval_function(C.byref(desc), C.byref(t), xt.data_ptr(), out.data_ptr(), len(x),
                                  C.c_void_p(torch.cuda.current_stream().cuda_stream))
        _lib.check(rc)
        assert_close(out.cpu().numpy(), f(x), 1e-12, 1e-13, name)


def test_cli_driver(tmp_path):
    import os
    import subprocess
    import sys
    out = tmp_path / "stars.csv"
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "bin/generate_stream.py", "config/toy2_config.yaml", "-o",
                        str(out), "--seed", "5"], cwd=root, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    df = pd.read_csv(out)
    assert len(df) == 110000 and set(df["flag"]) == {0, 1} and (df["flag"] == 1).sum() == 10000
    assert df["phi1"].between(-10, 10).all()


def test_standalone_noise_and_detection_entry_points(survey):
    """StreamInjector.sample_measured_magnitudes / detect_flag (reference observed.py:863-959)
    run on the GPU (ss_noisy_magnitudes, ss_efficiency + ss_bernoulli): with the reference's
    own draws (a NumPy Generator) they equal the oracle's NumPy arithmetic -- BAD_MAG pattern
    and flags exact, magnitudes to 1e-12 -- and free-running they use the stars' own Philox
    draws."""
    from stream_sim_b200.observed import StreamInjector
    inj = StreamInjector(survey)
    n = 200000
    rng = np.random.default_rng(4)
    mag = rng.uniform(17.0, 29.0, n)
    mag[:3] = [np.nan, 250.0, -250.0]
    err = 10 ** rng.uniform(-3, 0.7, n)              # up to 5 mag: some fluxes go negative
    z = np.random.default_rng(99).standard_normal(n)
    got = inj.sample_measured_magnitudes(mag, err, rng=np.random.default_rng(99))
    with np.errstate(invalid="ignore", divide="ignore", over="ignore"):
        flux = 3631.0 * 10 ** (-0.4 * mag)
        fobs = flux + (flux * err / 1.0857362) * z
        ref = np.where(fobs > 0.0, -2.5 * np.log10(fobs / 3631.0), np.nan)
    bad = np.array([isinstance(v, str) and v == "BAD_MAG" for v in got])
    assert np.array_equal(bad, np.isnan(ref)) and 100 < bad.sum() < n / 2
    assert_close(np.array([float(v) for v in got[~bad]]), ref[~bad], 1e-12, 1e-12, "noisy magnitudes")
    # free-running: the stars' own z_r (fp32 Box-Muller of Philox OBS block)
    d = oracle.device_draws(5, 0, n)
    free = inj.sample_measured_magnitudes(mag, err, seed=5)
    with np.errstate(invalid="ignore", divide="ignore", over="ignore"):
        fobs = flux + (flux * err / 1.0857362) * d["z_r"]
    ok = (free != "BAD_MAG") & (fobs > 1e-3 * flux)
    assert ok.sum() > n / 2
    assert_close(np.array([float(v) for v in free[ok]]), -2.5 * np.log10(fobs[ok] / 3631.0), 0, 5e-3,
                 "free-running noisy magnitudes")
    # detect_flag
    sv = oracle_survey()
    pix = rng.integers(0, 12 * 128 * 128, n)
    magr = rng.uniform(15.5, 28.0, n)
    u = np.random.default_rng(7).uniform(size=n)
    for perfect, table in ((False, "completeness"), (True, "efficiency_detection")):
        flag = inj.detect_flag(pix, magr, "r", rng=np.random.default_rng(7), perfect_galstarsep=perfect)
        ref_flag = u <= oracle.get_efficiency(sv, "r", magr, sv["maglim_maps"]["r"][pix], table)
        assert flag.dtype == bool and np.array_equal(flag, ref_flag), table
    free_flag = inj.detect_flag(pix, magr, "r", seed=5)
    assert np.array_equal(free_flag, d["u_det"] <= oracle.get_efficiency(
        sv, "r", magr, sv["maglim_maps"]["r"][pix]))


def test_inject_columnar_outputs(survey, tmp_path):
    """inject(output="table" | "arrow" | "parquet"): same numbers as the reference-layout
    DataFrame, no object columns, new columns left on the GPU until used."""
    import pyarrow.parquet as pq
    import torch
    from stream_sim_b200.observed import StreamInjector
    from stream_sim_b200.table import StarTable
    inj = StreamInjector(survey)
    n = 50000
    rng = np.random.default_rng(8)
    cat = pd.DataFrame(dict(ra=rng.uniform(0, 360, n), dec=np.degrees(np.arcsin(rng.uniform(-1, 1, n))),
                            mag_g=rng.uniform(16, 30, n), mag_r=rng.uniform(16, 30, n)))
    ref = inj.inject(cat, seed=3, verbose=False, perfect_galstarsep=True)
    tab = inj.inject(cat, seed=3, verbose=False, perfect_galstarsep=True, output="table")
    assert isinstance(tab, StarTable) and tab.columns == list(ref.columns) and len(tab) == n
    assert isinstance(tab.device_column("mag_r_obs"), torch.Tensor) and tab.device_column("mag_r_obs").is_cuda
    same = tab.to_pandas()
    for c in ref.columns:
        if ref[c].dtype == object:
            assert (ref[c].astype(str) == same[c].astype(str)).all(), c
        else:
            assert np.array_equal(ref[c].to_numpy(), same[c].to_numpy(), equal_nan=True), c
    bad = (ref["mag_r_obs"] == "BAD_MAG").to_numpy()
    assert 0 < bad.sum() < n and np.array_equal(np.isnan(tab["mag_r_obs"]), bad)
    arrow = inj.inject(cat, seed=3, verbose=False, perfect_galstarsep=True, output="arrow")
    assert arrow.column("mag_r_obs").null_count == int(bad.sum())
    path = inj.inject(cat, seed=3, verbose=False, perfect_galstarsep=True, output="parquet",
                      path=str(tmp_path / "inj.parquet"))
    back = pq.read_table(path).to_pandas()
    assert np.array_equal(back["flag_observed"].to_numpy(), ref["flag_observed"].to_numpy())
    assert np.array_equal(back["magerr_g"].to_numpy(), ref["magerr_g"].to_numpy(), equal_nan=True)
    with pytest.raises(ValueError, match="needs path"):
        inj.inject(cat, seed=3, verbose=False, output="parquet")
