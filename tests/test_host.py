"""Host-side logic: C-ABI library loads and exports every declared symbol, table
packing, configuration schema, function evaluators, error behaviour.  CPU only --
no compute call is made on the library here.
"""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT
from stream_sim_b200 import _lib


# ------------------------------------------------------------------ C ABI
def test_library_loads_and_exports_every_declared_symbol():
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "streamsim_b200.h")).read()
    declared = set(re.findall(r"^\s*(?:int|int64_t|const char\*)\s+(ss_[a-z0-9_]+)\s*\(", header, re.M))
    assert declared, "no declarations parsed from the header"
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.ss_abi_version() == _lib.ABI_VERSION
    m = re.search(r"#define SS_ABI_VERSION (\d+)", header)
    assert int(m.group(1)) == _lib.ABI_VERSION


def test_struct_mirrors_match_library():
    lib = _lib.load()
    sizes = (ctypes.c_int32 * 7)()
    assert lib.ss_struct_sizes(sizes) == 0
    assert [ctypes.sizeof(s) for s in _lib._STRUCTS] == list(sizes)
    assert lib.ss_struct_sizes(None) == -1
    assert b"NULL" in lib.ss_last_error()


def test_header_constants_match_binding():
    header = open(os.path.join(ROOT, "include", "streamsim_b200.h")).read()
    defs = dict(re.findall(r"#define (SS_[A-Z0-9_]+) (-?\d+)u?\b", header))
    assert int(defs["SS_N_COUNTERS"]) == _lib.N_COUNTERS
    assert int(defs["SS_HIST_BINS"]) == _lib.HIST_BINS
    assert int(defs["SS_FN_SPLINE_PRODUCT"]) == _lib.FN_SPLINE_PRODUCT
    assert int(defs["SS_DENSITY_NONE"]) == _lib.DENSITY_NONE
    assert int(defs["SS_CNT_INSIDE"]) == _lib.CNT_INSIDE
    assert int(defs["SS_STAGE_OBSERVE"]) == _lib.STAGE_OBSERVE


def test_no_gpu_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from stream_sim_b200.model import StreamModel
    from stream_sim_b200.utils import parse_config
    cfg = parse_config("config/toy1_config.yaml")["background"]
    with pytest.raises(_lib.StreamSimLibraryError, match="no CPU fallback"):
        StreamModel(cfg).sample(10)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "stream_sim_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "streamsim_oracle" not in text and "oracle." not in text, f


# ------------------------------------------------------------------ functions (host face)
def test_function_known_answers_from_reference_tests(known_answers):
    from stream_sim_b200.functions import (CubicSplineInterpolation, FileCubicSplineInterpolation,
                                           LinearDensityCubicSplineInterpolation,
                                           FileLinearDensityCubicSplineInterpolation)
    ka = known_answers
    x = np.array(ka["x"])
    nan = lambda v: np.array([np.nan if e is None else e for e in v])
    sp = CubicSplineInterpolation(np.array(ka["spread_nodes"]), np.array(ka["spread_values"]))
    np.testing.assert_allclose(sp(x), nan(ka["spread_expected"]))
    fsp = FileCubicSplineInterpolation("./data/patrick_2022_splines.csv", spline_type="spread",
                                       stream_name="atlas")
    np.testing.assert_allclose(fsp(x), nan(ka["spread_expected"]))
    ld = LinearDensityCubicSplineInterpolation(
        spread_nodes=np.array(ka["spread_nodes"]), spread_node_values=np.array(ka["spread_values"]),
        intensity_nodes=np.array(ka["intensity_nodes"]),
        intensity_node_values=np.array(ka["intensity_values"]))
    np.testing.assert_allclose(ld(x), nan(ka["linear_density_expected"]), atol=1e-7)
    fld = FileLinearDensityCubicSplineInterpolation(filename="./data/patrick_2022_splines.csv",
                                                    stream_name="atlas")
    np.testing.assert_allclose(fld(x), nan(ka["linear_density_expected"]), atol=1e-7)


def test_reference_test_file_runs_unchanged_against_this_package():
    """The reference's tests/test_functions.py, imported as-is, passes against the
    drop-in ``stream_sim`` alias (only where /root/reference is mounted)."""
    ref = "/root/reference/tests/test_functions.py"
    if not os.path.exists(ref):
        pytest.skip("reference not mounted on this box")
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_test_functions", ref)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    import stream_sim.functions as F
    assert mod.CubicSplineInterpolation is F.CubicSplineInterpolation
    for name in dir(mod):
        if name.startswith("test_"):
            getattr(mod, name)()


def test_functions_match_oracle_and_bounds():
    import streamsim_oracle as oracle
    from stream_sim_b200.functions import function_factory
    x = np.linspace(-12, 20, 501)
    cases = [dict(type="constant", value=2.5), dict(type="line", slope=0.1, intercept=16.5),
             dict(type="sinusoid", amplitude=0.75, period=9.0, phase=0.3, offset=0.1,
                  xmin=-9.0, xmax=9.0),
             dict(type="fileinterpolation", filename="./data/erkal_2016_pal_5_input.csv",
                  columns=["phi1", "width"]),
             dict(type="filecubicsplineinterpolation", filename="./data/patrick_2022_splines.csv",
                  stream_name="phoenix", spline_type="center")]
    for cfg in cases:
        kw = dict(cfg)
        f = function_factory(kw.pop("type"), **kw)
        assert np.array_equal(f(x), oracle.OFunc(cfg, ROOT)(x), equal_nan=True), cfg["type"]
    f = function_factory("line", slope=2.0, intercept=1.0, xmin=0.0, xmax=1.0)
    assert f(0.5) == 2.0 and np.isnan(f(1.5)) and isinstance(f(0.5), float)
    with pytest.raises(Exception, match="Unrecognized function: nope"):
        function_factory("nope")
    from stream_sim_b200.samplers import sampler_factory
    with pytest.raises(Exception, match="Unrecognized sampler: nope"):
        sampler_factory("nope")


def test_c_abi_validates_before_touching_cuda_and_fails_loudly_without_a_gpu():
    """Argument validation happens before any CUDA call (so it is testable here): bucket
    counts of the direct look-up tables must be powers of two (u * buckets has to be exact).
    With valid arguments the call goes on to CUDA and, in this GPU-less container, returns
    the runtime's error -- there is no CPU fallback behind the ABI."""
    import ctypes as C
    lib = _lib.load()

    def attempt(lut_n, iso_lut_n=None):
        p, t, o = _lib.SSParams(), _lib.SSTables(), _lib.SSOut()
        p.density_kind, p.grid_nthr, p.grid_lut_n = _lib.DENSITY_GRID, 10, lut_n
        for f in (p.track_center, p.track_spread, p.dist_center, p.dist_spread):
            f.kind = _lib.FN_CONSTANT
        fake = 0x1000                                   # never dereferenced on these paths
        t.grid_rows = t.grid_lut = t.grid_thr = t.blob = fake
        t.blob_bytes = 16
        if iso_lut_n:
            p.has_iso = p.has_dist = p.iso_direct = 1
            p.iso_n, p.iso_lut_n = 5, iso_lut_n
            t.iso_rows = t.iso_lut = t.iso_thr = fake
        o.phi1 = fake
        rc = lib.ss_generate(C.byref(p), C.byref(t), None, C.byref(o), C.c_uint64(5),
                             C.c_uint64(0), C.c_uint64(1), None)
        return rc, lib.ss_last_error().decode()

    assert attempt(1000) == (-1, "grid_lut_n must be a power of two")
    assert attempt(1024, 1000) == (-1, "iso_lut_n must be a power of two")
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except ImportError:
        has_gpu = False
    if not has_gpu:
        rc, msg = attempt(1024, 1024)
        assert rc == -2 and "cuda" in msg.lower()


# ------------------------------------------------------------------ packing
def test_blob_packer_layout_and_guides():
    from stream_sim_b200.engine import BlobPacker, search_guide, snr_error_threshold
    pk = BlobPacker()
    x = np.array([3.0, 1.0, 2.0, 2.0, 5.0, 4.0, 6.0, 7.0, 9.0, 8.0])
    y = np.arange(10.0)
    f1 = pk.add_linear(x, y)
    f2 = pk.add_linear(x, 2 * y)
    assert f1["x_off"] == f2["x_off"], "shared abscissae are stored once"
    blob = pk.finalize()
    assert blob.dtype == np.uint8 and blob.size % 16 == 0
    d = blob.view(np.float64)
    xs = d[f1["x_off"]:f1["x_off"] + 10]
    assert np.array_equal(xs, np.sort(x))
    ys = d[f1["c_off"]:f1["c_off"] + 10]
    assert np.array_equal(ys, y[np.argsort(x, kind="mergesort")])       # stable: 2.0 twice keeps order
    from stream_sim_b200.engine import GUIDE_ONE, GUIDE_MANY
    guide = blob.view(np.uint32)[2 * f1["g_off"]:2 * f1["g_off"] + f1["g_n"]].astype(np.int64)
    edges = f1["g_x0"] + np.arange(f1["g_n"] + 1) / f1["g_inv"]
    for b, g in enumerate(guide):
        j, one, many = g & 0x3fffffff, bool(g & GUIDE_ONE), bool(g & GUIDE_MANY)
        # every x of the bucket has its knot interval at j (or j+1 if flagged ONE)
        for x in np.linspace(edges[b], edges[b + 1], 7):
            true_j = min(np.searchsorted(xs, x, side="right") - 1, len(xs) - 1)
            if many:
                assert true_j >= j
            elif one:
                assert true_j in (j, j + 1) and (true_j == j + 1) == (x >= xs[j + 1])
            else:
                assert true_j == j
    assert sum(bool(g & GUIDE_MANY) for g in guide) <= 4        # only around the repeated knot
    cs = np.sort(np.random.default_rng(0).uniform(size=1000))
    gd = search_guide(cs, 64)
    for g in range(65):
        assert gd[g] == max(np.searchsorted(cs, g / 64, side="right") - 1, 0)
    e = snr_error_threshold(5.0)
    assert 1.0 / e >= 5.0 and not (1.0 / np.nextafter(e, 1.0) >= 5.0)


def test_flagged_bucket_guide_fuzz():
    """The loop-free interval look-up of the survey curves (csrc lin_eval: guide entry, at
    most one compare) emulated in NumPy against searchsorted, on random / uniform / clustered
    / near-duplicate knot tables, probed at random points, at every knot and bucket edge and
    at their neighbouring doubles."""
    from stream_sim_b200.engine import BlobPacker, GUIDE_MANY, GUIDE_ONE
    rng = np.random.default_rng(1)
    for trial in range(60):
        n = int(rng.integers(8, 600))
        kind = trial % 5
        if kind == 0:
            xs = np.sort(rng.uniform(-10, 2, n))
        elif kind == 1:
            xs = np.linspace(-10.4, 1.5, n)
        elif kind == 2:
            xs = np.sort(np.r_[rng.uniform(-10, 2, n - 4), [0.5, 0.5, 0.5 + 1e-13, 0.5 - 1e-13]])
        elif kind == 3:
            xs = np.sort(np.r_[np.linspace(-10, 0, n // 2), np.linspace(0, 1e-6, n - n // 2)])
        else:
            xs = np.cumsum(rng.exponential(1.0, n)) - 5
        pk = BlobPacker()
        f = pk.add_linear(xs, rng.normal(size=n))
        blob = pk.finalize()
        assert f["g_n"] > 0 and f["p"][2] == xs[0] and f["p"][3] == xs[-1]
        xp = blob.view(np.float64)[f["x_off"]:f["x_off"] + n]
        guide = blob.view(np.uint32)[2 * f["g_off"]:2 * f["g_off"] + f["g_n"]].astype(np.int64)
        edges = f["g_x0"] + np.arange(f["g_n"] + 1) / f["g_inv"]
        x = np.concatenate([rng.uniform(xp[0], xp[-1], 5000), xp, np.nextafter(xp, -np.inf),
                            np.nextafter(xp, np.inf), edges, np.nextafter(edges, -np.inf),
                            np.nextafter(edges, np.inf)])
        x = x[(x >= xp[0]) & (x <= xp[-1])]
        b = np.clip(((x - f["g_x0"]) * f["g_inv"]).astype(np.int64), 0, f["g_n"] - 1)
        g = guide[b]
        j = g & 0x3fffffff
        true_j = np.minimum(np.searchsorted(xp, x, side="right") - 1, n - 1)
        jj = np.where(((g & GUIDE_ONE) != 0) & (x >= xp[np.minimum(j + 1, n - 1)]), j + 1, j)
        scan = (g & GUIDE_MANY) != 0                 # the device scans: right from any start
        assert np.all(j[scan] <= true_j[scan] + 1)
        assert np.array_equal(jj[~scan], true_j[~scan]), (trial, kind)


def test_inverse_cdf_tables_match_scipy_interp1d():
    """The (sorted cdf, order) pair reproduces interp1d's stable sort, including the
    non-monotone CDFs of spline densities that dip below zero (atlas, phoenix)."""
    import scipy.interpolate
    from stream_sim_b200.samplers import FileLinearDensityCubicSplineInterpolationSampler as S
    from stream_sim_b200.samplers import inverse_cdf_tables
    s = S("./data/patrick_2022_splines.csv", "phoenix")
    xg, pdf = s.grid()
    cs, order = inverse_cdf_tables(xg, pdf)
    assert not np.array_equal(order, np.arange(len(order))), "phoenix CDF is non-monotone"
    cdf = np.cumsum(pdf)
    cdf /= cdf[-1]
    fn = scipy.interpolate.interp1d(cdf, list(range(len(cdf))), bounds_error=False, fill_value=0.0)
    u = np.random.default_rng(0).uniform(size=5000)
    mine = np.where((u < cs[0]) | (u > cs[-1]), 0.0, np.interp(u, cs, order.astype(float)))
    assert np.array_equal(mine, fn(u))


def test_inverse_cdf_step_function_is_exact():
    """engine.invcdf_segments: index(u) = val[#thr <= u] reproduces rint(interp1d(...)(u))
    exactly -- for the monotone Pal 5 CDF (one step up per grid point) and for the
    non-monotone atlas / phoenix spline CDFs (the index also steps down) -- at random
    draws, densely inside the non-monotone region, at every threshold and at the double
    just below it."""
    from stream_sim_b200.engine import invcdf_index, invcdf_segments, invcdf_thresholds
    from stream_sim_b200.samplers import FileLinearDensityCubicSplineInterpolationSampler as S
    from stream_sim_b200.samplers import inverse_cdf_tables
    rng = np.random.default_rng(21)
    for name in ("atlas", "phoenix"):
        xg, pdf = S("./data/patrick_2022_splines.csv", name).grid()
        cs, order = inverse_cdf_tables(xg, pdf)
        thr, val = invcdf_segments(cs, order)
        assert np.all(np.diff(thr) >= 0) and len(val) == len(thr) + 1
        assert (np.diff(val) < 0).sum() > 100, "the index steps down inside the negative-pdf region"
        assert invcdf_thresholds(cs) is not None      # (monotone helper ignores the ordering)
        bad = np.flatnonzero(order != np.arange(len(order)))
        probes = [rng.uniform(size=300000),
                  rng.uniform(cs[bad.min() - 3], cs[bad.max() + 3], size=300000),
                  thr[thr < 1.0], np.nextafter(thr[(thr > 0) & (thr < 1.0)], -1.0)]
        for u in probes:
            assert np.array_equal(val[np.searchsorted(thr, u, side="right")],
                                  invcdf_index(cs, u, order)), name
    # monotone case: steps of +1, thresholds equal to the dedicated helper's
    pdf = np.abs(np.sin(np.linspace(0, 7, 5000))) + 0.01
    cs, order = inverse_cdf_tables(np.linspace(-1, 1, 5000), pdf)
    thr, val = invcdf_segments(cs, order)
    assert np.all(np.diff(val) == 1) and val[0] == 0 and val[-1] == 4999
    assert np.array_equal(invcdf_thresholds(cs), thr)


def test_inverse_cdf_step_function_fuzz():
    """Random pdfs -- positive, with negative stretches (non-monotone CDF), with zeros
    (duplicate CDF knots), leading / trailing zero tails: invcdf_segments either reproduces
    the reference index exactly or declines (None -> generic device path); it never raises
    and never returns a wrong table."""
    from stream_sim_b200.engine import invcdf_index, invcdf_segments
    from stream_sim_b200.samplers import inverse_cdf_tables
    rng = np.random.default_rng(0)
    ok = 0
    for trial in range(60):
        n = int(rng.integers(5, 2000))
        kind = trial % 6
        if kind == 0:
            pdf = rng.uniform(0, 1, n)
        elif kind == 1:
            pdf = rng.normal(0.3, 0.5, n)
        elif kind == 2:
            pdf = np.where(rng.uniform(size=n) < 0.3, 0.0, rng.uniform(0, 1, n))
        elif kind == 3:
            pdf = np.sin(np.linspace(0, 9, n)) + 0.2
        elif kind == 4:
            pdf = np.r_[np.zeros(n // 3), rng.uniform(0, 1, n - n // 3)]
        else:
            pdf = np.r_[rng.uniform(0, 1, n - n // 4), np.zeros(n // 4)]
        if pdf.sum() <= 0:
            pdf = pdf - pdf.sum() / n + 1.0 / n
        cs, order = inverse_cdf_tables(np.linspace(-1, 1, n), pdf.copy())
        seg = invcdf_segments(cs, order)
        if seg is None:
            continue
        thr, val = seg
        inner = thr[(thr > 0) & (thr < 1)]
        u = np.concatenate([rng.uniform(size=5000), inner, np.nextafter(inner, -1.0)])
        assert np.array_equal(val[np.searchsorted(thr, u, side="right")],
                              invcdf_index(cs, u, order)), (trial, kind)
        ok += 1
    assert ok >= 30


# ------------------------------------------------------------------ configuration + survey schema
def test_word_tables_reproduce_the_step_function():
    """engine.word_thresholds / bucket_wlut (SSTables.grid_wlut, iso_wlut): for every 32-bit
    word w the integer look-up gives #{k : thr[k] <= w * 2**-32}, incl. thresholds at or below
    zero, at or above one, runs of equal thresholds, buckets fuller than three thresholds and
    the saturated k field."""
    from stream_sim_b200.engine import (WLUT_K_CAP, bucket_wlut, invcdf_segments, wlut_step,
                                        word_thresholds)
    from stream_sim_b200.model import StreamModel
    from stream_sim_b200 import samplers as S
    rng = np.random.default_rng(7)

    def check(thr, buckets):
        wthr, n_below = word_thresholds(thr)
        lut = bucket_wlut(wthr, n_below, buckets)
        shift = 32 - buckets.bit_length() + 1
        w64 = wthr.astype(np.uint64)
        cand = np.concatenate([
            rng.integers(0, 1 << 32, 100000, dtype=np.uint64), w64,
            np.minimum(w64 + 1, 0xFFFFFFFF), np.maximum(w64.astype(np.int64) - 1, 0).astype(np.uint64),
            np.array([0, 1, 0xFFFFFFFE, 0xFFFFFFFF], dtype=np.uint64),
            np.arange(buckets, dtype=np.uint64) << np.uint64(shift)])
        want = np.searchsorted(thr, cand.astype(np.float64) / 4294967296.0, side="right")
        assert np.array_equal(wlut_step(lut, wthr, n_below, cand), want)
        return lut[:, 0] >> 20

    from helpers import stream_config
    st = StreamModel(stream_config("pal5"))
    xg, pdf = st.density.density.grid(100000)
    thr, _ = invcdf_segments(*S.inverse_cdf_tables(xg, pdf))
    k = check(thr, 1 << 18)
    assert (k > 3).mean() < 0.01                      # the scan is the rare path
    k = check(st.isochrone.table.composite()[0][1:], 1 << 15)
    assert k.max() <= 3                               # the isochrone table never scans
    check(np.array([-1.0, 0.0, 0.0, 0.25, 0.25, 0.25, 0.25, 0.25, 0.5, 1.0, 1.5, np.inf]), 4)
    k = check(np.sort(rng.uniform(size=20000)) ** 8, 1 << 8)
    assert k.max() == WLUT_K_CAP
    check(np.array([]), 2)
    with pytest.raises(ValueError):
        word_thresholds([0.5, 0.25])


def test_configs_parse_and_models_build():
    from stream_sim_b200.model import (BackgroundModel, SplineStreamModel, StreamModel,
                                       TrackModel)
    from stream_sim_b200.utils import parse_config
    for name in ("toy1", "toy2", "pal5"):
        cfg = parse_config(f"config/{name}_config.yaml")
        st = dict(cfg["stream"])
        st.pop("isochrone", None)
        m = StreamModel(st)
        assert isinstance(m.track, TrackModel) and m.velocity is None
        BackgroundModel(cfg["background"])
        assert float(cfg["stream"]["nstars"]) > 0
    cfg = parse_config("config/atlas_spline_config.yaml")
    m = SplineStreamModel(cfg["stream"])           # upstream's raises AttributeError here
    assert m.distance_modulus is None and m.density is not None
    assert parse_config("a: {b: 1}") == {"a": {"b": 1}}
    with pytest.raises(Exception, match="Unrecognized sampler"):
        StreamModel(dict(density=dict(type="Bogus"), track=st["track"]))
    # isochrone needs ugali unless a table is given
    with pytest.raises(ImportError, match="ugali"):
        StreamModel(parse_config("config/toy1_config.yaml")["stream"])


def test_survey_curves_match_oracle_tables():
    import streamsim_oracle as oracle
    from stream_sim_b200 import synthetic as syn
    from stream_sim_b200.surveys import SurveyFactory, TabulatedFunction
    for first in (-8.0, -10.4, -12.0):
        eff = syn.efficiency_csv_columns(first=first)
        err = syn.photo_error_csv_columns(first=first)
        c = SurveyFactory.completeness_from_arrays(eff["delta_mag"], eff["detection_eff"], -10.4)
        ox, oy = oracle.completeness_table(eff["delta_mag"], eff["detection_eff"], -10.4)
        assert np.array_equal(c.x, ox) and np.array_equal(c.y, oy) and c.fill_value == 0.0
        e = SurveyFactory.photo_error_from_arrays(err["delta_mag"], err["log_mag_err"], -10.4)
        ox, oy = oracle.photo_error_table(err["delta_mag"], err["log_mag_err"], -10.4)
        assert np.array_equal(e.x, ox) and np.array_equal(e.y, oy) and e.fill_value == 1.0
        d = np.linspace(-15, 5, 999)
        assert np.array_equal(e(d), oracle._interp1d_linear(ox, oy, d, 1.0))
    import scipy.interpolate
    f = scipy.interpolate.interp1d([0.0, 1.0], [1.0, 3.0], bounds_error=False, fill_value=0.0)
    t = TabulatedFunction.coerce(f)
    assert np.array_equal(t(np.array([-1.0, 0.5, 2.0])), [0.0, 2.0, 0.0])


def test_survey_csv_loaders(tmp_path):
    import pandas as pd
    from stream_sim_b200 import synthetic as syn
    from stream_sim_b200.surveys import SurveyFactory
    p = tmp_path / "eff.csv"
    pd.DataFrame(syn.efficiency_csv_columns()).to_csv(p, index=False)
    both = SurveyFactory.set_completeness(str(p), -10.4)
    det = SurveyFactory.set_completeness(str(p), -10.4, selection="detected")
    assert both(-5.0) < det(-5.0) <= 1.0
    with pytest.raises(ValueError, match="Invalid selection"):
        SurveyFactory.set_completeness(str(p), selection="nope")
    q = tmp_path / "err.csv"
    pd.DataFrame(syn.photo_error_csv_columns()).to_csv(q, index=False)
    e = SurveyFactory.set_photo_error(str(q), -10.4)
    assert float(e(10.0)) == 1.0 and float(e(-20.0)) == 1.0


def test_survey_factory_loads_npy_survey(tmp_path):
    """YAML-driven loader + cache on a self-contained synthetic survey directory."""
    import pandas as pd
    from stream_sim_b200 import synthetic as syn
    from stream_sim_b200.surveys import Survey, SurveyFactory
    for kind, fn in (("maglim_g", "ml_g.npy"), ("maglim_r", "ml_r.npy"), ("ebv", "ebv.npy")):
        np.save(tmp_path / fn, syn.healpix_map(kind, 8))
    pd.DataFrame(syn.efficiency_csv_columns()).to_csv(tmp_path / "eff.csv", index=False)
    pd.DataFrame(syn.photo_error_csv_columns()).to_csv(tmp_path / "err.csv", index=False)
    cfg = dict(name="toy", release="r1",
               survey_files=dict(file_path=str(tmp_path), maglim_map_g="ml_g.npy",
                                 maglim_map_r="ml_r.npy", ebv_map="ebv.npy",
                                 completeness="eff.csv", completeness_band="r",
                                 log_photo_error="err.csv"),
               survey_properties=dict(bands=["g", "r"], coeff_extinc_g=3.66, sys_error=0.005,
                                      saturation=16.0, delta_saturation=-10.4))
    SurveyFactory.clear_cache(verbose=False)
    sv = Survey.load("toy", "r1", config_file=cfg, verbose=False)
    assert sv.bands == ["g", "r"] and sv.coeff_extinc == {"g": 3.66, "r": 2.285}
    assert sv.completeness_band == "r" and sv.delta_saturation == -10.4
    assert len(sv.maglim_maps["g"]) == 768 and np.isnan(sv.maglim_maps["g"]).any()
    assert sv.efficiency_detection is not None and sv.log_photo_error is not None
    assert sv.coverage is not None and set(np.unique(sv.coverage)) <= {0.0, 1.0}
    assert Survey.load("toy", "r1", verbose=False) is sv            # cached
    assert SurveyFactory.list_cached_surveys() == ["toy_r1"]
    Survey.load("toy", "r1", verbose=False, uniform_survey=True,
                uniform_survey_mask_type=["nan"])
    flat = SurveyFactory._cached_surveys["toy_r1_uniform"]
    vals = flat.maglim_maps["r"][~np.isnan(flat.maglim_maps["r"])]
    assert np.all(vals == vals[0])
    SurveyFactory.clear_cache("toy", verbose=False)
    assert SurveyFactory.list_cached_surveys() == []
    with pytest.raises(FileNotFoundError):
        Survey.load("nosuch", verbose=False)
    with pytest.raises(ValueError, match="Band 'i' not available"):
        sv.get_maglim("i")


def test_shipped_lsst_config_loads_hsp_and_fits_natively(tmp_path):
    """config/surveys/lsst_yr1.yaml as shipped (HealSparse depth maps + a HEALPix FITS dust
    map + the two CSV curves), pointed at a directory holding synthetic files of those names:
    the YAML-driven loader (reference surveys.py:891-1148) goes through the native .hsp and
    FITS readers -- neither healsparse nor healpy is installed here."""
    import pandas as pd
    import yaml
    from stream_sim_b200 import fitsio, healpix as hp, synthetic as syn
    from stream_sim_b200.surveys import Survey, SurveyFactory
    cfg = yaml.safe_load(open(os.path.join(ROOT, "config", "surveys", "lsst_yr1.yaml")))
    files = cfg["survey_files"]
    files["file_path"] = str(tmp_path)
    ml_g, ml_r, ebv = (syn.healpix_map(k, 32) for k in ("maglim_g", "maglim_r", "ebv"))
    fitsio.write_healsparse(str(tmp_path / files["maglim_map_g"]), ml_g, nside_coverage=8)
    fitsio.write_healsparse(str(tmp_path / files["maglim_map_r"]), ml_r, nside_coverage=8,
                            compress="GZIP_1")
    fitsio.write_healpix_map(str(tmp_path / files["ebv_map"]), ebv, dtype=np.float64)
    pd.DataFrame(syn.efficiency_csv_columns()).to_csv(tmp_path / files["completeness"], index=False)
    pd.DataFrame(syn.photo_error_csv_columns()).to_csv(tmp_path / files["log_photo_error"], index=False)
    SurveyFactory.clear_cache(verbose=False)
    sv = Survey.load("lsst", "yr1", config_file=cfg, verbose=False)
    for got, want in ((sv.maglim_maps["g"], ml_g), (sv.maglim_maps["r"], ml_r)):
        # .hsp depth maps are clipped at the saturation magnitude (reference :1013-1016)
        ok = ~np.isnan(want) & (want >= 16.0)
        assert np.array_equal(np.isnan(got), ~ok)
        assert np.array_equal(got[ok], want[ok].astype(np.float32))      # .hsp stores float32
    assert np.array_equal(np.asarray(sv.ebv_map), ebv)
    assert sorted(sv.bands) == sorted(["u", "g", "r", "i", "z", "y"]) and sv.completeness_band == "r"
    assert sv.coeff_extinc["r"] == 2.70136780871597 and sv.delta_saturation == -10.4
    assert sv.completeness is not None and sv.log_photo_error is not None
    assert hp.get_nside(sv.coverage) == 32
    SurveyFactory.clear_cache(verbose=False)


def test_native_fits_healpix_reader(tmp_path):
    """Survey ingestion without healpy: HEALPix BINTABLE files, RING/NESTED,
    1 or 1024 pixels per row, float32/float64, gzip."""
    from stream_sim_b200 import fitsio, healpix as hp
    from stream_sim_b200 import synthetic as syn
    from stream_sim_b200.surveys import SurveyFactory
    m = syn.healpix_map("maglim_r", 16)
    m = np.where(np.isnan(m), hp.UNSEEN, m)
    for name, kw in (("a.fits", {}), ("b.fits.gz", dict(dtype=np.float64, per_row=1)),
                     ("c.fits", dict(nest=True))):
        path = str(tmp_path / name)
        fitsio.write_healpix_map(path, hp.reorder_ring_to_nest(m) if kw.get("nest") else m, **kw)
        r = SurveyFactory.read_map(path)
        assert r.dtype == np.float64 and len(r) == len(m)
        if kw.get("dtype") is np.float64:
            assert np.array_equal(r, m)
        else:
            np.testing.assert_allclose(r, m, rtol=1e-6)
        assert np.array_equal(fitsio.read_healpix_map(path, nest=True),
                              hp.reorder_ring_to_nest(r))
    with pytest.raises(ValueError):
        bad = tmp_path / "bad.fits"
        bad.write_bytes(b"not a fits file".ljust(2880))
        fitsio.read_healpix_map(str(bad))


def test_native_healsparse_reader(tmp_path):
    """``.hsp`` maps (reference surveys.py:1255-1284, healsparse) are read natively: plain
    and tile-compressed (GZIP_1 / GZIP_2) files, empty coverage pixels, UNSEEN -> NaN and
    the map_min / map_max clips of ``open_map_sparse``.  No .hsp fixture ships with the
    reference: round trip through our own writer (parity unpinned, see fitsio.py)."""
    from stream_sim_b200 import fitsio, healpix as hp
    from stream_sim_b200.surveys import SurveyFactory
    rng = np.random.default_rng(3)
    nside = 64
    m = np.full(hp.nside2npix(nside), np.nan)
    idx = rng.choice(m.size, 9000, replace=False)
    m[idx] = rng.normal(25.0, 0.4, idx.size).astype(np.float32)
    m[: m.size // 3] = np.nan                       # whole coverage pixels without data
    seen = ~np.isnan(m)
    for compress in (None, "GZIP_1", "GZIP_2"):
        path = str(tmp_path / f"m_{compress}.hsp")
        fitsio.write_healsparse(path, m, nside_coverage=8, compress=compress)
        nside_cov, nside_sparse, cov, vals, sentinel = fitsio.read_healsparse(path)
        assert (nside_cov, nside_sparse, sentinel) == (8, nside, hp.UNSEEN)
        assert cov.size == hp.nside2npix(8) and vals.size % (nside // 8) ** 2 == 0
        out = SurveyFactory.read_map(path)
        assert np.array_equal(np.isnan(out), ~seen)
        assert np.array_equal(out[seen], m[seen])
        nest = fitsio.read_healsparse_map(path, nest=True)
        assert np.array_equal(hp.reorder_nest_to_ring(nest)[seen], m[seen])
    clipped = SurveyFactory.open_map_sparse(path, map_min=24.8, map_max=25.5)
    assert np.array_equal(np.isnan(clipped), ~(seen & (m >= 24.8) & (m <= 25.5)))
    with pytest.raises(ValueError):
        bad = tmp_path / "bad.hsp"
        fitsio.write_healpix_map(str(bad), np.zeros(hp.nside2npix(4)))
        fitsio.read_healsparse(str(bad))


def test_api_surface_covers_the_reference():
    """Every public class / function / method of the reference's six path modules exists
    here with the reference's parameters first, in order (scripts/api_diff.py imports the
    unmodified reference with its third-party imports stubbed).  Build container only."""
    import subprocess
    import sys
    if not os.path.isdir("/root/reference/stream_sim"):
        pytest.skip("the reference tree is only present in the build container")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "scripts", "api_diff.py")],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "SIGNATURE ORDER DIFFERS:\n" in r.stdout + "\n" and r.stdout.rstrip().endswith("DIFFERS:")


def test_plot_stream_in_mask_error_paths():
    """reference observed.py:1078-1134: ValueError when no mask can be built; drawing needs
    matplotlib (absent here: ImportError, after the mask was built on the host)."""
    import pandas as pd
    from stream_sim_b200.observed import StreamInjector
    from stream_sim_b200.surveys import Survey, SurveyFactory
    sv = Survey.synthetic(16, 16)
    SurveyFactory._build_coverage_map(sv, verbose=False)
    inj = StreamInjector(sv)
    df = pd.DataFrame(dict(ra=[10.0, 11.0], dec=[0.0, 1.0]))
    with pytest.raises(ValueError, match="Could not create mask"):
        inj.plot_stream_in_mask(df, None)
    try:
        import matplotlib  # noqa: F401
    except ImportError:
        with pytest.raises(ImportError, match="matplotlib"):
            inj.plot_stream_in_mask(df, ["footprint"])


def test_frames_and_ud_grade():
    import streamsim_oracle as oracle
    from stream_sim_b200 import healpix as hp
    from stream_sim_b200.frames import GreatCircleFrame
    from stream_sim_b200.synthetic import NOTEBOOK_ENDPOINTS
    fr = GreatCircleFrame.from_endpoints(*NOTEBOOK_ENDPOINTS)
    R = oracle.frame_from_endpoints(*NOTEBOOK_ENDPOINTS[0], *NOTEBOOK_ENDPOINTS[1])
    np.testing.assert_allclose(fr.R, R, atol=1e-15)
    np.testing.assert_allclose(fr.R @ fr.R.T, np.eye(3), atol=1e-14)
    assert np.allclose(GreatCircleFrame.coerce(fr.R).R, fr.R)
    # ud_grade: degrading a map that is constant per parent returns the parent value;
    # unseen children are ignored; all-unseen parents become UNSEEN
    coarse = np.arange(48, dtype=float)
    fine = hp.ud_grade(coarse, 4)
    assert len(fine) == 192 and np.array_equal(hp.ud_grade(fine, 2), coarse)
    fine2 = fine.copy()
    nest = hp.ring2nest(4, np.arange(192))
    fine2[nest < 4] = np.nan                     # all children of NEST parent 0
    back = hp.ud_grade(fine2, 2)
    assert (back == hp.UNSEEN).sum() == 1
    assert hp.npix2nside(12 * 64 ** 2) == 64
    with pytest.raises(ValueError):
        hp.npix2nside(100)


def test_shard_ranges_cover_exactly():
    from stream_sim_b200.distributed import shard_range
    for n in (0, 1, 7, 8, 1000003):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert all(hi >= lo for lo, hi in spans)


def test_snr_threshold_on_the_log_error():
    """The SNR cut decided on log10(statistical error): the host threshold reproduces the
    reference arithmetic 1/sqrt((10**x)**2 + sys**2) >= 5 on both sides of it."""
    from stream_sim_b200.engine import snr_loge_threshold
    for sys_err in (0.0, 0.005, 0.03, 0.15):
        x = snr_loge_threshold(sys_err)
        probe = x + np.array([-1e-3, -1e-9, -1e-13, 0.0]) 
        above = x + np.array([1e-13, 1e-9, 1e-3])
        ok = lambda v: 1.0 / np.sqrt((10 ** v) ** 2 + sys_err ** 2) >= 5.0
        assert ok(probe).all() and not ok(above).any(), sys_err
    assert np.isnan(snr_loge_threshold(0.3))          # sys error alone fails the cut


def test_direct_lookup_curves_match_interp1d():
    """The direct look-up form of the survey curves (SSFunc.q_*): its host model equals
    np.interp with fill values bit for bit -- at random points, at every knot and bucket edge
    and their neighbouring doubles, outside the table and at NaN."""
    from stream_sim_b200 import synthetic as syn
    from stream_sim_b200.engine import BlobPacker, direct_lookup_eval

    def ref(xs, ys, fill, x):
        out = np.interp(x, xs, ys)
        out = np.where((x < xs[0]) | (x > xs[-1]), fill, out)
        return np.where(np.isnan(x), np.nan, out)

    rng = np.random.default_rng(0)
    err, eff = syn.photo_error_csv_columns(), syn.efficiency_csv_columns()
    tables = [(err["delta_mag"], err["log_mag_err"], 1.0),
              (eff["delta_mag"], eff["classification_detection_eff"], 0.0)]
    for _ in range(10):
        n = int(rng.integers(2, 150))
        tables.append((np.sort(rng.uniform(-10, 3, n)), rng.normal(size=n), 0.25))
    built = 0
    for xs, ys, fill in tables:
        pk = BlobPacker()
        pk.add_doubles(np.zeros(int(rng.integers(0, 3))))          # odd / even starting offsets
        q = pk.add_direct_lookup(xs, ys, fill)
        if not q:
            continue
        built += 1
        assert q["q_off"] % 2 == 0 and q["q_seg_off"] % 2 == 0
        edges = xs[0] + np.arange(q["q_n"] + 1) / q["q_inv"]
        x = np.concatenate([rng.uniform(xs[0] - 3, xs[-1] + 3, 200000), xs, np.nextafter(xs, np.inf),
                            np.nextafter(xs, -np.inf), edges, np.nextafter(edges, np.inf),
                            np.nextafter(edges, -np.inf), [np.nan, -1e300, 1e300]])
        got, exp = direct_lookup_eval(pk.finalize(), q, x), ref(xs, ys, fill, x)
        assert np.array_equal(got, exp, equal_nan=True)
    assert built >= 8
    assert BlobPacker().add_direct_lookup([0.0, 0.0, 1.0], [1.0, 2.0, 3.0], 0.0) == {}   # duplicate knots


def test_star_table_sinks(tmp_path):
    """StarTable (the columnar result of inject(output="table")) on host arrays: the pandas
    view in the reference's layout ("BAD_MAG" object columns) and as plain floats, Arrow with
    nulls, a Parquet round trip and the CSV writer in both flavours."""
    import pandas as pd
    import pyarrow.parquet as pq
    from stream_sim_b200.table import StarTable
    n = 1000
    rng = np.random.default_rng(3)
    mobs = rng.uniform(18, 28, n)
    mobs[::7] = np.nan
    t = StarTable(dict(ra=rng.uniform(0, 360, n), mag_r_obs=mobs, magerr_r=rng.uniform(0, 1, n),
                       flag_observed=(rng.uniform(size=n) < 0.3).astype(np.uint8)),
                  bad_mag_columns=["mag_r_obs", "not_there"])
    assert len(t) == n and t.columns == ["ra", "mag_r_obs", "magerr_r", "flag_observed"]
    df = t.to_pandas()
    assert df["mag_r_obs"].dtype == object and (df["mag_r_obs"][::7] == "BAD_MAG").all()
    assert df["flag_observed"].dtype == bool
    assert float(df["mag_r_obs"][1]) == mobs[1]
    plain = t.to_pandas(compat=False)
    assert plain["mag_r_obs"].dtype == np.float64 and np.isnan(plain["mag_r_obs"][0])
    at = t.to_arrow()
    assert at.column("mag_r_obs").null_count == len(mobs[::7]) and at.num_rows == n
    back = pq.read_table(t.to_parquet(str(tmp_path / "t.parquet"))).to_pandas()
    assert np.array_equal(back["ra"].to_numpy(), t["ra"])
    assert np.array_equal(np.isnan(back["mag_r_obs"].to_numpy(dtype=float)), np.isnan(mobs))
    csv = pd.read_csv(t.to_csv(str(tmp_path / "t.csv")))
    assert (csv["mag_r_obs"][::7] == "BAD_MAG").all() and float(csv["mag_r_obs"][1]) == mobs[1]
    csv2 = pd.read_csv(t.to_csv(str(tmp_path / "t2.csv"), compat=False), float_precision="round_trip")
    assert np.isnan(csv2["mag_r_obs"][0]) and csv2["mag_r_obs"][1] == mobs[1]
    with pytest.raises(ValueError, match="different lengths"):
        StarTable(dict(a=np.zeros(3), b=np.zeros(4)))


def _fits_card(key, value=None, comment=""):
    """One 80-character header card, laid out by hand from the FITS standard (v4.0 section 4.2):
    keyword in columns 1-8, '= ' in 9-10, fixed-format values right-justified to column 30,
    strings quoted from column 11 and padded to at least 8 characters."""
    if value is None:
        return f"{key:<8}".ljust(80).encode("ascii")
    if isinstance(value, bool):
        field = f"{'T' if value else 'F':>20}"
    elif isinstance(value, int):
        field = f"{value:>20d}"
    else:
        field = "'" + f"{value:<8}" + "'"
        field = f"{field:<20}"
    text = f"{key:<8}= {field}" + (f" / {comment}" if comment else "")
    assert len(text) <= 80
    return text.ljust(80).encode("ascii")


def _fits_block(cards):
    raw = b"".join(cards) + _fits_card("END")
    return raw + b" " * (-len(raw) % 2880)


def test_healpix_fits_file_built_from_the_standard(tmp_path):
    """A HEALPix FITS map assembled byte by byte from the FITS standard, the way healpy's
    write_map lays it out (primary HDU without data, one BINTABLE with 1024 big-endian
    float32 values per row, PIXTYPE/ORDERING/NSIDE/INDXSCHM cards) -- nothing of
    stream_sim_b200.fitsio's own writer is involved -- is read back value for value, in RING
    and NESTED order; likewise a partial-sky file with an EXPLICIT pixel index column."""
    import struct
    from stream_sim_b200 import fitsio, healpix as hp
    from stream_sim_b200.surveys import SurveyFactory
    nside = 16
    npix = 12 * nside * nside
    values = (np.arange(npix, dtype=np.float32) * np.float32(0.25) - np.float32(100.0))
    values[5] = np.float32(hp.UNSEEN)
    primary = _fits_block([_fits_card("SIMPLE", True, "conforms to FITS standard"),
                           _fits_card("BITPIX", 8, "array data type"),
                           _fits_card("NAXIS", 0, "number of array dimensions"),
                           _fits_card("EXTEND", True)])
    per_row = 1024
    table_hdr = _fits_block([
        _fits_card("XTENSION", "BINTABLE", "binary table extension"),
        _fits_card("BITPIX", 8, "array data type"), _fits_card("NAXIS", 2),
        _fits_card("NAXIS1", 4 * per_row, "length of dimension 1"),
        _fits_card("NAXIS2", npix // per_row, "length of dimension 2"),
        _fits_card("PCOUNT", 0), _fits_card("GCOUNT", 1), _fits_card("TFIELDS", 1),
        _fits_card("TTYPE1", "TEMPERATURE", "Temperature map"), _fits_card("TFORM1", "1024E"),
        _fits_card("TUNIT1", "mag"),
        _fits_card("PIXTYPE", "HEALPIX", "HEALPIX pixelisation"),
        _fits_card("ORDERING", "NESTED", "Pixel ordering scheme, either RING or NESTED"),
        _fits_card("COMMENT"),                                     # a blank commentary card
        _fits_card("NSIDE", nside, "Resolution parameter of HEALPIX"),
        _fits_card("FIRSTPIX", 0), _fits_card("LASTPIX", npix - 1),
        _fits_card("INDXSCHM", "IMPLICIT", "Indexing: IMPLICIT or EXPLICIT"),
        _fits_card("OBJECT", "FULLSKY", "Sky coverage, either FULLSKY or PARTIAL")])
    data = struct.pack(f">{npix}f", *values.tolist())
    data += b"\\0" * (-len(data) % 2880)
    path = str(tmp_path / "by_hand.fits")
    with open(path, "wb") as f:
        f.write(primary + table_hdr + data)
    nest = fitsio.read_healpix_map(path, nest=True)
    assert nest.dtype == np.float64 and np.array_equal(nest, values.astype(np.float64))
    ring = fitsio.read_healpix_map(path)
    assert np.array_equal(ring, hp.reorder_nest_to_ring(values.astype(np.float64)))
    loaded = SurveyFactory.read_map(path)        # FITS maps come back as hp.read_map gives them
    assert np.array_equal(loaded, ring) and loaded[hp.ring2nest(nside, np.arange(npix)) == 5][0] < -1e30

    # partial sky: EXPLICIT indexing, columns PIXEL (32-bit integer) and SIGNAL (float64), RING
    pix = np.array([3, 77, 1000, npix - 1], dtype=np.int32)
    sig = np.array([21.5, 22.25, -3.0, 1e-3])
    hdr = _fits_block([
        _fits_card("XTENSION", "BINTABLE"), _fits_card("BITPIX", 8), _fits_card("NAXIS", 2),
        _fits_card("NAXIS1", 12), _fits_card("NAXIS2", len(pix)), _fits_card("PCOUNT", 0),
        _fits_card("GCOUNT", 1), _fits_card("TFIELDS", 2),
        _fits_card("TTYPE1", "PIXEL"), _fits_card("TFORM1", "1J"),
        _fits_card("TTYPE2", "SIGNAL"), _fits_card("TFORM2", "1D"),
        _fits_card("PIXTYPE", "HEALPIX"), _fits_card("ORDERING", "RING"),
        _fits_card("NSIDE", nside), _fits_card("INDXSCHM", "EXPLICIT"),
        _fits_card("OBJECT", "PARTIAL")])
    rows = b"".join(struct.pack(">id", int(p), float(s)) for p, s in zip(pix, sig))
    rows += b"\\0" * (-len(rows) % 2880)
    path2 = str(tmp_path / "partial.fits")
    with open(path2, "wb") as f:
        f.write(primary + hdr + rows)
    part = fitsio.read_healpix_map(path2)
    assert part.size == npix and np.array_equal(part[pix], sig)
    assert (part == hp.UNSEEN).sum() == npix - len(pix)
