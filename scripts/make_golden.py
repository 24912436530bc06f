#!/usr/bin/env python
"""Generate tests/golden/* by running the UNMODIFIED reference from /root/reference.

Run once in the build container (the reference cannot travel to the GPU box):

    python scripts/make_golden.py

What is produced and how it was obtained is recorded inside every fixture
(key ``provenance``).  Third-party packages that the reference imports but that
are not installed here (healpy, healsparse, astropy, gala, pylab, skyproj) are
replaced by stub modules *only so that ``import stream_sim.observed`` succeeds*;
the only stubbed function that is ever executed is ``healpy.ang2pix`` /
``get_nside``, for which the oracle's NumPy restatement is used -- so these
fixtures pin the reference's photometry/selection arithmetic, not healpy.
"""
import importlib.util
import json
import os
import sys
import tempfile
import types

import numpy as np
import pandas as pd

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
GOLD = os.path.join(REPO, "tests", "golden")


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


oracle = _load("streamsim_oracle", os.path.join(REPO, "oracle", "streamsim_oracle.py"))
synth = _load("ss_synthetic", os.path.join(REPO, "stream_sim_b200", "synthetic.py"))

# the reference must win the name ``stream_sim``
sys.path = [p for p in sys.path if os.path.abspath(p or ".") != REPO]
sys.path.insert(0, REF)


def install_stubs():
    hp = types.ModuleType("healpy")
    hp.UNSEEN = -1.6375e30
    hp.ang2pix = lambda nside, lon, lat, lonlat=True, nest=False: oracle.ang2pix_ring(
        nside, np.asarray(lon, dtype=float), np.asarray(lat, dtype=float))
    hp.get_nside = lambda m: oracle.npix2nside(len(m))
    hp.nside2npix = lambda ns: 12 * ns * ns
    sys.modules["healpy"] = hp
    for name in ("healsparse", "astropy", "astropy.coordinates", "astropy.units",
                 "gala", "gala.coordinates", "pylab", "skyproj"):
        sys.modules[name] = types.ModuleType(name)
    sys.modules["astropy"].coordinates = sys.modules["astropy.coordinates"]
    sys.modules["astropy"].units = sys.modules["astropy.units"]
    sys.modules["gala"].coordinates = sys.modules["gala.coordinates"]


install_stubs()
os.chdir(REF)  # the reference configs use ./data/... relative paths
from stream_sim.model import StreamModel            # noqa: E402
from stream_sim.utils import parse_config           # noqa: E402
from stream_sim import surveys as ref_surveys       # noqa: E402
from stream_sim import observed as ref_observed     # noqa: E402

PROV = ("made by scripts/make_golden.py from the unmodified reference at /root/reference "
        "(numpy %s)" % np.__version__)


# ------------------------------------------------------------------ generate
def golden_generate():
    n = 4096
    cases = {}
    for name in ("toy1", "toy2", "pal5"):
        cfg = parse_config(os.path.join(REF, "config", f"{name}_config.yaml"))["stream"]
        cfg.pop("isochrone", None)          # ugali absent
        cases[name] = cfg
    bkg = parse_config(os.path.join(REF, "config", "toy1_config.yaml"))["background"]
    cases["toy1_background"] = bkg
    # the spline configuration through plain StreamModel (SplineStreamModel is broken
    # upstream, model.py:670); same tables as config/atlas_spline_config.yaml
    fn = "./data/patrick_2022_splines.csv"
    for stream in ("atlas", "phoenix"):
        cases[f"{stream}_spline"] = dict(
            density=dict(type="FileLinearDensityCubicSplineInterpolation", filename=fn,
                         stream_name=stream),
            track=dict(center=dict(type="FileCubicSplineInterpolation", filename=fn,
                                   spline_type="center", stream_name=stream),
                       spread=dict(type="FileCubicSplineInterpolation", filename=fn,
                                   spline_type="spread", stream_name=stream),
                       sampler="Gaussian"),
            distance_modulus=dict(center=dict(type="FileCubicSplineInterpolation", filename=fn,
                                              spline_type="dist_mod", stream_name=stream),
                                  spread=dict(type="Constant", value=0.05),
                                  sampler="Gaussian"))
    for i, (name, cfg) in enumerate(cases.items()):
        seed = 1000 + i
        np.random.seed(seed)
        df = StreamModel(cfg).sample(n)
        # replay the global MT19937 stream in the reference's consumption order
        np.random.seed(seed)
        draws = {}
        draws["u_phi1"] = np.random.uniform(size=n)
        for key, sec in (("phi2", "track"), ("dist", "distance_modulus")):
            if not cfg.get(sec):
                continue
            if cfg[sec].get("sampler", "Gaussian").lower() == "gaussian":
                draws["z_" + key] = np.random.standard_normal(n)
            else:
                draws["u_" + key] = np.random.uniform(size=n)
        out = {k: df[k].to_numpy(dtype=float) for k in ("phi1", "phi2", "dist")
               if df[k].notna().any() or k != "dist"}
        if not cfg.get("distance_modulus"):
            out.pop("dist", None)
        # the oracle must agree bit for bit before the fixture is written
        mine = oracle.generate(cfg, draws, None, data_root=REF)
        for k, v in out.items():
            assert np.array_equal(mine[k], v, equal_nan=True), (name, k)
        np.savez_compressed(os.path.join(GOLD, f"generate_{name}.npz"),
                            config=json.dumps(cfg), provenance=PROV, seed=seed,
                            **{"draw_" + k: v for k, v in draws.items()},
                            **{"out_" + k: v for k, v in out.items()})
        print("generate", name, "ok", {k: float(np.nanmean(v)) for k, v in out.items()})


# -------------------------------------------------------------------- inject
def make_catalog(n, rng):
    """Stars over the whole sky, every ang2pix branch, every photometry branch."""
    ra = rng.uniform(0.0, 360.0, n)
    dec = np.degrees(np.arcsin(rng.uniform(-1.0, 1.0, n)))
    edge = np.degrees(np.arcsin(2.0 / 3.0))
    special = [(0.0, 90.0), (0.0, -90.0), (360.0, 0.0), (45.0, edge), (135.0, -edge),
               (359.999999, 89.9), (180.0, -89.95), (90.0, 0.0), (0.0, 0.0), (270.0, 66.0),
               (12.5, 89.5), (200.0, -89.6)]
    for i, (a, d) in enumerate(special):
        ra[i], dec[i] = a, d
    mag_r = rng.uniform(14.0, 29.5, n)      # saturated (<16) ... far below the depth
    mag_g = mag_r + rng.uniform(0.1, 1.5, n)
    mag_r[20:24] = 16.0                     # exactly at saturation
    mag_g[24:28] = 15.999
    return pd.DataFrame(dict(ra=ra, dec=dec, mag_g=mag_g, mag_r=mag_r))


def golden_inject():
    n = 3000
    tmp = tempfile.mkdtemp()
    variants = dict(A=dict(first=-8.0, compl_band="r"),       # table starts fainter than dsat
                    B=dict(first=-10.4, compl_band="g"),      # starts exactly at dsat
                    C=dict(first=-12.0, compl_band="r"))      # starts brighter than dsat
    maps = dict(ebv=synth.healpix_map("ebv", 512),
                g=synth.healpix_map("maglim_g", 128),
                r=synth.healpix_map("maglim_r", 128))
    for i, (tag, v) in enumerate(variants.items()):
        eff = synth.efficiency_csv_columns(first=v["first"])
        err = synth.photo_error_csv_columns(first=v["first"])
        f_eff = os.path.join(tmp, f"eff_{tag}.csv")
        f_err = os.path.join(tmp, f"err_{tag}.csv")
        pd.DataFrame(eff).to_csv(f_eff, index=False)
        pd.DataFrame(err).to_csv(f_err, index=False)
        props = synth.LSST_PROPS
        sv = ref_surveys.Survey(name="synthetic", release=tag, bands=["g", "r"],
                                maglim_maps=dict(g=maps["g"], r=maps["r"]),
                                coeff_extinc=dict(props["coeff_extinc"]),
                                saturation=dict(props["saturation"]),
                                sys_error=dict(props["sys_error"]),
                                completeness_band=v["compl_band"],
                                delta_saturation=props["delta_saturation"],
                                ebv_map=maps["ebv"])
        sv.completeness = ref_surveys.SurveyFactory.set_completeness(f_eff, props["delta_saturation"])
        sv.efficiency_detection = ref_surveys.SurveyFactory.set_completeness(
            f_eff, props["delta_saturation"], selection="detected")
        sv.log_photo_error = ref_surveys.SurveyFactory.set_photo_error(f_err, props["delta_saturation"])
        cat = make_catalog(n, np.random.default_rng(500 + i))
        seed = 42 + i
        for dust in (True, False):
            inj = ref_observed.StreamInjector(sv)
            res = inj.inject(cat, seed=seed, perfect_galstarsep=True, verbose=False,
                             dust_correction=dust)
            # replay PCG64(seed) in the reference's consumption order (observed.py:154-243)
            rng = np.random.default_rng(seed)
            draws = {}
            for band in ("r", "g"):
                draws["z_" + band] = rng.standard_normal(n)
                if band == v["compl_band"]:
                    draws["u_det"] = rng.uniform(size=n)
                    draws["u_det2"] = rng.uniform(size=n)
            out = {}
            for b in ("r", "g"):
                col = res[f"mag_{b}_obs"].to_numpy()
                out[f"mag_{b}_obs"] = np.array([np.nan if isinstance(x, str) and x == "BAD_MAG"
                                                else float(x) for x in col])
                out[f"magerr_{b}"] = res[f"magerr_{b}"].to_numpy(dtype=float)
            out["flag_observed"] = res["flag_observed"].to_numpy(dtype=bool)
            out["flag_perfect_galstarsep"] = res["flag_perfect_galstarsep"].to_numpy(dtype=bool)
            # oracle check (bit-equal) before writing
            osv = dict(ebv_map=maps["ebv"], maglim_maps=dict(g=maps["g"], r=maps["r"]),
                       coeff_extinc=props["coeff_extinc"], saturation=props["saturation"],
                       sys_error=props["sys_error"], delta_saturation=props["delta_saturation"],
                       completeness_band=v["compl_band"],
                       photo_error=oracle.photo_error_table(err["delta_mag"], err["log_mag_err"]),
                       completeness=oracle.completeness_table(
                           eff["delta_mag"], eff["classification_detection_eff"]),
                       efficiency_detection=oracle.completeness_table(
                           eff["delta_mag"], eff["detection_eff"]))
            mine = oracle.observe(osv, cat["ra"].to_numpy(), cat["dec"].to_numpy(),
                                  cat["mag_g"].to_numpy(), cat["mag_r"].to_numpy(), draws,
                                  perfect_galstarsep=True, dust_correction=dust)
            for k, val in out.items():
                assert np.array_equal(mine[k], val, equal_nan=True), (tag, dust, k,
                                                                      np.flatnonzero(mine[k] != val)[:5])
            np.savez_compressed(
                os.path.join(GOLD, f"inject_{tag}_dust{int(dust)}.npz"),
                provenance=PROV, seed=seed, compl_band=v["compl_band"], first=v["first"],
                dust_correction=dust, nside_maglim=128, nside_ebv=512, nside=4096,
                **{"cat_" + k: cat[k].to_numpy() for k in cat.columns},
                **{"draw_" + k: x for k, x in draws.items()},
                **{"out_" + k: x for k, x in out.items()})
            print("inject", tag, "dust", dust, "ok: detected", int(out["flag_observed"].sum()),
                  "perfect", int(out["flag_perfect_galstarsep"].sum()),
                  "bad_r", int(np.isnan(out["mag_r_obs"]).sum()),
                  "nan_err", int(np.isnan(out["magerr_r"]).sum()))


# ---------------------------------------------------- known answers (numbers only)
def golden_known_answers():
    ref_test = _load("ref_test_functions", os.path.join(REF, "tests", "test_functions.py"))
    ka = dict(
        provenance="numbers transcribed by scripts/make_golden.py from the reference's own "
                   "tests/test_functions.py:8-35 and notebook cell outputs",
        x=list(np.linspace(-10, 10, 20)),
        spread_nodes=list(ref_test.SPREAD_NODES), spread_values=list(ref_test.SPREAD_NODE_VALUES),
        spread_expected=[None if np.isnan(v) else v for v in ref_test.SPREAD_INTERP_ARRAY],
        intensity_nodes=list(ref_test.INTENSITY_NODES),
        intensity_values=list(ref_test.INTENSITY_NODE_VALUES),
        linear_density_expected=[None if np.isnan(v) else v
                                 for v in ref_test.LINEAR_DENSITY_INTERP_ARRAY],
        # notebooks/tutorial_inject_stream.ipynb cells 13-15
        notebook_inject=dict(seed=42, n=1000, accepted_trial=2,
                             radec_head=[[309.295394, -10.749590], [310.268379, -14.058956],
                                         [308.757294, -10.075699], [309.516004, -11.506027],
                                         [311.724304, -17.227233]],
                             phi_head=[[2.739560, -0.875874], [-0.611216, -0.083476],
                                       [3.585979, -0.741940], [1.973680, -0.695347],
                                       [-4.058227, 0.264566]],
                             saturated_magerr=10.000001),
        # notebooks/injection_lsst_example.ipynb cell 7: describe() of 115000 Pal-5 stars
        pal5_describe=dict(phi1=dict(mean=2.334, std=5.396, min=-7.099, max=16.278),
                           phi2=dict(mean=-0.509, std=0.649),
                           dist=dict(mean=16.733, std=0.540)))
    with open(os.path.join(GOLD, "known_answers.json"), "w") as f:
        json.dump(ka, f, indent=1)
    print("known answers written")


# ----------------------------------------------- distribution samples for KS tests
def golden_marginals():
    n = 20000
    out = {}
    for name in ("toy2", "pal5"):
        cfg = parse_config(os.path.join(REF, "config", f"{name}_config.yaml"))["stream"]
        cfg.pop("isochrone", None)
        np.random.seed(7)
        df = StreamModel(cfg).sample(n)
        for k in ("phi1", "phi2", "dist"):
            if df[k].notna().any():
                out[f"{name}_{k}"] = np.sort(df[k].to_numpy(dtype=float)).astype(np.float32)
    np.savez_compressed(os.path.join(GOLD, "marginals.npz"), provenance=PROV, **out)
    print("marginals written", list(out))


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    golden_known_answers()
    golden_generate()
    golden_inject()
    golden_marginals()
