"""Shared builders for the parity tests (oracle-side and product-side inputs)."""
import numpy as np

import streamsim_oracle as oracle
from stream_sim_b200 import synthetic as syn


class LazyMap:
    """A synthetic HEALPix map evaluated only at the pixels asked for (the maps are closed-form
    hashes of the pixel index): lets the oracle look up nside-4096 maps without building
    200-million-pixel arrays on the host.  Supports what the oracle uses: len() and
    indexing with an integer array."""

    def __init__(self, kind, nside):
        self.kind, self.nside = kind, nside

    def __len__(self):
        return 12 * self.nside * self.nside

    def __getitem__(self, pix):
        s_val, s_hole = syn._SALT[self.kind]
        pix = np.asarray(pix, dtype=np.uint64)
        return syn._map_from_hash(syn._mix_np(pix, s_val), syn._mix_np(pix, s_hole), self.kind, np)


def oracle_survey(compl_band="r", first=-10.4, nside_maglim=128, nside_ebv=512, lazy=False):
    """Survey dict for the oracle, from the same synthetic inputs as Survey.synthetic."""
    props = syn.LSST_PROPS
    eff = syn.efficiency_csv_columns(first=first)
    err = syn.photo_error_csv_columns(first=first)
    dsat = props["delta_saturation"]
    mk = LazyMap if lazy else syn.healpix_map
    return dict(ebv_map=mk("ebv", nside_ebv),
                maglim_maps=dict(g=mk("maglim_g", nside_maglim),
                                 r=mk("maglim_r", nside_maglim)),
                coeff_extinc=props["coeff_extinc"], saturation=props["saturation"],
                sys_error=props["sys_error"], delta_saturation=dsat,
                completeness_band=compl_band,
                photo_error=oracle.photo_error_table(err["delta_mag"], err["log_mag_err"], dsat),
                completeness=oracle.completeness_table(eff["delta_mag"],
                                                       eff["classification_detection_eff"], dsat),
                efficiency_detection=oracle.completeness_table(eff["delta_mag"],
                                                               eff["detection_eff"], dsat))


def stream_config(name):
    """Stream sections used by tests/bench: reference configs + synthetic isochrone.
    ``atlas_spline`` / ``phoenix_spline``: the Patrick et al. (2022) spline streams with
    their distance-modulus splines (the configs of the golden generate fixtures)."""
    from stream_sim_b200.utils import parse_config
    if name == "wholesky":
        # stress variant: stars over the whole sky -- flat density along phi1
        # in [-180, 180], a wide Gaussian across: both polar caps, every base face, the poles
        return dict(density=dict(type="Interpolation", xvals=[-180.0, 180.0], yvals=[1.0, 1.0]),
                    track=dict(center=dict(type="Constant", value=0.0),
                               spread=dict(type="Constant", value=40.0), sampler="Gaussian"),
                    distance_modulus=dict(center=dict(type="Line", slope=0.0, intercept=16.5),
                                          spread=dict(type="Constant", value=0.3), sampler="Gaussian"),
                    isochrone=dict(name="synthetic"))
    if name == "tails":
        # zero density at both ends: a quarter of the index thresholds sit at u = 0 (folded
        # into the first bucket's step), a quarter are never reached
        cfg = stream_config("gap")
        cfg["density"] = dict(type="Interpolation", xvals=[-10.0, -5.0, -4.99, 4.99, 5.0, 10.0],
                              yvals=[0.0, 0.0, 1.0, 1.0, 0.0, 0.0])
        return cfg
    if name == "gap":
        # a density with an almost empty stretch: ~10^5 index thresholds of the phi1 sampler
        # crowd into a few 32-bit words (one bucket of the word table saturates its count)
        return dict(density=dict(type="Interpolation", xvals=[-10.0, -9.0, -8.99, 8.99, 9.0, 10.0],
                                 yvals=[1.0, 1.0, 3e-6, 3e-6, 1.0, 1.0]),
                    track=dict(center=dict(type="Line", slope=0.05, intercept=0.0),
                               spread=dict(type="Constant", value=0.3), sampler="Gaussian"),
                    distance_modulus=dict(center=dict(type="Line", slope=0.1, intercept=16.5),
                                          spread=dict(type="Constant", value=0.05), sampler="Uniform"),
                    isochrone=dict(name="synthetic"))
    if name in ("atlas_spline", "phoenix_spline"):
        f = "./data/patrick_2022_splines.csv"
        sn = name.split("_")[0]
        spl = lambda t: dict(type="FileCubicSplineInterpolation", filename=f, spline_type=t,
                             stream_name=sn)
        return dict(density=dict(type="FileLinearDensityCubicSplineInterpolation", filename=f,
                                 stream_name=sn),
                    track=dict(center=spl("center"), spread=spl("spread"), sampler="Gaussian"),
                    distance_modulus=dict(center=spl("dist_mod"),
                                          spread=dict(type="Constant", value=0.05),
                                          sampler="Gaussian"),
                    isochrone=dict(name="synthetic"))
    cfg = parse_config(f"config/{name}_config.yaml")["stream"]
    cfg["isochrone"] = dict(name="synthetic")
    return cfg


def assert_close(a, b, rtol, atol=0.0, what=""):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    assert a.shape == b.shape, what
    nan_a, nan_b = np.isnan(a), np.isnan(b)
    assert np.array_equal(nan_a, nan_b), f"{what}: NaN pattern differs at {np.flatnonzero(nan_a != nan_b)[:5]}"
    ok = ~nan_a
    err = np.abs(a[ok] - b[ok])
    lim = atol + rtol * np.abs(b[ok])
    assert np.all(err <= lim), f"{what}: max err {err.max():.3e} (limit {lim[np.argmax(err - lim)]:.3e})"
