"""Shared builders for the parity tests (oracle-side and product-side inputs)."""
import numpy as np

import streamsim_oracle as oracle
from stream_sim_b200 import synthetic as syn


def oracle_survey(compl_band="r", first=-10.4, nside_maglim=128, nside_ebv=512):
    """Survey dict for the oracle, from the same synthetic inputs as Survey.synthetic."""
    props = syn.LSST_PROPS
    eff = syn.efficiency_csv_columns(first=first)
    err = syn.photo_error_csv_columns(first=first)
    dsat = props["delta_saturation"]
    return dict(ebv_map=syn.healpix_map("ebv", nside_ebv),
                maglim_maps=dict(g=syn.healpix_map("maglim_g", nside_maglim),
                                 r=syn.healpix_map("maglim_r", nside_maglim)),
                coeff_extinc=props["coeff_extinc"], saturation=props["saturation"],
                sys_error=props["sys_error"], delta_saturation=dsat,
                completeness_band=compl_band,
                photo_error=oracle.photo_error_table(err["delta_mag"], err["log_mag_err"], dsat),
                completeness=oracle.completeness_table(eff["delta_mag"],
                                                       eff["classification_detection_eff"], dsat),
                efficiency_detection=oracle.completeness_table(eff["delta_mag"],
                                                               eff["detection_eff"], dsat))


def stream_config(name):
    """Stream sections used by tests/bench: reference configs + synthetic isochrone."""
    from stream_sim_b200.utils import parse_config
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
