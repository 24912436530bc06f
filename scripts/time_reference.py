#!/usr/bin/env python
"""Time the UNMODIFIED reference (/root/reference) on this container's CPU: the parts of the
hot path that run here as written --

  * ``StreamModel.sample(N)``            reference stream_sim/model.py:106-157
  * ``StreamInjector.inject(catalog)``   reference stream_sim/observed.py:80-298

loaded exactly like scripts/make_golden.py does (third-party imports that are not installed
are stubbed; the one stub that executes is healpy.ang2pix = the oracle's NumPy restatement,
healpy's C++ would be faster).  ugali (isochrone photometry) and gala/astropy (the sky
rotation) are absent, so the catalog's magnitudes and ra/dec come from the oracle and are NOT
part of the timed region.  The reference cannot travel to the GPU box, hence this is measured
here, once, and committed:

    python scripts/time_reference.py [N] [nside] > profiles/r02_reference_as_written.json
"""
import json
import os
import platform
import sys
import time

import numpy as np
import pandas as pd

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
os.environ.setdefault("OMP_NUM_THREADS", "1")
os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
import make_golden as mg  # noqa: E402  (installs the stubs, imports the reference)


def main():
    n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000
    nside = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
    try:
        os.sched_setaffinity(0, {sorted(os.sched_getaffinity(0))[0]})
    except Exception:
        pass
    synth, oracle = mg.synth, mg.oracle
    cfg = mg.parse_config(os.path.join(mg.REF, "config", "pal5_config.yaml"))["stream"]
    cfg.pop("isochrone", None)
    model = mg.StreamModel(cfg)
    props = synth.LSST_PROPS
    maps = {k: synth.healpix_map(kind, nside) for k, kind in
            (("ebv", "ebv"), ("g", "maglim_g"), ("r", "maglim_r"))}
    import tempfile
    tmp = tempfile.mkdtemp()
    f_eff, f_err = os.path.join(tmp, "eff.csv"), os.path.join(tmp, "err.csv")
    pd.DataFrame(synth.efficiency_csv_columns()).to_csv(f_eff, index=False)
    pd.DataFrame(synth.photo_error_csv_columns()).to_csv(f_err, index=False)
    sv = mg.ref_surveys.Survey(name="synthetic", release="bench", bands=["g", "r"],
                               maglim_maps=dict(g=maps["g"], r=maps["r"]),
                               coeff_extinc=dict(props["coeff_extinc"]),
                               saturation=dict(props["saturation"]),
                               sys_error=dict(props["sys_error"]), completeness_band="r",
                               delta_saturation=props["delta_saturation"], ebv_map=maps["ebv"])
    sv.completeness = mg.ref_surveys.SurveyFactory.set_completeness(f_eff, props["delta_saturation"])
    sv.log_photo_error = mg.ref_surveys.SurveyFactory.set_photo_error(f_err, props["delta_saturation"])
    inj = mg.ref_observed.StreamInjector(sv)
    R = oracle.frame_from_endpoints(*synth.NOTEBOOK_ENDPOINTS[0], *synth.NOTEBOOK_ENDPOINTS[1])
    iso = synth.isochrone_table()

    t_sample, t_inject, frac = [], [], 0.0
    for rep in range(3):
        np.random.seed(rep)
        t0 = time.perf_counter()
        df = model.sample(n)
        t_sample.append(time.perf_counter() - t0)
        # not timed: sky positions and magnitudes (gala / ugali absent) from the oracle
        ra, dec = oracle.rotate(df["phi1"].to_numpy(float), df["phi2"].to_numpy(float), R)
        mg_, mr_ = oracle.sample_mags(iso, np.random.uniform(size=n), df["dist"].to_numpy(float))
        cat = pd.DataFrame(dict(ra=ra, dec=dec, mag_g=mg_, mag_r=mr_))
        t0 = time.perf_counter()
        res = inj.inject(cat, seed=rep, verbose=False, nside=nside)
        t_inject.append(time.perf_counter() - t0)
        frac = float(res["flag_observed"].mean())
    ts, ti = float(np.median(t_sample)), float(np.median(t_inject))
    print(json.dumps({
        "what": "unmodified reference timed in the build container (1 core): StreamModel.sample + "
                "StreamInjector.inject; rotation and isochrone photometry excluded (gala, astropy, "
                "ugali not installed), healpy.ang2pix = NumPy restatement",
        "kind": "reference", "cores": 1, "stars": n, "nside": nside,
        "sample_s": ts, "inject_s": ti, "stars_per_s_sample": n / ts, "stars_per_s_inject": n / ti,
        "stars_per_s": n / (ts + ti), "observed_fraction": frac,
        "cpu": platform.processor() or platform.machine(), "numpy": np.__version__,
        "pandas": pd.__version__, "repeats": 3, "statistic": "median"}))


if __name__ == "__main__":
    main()
