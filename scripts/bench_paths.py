#!/usr/bin/env python
"""Device-timed throughput of the individual entry points and configurations
(secondary numbers for DESIGN.md; the headline is bench.py).  One B200."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
os.chdir(ROOT)
import torch  # noqa: E402
from helpers import stream_config  # noqa: E402
from stream_sim_b200 import _lib  # noqa: E402
from stream_sim_b200.engine import StarEngine  # noqa: E402
from stream_sim_b200.frames import GreatCircleFrame  # noqa: E402
from stream_sim_b200.model import StreamModel  # noqa: E402
from stream_sim_b200.surveys import Survey  # noqa: E402
from stream_sim_b200.synthetic import NOTEBOOK_ENDPOINTS  # noqa: E402


def timed(fn, steps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps * 1e-3


def inject_paths(n=10_000_000):
    """Wall-clock of StreamInjector.inject through the Python API at n rows, per output mode:
    how much of a call is host work around the kernel (pandas, object columns, sinks)."""
    import tempfile
    import time
    import numpy as np
    import pandas as pd
    from stream_sim_b200.observed import StreamInjector
    dev = torch.device("cuda", 0)
    inj = StreamInjector(Survey.synthetic(4096, 4096, device=dev))
    rng = np.random.default_rng(0)
    cat = pd.DataFrame(dict(ra=rng.uniform(0, 360, n), dec=np.degrees(np.arcsin(rng.uniform(-1, 1, n))),
                            mag_g=rng.uniform(17, 29, n), mag_r=rng.uniform(17, 29, n)))
    tmp = tempfile.mkdtemp()
    modes = {
        "pandas, reference layout (BAD_MAG object columns)": dict(),
        "pandas, bad_mag='nan' (float columns)": dict(bad_mag="nan"),
        "table (columns stay on the GPU)": dict(output="table"),
        "arrow": dict(output="arrow"),
        "parquet (zstd)": dict(output="parquet", path=os.path.join(tmp, "inj.parquet")),
    }
    res = {}
    inj.inject(cat.iloc[:1000], seed=1, verbose=False)             # build the engine, warm up
    for name, kw in modes.items():
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = inj.inject(cat, seed=1, verbose=False, **kw)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        res[name] = dict(seconds=dt, rows_per_s=n / dt)
        del out
    t0 = time.perf_counter()
    cat.to_csv(os.path.join(tmp, "cat.csv"), index=False)
    res["(for scale) DataFrame.to_csv of the 4 input columns"] = dict(seconds=time.perf_counter() - t0)
    return dict(rows=n, modes=res)


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "inject":
        print(json.dumps(inject_paths(int(float(sys.argv[2])) if len(sys.argv) > 2 else 10_000_000), indent=1))
        return
    n = 1 << 25
    dev = torch.device("cuda", 0)
    frame = GreatCircleFrame.from_endpoints(*NOTEBOOK_ENDPOINTS)
    survey = Survey.synthetic(4096, 4096, device=dev)
    res = {}
    spline = dict(density=dict(type="FileLinearDensityCubicSplineInterpolation",
                               filename="./data/patrick_2022_splines.csv", stream_name="atlas"),
                  track=dict(center=dict(type="FileCubicSplineInterpolation",
                                         filename="./data/patrick_2022_splines.csv",
                                         spline_type="center", stream_name="atlas"),
                             spread=dict(type="FileCubicSplineInterpolation",
                                         filename="./data/patrick_2022_splines.csv",
                                         spline_type="spread", stream_name="atlas"),
                             sampler="Gaussian"),
                  distance_modulus=dict(center=dict(type="FileCubicSplineInterpolation",
                                                    filename="./data/patrick_2022_splines.csv",
                                                    spline_type="dist_mod", stream_name="atlas"),
                                        spread=dict(type="Constant", value=0.05)),
                  isochrone=dict(name="synthetic"))
    cfgs = dict(pal5=stream_config("pal5"), toy1=stream_config("toy1"), atlas_spline=spline)
    for name, cfg in cfgs.items():
        eng = StarEngine(stream=StreamModel(cfg), survey=survey, frame=frame, device=dev, hist=True)
        kind = {_lib.DENSITY_GRID: "grid table", _lib.DENSITY_INVCDF: "generic inverse CDF",
                _lib.DENSITY_UNIFORM: "uniform"}[eng.params.density_kind]
        cnt = eng.new_counters()
        out = eng.generate_observe(n, seed=1, counters=cnt)
        dt = timed(lambda: eng.generate_observe(n, seed=1, counters=cnt, out=out))
        res[f"fused_{name}"] = dict(stars_per_s=n / dt, density=kind, bytes_per_star=89,
                                    hbm_gbs=n * 89 / dt / 1e9)
        if name == "pal5":
            gen = eng.generate(n, seed=1)
            dt = timed(lambda: eng.generate(n, seed=1, out=gen))
            res["generate_only_pal5"] = dict(stars_per_s=n / dt, bytes_per_star=40,
                                             hbm_gbs=n * 40 / dt / 1e9)
            cat = {k: out[k] for k in ("ra", "dec", "mag_g", "mag_r")}
            obs = eng.observe(cat, seed=1)
            dt = timed(lambda: eng.observe(cat, seed=1, out=obs))
            res["observe_only"] = dict(stars_per_s=n / dt, bytes_per_star=65,
                                       hbm_gbs=n * 65 / dt / 1e9)
            dt = timed(lambda: eng.run_summary(n, seed=1, counters=cnt))
            res["summary_only_pal5"] = dict(stars_per_s=n / dt, bytes_per_star=0)
            rot = eng.rotate(out["phi1"], out["phi2"])
            dt = timed(lambda: eng.rotate(out["phi1"], out["phi2"], out=rot))
            res["rotate_only"] = dict(stars_per_s=n / dt, bytes_per_star=32,
                                      hbm_gbs=n * 32 / dt / 1e9)
        del eng, out
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
