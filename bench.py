#!/usr/bin/env python
"""Benchmark of the per-star generate-and-observe path (BASELINE.json metric:
observed stars/sec).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--stars S] [--dtype f64|f32]
  python bench.py --impl reference ...      # the CPU arm (oracle port, all host cores)

One step = one fused generate -> rotate -> observe pass over S synthetic stars per
GPU of the workload "LSST-like injection, nside-4096 depth + E(B-V) maps, Pal 5
stream model (reference config/pal5_config.yaml) + synthetic isochrone, stream
placed with the tutorial notebook's great-circle frame".  Under torchrun each
rank takes the index range [rank*S, (rank+1)*S) of every step (weak scaling);
the only collective is the SUM all-reduce of the counters.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)
os.chdir(ROOT)

import numpy as np  # noqa: E402

WORKLOAD = "lsst_injection_nside4096_pal5"
NSIDE = 4096
BYTES_PER_STAR = {"f64": 11 * 8 + 1, "f32": 11 * 4 + 1}
COLUMNS = ["phi1", "phi2", "dist", "mag_g", "mag_r", "ra", "dec", "mag_g_obs", "mag_r_obs",
           "magerr_g", "magerr_r", "flag_observed"]


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def kernel_metrics(dtype, math):
    """Counters of the production kernel from the committed ncu capture (profiles/
    kernel_metrics.json: DRAM bytes per star, issue-slot and pipe utilisation), or {}.  They
    are citations of a profile of this kernel, not measurements of this run."""
    try:
        with open(os.path.join(ROOT, "profiles", "kernel_metrics.json")) as f:
            return json.load(f).get(f"{dtype}_{math}", {})
    except Exception:
        return {}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region.  The
    sampler is started before the warm-up (nvidia-smi takes a while to start) and
    every row is time-stamped on arrival; stop(t0, t1) keeps the rows that fall
    inside the timed window, widened to the neighbouring samples if the window was
    shorter than a few sampling periods."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        pad = 0.0
        rows = [r for t, r in self.rows if t0 <= t <= t1]
        while len(rows) < 3 and pad < 0.5:          # short region: include the neighbours
            pad += 0.05
            rows = [r for t, r in self.rows if t0 - pad <= t <= t1 + pad]
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for r in rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for nm, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm), "window_pad_s": pad}


# ------------------------------------------------------------------ CPU arm
_CPU = {}


def _cpu_setup(maps):
    import streamsim_oracle as oracle
    from helpers import stream_config
    from stream_sim_b200 import synthetic as syn
    props = syn.LSST_PROPS
    eff, err = syn.efficiency_csv_columns(), syn.photo_error_csv_columns()
    dsat = props["delta_saturation"]
    _CPU["sv"] = dict(
        ebv_map=maps["ebv"], maglim_maps=dict(g=maps["g"], r=maps["r"]),
        coeff_extinc=props["coeff_extinc"], saturation=props["saturation"],
        sys_error=props["sys_error"], delta_saturation=dsat, completeness_band="r",
        photo_error=oracle.photo_error_table(err["delta_mag"], err["log_mag_err"], dsat),
        completeness=oracle.completeness_table(eff["delta_mag"],
                                               eff["classification_detection_eff"], dsat))
    _CPU["cfg"] = stream_config("pal5")
    _CPU["iso"] = syn.isochrone_table()
    _CPU["R"] = oracle.frame_from_endpoints(*syn.NOTEBOOK_ENDPOINTS[0], *syn.NOTEBOOK_ENDPOINTS[1])


def _cpu_chunk(args):
    """One worker call = what one reference call does for `n` stars: draw with
    NumPy's generators, then the whole NumPy chain (tables rebuilt per call, as
    the reference does)."""
    import streamsim_oracle as oracle
    seed, n = args
    rng = np.random.default_rng(seed)
    draws = dict(u_phi1=rng.uniform(size=n), z_phi2=rng.standard_normal(n),
                 u_dist=rng.uniform(size=n), u_mass=rng.uniform(size=n),
                 z_r=rng.standard_normal(n), u_det=rng.uniform(size=n),
                 z_g=rng.standard_normal(n))
    out = oracle.generate_observe(_CPU["cfg"], _CPU["iso"], _CPU["R"], _CPU["sv"], draws,
                                  nside=NSIDE)
    return int(out["flag_observed"].sum())


def cpu_maps(nside):
    """nside-4096 maps as NumPy arrays, hashed with torch's multi-threaded CPU ops."""
    from stream_sim_b200 import synthetic as syn
    return {k: syn.healpix_map(kind, nside, device="cpu").numpy()
            for k, kind in (("ebv", "ebv"), ("g", "maglim_g"), ("r", "maglim_r"))}


CPU_CHUNK = 1 << 20          # stars per reference-sized call (1e6 .. 1e7)


def host_cores():
    """The cores this process may use (its affinity mask), as a sorted list."""
    try:
        return sorted(os.sched_getaffinity(0))
    except Exception:
        return list(range(os.cpu_count() or 1))


def cpu_steps(maps, chunk, cores, steps, warmup):
    """Per-step wall times of the oracle port: every step = one `chunk`-star call on each of
    `cores` forked workers (all cores busy at once, like a user running one process per core)."""
    import multiprocessing as mp
    _cpu_setup(maps)
    ctx = mp.get_context("fork")
    times = []
    with ctx.Pool(cores) as pool:
        pool.map(_cpu_chunk, [(1, 1024)] * cores)          # fork + import + page in the maps
        for step in range(warmup + steps):
            jobs = [(1000 * step + 7 + i, chunk) for i in range(cores)]
            t0 = time.perf_counter()
            pool.map(_cpu_chunk, jobs, chunksize=1)
            dt = time.perf_counter() - t0
            if step >= warmup:
                times.append(dt)
    return times


def reference_as_written():
    """The unmodified reference timed in the build container (scripts/time_reference.py): it
    cannot run on the GPU box (needs /root/reference), so the committed record is cited."""
    try:
        with open(os.path.join(ROOT, "profiles", "r02_reference_as_written.json")) as f:
            d = json.load(f)
        return {"value": d["stars_per_s"], "unit": "stars/s", "cores": d["cores"], "kind": "reference",
                "sample": f"{d['stars']} stars, StreamModel.sample + StreamInjector.inject as written, "
                          f"nside {d['nside']}, median of {d['repeats']}",
                "source": "profiles/r02_reference_as_written.json (measured in the build "
                          "container, not in this run)"}
    except Exception:
        return None


def cpu_baseline_leg(maps, steps=1, warmup=0):
    cores = host_cores()
    times = cpu_steps(maps, CPU_CHUNK, len(cores), steps, warmup)
    per_step = len(cores) * CPU_CHUNK
    value = per_step / float(np.median(times))
    return {"value": value, "unit": "stars/s", "cores": len(cores), "kind": "port",
            "sample": f"{per_step * len(times)} stars of the same workload ({CPU_CHUNK} per call, one "
                      f"call per core and step, {len(times)} step(s), median)",
            "affinity": f"{cores[0]}-{cores[-1]} ({len(cores)} cores)",
            "per_core": value / len(cores)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    maps = cpu_maps(NSIDE)
    cores = host_cores()
    times = cpu_steps(maps, CPU_CHUNK, len(cores), args.steps, max(args.warmup, 1))
    per_step = len(cores) * CPU_CHUNK
    med = float(np.median(times))
    value = per_step / med
    line = {
        "impl": "reference", "metric": "observed stars/sec", "value": value, "unit": "stars/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * med, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "stars_per_step": per_step, "nside": NSIDE,
                   "statistic": "median over steps",
                   "note": "CPU arm: NumPy restatement of the reference chain (oracle port) on every "
                           "host core; the Python reference needs healpy/astropy/gala/ugali, absent "
                           "offline (its as-written timing: cpu_baseline_reference)"},
        "cpu_baseline": {"value": value, "unit": "stars/s", "cores": len(cores), "kind": "port",
                         "sample": f"{per_step} stars per step ({CPU_CHUNK} per call and core), "
                                   f"{args.steps} steps, median",
                         "affinity": f"{cores[0]}-{cores[-1]} ({len(cores)} cores)"},
        "cpu_baseline_reference": reference_as_written(),
        "e2e": {"value": value, "unit": "stars/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------ GPU arm
MATH_NOTE = {
    "f64": "every value computed in double",
    "fast": "same pixels, flags, positions and true magnitudes as f64 (all decisions in double); "
            "magerr / mag_obs values from fp32 MUFU arithmetic with a double redo inside a guard "
            "band: <= 1e-5 relative (north star bar for fp64 outputs)",
}


def wholesky_config():
    """Stress variant: stars spread over the whole sky, so that every map
    gather goes to a different part of the 200-million-pixel maps and misses L2."""
    from helpers import stream_config
    return stream_config("wholesky")


def build_engine(device, nside, survey=None, math="f64", fused=1, workload="pal5", packed="auto",
                 normals="f32", map_dtype="f64"):
    from helpers import stream_config
    from stream_sim_b200.engine import StarEngine
    from stream_sim_b200.frames import GreatCircleFrame
    from stream_sim_b200.model import StreamModel
    from stream_sim_b200.surveys import Survey
    from stream_sim_b200.synthetic import NOTEBOOK_ENDPOINTS
    survey = Survey.synthetic(nside, nside, device=device) if survey is None else survey
    cfg = wholesky_config() if workload == "wholesky" else stream_config("pal5")
    eng = StarEngine(stream=StreamModel(cfg), survey=survey,
                     frame=GreatCircleFrame.from_endpoints(*NOTEBOOK_ENDPOINTS), device=device,
                     nside=nside, hist=True, math=math, fused_kernel=bool(fused), packed_maps=packed,
                     normals=normals, map_dtype=map_dtype)
    return eng, survey


def run_gpu(args):
    import torch
    import torch.distributed as dist
    from stream_sim_b200 import distributed as ssd
    rank, world, local = ssd.init_from_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200; there is no CPU fallback for the product path "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    # host buffers of the end-to-end leg on the GPU's own NUMA node
    prev_affinity = ssd.bind_host_to_gpu(local)
    packed = "auto" if args.packed_maps else False
    eng, survey = build_engine(device, NSIDE, math=args.math, fused=args.fused,
                               workload=args.workload, packed=packed, normals=args.normals,
                               map_dtype=args.maps)
    n = args.stars
    dtype = args.dtype
    out = eng.allocate(n, COLUMNS, dtype)
    counters = eng.new_counters()
    bps = BYTES_PER_STAR[dtype]

    def step(k):
        # counters accumulate LOCALLY; the one collective of the path (SUM all-reduce of the
        # 10 KB counter block) runs once, after the timed region
        eng.generate_observe(n, seed=args.seed, offset=(k * world + rank) * n, counters=counters,
                             dtype=dtype, out=out)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    for k in range(args.warmup):
        step(k)
    barrier()
    # ---- timed region: exactly K steps; one star-kernel launch per step
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * args.steps + 2)]
    barrier()
    t_wall0 = time.time()
    ev[0].record()
    for k in range(args.steps):
        ev[2 + 2 * k].record()
        eng.generate_observe(n, seed=args.seed, offset=((args.warmup + k) * world + rank) * n,
                             counters=counters, dtype=dtype, out=out)
        ev[3 + 2 * k].record()
    ev[1].record()
    barrier()
    t_wall1 = time.time()
    total_ms = ssd.max_over_ranks(ev[0].elapsed_time(ev[1]), device)
    kernel_ms = float(np.mean([ev[2 + 2 * k].elapsed_time(ev[3 + 2 * k]) for k in range(args.steps)]))
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    eng_packed = bool(eng.maps.packed)
    value = world * n * args.steps / (total_ms * 1e-3)
    total = ssd.reduce_counters(counters.clone())       # reduce ONCE, into a copy
    summary = eng.summarize(total)
    expect_total = world * n * (args.warmup + args.steps)
    if summary["n_total"] != expect_total or not (0 < summary["n_observed"] < summary["n_total"]):
        raise SystemExit(f"counter check failed: n_total {summary['n_total']} (expected "
                         f"{expect_total}), n_observed {summary['n_observed']}")

    # ---- supplementary: the other arithmetic mode of the production kernel, same workload
    other = None
    if not args.no_other_math:
        other_math = "fast" if args.math == "f64" else "f64"
        eng_m, _ = build_engine(device, NSIDE, survey=survey, math=other_math, fused=args.fused,
                                workload=args.workload, packed=packed)
        cm = eng_m.new_counters()
        for k in range(3):
            eng_m.generate_observe(n, seed=args.seed, offset=k * n, counters=cm, dtype=dtype, out=out)
        barrier()
        m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        m0.record()
        for k in range(args.steps):
            eng_m.generate_observe(n, seed=args.seed, offset=((3 + k) * world + rank) * n,
                                   counters=cm, dtype=dtype, out=out)
        m1.record()
        barrier()
        other_ms = ssd.max_over_ranks(m0.elapsed_time(m1), device) / args.steps
        other = {"math": other_math, "value": world * n / (other_ms * 1e-3), "unit": "stars/s",
                 "ms_per_step": other_ms, "note": MATH_NOTE[other_math]}
        del eng_m

    # ---- end to end through the C ABI with HOST buffers (pinned), D2H inside; same stars
    # per step as the device leg
    n_e2e = min(n, args.e2e_stars) if args.e2e_stars else n
    if args.no_e2e:
        n_e2e = 1 << 16
    chunk = min(n_e2e, args.e2e_chunk)
    host = eng.allocate_host(n_e2e, COLUMNS, dtype)
    ws = eng.host_workspace(chunk, dtype)
    dev_tabs = eng.model_tensors()
    host_tabs = [t.cpu().pin_memory() for t in dev_tabs]

    def e2e_step(k):
        # H2D of this step's inputs: the model tables (the path has no per-star input)
        for d, h in zip(dev_tabs, host_tabs):
            d.copy_(h, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        eng.generate_observe_host(n_e2e, host, ws, chunk, seed=args.seed,
                                  offset=(k * world + rank) * n_e2e, dtype=dtype)

    e2e_step(0)
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(1, min(args.steps, 3))
    for k in range(e2e_steps):
        e2e_step(1 + k)
    barrier()
    e2e_s = ssd.max_over_ranks(time.perf_counter() - t0, device)
    e2e_value = world * n_e2e * e2e_steps / e2e_s
    h2d = sum(h.numel() * h.element_size() for h in host_tabs)
    d2h = n_e2e * bps + 8 * int(host["counters"].numel())

    # ---- what the copies alone can do: the same 12 columns, device -> the same pinned
    # buffers, one cudaMemcpyAsync per column, (a) this rank alone, (b) all ranks at once
    def copy_columns():
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for c in COLUMNS:
            host[c].copy_(out[c][:n_e2e], non_blocking=True)
        e1.record()
        e1.synchronize()
        return e0.elapsed_time(e1) * 1e-3

    copy_columns()
    alone_s = None
    for r in range(world):                  # ranks take turns
        barrier()
        if r == rank:
            alone_s = min(copy_columns() for _ in range(2))
    # all ranks at once: every round starts at a barrier and counts as its slowest rank (a
    # rank's own best time would credit it with the bandwidth the finished ranks left free)
    together_s = None
    for _ in range(3):
        barrier()
        round_s = ssd.max_over_ranks(copy_columns(), device)
        together_s = round_s if together_s is None else min(together_s, round_s)
    barrier()
    alone_gbs = n_e2e * bps / alone_s / 1e9
    alone_min_gbs = -ssd.max_over_ranks(-alone_gbs, device)
    ceiling_gbs = world * n_e2e * bps / together_s / 1e9

    # a separate, labelled leg: the same call with fp32 output columns (45 B per star instead of
    # 89) -- NOT the headline, which stays on the reference's float64 columns
    e2e_f32 = None
    if dtype == "f64" and not args.no_e2e:
        del host, ws                      # one set of pinned buffers at a time
        host32 = eng.allocate_host(n_e2e, COLUMNS, "f32")
        ws32 = eng.host_workspace(chunk, "f32")

        def e2e32_step(k):
            eng.generate_observe_host(n_e2e, host32, ws32, chunk, seed=args.seed,
                                      offset=(k * world + rank) * n_e2e, dtype="f32")
        e2e32_step(0)
        barrier()
        t0 = time.perf_counter()
        for k in range(e2e_steps):
            e2e32_step(1 + k)
        barrier()
        s32 = ssd.max_over_ranks(time.perf_counter() - t0, device)
        v32 = world * n_e2e * e2e_steps / s32
        e2e_f32 = {"value": v32, "unit": "stars/s", "gbs": v32 * 45 / 1e9,
                   "note": "same call, fp32 output columns (45 B/star): a labelled extra, not the headline"}
        del host32, ws32

    if world > 1:
        dist.destroy_process_group()
    if rank != 0:
        return
    peak, peak_src = measured_peak()
    achieved = n * bps / (kernel_ms * 1e-3) / 1e9
    km = kernel_metrics(dtype, args.math) if args.fused else {}
    traffic = km.get("dram_bytes_per_star")
    line = {
        "metric": "observed stars/sec", "value": value, "unit": "stars/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": dtype,
        "data": "synthetic",
        "config": {"workload": WORKLOAD if args.workload == "pal5" else
                   "lsst_injection_nside4096_wholesky_stress", "stars_per_step_per_gpu": n,
                   "nside": NSIDE,
                   "stream": "config/pal5_config.yaml + synthetic isochrone" if args.workload == "pal5"
                   else "stress variant: flat density over phi1 in [-180,180], 40-deg Gaussian width",
                   "maps": "NON-PARITY OPTION: survey maps rounded to fp32, packed records, 16 B per pixel"
                   if args.maps == "f32" else
                   "packed (ebv, maglim_g, maglim_r) records, 32 B per pixel" if bool(eng_packed)
                   else "three separate float64 maps",
                   "rng": "philox4x32-10 (device)", "seed": args.seed,
                   "normals": "fp32 Box-Muller (MUFU), like curand_normal" if args.normals == "f32"
                   else "double-precision Box-Muller",
                   "math": args.math, "math_note": MATH_NOTE[args.math],
                   "l2": "no per-star inputs; outputs (%.1f GB/step) exceed L2 and are "
                         "rewritten every step; map/table gathers are L2-resident by design"
                         % (n * bps / 1e9),
                   "parallelism": f"stars sharded over {world} GPU(s), maps replicated",
                   "n_total": summary["n_total"], "n_observed": summary["n_observed"],
                   "observed_fraction": summary["n_observed"] / max(1, summary["n_total"])},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": None if traffic is None else traffic * n,
                     "traffic_source": "profiles/kernel_metrics.json (ncu capture of this kernel: DRAM "
                                       "bytes per star x stars per launch; a citation, not measured "
                                       "in this run)",
                     "issue_frac": km.get("issue_active_frac"),
                     "fp64_pipe_frac": km.get("fp64_pipe_frac"),
                     "lsu_data_pipe_frac": km.get("lsu_data_pipe_frac"),
                     "instructions_per_star": km.get("instructions_per_star"),
                     "l2_read_hit_rate": km.get("l2_read_hit_rate"),
                     "profile": km.get("source"),
                     "peak_source": peak_src,
                     "kernel": "k_fused" if args.fused else "k_stars<GENERATE|ROTATE|OBSERVE>",
                     "algorithmic_bytes_per_star": bps, "kernel_ms": kernel_ms},
        "e2e": {"value": e2e_value, "unit": "stars/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "stars_per_step_per_gpu": n_e2e,
                "steps": e2e_steps, "chunk_stars": chunk,
                "api": "ss_generate_observe_host (C ABI, pinned host buffers)",
                "gbs": e2e_value * bps / 1e9,
                "pcie_gbs": alone_min_gbs,
                "copy_ceiling_gbs": ceiling_gbs,
                "frac_of_ceiling": e2e_value * bps / 1e9 / ceiling_gbs,
                "ceiling_note": "plain cudaMemcpyAsync of the same 12 columns into the same pinned "
                                "buffers: pcie_gbs = slowest rank copying alone, copy_ceiling_gbs = "
                                "all ranks copying at once (whole job)",
                "fp32_columns": e2e_f32,
                "host_numa_bound": prev_affinity is not None},
        "gpu_launches": args.steps,
        "clocks": clocks,
    }
    if other is not None:
        other["roofline_frac"] = (other["value"] / world) * bps / 1e9 / peak
        line["other_math"] = other
    if not args.no_cpu:
        if prev_affinity is not None:
            os.sched_setaffinity(0, prev_affinity)      # the CPU arm gets every host core
        maps = {k: v.cpu().numpy() for k, v in (("ebv", survey.ebv_map), ("g", survey.maglim_maps["g"]),
                                                ("r", survey.maglim_maps["r"]))}
        del eng, survey, out
        torch.cuda.empty_cache()
        line["cpu_baseline"] = cpu_baseline_leg(maps, steps=3 if world == 1 else 1)
        line["cpu_baseline_reference"] = reference_as_written()
    emit(line)


# ------------------------------------------------------------------ census (parity at scale)
def _census_chunk(path):
    """Worker: the oracle on one chunk of exported draws, compared with the GPU's columns."""
    import streamsim_oracle as oracle
    z = np.load(path)
    draws = {k[2:]: z[k] for k in z.files if k.startswith("d_")}
    draws["u_phi2"], draws["z_dist"] = draws["z_phi2"], draws["u_dist"]
    ref = oracle.generate_observe(_CPU["cfg"], _CPU["iso"], _CPU["R"], _CPU["sv"], draws, nside=NSIDE)
    res = {"n": int(len(z["pix"])),
           "pix_diff": int((ref["pix"] != z["pix"]).sum()),
           "flag_diff": int((ref["flag_observed"] != z["flag_observed"].astype(bool)).sum()),
           "phi1_diff": int((ref["phi1"] != z["phi1"]).sum())}
    worst = 0.0
    for k in ("phi2", "dist", "mag_g", "mag_r", "ra", "dec", "mag_g_obs", "mag_r_obs", "magerr_g",
              "magerr_r"):
        a, b = z[k], ref[k]
        nan_a, nan_b = np.isnan(a), np.isnan(b)
        res.setdefault("nan_pattern_diff", 0)
        res["nan_pattern_diff"] += int((nan_a != nan_b).sum())
        ok = ~(nan_a | nan_b)
        if ok.any():
            worst = max(worst, float(np.max(np.abs(a[ok] - b[ok]) / np.maximum(np.abs(b[ok]), 1e-300))))
    res["max_rel_err"] = worst
    os.unlink(path)
    return res


def run_census(args):
    """Parity census at BASELINE size: N stars through (1) the production kernel and (2) the
    generic kernel with its draws exported -- byte-compared on the device -- and (3) the NumPy
    oracle fed with those draws on all host cores: the COUNT of differing pixel indices and
    flags (expected 0) is the evidence for the kernel's rewrites of the reference's chain
    (z = w_z / tt = lon * 2/pi instead of the degree round trip; one log instead of
    exp10 + log10; SNR cut decided on log10(error))."""
    import multiprocessing as mp
    import torch
    from helpers import LazyMap
    torch.cuda.set_device(0)
    device = torch.device("cuda", 0)
    eng, survey = build_engine(device, NSIDE, math=args.math, fused=1)
    gen, _ = build_engine(device, NSIDE, survey=survey, math=args.math, fused=0)
    _cpu_setup(dict(ebv=LazyMap("ebv", NSIDE), g=LazyMap("maglim_g", NSIDE), r=LazyMap("maglim_r", NSIDE)))
    cores = host_cores()
    chunk, total = 1 << 22, args.census
    shm = "/dev/shm" if os.path.isdir("/dev/shm") else "/tmp"
    ctx = mp.get_context("fork")
    pending, agg, kernel_diff, t0 = [], [], 0, time.perf_counter()
    with ctx.Pool(len(cores)) as pool:
        for lo in range(0, total, chunk):
            m = min(chunk, total - lo)
            a = eng.generate_observe(m, seed=args.seed, offset=lo, pix=True)
            b = gen.generate_observe(m, seed=args.seed, offset=lo, pix=True, export_draws=True)
            for k in a:
                kernel_diff += int((a[k].view(torch.uint8) != b[k].view(torch.uint8)).sum())
            d = gen.exported_draws(b)
            path = os.path.join(shm, f"ss_census_{os.getpid()}_{lo}.npz")
            np.savez(path, **{"d_" + k: d[k] for k in ("u_phi1", "z_phi2", "u_dist", "u_mass", "z_r",
                                                      "z_g", "u_det", "u_det2")},
                     **{k: v.cpu().numpy() for k, v in a.items()})
            pending.append(pool.apply_async(_census_chunk, (path,)))
            while len(pending) > 2 * len(cores):
                agg.append(pending.pop(0).get())
        agg += [p.get() for p in pending]
    out = {"census": {"stars": sum(r["n"] for r in agg), "math": args.math,
                      "fused_vs_generic_bytes_differing": kernel_diff,
                      "pix_differing": sum(r["pix_diff"] for r in agg),
                      "flag_observed_differing": sum(r["flag_diff"] for r in agg),
                      "phi1_differing": sum(r["phi1_diff"] for r in agg),
                      "nan_pattern_differing": sum(r["nan_pattern_diff"] for r in agg),
                      "max_rel_err_float_columns": max(r["max_rel_err"] for r in agg),
                      "oracle_cores": len(cores), "seconds": time.perf_counter() - t0,
                      "workload": WORKLOAD, "nside": NSIDE, "seed": args.seed}}
    emit(out)


_RESULT_FD = None


def emit(line):
    """The one JSON line, on the process's real stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_RESULT_FD, data)


def main():
    global _RESULT_FD
    # stdout carries exactly one JSON line.  Libraries write there too (NCCL prints its
    # version banner to fd 1, warnings from forked workers ...): keep the real stdout
    # aside for the result and point fd 1 at stderr for everything else.
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser(description=__doc__)
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--stars", type=int, default=100_000_000,
                    help="stars per step per GPU (BASELINE config 4: 1e8 stars)")
    ap.add_argument("--e2e-stars", type=int, default=0,
                    help="stars per step per GPU of the end-to-end leg (0 = same as --stars)")
    ap.add_argument("--e2e-chunk", type=int, default=1 << 22,
                    help="stars per chunk of the host pipeline (double-buffered)")
    ap.add_argument("--dtype", choices=("f64", "f32"), default="f64")
    ap.add_argument("--seed", type=int, default=12345)
    ap.add_argument("--impl", choices=("b200", "reference"), default="b200")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--math", choices=("f64", "fast"), default="f64",
                    help="arithmetic of the main leg (see MATH_NOTE)")
    ap.add_argument("--maps", choices=("f64", "f32"), default="f64",
                    help="f32: survey maps rounded to fp32 and gathered as 16-byte records -- a "
                         "labelled non-parity option (the reference's maps are float64)")
    ap.add_argument("--normals", choices=("f32", "f64"), default="f32",
                    help="f64: the stars' normal draws from a double-precision Box-Muller "
                         "(SSParams.normals_f64); the default is the fp32 one, like cuRAND's")
    ap.add_argument("--no-other-math", action="store_true",
                    help="skip the supplementary leg that times the other arithmetic mode")
    ap.add_argument("--census", type=int, default=0,
                    help="instead of timing: run N stars through the production kernel, the generic "
                         "kernel and the NumPy oracle and print the number of differing pixels/flags")
    ap.add_argument("--workload", choices=("pal5", "wholesky"), default="pal5",
                    help="pal5 = the BASELINE configuration (default); wholesky = the gather stress "
                         "variant of SURVEY 8d (supplementary, never the headline)")
    ap.add_argument("--packed-maps", type=int, default=1,
                    help="1: interleaved (ebv, maglim_g, maglim_r) map records (default); 0: separate maps")
    ap.add_argument("--fused", type=int, default=1,
                    help="1: production kernel k_fused (default); 0: the generic k_stars")
    ap.add_argument("--no-e2e", action="store_true",
                    help="profiling runs: shrink the end-to-end leg to a token size")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = max(args.warmup, 1) if args.impl == "reference" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    elif args.census:
        run_census(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
