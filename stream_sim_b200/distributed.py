"""Multi-GPU sharding: one process per GPU, stars split by contiguous global
index range, maps and tables replicated, and a single SUM all-reduce of the
summary counters + histograms (about 10 KB) as the only collective
(SURVEY.md section 8e).  Because every draw is keyed by the global star index,
per-star results are identical for any world size.
"""
from __future__ import annotations

import os


def shard_range(n_total, rank, world_size):
    """Contiguous index range [lo, hi) of ``rank``: ceil(n/world) stars per rank,
    the last ranks possibly short or empty."""
    n_total, rank, world_size = int(n_total), int(rank), int(world_size)
    per = -(-n_total // world_size)
    lo = min(n_total, rank * per)
    return lo, min(n_total, lo + per)


def init_from_env(backend=None):
    """Initialise torch.distributed from RANK / WORLD_SIZE / MASTER_* (torchrun).
    Returns (rank, world_size, local_rank); a no-op single-process world if the
    variables are absent."""
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def bind_host_to_gpu(device_index):
    """Pin this process to the CPUs next to GPU ``device_index`` (NVML's CPU affinity of
    the device), so that the pinned host buffers it allocates afterwards land on the GPU's
    own NUMA node: with one process per GPU the device<->host copies of the ``*_host``
    entry points then stay off the inter-socket link.  Returns the previous affinity set
    (pass it to ``os.sched_setaffinity(0, ...)`` to undo), or None if nothing was changed
    (no NVML, no affinity information, not Linux)."""
    try:
        import pynvml
        import torch
        prev = os.sched_getaffinity(0)
        pynvml.nvmlInit()
        try:
            uuid = str(torch.cuda.get_device_properties(device_index).uuid)
            handle = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid if not uuid.startswith("GPU-") else uuid).encode())
        except Exception:
            handle = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (ncpu + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        cpus &= prev
        if not cpus or cpus == prev:
            return None
        os.sched_setaffinity(0, cpus)
        return prev
    except Exception:
        return None


def reduce_counters(counters):
    """In-place SUM all-reduce of a counters tensor (int64) over all ranks.

    Call it ONCE per run on locally accumulated counters (or on a clone): reducing a
    running total in place and then accumulating into it again multiplies the earlier
    steps by the world size at every further reduction."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        if counters.is_cuda and dist.get_backend() == "gloo":
            host = counters.cpu()                 # gloo reduces on the host
            dist.all_reduce(host, op=dist.ReduceOp.SUM)
            counters.copy_(host)
        else:
            dist.all_reduce(counters, op=dist.ReduceOp.SUM)
    return counters


def max_over_ranks(value, device=None):
    """Max of a Python float over ranks (used for device-timed durations)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64,
                     device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def generate_observe_sharded(engine, n_total, seed=0, counters=None, **kw):
    """Run the fused path for this rank's share of ``n_total`` stars and add the
    GLOBAL (all-rank) counters of this call to ``counters``.  The launch accumulates into
    a fresh local block which is reduced once and then added, so a ``counters`` tensor
    that is reused across calls stays a correct global running total on every rank.
    Returns (columns of the local shard, (lo, hi))."""
    import torch.distributed as dist
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    lo, hi = shard_range(n_total, rank, world)
    local = engine.new_counters() if counters is not None else None
    out = engine.generate_observe(hi - lo, seed=seed, offset=lo, counters=local, **kw)
    if counters is not None:
        counters += reduce_counters(local)
    return out, (lo, hi)
