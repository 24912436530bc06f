"""Multi-rank host logic on CPU: world_size-2 gloo process group.  The per-star
kernel itself needs a GPU; what is checked here is the sharding arithmetic and the
one collective of the path -- the SUM all-reduce of the counters/histograms."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from stream_sim_b200 import _lib


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n_total, out_dir):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                      MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from stream_sim_b200 import distributed as ssd
    r, w, _ = ssd.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    lo, hi = ssd.shard_range(n_total, r, w)
    # stand-in for the kernel's counter block: n_total and a histogram of the
    # global star index, which is what makes the reduction checkable
    counters = torch.zeros(_lib.COUNTERS_LEN, dtype=torch.int64)
    counters[_lib.CNT_TOTAL] = hi - lo
    idx = np.arange(lo, hi)
    hist = np.bincount(idx * _lib.HIST_BINS // max(n_total, 1), minlength=_lib.HIST_BINS)
    counters[_lib.N_COUNTERS:_lib.N_COUNTERS + _lib.HIST_BINS] = torch.from_numpy(hist)
    ssd.reduce_counters(counters)
    t = ssd.max_over_ranks(1.0 + rank)
    np.save(os.path.join(out_dir, f"rank{rank}.npy"), counters.numpy())
    assert t == float(world)
    dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [1000003, 5])
def test_two_rank_gloo_reduce(tmp_path, n_total):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), n_total, str(tmp_path)), nprocs=world, join=True)
    a = np.load(tmp_path / "rank0.npy")
    b = np.load(tmp_path / "rank1.npy")
    assert np.array_equal(a, b), "all ranks hold the reduced counters"
    assert a[_lib.CNT_TOTAL] == n_total
    hist = a[_lib.N_COUNTERS:_lib.N_COUNTERS + _lib.HIST_BINS]
    expect = np.bincount(np.arange(n_total) * _lib.HIST_BINS // n_total, minlength=_lib.HIST_BINS)
    assert np.array_equal(hist, expect)


class _FakeEngine:
    """Stands in for StarEngine on a CPU box: `generate_observe` only fills the counter
    block, with n_total and a histogram of the global star index."""

    def new_counters(self):
        return torch.zeros(_lib.COUNTERS_LEN, dtype=torch.int64)

    def generate_observe(self, n, seed=0, offset=0, counters=None, **kw):
        idx = np.arange(offset, offset + n)
        counters[_lib.CNT_TOTAL] += n
        counters[_lib.CNT_OBSERVED] += int(((idx + seed) % 5 == 0).sum())
        h = np.bincount(idx % _lib.HIST_BINS, minlength=_lib.HIST_BINS)
        counters[_lib.N_COUNTERS:_lib.N_COUNTERS + _lib.HIST_BINS] += torch.from_numpy(h)
        return {}


def _sharded_worker(rank, world, port, n_total, steps, out_dir):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                      MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from stream_sim_b200 import distributed as ssd
    ssd.init_from_env(backend="gloo")
    eng = _FakeEngine()
    total = eng.new_counters()
    for step in range(steps):         # ONE counters tensor reused across calls
        ssd.generate_observe_sharded(eng, n_total, seed=step, counters=total)
    np.save(os.path.join(out_dir, f"rank{rank}.npy"), total.numpy())
    dist.destroy_process_group()


def test_sharded_running_total_over_many_steps(tmp_path):
    """Regression for the round-1 bench: a counters tensor that is reused over many steps
    must stay the plain global sum (not grow by the world size at every reduction)."""
    world, n_total, steps = 2, 100001, 25
    mp.spawn(_sharded_worker, args=(world, _free_port(), n_total, steps, str(tmp_path)),
             nprocs=world, join=True)
    a, b = np.load(tmp_path / "rank0.npy"), np.load(tmp_path / "rank1.npy")
    assert np.array_equal(a, b)
    single = _FakeEngine()
    ref = single.new_counters()
    for step in range(steps):
        single.generate_observe(n_total, seed=step, offset=0, counters=ref)
    assert np.array_equal(a, ref.numpy())
    assert a[_lib.CNT_TOTAL] == steps * n_total and 0 < a[_lib.CNT_OBSERVED] < a[_lib.CNT_TOTAL]
