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
