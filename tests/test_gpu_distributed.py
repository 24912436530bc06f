"""The one collective of the path on real GPUs (run with -m gpu): ranks take contiguous
index ranges of the same global run, accumulate counters + histograms locally, and reduce
them once per call -- the reduced block must be byte-identical to a single-GPU
``run_summary`` over the whole index range, and the per-star columns of each shard to the
matching slice of a single-GPU run.  With >= 2 GPUs the ranks sit on different devices and
reduce with NCCL; on a one-GPU box both ranks share cuda:0 and reduce through gloo (no rank
ever waits for the other on the device: the reduction runs on the host)."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

N_TOTAL, STEPS, SEED = 3_000_001, 3, 77


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _engine(device):
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__))))
    from helpers import stream_config
    from stream_sim_b200.engine import StarEngine
    from stream_sim_b200.frames import GreatCircleFrame
    from stream_sim_b200.model import StreamModel
    from stream_sim_b200.surveys import Survey
    from stream_sim_b200.synthetic import NOTEBOOK_ENDPOINTS
    return StarEngine(stream=StreamModel(stream_config("pal5")), survey=Survey.synthetic(128, 512),
                      frame=GreatCircleFrame.from_endpoints(*NOTEBOOK_ENDPOINTS), hist=True,
                      perfect_galstarsep=True, device=device)


def _rank(rank, world, port, backend, out_dir):
    import torch
    import torch.distributed as dist
    os.chdir(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    local = rank if backend == "nccl" else 0
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(local),
                      MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from stream_sim_b200 import distributed as ssd
    torch.cuda.set_device(local)
    ssd.init_from_env(backend=backend)
    eng = _engine(f"cuda:{local}")
    total = eng.new_counters()
    for step in range(STEPS):                  # one counters tensor reused across steps
        out, (lo, hi) = ssd.generate_observe_sharded(eng, N_TOTAL, seed=SEED + step, counters=total,
                                                     columns=["phi1", "mag_r_obs", "flag_observed"])
    torch.cuda.synchronize()
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), counters=total.cpu().numpy(), lo=lo, hi=hi,
             phi1=out["phi1"].cpu().numpy(), mag_r_obs=out["mag_r_obs"].cpu().numpy(),
             flag=out["flag_observed"].cpu().numpy())
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_counters_equal_single_gpu(tmp_path, world):
    import torch
    import torch.multiprocessing as mp
    assert torch.cuda.is_available(), "these tests need a B200"
    ngpu = torch.cuda.device_count()
    backend = "nccl" if ngpu >= world else "gloo"
    if backend == "gloo" and world > 2:
        pytest.skip("the 4-rank case needs 4 GPUs (NCCL)")
    mp.spawn(_rank, args=(world, _free_port(), backend, str(tmp_path)), nprocs=world, join=True)
    eng = _engine("cuda:0")
    ref = eng.new_counters()
    for step in range(STEPS):
        eng.run_summary(N_TOTAL, seed=SEED + step, counters=ref)
    last = eng.generate_observe(N_TOTAL, seed=SEED + STEPS - 1,
                                columns=["phi1", "mag_r_obs", "flag_observed"])
    ref = ref.cpu().numpy()
    covered = 0
    for r in range(world):
        z = np.load(tmp_path / f"rank{r}.npz")
        assert np.array_equal(z["counters"], ref), f"rank {r}: reduced counters + histograms"
        lo, hi = int(z["lo"]), int(z["hi"])
        covered += hi - lo
        for key, col in (("phi1", "phi1"), ("mag_r_obs", "mag_r_obs"), ("flag", "flag_observed")):
            mine, whole = z[key], last[col][lo:hi].cpu().numpy()
            assert np.array_equal(mine.view(np.uint8), whole.view(np.uint8)), (r, key)
    assert covered == N_TOTAL
    from stream_sim_b200 import _lib
    assert ref[_lib.CNT_TOTAL] == STEPS * N_TOTAL
    assert 0 < ref[_lib.CNT_OBSERVED] < ref[_lib.CNT_TOTAL]
    assert ref[_lib.N_COUNTERS:_lib.N_COUNTERS + _lib.HIST_BINS].sum() == STEPS * N_TOTAL
