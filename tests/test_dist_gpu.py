"""Two-rank NCCL test of the side-stream path of utils.comm.reduce_dict (round-1 advisor finding: the returned
tensors must be ordered after the collective on the caller's stream).  Needs two GPUs: skipped on a single-GPU box."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", device_id=dev)
    from simpledepthestimation_b200.utils.comm import reduce_dict

    side = torch.cuda.Stream(device=dev)
    ok = True
    for it in range(20):
        # a long-running producer on the current stream, so that an unordered read would see stale values
        x = torch.randn(2048, 2048, device=dev)
        for _ in range(4):
            x = x @ x.t() * 1e-3
        losses = {"rec_loss": (x.sum() * 0 + (rank + 1.0) * (it + 1)), "smooth_loss": (x.sum() * 0 + 10.0 * (rank + 1))}
        red = reduce_dict(losses, average=True, stream=side)
        if rank == 0:
            want = {"rec_loss": (it + 1) * (1.0 + 2.0) / 2, "smooth_loss": 10.0 * (1.0 + 2.0) / 2}
            ok &= all(abs(float(red[k]) - want[k]) < 1e-6 for k in want)   # .item() on the caller's stream
    out[rank] = ok
    dist.barrier()
    dist.destroy_process_group()


def test_reduce_dict_side_stream_is_ordered_before_the_callers_read():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    mgr = mp.get_context("spawn").Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, 29517, out), nprocs=2, join=True)
    assert out[0] and out[1]
