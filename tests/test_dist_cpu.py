"""World-size-2 (gloo, CPU) check of the N>1 host logic: contiguous batch slices, per-rank mean
losses, `reduce_dict` averaging == the single-process batch loss.  The per-rank loss is the CPU
oracle here (no GPU in this tier); the arithmetic under test is the sharding and the reduction."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from simpledepthestimation_b200.utils import comm


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, B, out):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from helpers import port_mono_from_vec
    from simpledepthestimation_b200.synthetic import mono_inputs

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    inp = mono_inputs(B, 32, 64, seed=0)  # every rank builds the global batch and takes its slice
    sl = comm.shard_batch(B)
    mine = dict(img=inp["img"][sl], ctx=[c[sl] for c in inp["ctx"]], K=inp["K"][sl],
                depth=[d[sl] for d in inp["depth"]], pose_vec=[v[sl] for v in inp["pose_vec"]])
    res = port_mono_from_vec(mine, torch.float64)
    red = comm.reduce_dict({"rec_loss": res["rec_loss"].detach(), "smooth_loss": res["smooth_loss"].detach()})
    # gradient averaging as DDP does it: all-reduce(sum) / world on the (here: pose) gradients
    g = torch.cat([v.flatten() for v in res["grad_pose_vec"]]).clone()
    full_g = torch.zeros(B * 12, dtype=torch.float64)
    full_g.view(2, B, 6)[:, sl] = g.view(2, -1, 6) / world
    dist.all_reduce(full_g)
    if rank == 0:
        torch.save({"red": {k: float(v) for k, v in red.items()}, "grad": full_g, "slice": (sl.start, sl.stop)}, out)
    dist.destroy_process_group()


def test_shard_batch_slices():
    assert comm.shard_batch(96, 8, 3) == slice(36, 48)
    assert comm.shard_batch(12, 1, 0) == slice(0, 12)
    with pytest.raises(ValueError):
        comm.shard_batch(10, 4, 0)
    assert comm.get_world_size() == 1 and comm.get_rank() == 0 and comm.is_main_process()
    d = {"a": torch.tensor(1.0)}
    assert comm.reduce_dict(d) is d  # single process: passthrough, as in the reference


def test_two_rank_slices_reproduce_the_global_batch(tmp_path):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from helpers import port_mono_from_vec, rel_err
    from simpledepthestimation_b200.synthetic import mono_inputs

    B, world = 4, 2
    out = str(tmp_path / "rank0.pt")
    mp.spawn(_worker, args=(world, _free_port(), B, out), nprocs=world, join=True)
    got = torch.load(out)
    ref = port_mono_from_vec(mono_inputs(B, 32, 64, seed=0), torch.float64)
    assert got["slice"] == (0, 2)
    # reduce_dict carries the scalars in fp32, like the reference's loss dict
    assert abs(got["red"]["rec_loss"] - float(ref["rec_loss"])) < 2e-7 * float(ref["rec_loss"])
    assert abs(got["red"]["smooth_loss"] - float(ref["smooth_loss"])) < 2e-7 * float(ref["smooth_loss"])
    full = torch.cat([v.flatten() for v in ref["grad_pose_vec"]])
    assert rel_err(got["grad"], full) < 1e-10
