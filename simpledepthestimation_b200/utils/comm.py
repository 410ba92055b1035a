"""Rank helpers and the loss-scalar reduction of the data-parallel trainer, with the reference's
names and semantics (detectron2/utils/comm.py:27-50,235-263).  The loss path itself needs no
collective: every rank runs the fused kernels on its batch slice; only the logged loss scalars
(and DDP's gradient all-reduce, unchanged) cross GPUs."""
from __future__ import annotations

import torch
import torch.distributed as dist


def get_world_size() -> int:
    if not dist.is_available() or not dist.is_initialized():
        return 1
    return dist.get_world_size()


def get_rank() -> int:
    if not dist.is_available() or not dist.is_initialized():
        return 0
    return dist.get_rank()


def is_main_process() -> bool:
    return get_rank() == 0


def shard_batch(batch_size: int, world_size: int = None, rank: int = None) -> slice:
    """Contiguous slice of the global batch owned by `rank` (IMS_PER_BATCH // world_size per rank,
    divisibility enforced as in detectron2/data/build.py:74-81)."""
    world_size = get_world_size() if world_size is None else world_size
    rank = get_rank() if rank is None else rank
    if batch_size % world_size != 0:
        raise ValueError(f"batch size {batch_size} is not divisible by the number of ranks {world_size}")
    per = batch_size // world_size
    return slice(rank * per, (rank + 1) * per)


def reduce_dict(input_dict, average=True, stream=None):
    """Reduce a dict of scalar tensors to rank 0 (comm.py:235-263): keys are sorted, values stacked,
    one `dist.reduce`, divided by the world size on rank 0 when `average`.

    With `stream` (a torch.cuda.Stream) the 8-byte collective is enqueued on that side stream after the producing
    kernels, so it does not sit between two compute kernels (the reduction is logging-only,
    projects/MonoDepth2/train.py:95-98).  The current stream then waits for the side stream, so whatever the caller does
    with the returned tensors next (`.item()`, logging, arithmetic) is ordered after the reduction, and the stacked
    buffer is recorded on the side stream so the caching allocator cannot reuse it while the collective is pending."""
    world_size = get_world_size()
    if world_size < 2:
        return input_dict
    with torch.no_grad():
        names = sorted(input_dict.keys())
        values = torch.stack([input_dict[k].detach().float().reshape(()) for k in names], dim=0)
        if stream is not None:
            current = torch.cuda.current_stream()
            stream.wait_stream(current)
            with torch.cuda.stream(stream):
                dist.reduce(values, dst=0)
                if dist.get_rank() == 0 and average:
                    values /= world_size
            values.record_stream(stream)
            current.wait_stream(stream)
        else:
            dist.reduce(values, dst=0)
            if dist.get_rank() == 0 and average:
                values /= world_size
        return {k: v for k, v in zip(names, values)}
