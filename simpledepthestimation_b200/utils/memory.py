"""Batch-dict device transfer with the reference's semantics (detectron2/utils/memory.py:13-25):
tensors and numpy arrays move to the device, containers recurse, everything else passes."""
import numpy as np
import torch


def to_cuda(data, device="cuda"):
    if isinstance(data, torch.Tensor):
        return data.to(device, non_blocking=True)
    if isinstance(data, np.ndarray):
        return torch.from_numpy(data).to(device, non_blocking=True)
    if isinstance(data, list):
        return [to_cuda(d, device) for d in data]
    if isinstance(data, tuple):
        return tuple(to_cuda(d, device) for d in data)
    if isinstance(data, dict):
        return {k: to_cuda(v, device) for k, v in data.items()}
    return data
