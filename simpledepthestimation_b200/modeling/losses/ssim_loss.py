"""SSIM / WeightedSSIM modules with the reference's constructor and call signatures
(detectron2/modeling/losses/ssim_loss.py:6-53, 56-111), backed by the sde_ssim_* CUDA entry points."""
from __future__ import annotations

import torch.nn as nn

from ...ops import ssim, weighted_ssim


class SSIM(nn.Module):
    """SSIM(C1, C2)(x, y) -> clamp((1 - ssim) / 2, 0, 1) per channel; 3x3 window on the reflect-padded
    images (ssim_loss.py:34-53).  kernel_size / stride other than 3 / 1 are not used by the reference."""

    def __init__(self, C1=1e-4, C2=9e-4, kernel_size=3, stride=1):
        super().__init__()
        if kernel_size != 3 or stride != 1:
            raise NotImplementedError("SSIM: the CUDA path implements the 3x3, stride-1 window the reference uses")
        self.C1, self.C2 = C1, C2

    def forward(self, x, y):
        return ssim(x, y, self.C1, self.C2)


class WeightedSSIM(nn.Module):
    """WeightedSSIM(C1, C2)(x, y, w) -> (clamp((1 - ssim) / 2, 0, 1), avg_w) (ssim_loss.py:84-111).
    C1 or C2 may be 'inf' / float('inf') to drop that factor (ssim_loss.py:97-105)."""

    def __init__(self, C1=1e-4, C2=9e-4, kernel_size=3, stride=1):
        super().__init__()
        if kernel_size != 3 or stride != 1:
            raise NotImplementedError("WeightedSSIM: the CUDA path implements the 3x3, stride-1 window the reference uses")
        self.C1, self.C2 = float(C1), float(C2)

    def forward(self, x, y, w):
        return weighted_ssim(x, y, w, self.C1, self.C2)
