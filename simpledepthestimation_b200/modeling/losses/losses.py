"""silog_loss / variance_loss with the reference's names (detectron2/modeling/losses/losses.py:5-18), backed by the
sde_silog_loss_* / sde_variance_loss_* CUDA entry points."""
from __future__ import annotations

import torch.nn as nn

from ...ops import silog, variance


class silog_loss(nn.Module):
    """sqrt(mean(d^2) - variance_focus * mean(d)^2) * 10 with d = log(est) - log(gt) where gt > 1 (losses.py:5-13)."""

    def __init__(self, variance_focus):
        super().__init__()
        self.variance_focus = variance_focus

    def forward(self, depth_est, depth_gt):
        return silog(depth_est, depth_gt, self.variance_focus)


def variance_loss(depth):
    """1 / mean((depth / mean(depth) - 1)^2) over the whole tensor."""
    return variance(depth)
