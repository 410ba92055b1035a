"""variance_loss with the reference's name (detectron2/modeling/losses/losses.py:16-18), backed by the
sde_variance_loss_* CUDA entry points.  silog_loss (losses.py:5-13) is the supervised loss; LOSS.SUPERVISED_WEIGHT
is 0 in every shipped config and the models reject it."""
from __future__ import annotations

from ...ops import variance


def variance_loss(depth):
    """1 / mean((depth / mean(depth) - 1)^2) over the whole tensor."""
    return variance(depth)
