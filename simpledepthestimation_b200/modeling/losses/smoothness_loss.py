"""smoothness_loss with the reference's signature (detectron2/modeling/losses/smoothness_loss.py:42-80),
backed by the sde_smoothness_* CUDA entry points."""
from __future__ import annotations

from ...ops import smoothness


def smoothness_loss(depth, image, reversed=False):
    """Edge-aware smoothness of the mean-normalised inverse depth.  depth [B,1,H,W], image [B,C,H,W] ->
    scalar.  `reversed` flips the sign of the finite differences (smoothness_loss.py:17-20,36-39), which the
    absolute values remove again, so it does not change the result."""
    return smoothness(depth, image)
