"""MotionLearning regularisers with the reference's names and argument meaning
(detectron2/modeling/losses/motion_loss.py:7-64), backed by the sde_motion_* CUDA entry points.  They consume the
coords / occlusion maps the fused rgbd-consistency kernel emits."""
from __future__ import annotations

import torch

from ...ops import (motion_field_regularizers, motion_smoothness, motion_sparsity, motion_translation_consistency,
                    motion_translation_consistency_split)


def motion_consistency_loss(coords_A_in_B, mask, R_A2B, R_B2A, t_A2B, t_B2A):
    """Rotation / translation cycle consistency (motion_loss.py:7-48).

    coords_A_in_B [B,H,W,2] normalised (no gradient, :11), mask [B,1,H,W], R [B,3,3], t [B,3,H,W].
    Returns (rot_error, trans_error) scalars.  The translation term -- sample t_B2A where each A pixel lands in B,
    compose with (R_A2B, t_A2B), normalise -- is one CUDA launch forward and one backward (deterministic scatter
    into t_B2A); the rotation term is [B,3,3] arithmetic and stays in torch."""
    eye = torch.eye(3, device=R_A2B.device, dtype=R_A2B.dtype)[None]
    rot_err = ((R_A2B @ R_B2A - eye) ** 2).mean(dim=[1, 2])
    rot_scale = ((R_A2B - eye) ** 2).mean(dim=[1, 2]) + ((R_B2A - eye) ** 2).mean(dim=[1, 2])
    rot_error = (rot_err / (rot_scale + 1e-24)).mean()
    trans_error = motion_translation_consistency(coords_A_in_B, mask, R_A2B, t_A2B, t_B2A)
    return rot_error, trans_error


def motion_smoothness_loss_fn(motion_field, warp_around=False):
    """mean sqrt(1e-24 + dx^2 + dy^2) over [B,C,H-1,W-1] (motion_loss.py:51-55)."""
    return motion_smoothness(motion_field)


def motion_sparsity_loss_fn(motion_map):
    """L1/2-style sparsity with a detached per-(sample, channel) mean (motion_loss.py:58-64)."""
    return motion_sparsity(motion_map)


def _rotation_cycle(R_A2B, R_B2A):
    eye = torch.eye(3, device=R_A2B.device, dtype=R_A2B.dtype)[None]
    rot_err = ((R_A2B @ R_B2A - eye) ** 2).mean(dim=[1, 2])
    rot_scale = ((R_A2B - eye) ** 2).mean(dim=[1, 2]) + ((R_B2A - eye) ** 2).mean(dim=[1, 2])
    return (rot_err / (rot_scale + 1e-24)).mean()


def motion_consistency_loss_split(coords_A_in_B, mask, pose_A2B, pose_B2A, field_A2B=None, field_B2A=None):
    """motion_consistency_loss (motion_loss.py:7-48) on the pieces MotionLearningModel holds -- the [B,4,4] poses and
    the residual translation fields (or None) -- instead of the overall fields t = pose[:, :3, [3], None] + field of
    MotionLearning.py:143-147: the sum is formed per pixel inside the kernels.  Returns (rot_error, trans_error)."""
    R_ab = pose_A2B[:, :3, :3]
    trans = motion_translation_consistency_split(coords_A_in_B, mask, R_ab, pose_A2B, field_A2B, pose_B2A, field_B2A)
    return _rotation_cycle(R_ab, pose_B2A[:, :3, :3]), trans


def motion_field_regularizer_losses(pose_A2B, field_A2B):
    """(motion_smoothness_loss_fn(mn), motion_sparsity_loss_fn(mn)) for mn = field / sqrt(3 mean(t^2) + 1e-12) with
    t = pose[:, :3, 3] + field: lines 203-220 of MotionLearning.py for one direction, two fused launches."""
    return motion_field_regularizers(pose_A2B, field_A2B)
