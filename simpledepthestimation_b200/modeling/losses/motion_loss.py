"""MotionLearning regularisers with the reference's names and argument meaning
(detectron2/modeling/losses/motion_loss.py:7-64).  Row N1 of SURVEY.md section 8(f): they consume the
coords / occlusion maps the fused rgbd-consistency kernel emits; the arithmetic here is plain
torch on the GPU (autograd) until the fused kernels for them land."""
from __future__ import annotations

import torch
import torch.nn.functional as F


def motion_consistency_loss(coords_A_in_B, mask, R_A2B, R_B2A, t_A2B, t_B2A):
    """Rotation / translation cycle consistency (motion_loss.py:7-48).

    coords_A_in_B [B,H,W,2] normalised, mask [B,1,H,W], R [B,3,3], t [B,3,H,W].
    Returns (rot_error, trans_error) scalars."""
    # the B->A translation seen from where each A pixel lands in B (motion_loss.py:11-12)
    t_hat = F.grid_sample(t_B2A, coords_A_in_B.detach(), mode="bilinear", padding_mode="zeros", align_corners=True)
    eye = torch.eye(3, device=R_A2B.device, dtype=R_A2B.dtype)[None]
    # composing A->B after B->A should give the identity: rotation R_A2B R_B2A, translation R_A2B t_hat + t_A2B
    rot_err = ((R_A2B @ R_B2A - eye) ** 2).mean(dim=[1, 2])
    rot_scale = ((R_A2B - eye) ** 2).mean(dim=[1, 2]) + ((R_B2A - eye) ** 2).mean(dim=[1, 2])
    rot_error = (rot_err / (rot_scale + 1e-24)).mean()
    residual = torch.einsum("bij,bjhw->bihw", R_A2B, t_hat) + t_A2B
    trans_err = (residual ** 2).sum(1) / ((t_A2B ** 2).sum(1) + (t_hat ** 2).sum(1) + 1e-24)
    trans_error = (mask[:, 0] * trans_err).mean()
    return rot_error, trans_error


def motion_smoothness_loss_fn(motion_field, warp_around=False):
    """mean sqrt(1e-24 + dx^2 + dy^2) over [B,3,H-1,W-1] (motion_loss.py:51-55)."""
    dx = (motion_field[:, :, :, 1:] - motion_field[:, :, :, :-1])[:, :, 1:, :]
    dy = (motion_field[:, :, 1:, :] - motion_field[:, :, :-1, :])[:, :, :, 1:]
    return torch.sqrt(1e-24 + dx ** 2 + dy ** 2).mean()


def motion_sparsity_loss_fn(motion_map):
    """L1/2-style sparsity with a detached per-(sample, channel) mean (motion_loss.py:58-64)."""
    a = motion_map.abs()
    mean = a.mean([2, 3], keepdim=True).detach()
    return (2 * mean * torch.sqrt(a / (mean + 1e-24) + 1)).mean()
