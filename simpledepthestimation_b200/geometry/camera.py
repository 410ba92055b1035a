"""Pinhole-camera helpers with the reference's names and argument meaning
(detectron2/geometry/camera.py).  The per-pixel work (back-project / project / warp) lives in
the CUDA kernels; only the tiny [B,3,3] bookkeeping and the pyramid resize stay in torch."""
from __future__ import annotations

import torch
import torch.nn.functional as F


def scale_intrinsics(K, x_scale, y_scale):
    """In-place fx,cx *= x_scale ; fy,cy *= y_scale, returns K (camera.py:14-22)."""
    K[..., 0, 0] *= x_scale
    K[..., 1, 1] *= y_scale
    K[..., 0, 2] *= x_scale
    K[..., 1, 2] *= y_scale
    return K


def inv_intrinsics(K):
    """Formula inverse written into a clone of K (camera.py:25-37)."""
    assert K.dim() == 3
    Ki = K.clone()
    Ki[:, 0, 0] = 1.0 / K[:, 0, 0]
    Ki[:, 1, 1] = 1.0 / K[:, 1, 1]
    Ki[:, 0, 2] = -1.0 * K[:, 0, 2] / K[:, 0, 0]
    Ki[:, 1, 2] = -1.0 * K[:, 1, 2] / K[:, 1, 1]
    return Ki


def resize_img(image, dst_size, mode="bilinear"):
    """align_corners bilinear resize, identity when the size already matches (camera.py:40-46).

    CUDA fp32 data tensors (the image pyramid of the loss path, MonoDepth2.py:82,88) go through the
    sde_resize_bilinear kernel; tensors that need a gradient, other modes and host tensors (data
    loading) use F.interpolate, which is what the reference calls."""
    if image.shape[-2] == dst_size[-2] and image.shape[-1] == dst_size[-1]:
        return image
    if mode == "bilinear" and image.is_cuda and image.dtype == torch.float32 and not image.requires_grad:
        from ..ops import resize_bilinear
        return resize_bilinear(image, dst_size)
    return F.interpolate(image, size=tuple(dst_size), mode=mode,
                         align_corners=True if mode != "nearest" else None)


def resize_img_avgpool(image, dst_size):
    """adaptive average pooling resize, identity when the size already matches (camera.py:49-54).  CUDA fp32 tensors
    (frames, depths and motion fields of MotionLearning at NUM_SCALES > 1, MotionLearning.py:126-144) go through the
    sde_resize_avgpool kernels (differentiable); host tensors use F.adaptive_avg_pool2d as the reference does."""
    if image.shape[-2] == dst_size[-2] and image.shape[-1] == dst_size[-1]:
        return image
    if image.is_cuda and image.dtype == torch.float32:
        from ..ops import resize_avgpool
        return resize_avgpool(image, dst_size)
    return F.adaptive_avg_pool2d(image, tuple(dst_size))


def view_synthesis(image_B, depth_A, intrinsics, R_A_to_B, t_A_to_B):
    """Warps image_B into frame A (camera.py:166-202): back-project depth_A with K^-1, move by (R, t),
    project with K, clamp, bilinear sample.  image_B [B,C,H,W], depth_A [B,1,H,W], intrinsics [B,3,3] (at
    this size), R_A_to_B [B,3,3], t_A_to_B [B,3,1,1] or [B,3,H,W].  Returns (sampled [B,C,H,W],
    depth_in_B [B,1,H,W], coords [B,H,W,2] normalised (x,y), valid [B,1,H,W] bool).  Differentiable
    w.r.t. depth_A, R, t and image_B (deterministic scatter).  One CUDA launch forward, one backward."""
    from ..ops import view_synthesis as _vs
    return _vs(image_B, depth_A, intrinsics, R_A_to_B, t_A_to_B)
