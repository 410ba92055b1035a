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
    """align_corners bilinear resize, identity when the size already matches (camera.py:40-46)."""
    if image.shape[-2] == dst_size[-2] and image.shape[-1] == dst_size[-1]:
        return image
    return F.interpolate(image, size=tuple(dst_size), mode=mode,
                         align_corners=True if mode != "nearest" else None)


def resize_img_avgpool(image, dst_size):
    """adaptive average pooling resize (camera.py:49-54)."""
    if image.shape[-2] == dst_size[-2] and image.shape[-1] == dst_size[-1]:
        return image
    return F.adaptive_avg_pool2d(image, tuple(dst_size))
