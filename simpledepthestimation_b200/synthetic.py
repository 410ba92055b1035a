"""Seeded synthetic KITTI-/Waymo-shaped inputs for the view-synthesis loss path.

Follows the spec of SURVEY.md Appendix D / section 8(d).  Everything is produced
with numpy's PCG64 generator and plain float64 arithmetic (then cast to fp32) so
that the CPU oracle and the CUDA path see bit-identical tensors; nothing here
depends on a GPU.

Shapes follow the reference's batch-dict contract
(detectron2/modeling/meta_arch/MonoDepth2.py:68-71): ``img_orig`` [B,3,H,W] in
[0,1], ``ctx_img_orig`` list of S such tensors, ``intrinsics`` [B,3,3],
``depth_pred`` list (finest first) of [B,1,h_i,w_i] and ``pose_pred`` list of S
[B,4,4] matrices (detectron2/modeling/pose_net/PoseNet.py:59-63).
"""
from __future__ import annotations

import numpy as np
import torch

__all__ = ["euler_pose", "mono_inputs", "motion_inputs", "kitti_intrinsics", "CONFIGS"]

# BASELINE.json configs -> concrete shapes
CONFIGS = {
    "cfg1": dict(kind="mono", B=1, H=192, W=640, scales=4, S=2),
    "cfg2": dict(kind="mono", B=12, H=192, W=640, scales=4, S=2),
    "cfg3": dict(kind="mono", B=8, H=320, W=1024, scales=4, S=2),
    "cfg4": dict(kind="motion", B=4, H=1280, W=1920, scales=1, S=1),
    "cfg5": dict(kind="mono", B=96, H=192, W=640, scales=4, S=2),
}


def _upsample(a: np.ndarray, H: int, W: int) -> np.ndarray:
    """Separable linear interpolation of a [...,h,w] float64 array to [...,H,W]
    (align-corners convention).  Written in numpy so the bits do not depend on
    a torch kernel version."""
    h, w = a.shape[-2:]
    ys = np.linspace(0.0, h - 1.0, H)
    xs = np.linspace(0.0, w - 1.0, W)
    y0 = np.clip(np.floor(ys).astype(np.int64), 0, h - 2)
    x0 = np.clip(np.floor(xs).astype(np.int64), 0, w - 2)
    fy = (ys - y0)[:, None]
    fx = xs - x0
    top = a[..., y0, :]
    bot = a[..., y0 + 1, :]
    rows = top * (1.0 - fy) + bot * fy
    left = rows[..., x0]
    right = rows[..., x0 + 1]
    return left * (1.0 - fx) + right * fx


def _smooth(a: np.ndarray) -> np.ndarray:
    """3-tap binomial blur along the last two axes (edge replicated)."""
    p = np.pad(a, [(0, 0)] * (a.ndim - 2) + [(1, 1), (1, 1)], mode="edge")
    v = 0.25 * p[..., :-2, :] + 0.5 * p[..., 1:-1, :] + 0.25 * p[..., 2:, :]
    return 0.25 * v[..., :, :-2] + 0.5 * v[..., :, 1:-1] + 0.25 * v[..., :, 2:]


def kitti_intrinsics(B: int, H: int, W: int) -> torch.Tensor:
    """KITTI P_rect after the reference's Resize."""
    K = np.array([[0.58 * W, 0.0, 0.49 * W], [0.0, 1.92 * H, 0.46 * H], [0.0, 0.0, 1.0]])
    return torch.from_numpy(np.broadcast_to(K, (B, 3, 3)).astype(np.float32).copy())


def euler_pose(vec: torch.Tensor) -> torch.Tensor:
    """[tx,ty,tz,rx,ry,rz] -> [B,4,4] with R = Rx @ Ry @ Rz.

    Same convention as the reference's pose_vec2mat/euler2mat
    (detectron2/geometry/pose_utils.py:98-137); differentiable, dtype-agnostic.
    """
    t, r = vec[:, :3], vec[:, 3:]
    cx, cy, cz = torch.cos(r[:, 0]), torch.cos(r[:, 1]), torch.cos(r[:, 2])
    sx, sy, sz = torch.sin(r[:, 0]), torch.sin(r[:, 1]), torch.sin(r[:, 2])
    o, z = torch.ones_like(cx), torch.zeros_like(cx)
    Rx = torch.stack([o, z, z, z, cx, -sx, z, sx, cx], 1).view(-1, 3, 3)
    Ry = torch.stack([cy, z, sy, z, o, z, -sy, z, cy], 1).view(-1, 3, 3)
    Rz = torch.stack([cz, -sz, z, sz, cz, z, z, z, o], 1).view(-1, 3, 3)
    R = Rx @ Ry @ Rz
    top = torch.cat([R, t[:, :, None]], 2)
    bottom = torch.tensor([0.0, 0.0, 0.0, 1.0], dtype=vec.dtype, device=vec.device).expand(len(vec), 1, 4)
    return torch.cat([top, bottom], 1)


def _frames(rng, B, H, W, shifts):
    coarse = rng.random((B, 3, H // 16 + 2, W // 16 + 2))
    target = np.clip(0.8 * _upsample(_smooth(coarse), H, W) + 0.2 * rng.random((B, 3, H, W)), 0.0, 1.0)
    sources = []
    for s in shifts:
        src = np.clip(0.9 * np.roll(target, s, axis=-1) + 0.1 * rng.random((B, 3, H, W)), 0.0, 1.0)
        sources.append(src)
    return target, sources


def _depth(rng, B, h, w, lo=None):
    base = rng.standard_normal((B, 1, h // 8 + 2, w // 8 + 2))
    logits = 1.5 * _upsample(base, h, w) - 2.0 + 0.05 * rng.standard_normal((B, 1, h, w))
    disp = 1.0 / 80.0 + (10.0 - 1.0 / 80.0) / (1.0 + np.exp(-logits))
    d = 1.0 / disp
    if lo is not None:
        d = np.maximum(d * (lo / 0.1), lo)
    return d


def _t32(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))


def mono_inputs(B=1, H=192, W=640, scales=4, S=2, seed=0, pose_scale=1.0, min_depth=None):
    """MonoDepth2-path inputs.  Returns a dict of CPU fp32 tensors:
    img, ctx (list S), K [B,3,3], depth (list `scales`), pose_vec (list S of [B,6])."""
    rng = np.random.Generator(np.random.PCG64(seed))
    shifts = [3, -3, 5, -5][:S]
    target, sources = _frames(rng, B, H, W, shifts)
    depth = [_depth(rng, B, H >> i, W >> i, min_depth) for i in range(scales)]
    amp = 0.01 * pose_scale * np.array([10.0, 2.0, 10.0, 1.0, 1.0, 1.0])
    vecs = [rng.standard_normal((B, 6)) * amp for _ in range(S)]
    return dict(
        img=_t32(target), ctx=[_t32(s) for s in sources], K=kitti_intrinsics(B, H, W),
        depth=[_t32(d) for d in depth], pose_vec=[_t32(v) for v in vecs],
    )


def motion_inputs(B=1, H=128, W=416, seed=0, pose_scale=1.0):
    """MotionLearning-path inputs (two frames, both directions).  Returns CPU fp32:
    img1, img2 [B,3,H,W]; depth1, depth2 [B,1,H,W]; K; pose_vec [2B,6] (1->2 then
    2->1, the reference's chunk order, MotionLearning.py:103); motion [2B,3,H,W]."""
    rng = np.random.Generator(np.random.PCG64(seed))
    target, sources = _frames(rng, B, H, W, [3])
    d1 = _depth(rng, B, H, W)
    d2 = d1 * (1.0 + 0.05 * rng.standard_normal((B, 1, H, W)))
    amp = 0.01 * pose_scale * np.array([10.0, 2.0, 10.0, 1.0, 1.0, 1.0])
    vec = rng.standard_normal((2 * B, 6)) * amp
    mo = 0.02 * _upsample(rng.standard_normal((2 * B, 3, H // 8 + 2, W // 8 + 2)), H, W)
    return dict(
        img1=_t32(target), img2=_t32(sources[0]), depth1=_t32(d1), depth2=_t32(d2),
        K=kitti_intrinsics(B, H, W), pose_vec=_t32(vec), motion=_t32(mo),
    )
