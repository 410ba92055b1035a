"""CPU oracle ("port"): the reference's view-synthesis loss path restated with the
same ATen op sequence, in torch, dtype-agnostic (fp32 or fp64), autograd for grads.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Citations are file:line in
the reference tree (zzzxxxttt/SimpleDepthEstimation).  Validated against the
real reference by oracle/validate_port.py and tests/test_oracle_golden.py.

Differences from the reference, all deliberate and documented in SURVEY.md:
  * translation given as [B,3,1,1] is broadcast to [B,3,h,w]
    (MonoDepth2.py:94 vs camera.py:169 -- NaN at HEAD otherwise);
  * the dead PhotometricLoss module (photometric_loss.py) is not restated.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------- geometry


def scale_K(K, sx, sy):
    """camera.py:14-22 -- fx,cx *= sx ; fy,cy *= sy (no half-pixel shift)."""
    K = K.clone()
    K[:, 0, 0] = K[:, 0, 0] * sx
    K[:, 0, 2] = K[:, 0, 2] * sx
    K[:, 1, 1] = K[:, 1, 1] * sy
    K[:, 1, 2] = K[:, 1, 2] * sy
    return K


def inv_K(K):
    """camera.py:25-37 -- formula inverse written into a clone of K (skew kept)."""
    Ki = K.clone()
    Ki[:, 0, 0] = 1.0 / K[:, 0, 0]
    Ki[:, 1, 1] = 1.0 / K[:, 1, 1]
    Ki[:, 0, 2] = -1.0 * K[:, 0, 2] / K[:, 0, 0]
    Ki[:, 1, 2] = -1.0 * K[:, 1, 2] / K[:, 1, 1]
    return Ki


def resize_bilinear(img, hw):
    """camera.py:40-46 -- align_corners bilinear, identity when sizes match."""
    if tuple(img.shape[-2:]) == tuple(hw):
        return img
    return F.interpolate(img, size=tuple(hw), mode="bilinear", align_corners=True)


def resize_avgpool(img, hw):
    """camera.py:49-54."""
    if tuple(img.shape[-2:]) == tuple(hw):
        return img
    return F.adaptive_avg_pool2d(img, tuple(hw))


def pixel_grid(B, h, w, like):
    """camera.py:57-122 -- [B,3,h,w] of (x, y, 1)."""
    ys, xs = torch.meshgrid(torch.arange(h, dtype=like.dtype, device=like.device),
                            torch.arange(w, dtype=like.dtype, device=like.device), indexing="ij")
    return torch.stack([xs, ys, torch.ones_like(xs)], 0).expand(B, 3, h, w)


def project(depth, K, R, t):
    """camera.py:125-163,172-178 -- back-project with K^-1, move by (R,t), project with K.

    depth [B,1,h,w]; K,R [B,3,3]; t [B,3,1,1] or [B,3,h,w].
    Returns X,Y (pixel units, unclamped), Z (unclamped), valid (bool), all [B,h,w]."""
    B, _, h, w = depth.shape
    pts = inv_K(K).bmm((pixel_grid(B, h, w, depth) * depth).reshape(B, 3, -1))
    tt = t.expand(-1, -1, h, w).reshape(B, 3, -1)
    proj = K.bmm(R).bmm(pts) + K.bmm(tt)
    X = proj[:, 0] / (proj[:, 2] + 1e-6)
    Y = proj[:, 1] / (proj[:, 2] + 1e-6)
    Z = proj[:, 2]
    valid = X.isfinite() & (X >= 0) & (X < w - 1) & Y.isfinite() & (Y >= 0) & (Y < h - 1) & (Z > 0)
    return X.view(B, h, w), Y.view(B, h, w), Z.view(B, h, w), valid.view(B, h, w)


def view_synthesis(image_B, depth_A, K, R, t):
    """camera.py:166-202.  Returns (sampled [B,C,h,w], depth_in_B [B,1,h,w] clamped at
    1e-5, coords [B,h,w,2] normalised (x,y), valid [B,1,h,w] bool)."""
    B, _, h, w = depth_A.shape
    X, Y, Z, valid = project(depth_A, K, R, t)
    Xs = torch.clamp(X.nan_to_num(), 0, w - 1)
    Ys = torch.clamp(Y.nan_to_num(), 0, h - 1)
    coords = torch.stack([2 * Xs / (w - 1) - 1.0, 2 * Ys / (h - 1) - 1.0], -1)
    sampled = F.grid_sample(image_B, coords, mode="bilinear", padding_mode="zeros", align_corners=True)
    return sampled, torch.clamp(Z, min=1e-5)[:, None], coords, valid[:, None]


# --------------------------------------------------------------------------- losses


def _box3_reflect(x):
    return F.avg_pool2d(F.pad(x, (1, 1, 1, 1), mode="reflect"), 3, stride=1)


def ssim_map(x, y, C1=1e-4, C2=9e-4):
    """ssim_loss.py:34-53 -- clamp((1-SSIM)/2, 0, 1) per channel, 3x3 reflect window."""
    mx, my = _box3_reflect(x), _box3_reflect(y)
    sx = _box3_reflect(x * x) - mx * mx
    sy = _box3_reflect(y * y) - my * my
    sxy = _box3_reflect(x * y) - mx * my
    num = (2 * mx * my + C1) * (2 * sxy + C2)
    den = (mx * mx + my * my + C1) * (sx + sy + C2)
    return torch.clamp((1.0 - num / den) / 2.0, 0.0, 1.0)


def weighted_ssim_map(x, y, wgt, C1=float("inf"), C2=9e-6):
    """ssim_loss.py:84-111 -- returns (clamp((1-ssim)/2,0,1), avg_w)."""
    C1, C2 = float(C1), float(C2)
    avg_w = F.avg_pool2d(wgt, 3, stride=1, padding=1)
    wp = wgt + 1e-2
    inv = 1.0 / (avg_w + 1e-2)

    def wavg(z):
        return _box3_reflect(z * wp) * inv

    mx, my = wavg(x), wavg(y)
    sx = wavg(x ** 2) - mx ** 2
    sy = wavg(y ** 2) - my ** 2
    sxy = wavg(x * y) - mx * my
    if C1 == float("inf"):
        num, den = 2 * sxy + C2, sx + sy + C2
    elif C2 == float("inf"):
        num, den = 2 * mx * my + C1, mx ** 2 + my ** 2 + C1
    else:
        num = (2 * sxy + C2) * (2 * mx * my + C1)
        den = (sx + sy + C2) * (mx ** 2 + my ** 2 + C1)
    return torch.clamp((1.0 - num / den) / 2.0, 0.0, 1.0), avg_w


def photometric_error(S, A, ssim_w=0.85, C1=1e-4, C2=9e-4, clip=0.0):
    """MonoDepth2.py:137-149 -- [B,1,h,w] = w*mean_c ssim + (1-w)*mean_c |S-A|."""
    pe = (S - A).abs().mean(1, True)
    if ssim_w > 0.0:
        pe = ssim_map(S, A, C1, C2).mean(1, True) * ssim_w + pe * (1 - ssim_w)
    if clip > 0.0:
        pe = torch.clamp(pe, max=float(pe.mean() + clip * pe.std()))
    return pe


def smoothness(depth, image):
    """smoothness_loss.py:42-80 -- edge-aware smoothness of mean-normalised inverse depth."""
    inv = 1.0 / depth.clamp(min=1e-6)
    n = inv / inv.mean(2, True).mean(3, True).clamp(min=1e-6)
    gx = n[..., :, :-1] - n[..., :, 1:]
    gy = n[..., :-1, :] - n[..., 1:, :]
    wx = torch.exp(-(image[..., :, :-1] - image[..., :, 1:]).abs().mean(1, True))
    wy = torch.exp(-(image[..., :-1, :] - image[..., 1:, :]).abs().mean(1, True))
    return (gx * wx).abs().mean() + (gy * wy).abs().mean()


def variance_loss(depth):
    """losses.py:16-18."""
    return 1.0 / ((depth / depth.mean() - 1.0) ** 2).mean()


def silog_loss(depth_est, depth_gt, variance_focus=0.85):
    """losses.py:5-13."""
    mask = depth_gt > 1.0
    d = torch.log(depth_est[mask]) - torch.log(depth_gt[mask])
    return torch.sqrt((d ** 2).mean() - variance_focus * (d.mean() ** 2)) * 10.0


def disp_to_depth(disp, min_depth, max_depth):
    """layers/depth_decoder.py:9-18: (scaled_disp, depth)."""
    min_disp = 1 / max_depth
    max_disp = 1 / min_depth
    scaled_disp = min_disp + (max_disp - min_disp) * disp
    return scaled_disp, 1 / scaled_disp


def pose_vec2mat(vec):
    """geometry/pose_utils.py:98-137: [B,6] (tx,ty,tz,rx,ry,rz) -> [B,4,4], R = Rx Ry Rz."""
    x, y, z = vec[:, 3], vec[:, 4], vec[:, 5]
    zero, one = torch.zeros_like(x), torch.ones_like(x)
    rz = torch.stack([z.cos(), -z.sin(), zero, z.sin(), z.cos(), zero, zero, zero, one], 1).view(-1, 3, 3)
    ry = torch.stack([y.cos(), zero, y.sin(), zero, one, zero, -y.sin(), zero, y.cos()], 1).view(-1, 3, 3)
    rx = torch.stack([one, zero, zero, zero, x.cos(), -x.sin(), zero, x.sin(), x.cos()], 1).view(-1, 3, 3)
    top = torch.cat([rx.bmm(ry).bmm(rz), vec[:, :3].unsqueeze(-1)], 2)
    bottom = torch.tensor([0.0, 0.0, 0.0, 1.0], dtype=vec.dtype, device=vec.device).expand(len(vec), 1, 4)
    return torch.cat([top, bottom], 1)


# --------------------------------------------------------------------------- MonoDepth2 loop


def mono_loss(img, ctx, K, depth, pose, ssim_w=0.85, C1=1e-4, C2=9e-4, clip=0.0, automask=True,
              smooth_w=1e-3, reduce="min", var_w=0.0, pyramid=None, want_maps=False):
    """MonoDepth2.py:68-125 -- multi-scale photometric + automask + min-reprojection + smoothness.

    img [B,3,H,W]; ctx list of S; K [B,3,3] (full res); depth list finest-first of
    [B,1,h_i,w_i]; pose list of S [B,4,4].  `pyramid` optionally supplies pre-built
    (target_i, [source_ij]) per scale.  Returns dict(rec_loss, smooth_loss[, var_loss],
    and with want_maps: argmin (list [B,h,w] int64), cand (list [B,2S,h,w]))."""
    n = len(depth)
    H, W = img.shape[-2:]
    out = {"rec_loss": 0.0}
    if smooth_w > 0.0:
        out["smooth_loss"] = 0.0
    if var_w > 0.0:
        out["var_loss"] = 0.0
    maps, cands = [], []
    for i, d in enumerate(depth):
        h, w = d.shape[-2:]
        scale_w = 1.0 / 2 ** (n - i - 1)
        if pyramid is None:
            A = resize_bilinear(img, (h, w))
            srcs = [resize_bilinear(c, (h, w)) for c in ctx]
        else:
            A, srcs = pyramid[i]
        Ki = scale_K(K, w / W, h / H)
        cand = []
        for src, T in zip(srcs, pose):
            S, _, _, _ = view_synthesis(src, d, Ki, T[:, :3, :3], T[:, :3, 3][:, :, None, None])
            cand.append(photometric_error(S, A, ssim_w, C1, C2, clip))
            if automask:
                cand.append(photometric_error(src, A, ssim_w, C1, C2, clip))
        if reduce == "min":
            stack = torch.cat(cand, 1)
            m, arg = stack.min(1, True)
            out["rec_loss"] = out["rec_loss"] + m.mean() / n
            if want_maps:
                maps.append(arg[:, 0])
                cands.append(stack.detach())
        elif reduce == "mean":
            out["rec_loss"] = out["rec_loss"] + sum(c.mean() for c in cand) / len(cand) / n
        else:
            raise NotImplementedError(reduce)
        if smooth_w > 0.0:
            out["smooth_loss"] = out["smooth_loss"] + smoothness(d, A) * scale_w * smooth_w / n
        if var_w > 0.0:
            out["var_loss"] = out["var_loss"] + variance_loss(d) * scale_w * var_w / n
    if want_maps:
        out["argmin"], out["cand"] = maps, cands
    return out


# --------------------------------------------------------------------------- MotionLearning


def rgbd_consistency(frame_A, frame_B, depth_A, depth_B, K, R, t, ssim_w=3.0, C1=float("inf"), C2=9e-6,
                     depth_l1_w=0.0):
    """MotionLearning.py:248-291."""
    out = {}
    sampled, z, coords, valid = view_synthesis(torch.cat([frame_B, depth_B], 1), depth_A, K, R, t)
    S_rgb, S_d = sampled[:, :3], sampled[:, 3:4]
    occ = (z < S_d).to(z.dtype) * valid.to(z.dtype)
    out["coords_A_in_B"], out["occlusion_mask"] = coords, occ
    norm = occ.sum([1, 2, 3]) + 1
    if depth_l1_w > 0:
        out["depth_l1_loss"] = (((S_d.detach() - z).abs() * occ).sum([1, 2, 3]) / norm).mean() * depth_l1_w
    out["rgb_l1_loss"] = ((S_rgb - frame_A).abs() * occ).mean()
    if ssim_w > 0.0:
        derr = (z - S_d) ** 2
        m2 = ((derr * occ).sum([1, 2, 3]) / norm + 1e-4).view(-1, 1, 1, 1)
        wgt = ((m2 / (derr + m2)) * valid.to(z.dtype)).detach()
        smap, avg_w = weighted_ssim_map(S_rgb, frame_A, wgt, C1, C2)
        out["depth_proximity_weight"] = wgt
        out["ssim_loss"] = (smap * avg_w).mean() * ssim_w * 0.5
    return out


def motion_consistency(coords, mask, R_ab, R_ba, t_ab, t_ba):
    """motion_loss.py:7-48 -- rotation / translation cycle consistency."""
    B, _, h, w = t_ab.shape
    t_hat = F.grid_sample(t_ba, coords.detach(), mode="bilinear", padding_mode="zeros", align_corners=True)
    eye = torch.eye(3, dtype=R_ab.dtype, device=R_ab.device)[None]
    rot_err = ((R_ab @ R_ba - eye) ** 2).mean([1, 2])
    rot_scale = ((R_ab - eye) ** 2).mean([1, 2]) + ((R_ba - eye) ** 2).mean([1, 2]) + 1e-24
    rot = (rot_err / rot_scale).mean()
    zero = torch.einsum("bij,bjhw->bihw", R_ab, t_hat) + t_ab
    terr = (zero ** 2).sum(1) / ((t_ab ** 2).sum(1) + (t_hat ** 2).sum(1) + 1e-24)
    return rot, (mask[:, 0] * terr).mean()


def motion_smoothness(m):
    """motion_loss.py:51-55."""
    gx = (m[..., :, 1:] - m[..., :, :-1])[:, :, 1:, :]
    gy = (m[..., 1:, :] - m[..., :-1, :])[:, :, :, 1:]
    return torch.sqrt(1e-24 + gx ** 2 + gy ** 2).mean()


def motion_sparsity(m):
    """motion_loss.py:58-64."""
    a = m.abs()
    mean = a.mean([2, 3], keepdim=True).detach()
    return (2 * mean * torch.sqrt(a / (mean + 1e-24) + 1)).mean()


def motion_loss(img1, img2, depth1, depth2, K, pose, motion=None, num_scales=1, ssim_w=3.0, C1=float("inf"),
                C2=9e-6, depth_l1_w=0.0, smooth_w=1e-3, motion_smooth_w=1.0, motion_sparsity_w=0.2,
                rot_cycle_w=1e-3, trans_cycle_w=5e-2, scale_normalize=False):
    """MotionLearning.py:118-241 -- two-frame bidirectional loss.  pose [2B,4,4] (1->2 then 2->1),
    motion [2B,3,H,W] or None.  Returns dict of scalar losses plus the last scale's maps."""
    B = img1.shape[0]
    P12, P21 = pose[:B], pose[B:]
    m12 = m21 = None
    if motion is not None:
        m12, m21 = motion[:B], motion[B:]
    L = {}

    def add(k, v):
        L[k] = L.get(k, 0.0) + v

    maps = {}
    for i in reversed(range(num_scales)):
        sw = 1.0 / 2 ** i
        h, w = int(depth1.shape[-2] * sw), int(depth1.shape[-1] * sw)
        f1, f2 = resize_avgpool(img1, (h, w)), resize_avgpool(img2, (h, w))
        Ki = scale_K(K, sw, sw)
        d1, d2 = resize_avgpool(depth1, (h, w)), resize_avgpool(depth2, (h, w))
        R12, R21 = P12[:, :3, :3], P21[:, :3, :3]
        t12, t21 = P12[:, :3, 3][:, :, None, None], P21[:, :3, 3][:, :, None, None]
        if motion is not None:
            r12, r21 = resize_avgpool(m12, (h, w)), resize_avgpool(m21, (h, w))
            t12, t21 = t12 + r12, t21 + r21
        else:
            t12, t21 = t12.expand(-1, -1, h, w), t21.expand(-1, -1, h, w)
        if scale_normalize:
            dm = torch.cat([d1, d2], 0).mean()
            d1n, d2n, t12, t21 = d1 / dm, d2 / dm, t12 / dm, t21 / dm
            if motion is not None:
                r12, r21 = r12 / dm, r21 / dm
        else:
            d1n, d2n = d1, d2
        o12 = rgbd_consistency(f1, f2, d1n, d2n, Ki, R12, t12, ssim_w, C1, C2, depth_l1_w)
        o21 = rgbd_consistency(f2, f1, d2n, d1n, Ki, R21, t21, ssim_w, C1, C2, depth_l1_w)
        for o in (o12, o21):
            for k, v in o.items():
                if "loss" in k:
                    add(k, v * sw)
        if rot_cycle_w > 0 or trans_cycle_w > 0:
            for (o, Ra, Rb, ta, tb) in ((o12, R12, R21, t12, t21), (o21, R21, R12, t21, t12)):
                rot, tr = motion_consistency(o["coords_A_in_B"], o["occlusion_mask"], Ra, Rb, ta, tb)
                add("rot_loss", rot * sw * rot_cycle_w)
                add("trans_loss", tr * sw * trans_cycle_w)
        if motion is not None:
            for r, t in ((r12, t12), (r21, t21)):
                mn = r / torch.sqrt(t.pow(2).mean([1, 2, 3], keepdim=True) * 3.0 + 1e-12)
                if motion_smooth_w > 0.0:
                    add("motion_smooth_loss", motion_smoothness(mn) * sw * motion_smooth_w)
                if motion_sparsity_w > 0.0:
                    add("motion_sparsity_loss", motion_sparsity(mn) * sw * motion_sparsity_w)
        if smooth_w > 0.0:
            add("smooth_loss", smoothness(d1n, f1) * sw * smooth_w)
            add("smooth_loss", smoothness(d2n, f2) * sw * smooth_w)
        maps = {"o12": o12, "o21": o21}
    L["_maps"] = maps
    return L
