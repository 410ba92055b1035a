"""Closed-form per-pixel statement of the MonoDepth2 loss path WITH analytic gradients.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  This is the math the CUDA
kernels implement, written without autograd so that
every backward formula used on the GPU has a CPU twin that tests can compare with
autograd of oracle/port.py (which itself is pinned to the reference).  Works in
fp64 (default) or fp32; vectorised torch, CPU.

Reference lines: projection camera.py:125-163,172-178; sampling camera.py:184-197;
SSIM ssim_loss.py:34-53; photometric MonoDepth2.py:137-144; min/automask
MonoDepth2.py:87-101,116-124; smoothness smoothness_loss.py:62-80.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def _shift_sum3(p, h, w):
    return sum(p[..., dy:dy + h, dx:dx + w] for dy in range(3) for dx in range(3))


def box_reflect(z):
    """Sum over the 3x3 window of the reflect-padded plane (not divided by 9)."""
    h, w = z.shape[-2:]
    return _shift_sum3(F.pad(z, (1, 1, 1, 1), mode="reflect"), h, w)


def box_reflect_adjoint(c):
    """Adjoint of box_reflect: scatter window-centre coefficients back to pixels,
    folding the pad rows/cols (-1 -> 1, h -> h-2) onto the image."""
    h, w = c.shape[-2:]
    g = _shift_sum3(F.pad(c, (2, 2, 2, 2)), h + 2, w + 2)  # gradient on the padded domain
    rows = g[..., 1:-1, :].clone()
    rows[..., 1, :] += g[..., 0, :]
    rows[..., h - 2, :] += g[..., h + 1, :]
    out = rows[..., :, 1:-1].clone()
    out[..., :, 1] += rows[..., :, 0]
    out[..., :, w - 2] += rows[..., :, w + 1]
    return out


def ssim_terms(S, A, C1, C2):
    """Window statistics and SSIM for planes S, A [...,h,w]."""
    mx, my = box_reflect(S) / 9, box_reflect(A) / 9
    exx, eyy, exy = box_reflect(S * S) / 9, box_reflect(A * A) / 9, box_reflect(S * A) / 9
    sx, sy, sxy = exx - mx * mx, eyy - my * my, exy - mx * my
    n1, n2 = 2 * mx * my + C1, 2 * sxy + C2
    d1, d2 = mx * mx + my * my + C1, sx + sy + C2
    ssim = (n1 * n2) / (d1 * d2)
    return mx, my, n1, n2, d1, d2, ssim


def photometric(S, A, ssim_w, C1, C2):
    """pe [B,h,w] for S, A [B,3,h,w]."""
    l1 = (S - A).abs().mean(1)
    if ssim_w <= 0:
        return l1
    ssim = ssim_terms(S, A, C1, C2)[-1]
    return torch.clamp((1 - ssim) / 2, 0, 1).mean(1) * ssim_w + l1 * (1 - ssim_w)


def photometric_grad_S(S, A, g_pe, ssim_w, C1, C2):
    """d(sum g_pe*pe)/dS, [B,3,h,w]; g_pe [B,h,w]."""
    l1w = (1 - ssim_w) if ssim_w > 0 else 1.0
    gS = (l1w / 3) * torch.sign(S - A) * g_pe[:, None]
    if ssim_w > 0:
        mx, my, n1, n2, d1, d2, ssim = ssim_terms(S, A, C1, C2)
        half = (1 - ssim) / 2
        gate = ((half >= 0) & (half <= 1)).to(S.dtype)
        g_ssim = -0.5 * (ssim_w / 3) * g_pe[:, None] * gate
        d_mx = 2 * my * (n2 - n1) / (d1 * d2) - 2 * mx * ssim * (1 / d1 - 1 / d2)
        d_exx = -ssim / d2
        d_exy = 2 * n1 / (d1 * d2)
        a, b, c = g_ssim * d_mx / 9, g_ssim * 2 * d_exx / 9, g_ssim * d_exy / 9
        gS = gS + box_reflect_adjoint(a) + S * box_reflect_adjoint(b) + A * box_reflect_adjoint(c)
    return gS


def project(d, K, R, t):
    """d [B,h,w]; K,R [B,3,3]; t [B,3] or [B,3,h,w].  Returns dict of per-pixel terms."""
    B, h, w = d.shape
    ys, xs = torch.meshgrid(torch.arange(h, dtype=d.dtype), torch.arange(w, dtype=d.dtype), indexing="ij")
    Ki = K.clone()
    Ki[:, 0, 0], Ki[:, 1, 1] = 1 / K[:, 0, 0], 1 / K[:, 1, 1]
    Ki[:, 0, 2], Ki[:, 1, 2] = -K[:, 0, 2] / K[:, 0, 0], -K[:, 1, 2] / K[:, 1, 1]
    pix = torch.stack([xs, ys, torch.ones_like(xs)], 0)[None]  # [1,3,h,w]
    ray = torch.einsum("bij,zjhw->bihw", Ki, pix)  # K^-1 [x,y,1]
    P = ray * d[:, None]
    M = K @ R
    tt = t[:, :, None, None] if t.dim() == 2 else t
    tau = torch.einsum("bij,bjhw->bihw", K, tt.expand(B, 3, h, w))
    p = torch.einsum("bij,bjhw->bihw", M, P) + tau
    q = 1 / (p[:, 2] + 1e-6)
    X, Y, Z = p[:, 0] * q, p[:, 1] * q, p[:, 2]
    return dict(ray=ray, P=P, M=M, q=q, X=X, Y=Y, Z=Z)


def bilinear(img, X, Y):
    """img [B,C,h,w]; X,Y [B,h,w] unclamped pixel coords.  Returns sampled and the terms
    needed by the backward (clamped coords, taps)."""
    B, C, h, w = img.shape
    ix = torch.clamp(torch.nan_to_num(X), 0, w - 1)
    iy = torch.clamp(torch.nan_to_num(Y), 0, h - 1)
    x0, y0 = torch.floor(ix), torch.floor(iy)
    ax, ay = ix - x0, iy - y0
    x0i, y0i = x0.long(), y0.long()
    x1i, y1i = torch.clamp(x0i + 1, max=w - 1), torch.clamp(y0i + 1, max=h - 1)
    inx = (x0i + 1 <= w - 1).to(img.dtype)  # out-of-range taps contribute zero
    iny = (y0i + 1 <= h - 1).to(img.dtype)
    flat = img.reshape(B, C, h * w)

    def tap(yi, xi):
        return torch.gather(flat, 2, (yi * w + xi).reshape(B, 1, -1).expand(B, C, -1)).view(B, C, h, w)

    v00, v01 = tap(y0i, x0i), tap(y0i, x1i) * inx[:, None]
    v10, v11 = tap(y1i, x0i) * iny[:, None], tap(y1i, x1i) * (inx * iny)[:, None]
    ax_, ay_ = ax[:, None], ay[:, None]
    S = v00 * (1 - ax_) * (1 - ay_) + v01 * ax_ * (1 - ay_) + v10 * (1 - ax_) * ay_ + v11 * ax_ * ay_
    dSdx = (v01 - v00) * (1 - ay_) + (v11 - v10) * ay_
    dSdy = (v10 - v00) * (1 - ax_) + (v11 - v01) * ax_
    gate_x = (torch.isfinite(X) & (X >= 0) & (X <= w - 1)).to(img.dtype)
    gate_y = (torch.isfinite(Y) & (Y >= 0) & (Y <= h - 1)).to(img.dtype)
    return S, dSdx * gate_x[:, None], dSdy * gate_y[:, None]


def smoothness_fwd_bwd(d, A, g=1.0):
    """Edge-aware smoothness of one scale and its gradient w.r.t. depth using the
    1-homogeneity shortcut.  d [B,h,w], A [B,3,h,w]."""
    B, h, w = d.shape
    inv = 1 / d.clamp(min=1e-6)
    mbar = inv.mean((1, 2)).clamp(min=1e-6)
    wx = torch.exp(-(A[..., :, :-1] - A[..., :, 1:]).abs().mean(1))
    wy = torch.exp(-(A[..., :-1, :] - A[..., 1:, :]).abs().mean(1))
    dx = inv[:, :, :-1] - inv[:, :, 1:]
    dy = inv[:, :-1, :] - inv[:, 1:, :]
    nx, ny = B * h * (w - 1), B * (h - 1) * w
    Lb = ((dx * wx).abs().sum((1, 2)) / nx + (dy * wy).abs().sum((1, 2)) / ny) / mbar  # per-image loss
    loss = Lb.sum()
    # G = dL/dn (local), then dL/dinv = G/mbar - Lb/(h*w*mbar)
    G = torch.zeros_like(d)
    sx, sy = torch.sign(dx) * wx / nx, torch.sign(dy) * wy / ny
    G[:, :, :-1] += sx
    G[:, :, 1:] -= sx
    G[:, :-1, :] += sy
    G[:, 1:, :] -= sy
    g_inv = G / mbar[:, None, None] - (Lb / (h * w * mbar))[:, None, None]
    g_d = -inv * inv * g_inv * (d >= 1e-6).to(d.dtype)
    return loss, g * g_d


def mono_loss_fwd_bwd(img_pyr, src_pyr, K, depth, pose, ssim_w=0.85, C1=1e-4, C2=9e-4, automask=True,
                      smooth_w=1e-3, reduce="min", g_rec=1.0, g_smooth=1.0):
    """Full MonoDepth2 loss on a pre-built pyramid, with analytic gradients.

    img_pyr: list over scales of [B,3,h,w]; src_pyr: list over scales of list over sources;
    K [B,3,3] full-res intrinsics (scale 0 resolution); depth list of [B,1,h,w]; pose list of [B,4,4].
    Returns dict(rec_loss, smooth_loss, argmin (list uint8 [B,h,w]), grad_depth (list [B,1,h,w]),
    grad_pose (list [B,4,4]))."""
    n = len(depth)
    S_n = len(pose)
    H, W = img_pyr[0].shape[-2:]
    dt = depth[0].dtype
    rec, smooth = torch.zeros((), dtype=dt), torch.zeros((), dtype=dt)
    g_depth, argmins = [], []
    g_pose = [torch.zeros_like(T) for T in pose]
    for i in range(n):
        A, srcs, d = img_pyr[i], src_pyr[i], depth[i][:, 0]
        B, h, w = d.shape
        Ki = K.clone()
        Ki[:, 0, 0] *= w / W
        Ki[:, 0, 2] *= w / W
        Ki[:, 1, 1] *= h / H
        Ki[:, 1, 2] *= h / H
        N = B * h * w
        geo, warped, cand = [], [], []
        for j in range(S_n):
            pr = project(d, Ki, pose[j][:, :3, :3], pose[j][:, :3, 3])
            Sj, dSdx, dSdy = bilinear(srcs[j], pr["X"], pr["Y"])
            geo.append((pr, dSdx, dSdy))
            warped.append(Sj)
            cand.append(photometric(Sj, A, ssim_w, C1, C2))
            if automask:
                cand.append(photometric(srcs[j], A, ssim_w, C1, C2))
        stack = torch.stack(cand, 1)
        stride = 2 if automask else 1
        if reduce == "min":
            m, arg = stack.min(1)
            rec = rec + m.mean() / n
            argmins.append(arg.to(torch.uint8))
        else:
            rec = rec + stack.mean((0, 2, 3)).sum() / stack.shape[1] / n
            argmins.append(torch.zeros(B, h, w, dtype=torch.uint8))
        gd = torch.zeros_like(d)
        for j in range(S_n):
            k = j * stride
            if reduce == "min":
                g_pe = (arg == k).to(dt) * (g_rec / (n * N))
            else:
                g_pe = torch.full_like(d, g_rec / (n * N * stack.shape[1]))
            gS = photometric_grad_S(warped[j], A, g_pe, ssim_w, C1, C2)
            pr, dSdx, dSdy = geo[j]
            gX, gY = (gS * dSdx).sum(1), (gS * dSdy).sum(1)
            q, X, Y = pr["q"], pr["X"], pr["Y"]
            Xz, Yz = torch.nan_to_num(X, posinf=0.0, neginf=0.0), torch.nan_to_num(Y, posinf=0.0, neginf=0.0)
            fx, fy, sk = Ki[:, 0, 0], Ki[:, 1, 1], Ki[:, 0, 1]
            cx, cy = Ki[:, 0, 2], Ki[:, 1, 2]
            v = lambda s: s[:, None, None]  # noqa: E731
            # gradient w.r.t. the projective point p (pixel units), then camera-space form K^T g_p
            u0, u1 = gX * q, gY * q
            gc0 = v(fx) * u0
            gc1 = v(sk) * u0 + v(fy) * u1
            gc2 = -(gX * (Xz - v(cx)) + gY * (Yz - v(cy))) * q
            gc = torch.stack([gc0, gc1, gc2], 1)  # = K^T g_p, [B,3,h,w]
            R = pose[j][:, :3, :3]
            g_pose[j][:, :3, :3] += torch.einsum("bihw,bjhw->bij", gc, pr["P"])
            g_pose[j][:, :3, 3] += gc.sum((2, 3))
            gP = torch.einsum("bji,bjhw->bihw", R, gc)  # R^T gc
            gd = gd + (gP * pr["ray"]).sum(1)
        if smooth_w > 0:
            sw = (1.0 / 2 ** (n - i - 1)) * smooth_w / n
            ls, gs = smoothness_fwd_bwd(d, A, g_smooth * sw)
            smooth = smooth + ls * sw
            gd = gd + gs
        g_depth.append(gd[:, None])
    return dict(rec_loss=rec, smooth_loss=smooth, argmin=argmins, grad_depth=g_depth, grad_pose=g_pose)
