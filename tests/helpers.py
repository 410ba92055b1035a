"""Shared helpers of the parity tests: the oracle side runs on CPU (oracle/port.py in fp64 or
fp32), the product side goes through the C ABI on cuda:0, both on identical bits."""
import torch

from oracle import port
from simpledepthestimation_b200.synthetic import euler_pose, mono_inputs


def build_pyramid(inp, dtype=torch.float32):
    """CPU pyramid exactly as the reference builds it (resize_img, camera.py:40-46)."""
    img = inp["img"].to(dtype)
    ctx = [c.to(dtype) for c in inp["ctx"]]
    tgt = [port.resize_bilinear(img, d.shape[-2:]).contiguous() for d in inp["depth"]]
    src = [[port.resize_bilinear(c, d.shape[-2:]).contiguous() for c in ctx] for d in inp["depth"]]
    return tgt, src


def oracle_mono(inp, dtype=torch.float64, tgt32=None, src32=None, **kw):
    """Oracle loss + grads on the fp32 pyramid bits (cast up), so only the loss path differs."""
    if tgt32 is None:
        tgt32, src32 = build_pyramid(inp)
    depth = [d.to(dtype).clone().requires_grad_() for d in inp["depth"]]
    pose = [euler_pose(v.float()).to(dtype).requires_grad_() for v in inp["pose_vec"]]
    pyr = [(t.to(dtype), [s.to(dtype) for s in ss]) for t, ss in zip(tgt32, src32)]
    out = port.mono_loss(inp["img"].to(dtype), None, inp["K"].to(dtype), depth, pose, pyramid=pyr,
                         want_maps=True, **kw)
    total = out["rec_loss"] + out.get("smooth_loss", 0.0)
    total.backward()
    out["grad_depth"] = [d.grad for d in depth]
    out["grad_pose"] = [p.grad for p in pose]
    return out


def rel_err(a, b):
    a = torch.as_tensor(a, dtype=torch.float64).cpu()
    b = torch.as_tensor(b, dtype=torch.float64).cpu()
    return float((a - b).abs().max() / b.abs().max().clamp(min=1e-300))


# ------------------------------------------------------------------------------------------------
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False))


def golden_mono_inputs(g):
    """Rebuilds the input dict of a mono golden fixture (stored tensors, or regenerated from the
    seed when only the checksum is stored -- returns None if the regenerated bits differ)."""
    import hashlib

    B, H, W, seed = (int(v) for v in g["meta"])
    if "img" in g:
        t = torch.from_numpy
        n_ctx = sum(1 for k in g if k.startswith("ctx"))
        n_depth = sum(1 for k in g if k.startswith("depth") and k[5:].isdigit())
        return dict(img=t(g["img"]), ctx=[t(g[f"ctx{j}"]) for j in range(n_ctx)], K=t(g["K"]),
                    depth=[t(g[f"depth{i}"]) for i in range(n_depth)],
                    pose_vec=[t(g[f"pose_vec{j}"]) for j in range(n_ctx)])
    inp = mono_inputs(B, H, W, seed=seed, pose_scale=float(g["pose_scale"]))
    h = hashlib.sha256()
    for x in [inp["img"], *inp["ctx"], inp["K"], *inp["depth"], *inp["pose_vec"]]:
        h.update(np.ascontiguousarray(x.numpy()).tobytes())
    return inp if h.hexdigest() == str(g["input_sha256"]) else None


def port_mono_from_vec(inp, dtype=torch.float64, **kw):
    """oracle/port.py on the raw inputs (pyramid built inside, pose from the 6-DoF vector) --
    the exact computation of the reference's MonoDepth2Model.forward."""
    depth = [d.to(dtype).clone().requires_grad_() for d in inp["depth"]]
    vecs = [v.to(dtype).clone().requires_grad_() for v in inp["pose_vec"]]
    out = port.mono_loss(inp["img"].to(dtype), [c.to(dtype) for c in inp["ctx"]], inp["K"].to(dtype), depth,
                         [euler_pose(v) for v in vecs], want_maps=kw.get("reduce", "min") == "min", **kw)
    (out["rec_loss"] + out["smooth_loss"]).backward()
    out["grad_depth"] = [d.grad for d in depth]
    out["grad_pose_vec"] = [v.grad for v in vecs]
    return out


def gpu_mono_from_vec(inp, dev, pyr=None, **plan_kw):
    """The product path on `dev`: torch builds the pyramid and the pose matrices (as the trainer
    does), the fused CUDA loss does the rest.  Returns losses, argmin and gradients."""
    from simpledepthestimation_b200.functional import MonoLossPlan, mono_photometric_smoothness_loss
    from simpledepthestimation_b200.geometry.camera import resize_img

    B, _, H, W = inp["img"].shape
    g = lambda t: t.to(dev).contiguous()  # noqa: E731
    depth = [g(d).requires_grad_() for d in inp["depth"]]
    vecs = [g(v).requires_grad_() for v in inp["pose_vec"]]
    sizes = [tuple(d.shape[-2:]) for d in depth]
    # pyramid from the CPU resize so the oracle and the kernel see identical pyramid bits
    if pyr is None:
        tgt = [g(port.resize_bilinear(inp["img"], s)) for s in sizes]
        src = [[g(port.resize_bilinear(c, s)) for c in inp["ctx"]] for s in sizes]
    else:
        tgt, src = [g(t) for t in pyr[0]], [[g(x) for x in row] for row in pyr[1]]
    plan = MonoLossPlan(B, sizes, len(vecs), (H, W), dev, **plan_kw)
    rec, sm, argmin = mono_photometric_smoothness_loss(plan, tgt, src, depth, g(inp["K"]), [euler_pose(v) for v in vecs])
    (rec + sm).backward()
    torch.cuda.synchronize()
    return dict(rec_loss=rec.detach().cpu(), smooth_loss=sm.detach().cpu(), argmin=[a.cpu() for a in argmin],
                grad_depth=[d.grad.cpu() for d in depth], grad_pose_vec=[v.grad.cpu() for v in vecs])


def stable_mask(inp, tgt, src, scale, automask=True, margin=1e-3):
    """Decision-stable pixels of one scale from the fp64 oracle (SURVEY.md App. C protocol):
    bilinear cell, coordinate clamp, sign(S-A), SSIM clamp and argmin are all away from their
    switching points.  Argmin / SSIM-clamp instabilities are dilated by the 3x3 window."""
    import torch.nn.functional as F

    dt = torch.float64
    d = inp["depth"][scale].to(dt)
    B, _, h, w = d.shape
    H, W = inp["img"].shape[-2:]
    Ki = port.scale_K(inp["K"].to(dt), w / W, h / H)
    A = tgt[scale].to(dt)
    local = torch.ones(B, h, w, dtype=torch.bool)
    window = torch.ones(B, h, w, dtype=torch.bool)
    cands = []
    for j, v in enumerate(inp["pose_vec"]):
        T = euler_pose(v.float()).to(dt)
        X, Y, Z, _ = port.project(d, Ki, T[:, :3, :3], T[:, :3, 3][:, :, None, None])
        for c, n in ((X, w), (Y, h)):
            inside = (c > margin) & (c < n - 1 - margin)
            outside = (c < -margin) | (c > n - 1 + margin)
            fr = c - torch.floor(c)
            local &= outside | (inside & (fr > margin) & (fr < 1 - margin))
        S = port.view_synthesis(src[scale][j].to(dt), d, Ki, T[:, :3, :3], T[:, :3, 3][:, :, None, None])[0]
        local &= ((S - A).abs() > 1e-5).all(1)
        mx, my = port._box3_reflect(S), port._box3_reflect(A)
        sx, sy = port._box3_reflect(S * S) - mx * mx, port._box3_reflect(A * A) - my * my
        sxy = port._box3_reflect(S * A) - mx * my
        ssim = ((2 * mx * my + 1e-4) * (2 * sxy + 9e-4)) / ((mx * mx + my * my + 1e-4) * (sx + sy + 9e-4))
        hv = (1 - ssim) / 2
        window &= ((hv > 1e-4) & (hv < 1 - 1e-4)).all(1)
        cands.append(port.photometric_error(S, A))
        if automask:
            cands.append(port.photometric_error(src[scale][j].to(dt), A))
    stack = torch.cat(cands, 1)
    if stack.shape[1] > 1:
        top2 = stack.topk(2, dim=1, largest=False)[0]
        window &= (top2[:, 1] - top2[:, 0]) > 1e-5
    eroded = -F.max_pool2d(-window.to(dt)[:, None], 3, stride=1, padding=1)[:, 0] > 0.5
    return local & eroded


def gpu_mono_from_pose(depth, K, pose, tgt, src, full_size, dev, **plan_kw):
    """Product path on explicit pyramid and [B,4,4] pose tensors (all CPU fp32 inputs);
    gradients w.r.t. depth and the pose matrices."""
    from simpledepthestimation_b200.functional import MonoLossPlan, mono_photometric_smoothness_loss

    g = lambda t: t.to(dev).contiguous()  # noqa: E731
    d = [g(x).requires_grad_() for x in depth]
    p = [g(x).requires_grad_() for x in pose]
    sizes = [tuple(x.shape[-2:]) for x in d]
    plan = MonoLossPlan(d[0].shape[0], sizes, len(p), full_size, dev, **plan_kw)
    rec, sm, argmin = mono_photometric_smoothness_loss(plan, [g(t) for t in tgt], [[g(x) for x in row] for row in src],
                                                       d, g(K), p)
    (rec + sm).backward()
    torch.cuda.synchronize()
    return dict(rec_loss=rec.detach().cpu(), smooth_loss=sm.detach().cpu(), argmin=[a.cpu() for a in argmin],
                grad_depth=[x.grad.cpu() for x in d], grad_pose=[x.grad.cpu() for x in p])
