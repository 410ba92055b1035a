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
