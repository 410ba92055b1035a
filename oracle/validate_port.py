"""Checks oracle/port.py against the real reference (build container only).

TEST INFRASTRUCTURE ONLY.  Usage: python -m oracle.validate_port
Prints relative differences of every loss and gradient in fp64 and fp32.
"""
import sys

import torch

from oracle import port, ref_import
from simpledepthestimation_b200.synthetic import euler_pose, mono_inputs, motion_inputs


def rel(a, b):
    a, b = torch.as_tensor(a, dtype=torch.float64), torch.as_tensor(b, dtype=torch.float64)
    return float((a - b).abs().max() / b.abs().max().clamp(min=1e-300))


def port_mono(inp, dtype, **kw):
    depth = [d.to(dtype).clone().requires_grad_() for d in inp["depth"]]
    vecs = [v.to(dtype).clone().requires_grad_() for v in inp["pose_vec"]]
    out = port.mono_loss(inp["img"].to(dtype), [c.to(dtype) for c in inp["ctx"]], inp["K"].to(dtype), depth,
                         [euler_pose(v) for v in vecs], **kw)
    sum(v for k, v in out.items() if "loss" in k).backward()
    res = {k: v.detach() for k, v in out.items() if "loss" in k}
    res["grad_depth"] = [d.grad for d in depth]
    res["grad_pose_vec"] = [v.grad for v in vecs]
    return res


def port_motion(inp, dtype, with_motion=True, **kw):
    d1 = inp["depth1"].to(dtype).clone().requires_grad_()
    d2 = inp["depth2"].to(dtype).clone().requires_grad_()
    vec = inp["pose_vec"].to(dtype).clone().requires_grad_()
    mo = inp["motion"].to(dtype).clone().requires_grad_()
    out = port.motion_loss(inp["img1"].to(dtype), inp["img2"].to(dtype), d1, d2, inp["K"].to(dtype),
                           euler_pose(vec), mo if with_motion else None, **kw)
    sum(v for k, v in out.items() if "loss" in k).backward()
    res = {k: v.detach() for k, v in out.items() if "loss" in k}
    res.update(grad_depth1=d1.grad, grad_depth2=d2.grad, grad_pose_vec=vec.grad,
               grad_motion=mo.grad if with_motion else None)
    return res


def main():
    worst = 0.0
    for dtype in (torch.float64, torch.float32):
        for (B, H, W, over) in [(2, 32, 64, {}), (1, 96, 320, {}), (2, 48, 80, dict(AUTOMASK=False)),
                                (2, 48, 80, dict(PHOTOMETRIC_REDUCE="mean"))]:
            inp = mono_inputs(B, H, W)
            ref = ref_import.run_mono(inp, dtype, **over)
            kw = dict(automask=over.get("AUTOMASK", True), reduce=over.get("PHOTOMETRIC_REDUCE", "min"))
            mine = port_mono(inp, dtype, **kw)
            for k in ("rec_loss", "smooth_loss"):
                r = rel(mine[k], ref[k]); worst = max(worst, r if dtype == torch.float64 else 0)
                print(f"mono {dtype} {B}x{H}x{W} {over} {k}: rel {r:.2e}")
            for i, (a, b) in enumerate(zip(mine["grad_depth"], ref["grad_depth"])):
                r = rel(a, b); worst = max(worst, r if dtype == torch.float64 else 0)
                print(f"   grad_depth[{i}] rel {r:.2e}")
            for i, (a, b) in enumerate(zip(mine["grad_pose_vec"], ref["grad_pose_vec"])):
                r = rel(a, b); worst = max(worst, r if dtype == torch.float64 else 0)
                print(f"   grad_pose[{i}] rel {r:.2e}")
        for wm in (True, False):
            inp = motion_inputs(2, 32, 64)
            ref = ref_import.run_motion(inp, dtype, with_motion=wm)
            mine = port_motion(inp, dtype, with_motion=wm)
            for k in ref:
                if k in mine and mine[k] is not None and torch.is_tensor(ref[k]):
                    r = rel(mine[k], ref[k]); worst = max(worst, r if dtype == torch.float64 else 0)
                    print(f"motion {dtype} with_motion={wm} {k}: rel {r:.2e}")
    print("worst fp64 rel diff:", worst)
    return 0 if worst < 1e-10 else 1


if __name__ == "__main__":
    sys.exit(main())
