"""Golden vectors of silog_loss / disp_to_depth / pose_vec2mat from the UNMODIFIED reference (build container only).

TEST INFRASTRUCTURE ONLY.  Usage: python -m oracle.make_golden_ops   -> tests/golden/depth_ops.npz
Inputs are seeded; outputs and gradients come from the reference's own functions in fp64
(detectron2/modeling/losses/losses.py:5-13, detectron2/layers/depth_decoder.py:9-18,
detectron2/geometry/pose_utils.py:98-137).
"""
import importlib.util
import os

import numpy as np
import torch

from oracle import ref_import

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _disp_to_depth_fn():
    """depth_decoder.py imports the conv layers of the decoder next to disp_to_depth; load the file as a plain module."""
    path = os.path.join(ref_import.REF_ROOT, "detectron2", "layers", "depth_decoder.py")
    spec = importlib.util.spec_from_file_location("_ref_depth_decoder", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.disp_to_depth


def main():
    ns = ref_import.load()
    g = torch.Generator().manual_seed(7)
    out = {}
    # silog_loss: depth maps with part of the ground truth below the gt > 1 mask
    est = (torch.rand(2, 1, 24, 40, generator=g, dtype=torch.float64) * 60 + 0.5).requires_grad_()
    gt = torch.rand(2, 1, 24, 40, generator=g, dtype=torch.float64) * 80
    loss = ns.losses.silog_loss(0.85)(est, gt)
    (loss * 0.7).backward()
    out.update(silog_est=est.detach().numpy(), silog_gt=gt.numpy(), silog_loss=loss.detach().numpy(),
               silog_grad=est.grad.numpy())
    # disp_to_depth (DepthResNet: min_depth 0.1, MAX_DEPTH 80)
    disp = torch.rand(2, 1, 16, 24, generator=g, dtype=torch.float64).requires_grad_()
    scaled, depth = _disp_to_depth_fn()(disp, 0.1, 80.0)
    w1 = torch.rand(scaled.shape, generator=g, dtype=torch.float64)
    w2 = torch.rand(scaled.shape, generator=g, dtype=torch.float64)
    ((scaled * w1).sum() + (depth * w2).sum()).backward()
    out.update(disp=disp.detach().numpy(), disp_scaled=scaled.detach().numpy(), disp_depth=depth.detach().numpy(),
               disp_w1=w1.numpy(), disp_w2=w2.numpy(), disp_grad=disp.grad.numpy())
    # pose_vec2mat
    vec = (torch.randn(5, 6, generator=g, dtype=torch.float64) * torch.tensor([0.3, 0.1, 0.3, 0.4, 0.4, 0.4],
                                                                              dtype=torch.float64)).requires_grad_()
    T = ns.pose_utils.pose_vec2mat(vec)
    wT = torch.rand(T.shape, generator=g, dtype=torch.float64)
    (T * wT).sum().backward()
    out.update(pose_vec=vec.detach().numpy(), pose_mat=T.detach().numpy(), pose_w=wT.numpy(), pose_grad=vec.grad.numpy())
    path = os.path.join(ROOT, "tests", "golden", "depth_ops.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
