"""Stress run of sde_mono_loss_step with tile-level dependencies: thousands of back-to-back steps on alternating inputs
with ONE set of kept planes; every step's outputs must carry the bits of the first step on the same input.  A missing
fence, a flag that is cleared too early / too late or a stale L1 line of the derivative planes would show up as a rare
mismatch.  Small shapes (warp grid smaller than the SM count: some SMs run no warp-kernel CTA between two backward
launches) and the bench shape.  usage: python tools/stress_flow.py [iterations]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from simpledepthestimation_b200.functional import MonoLossPlan  # noqa: E402
from simpledepthestimation_b200.geometry.camera import resize_img  # noqa: E402
from simpledepthestimation_b200.synthetic import euler_pose, mono_inputs  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
dev = torch.device("cuda", 0)
ones = torch.ones(2, device=dev)
for (B, H, W), streams in (((2, 96, 320), 1), ((1, 48, 160), 1), ((3, 50, 70), 1), ((12, 192, 640), 1), ((12, 192, 640), 2)):
    sets = []
    for seed in (71, 72, 73):
        inp = mono_inputs(B, H, W, 4, 2, seed=seed)
        sizes = [tuple(d.shape[-2:]) for d in inp["depth"]]
        mv = lambda t: t.to(dev).contiguous()  # noqa: E731
        sets.append(([mv(resize_img(inp["img"], s)) for s in sizes], [[mv(resize_img(c, s)) for c in inp["ctx"]] for s in sizes],
                     [mv(d) for d in inp["depth"]], mv(inp["K"]), [mv(euler_pose(v)) for v in inp["pose_vec"]]))
    plan = MonoLossPlan(B, sizes, 2, (H, W), dev, streams=streams)
    saved = plan.new_warped()
    outs = [[torch.empty(2, device=dev), [torch.empty(B, h, w, dtype=torch.uint8, device=dev) for h, w in sizes],
             [torch.empty_like(d) for d in s[2]], [torch.empty_like(p) for p in s[4]]] for s in sets]
    ref = []
    for k, s in enumerate(sets):
        plan.forward_backward(*s, ones, out=outs[k][0], argmin_out=outs[k][1], grad_depth=outs[k][2], grad_pose=outs[k][3], warped=saved)
        torch.cuda.synchronize()
        ref.append([t.clone() for t in [outs[k][0]] + outs[k][1] + outs[k][2] + outs[k][3]])
    bad = 0
    # check in chunks: steps are issued back to back, the outputs of the last step on every input are compared
    for it in range(iters):
        k = it % 3
        plan.forward_backward(*sets[k], ones, out=outs[k][0], argmin_out=outs[k][1], grad_depth=outs[k][2], grad_pose=outs[k][3], warped=saved)
        if it % 7 == 6 or it == iters - 1:
            torch.cuda.synchronize()
            for kk in range(3):
                cur = [outs[kk][0]] + outs[kk][1] + outs[kk][2] + outs[kk][3]
                if not all(torch.equal(a, b) for a, b in zip(cur, ref[kk])):
                    bad += 1
    print(f"{B}x{H}x{W} streams {streams} flow mask {os.environ.get('SDE_FLOW_MASK', 'default')}: {iters} steps, {bad} mismatching checks")
    assert bad == 0
print("ok")
