"""Same-box timing of one build of libsde_loss.so (SDE_LIB_PATH selects it): cfg2 step / forward / backward in
microseconds, median over blocks of CUDA-event-timed launches.  usage: python tools/ab_step.py [tag] [cfg2|cfg3|cfg4|cfg5]
(cfg5 = the 96-sample global batch of configs[4] on one GPU)"""
import os
import sys
import statistics

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from simpledepthestimation_b200.functional import MonoLossPlan, MotionLossPlan  # noqa: E402
from simpledepthestimation_b200.geometry.camera import resize_img  # noqa: E402
from simpledepthestimation_b200.synthetic import euler_pose, mono_inputs, motion_inputs  # noqa: E402

tag = sys.argv[1] if len(sys.argv) > 1 else "lib"
cfg = sys.argv[2] if len(sys.argv) > 2 else "cfg2"
dev = torch.device("cuda", 0)


def cached(name, fn):
    path = f"/tmp/sde_ab_{name}.pt"
    if os.path.exists(path):
        return torch.load(path)
    v = fn()
    torch.save(v, path)
    return v


def blocks(fn, n_blocks=9, iters=50):
    for _ in range(10):
        fn(0)
    torch.cuda.synchronize()
    out = []
    for _ in range(n_blocks):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(iters):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        out.append(e0.elapsed_time(e1) / iters * 1e3)
    return statistics.median(out), min(out)


if cfg in ("cfg2", "cfg3", "cfg5"):
    B, H, W = {"cfg2": (12, 192, 640), "cfg3": (8, 320, 1024), "cfg5": (96, 192, 640)}[cfg]
    nsets = 3 if cfg != "cfg5" else 1
    sets = []
    for k in range(nsets):
        inp = cached(f"{cfg}_{k}", lambda: mono_inputs(B, H, W, 4, 2, seed=k))
        sizes = [tuple(d.shape[-2:]) for d in inp["depth"]]
        mv = lambda t: t.to(dev).contiguous()  # noqa: E731
        sets.append(([mv(resize_img(inp["img"], s)) for s in sizes], [[mv(resize_img(c, s)) for c in inp["ctx"]] for s in sizes],
                     [mv(d) for d in inp["depth"]], mv(inp["K"]), [mv(euler_pose(v)) for v in inp["pose_vec"]]))
    plan = MonoLossPlan(B, sizes, 2, (H, W), dev)
    ones = torch.ones(2, device=dev)
    losses = [torch.empty(2, device=dev) for _ in sets]
    gd = [[torch.empty_like(d) for d in s[2]] for s in sets]
    gp = [[torch.empty_like(p) for p in s[4]] for s in sets]
    warped = [plan.new_warped() for _ in sets]
    argm = [plan.forward(*s, out=losses[k], warped=warped[k])[1] for k, s in enumerate(sets)]

    def fwd(i):
        k = i % nsets
        plan.forward(*sets[k], out=losses[k], argmin_out=argm[k], warped=warped[k])

    def bwd(i):
        k = i % nsets
        plan.backward(*sets[k], argm[k], ones, gd[k], gp[k], warped=warped[k])

    def step(i):
        k = i % nsets
        plan.forward_backward(*sets[k], ones, out=losses[k], argmin_out=argm[k], grad_depth=gd[k], grad_pose=gp[k], warped=warped[k])
    st, f, b = blocks(step), blocks(fwd), blocks(bwd)
    print(f"{tag:>14s} {cfg} step {st[0]:7.1f} (min {st[1]:7.1f})  fwd {f[0]:6.1f} (min {f[1]:6.1f})  bwd {b[0]:6.1f} (min {b[1]:6.1f}) us"
          f"  loss {float(losses[0][0]):.7f} gd0sum {float(gd[0][0].double().abs().sum()):.9e} gp {float(gp[0][0].double().abs().sum()):.9e}")
else:
    B, H, W = 4, 1280, 1920
    mi = cached("cfg4", lambda: motion_inputs(B, H, W, seed=0))
    mv = lambda t: t.to(dev).contiguous()  # noqa: E731
    f1, f2, d1, d2, K4 = mv(mi["img1"]), mv(mi["img2"]), mv(mi["depth1"]), mv(mi["depth2"]), mv(mi["K"])
    pose4, mo = mv(euler_pose(mi["pose_vec"])), mv(mi["motion"])
    mplan = MotionLossPlan(B, (H, W), dev, 2, with_field=True)
    args4 = ([f1, f2], [f2, f1], [d1, d2], [d2, d1], K4, [pose4[:B].contiguous(), pose4[B:].contiguous()],
             [mo[:B].contiguous(), mo[B:].contiguous()])
    ml, gl = torch.empty(2, 4, device=dev), torch.ones(2, 4, device=dev)
    mgd, mgp = [torch.empty_like(d1), torch.empty_like(d2)], [torch.empty(B, 4, 4, device=dev) for _ in range(2)]
    mgf = [torch.empty_like(args4[6][0]) for _ in range(2)]
    mw = mplan.new_warped()

    def mfwd(i):
        mplan.forward(*args4, want_maps=False, out=ml, warped=mw)

    def mbwd(i):
        mplan.backward(*args4, gl, mgd, mgp, mgf, warped=mw)

    def mstep(i):
        mfwd(i)
        mbwd(i)
    st, f, b = blocks(mstep, 7, 10), blocks(mfwd, 7, 10), blocks(mbwd, 7, 10)
    print(f"{tag:>14s} cfg4 step {st[0]:7.1f} (min {st[1]:7.1f})  fwd {f[0]:6.1f}  bwd {b[0]:6.1f} us  losses {ml.flatten().tolist()[:3]}"
          f" gd {float(mgd[0].double().abs().sum()):.9e} gf {float(mgf[0].double().abs().sum()):.9e}")
