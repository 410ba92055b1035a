"""Timing prototype: the cfg2 step as K sub-batches on K streams (tails of one sub-batch's kernels overlap the other's
heads) against the single launch sequence.  usage: python tools/ab_split.py [K]"""
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from simpledepthestimation_b200.functional import MonoLossPlan  # noqa: E402
from simpledepthestimation_b200.geometry.camera import resize_img  # noqa: E402
from simpledepthestimation_b200.synthetic import euler_pose, mono_inputs  # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 2
dev = torch.device("cuda", 0)
B, H, W = 12, 192, 640
nsets = 3


def cached(name, fn):
    path = f"/tmp/sde_ab_{name}.pt"
    if os.path.exists(path):
        return torch.load(path)
    v = fn()
    torch.save(v, path)
    return v


def build(Bsub, parts):
    sets = []
    for k in range(nsets):
        inp = cached(f"cfg2_{k}", lambda: mono_inputs(B, H, W, 4, 2, seed=k))
        sizes = [tuple(d.shape[-2:]) for d in inp["depth"]]
        mv = lambda t: t.to(dev).contiguous()  # noqa: E731
        tgt = [mv(resize_img(inp["img"], s)) for s in sizes]
        src = [[mv(resize_img(c, s)) for c in inp["ctx"]] for s in sizes]
        depth, Kc, pose = [mv(d) for d in inp["depth"]], mv(inp["K"]), [mv(euler_pose(v)) for v in inp["pose_vec"]]
        subs = []
        for q in range(parts):
            sl = slice(q * Bsub, (q + 1) * Bsub)
            c = lambda t: t[sl].contiguous()  # noqa: E731
            subs.append(([c(t) for t in tgt], [[c(x) for x in row] for row in src], [c(d) for d in depth], c(Kc), [c(p) for p in pose]))
        sets.append(subs)
    return sets, sizes


def run(parts):
    Bsub = B // parts
    sets, sizes = build(Bsub, parts)
    plans = [MonoLossPlan(Bsub, sizes, 2, (H, W), dev) for _ in range(parts)]
    streams = [torch.cuda.Stream(device=dev) for _ in range(parts)]
    ones = torch.ones(2, device=dev)
    st = []
    for k in range(nsets):
        row = []
        for q in range(parts):
            s = sets[k][q]
            losses, warped = torch.empty(2, device=dev), plans[q].new_warped()
            _, argm = plans[q].forward(*s, out=losses, warped=warped)
            row.append(dict(losses=losses, warped=warped, argm=argm, gd=[torch.empty_like(d) for d in s[2]],
                            gp=[torch.empty_like(p) for p in s[4]]))
        st.append(row)
    torch.cuda.synchronize()

    def step(i):
        k = i % nsets
        main = torch.cuda.current_stream()
        for q in range(parts):
            sq = streams[q] if parts > 1 else main
            if parts > 1:
                sq.wait_stream(main)
            with torch.cuda.stream(sq):
                s, t = sets[k][q], st[k][q]
                plans[q].forward(*s, out=t["losses"], argmin_out=t["argm"], warped=t["warped"])
                plans[q].backward(*s, t["argm"], ones, t["gd"], t["gp"], warped=t["warped"])
        if parts > 1:
            for q in range(parts):
                main.wait_stream(streams[q])

    graphs = []
    for k in range(nsets):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            step(k)
        graphs.append(g)

    def timeit(fn):
        for i in range(10):
            fn(i)
        torch.cuda.synchronize()
        out = []
        for _ in range(9):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(50):
                fn(i)
            e1.record()
            torch.cuda.synchronize()
            out.append(e0.elapsed_time(e1) / 50 * 1e3)
        return statistics.median(out)
    print(f"parts {parts}: eager {timeit(step):7.1f} us   graph {timeit(lambda i: graphs[i % nsets].replay()):7.1f} us")


run(1)
run(K)
if K != 3:
    run(3)
