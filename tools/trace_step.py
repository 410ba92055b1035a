"""Per-CTA timeline of one cfg2 step (warp -> forward -> backward) from a -DSDE_TRACE build of the library:
how many CTAs of each kernel are resident per 5 us bin, CTA lifetimes and the time spent waiting on flow flags.
usage: SDE_LIB_PATH=build/variants/libsde_trace.so [SDE_FLOW_MASK=..] python tools/trace_step.py [cfg2|cfg3]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from simpledepthestimation_b200 import _lib  # noqa: E402
from simpledepthestimation_b200.functional import MonoLossPlan  # noqa: E402
from simpledepthestimation_b200.geometry.camera import resize_img  # noqa: E402
from simpledepthestimation_b200.synthetic import euler_pose, mono_inputs  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
B, H, W = (12, 192, 640) if cfg == "cfg2" else (8, 320, 1024)
dev = torch.device("cuda", 0)
lib = _lib.load()
lib.sde_debug_trace.restype, lib.sde_debug_trace.argtypes = C.c_int, [C.c_void_p]
assert lib.sde_debug_trace(None) == 0
sets = []
for k in range(3):
    inp = mono_inputs(B, H, W, 4, 2, seed=k)
    sizes = [tuple(d.shape[-2:]) for d in inp["depth"]]
    mv = lambda t: t.to(dev).contiguous()  # noqa: E731
    sets.append(([mv(resize_img(inp["img"], s)) for s in sizes], [[mv(resize_img(c, s)) for c in inp["ctx"]] for s in sizes],
                 [mv(d) for d in inp["depth"]], mv(inp["K"]), [mv(euler_pose(v)) for v in inp["pose_vec"]]))
plan = MonoLossPlan(B, sizes, 2, (H, W), dev, streams=1)
ones = torch.ones(2, device=dev)
warped = [plan.new_warped() for _ in sets]
for i in range(12):
    plan.forward_backward(*sets[i % 3], ones, warped=warped[i % 3])
torch.cuda.synchronize()
buf = np.zeros((3, 8192, 4), dtype=np.uint64)
assert lib.sde_debug_trace(buf.ctypes.data) == 0
names = ["warp", "fwd", "bwd"]
t0 = min(int(buf[k, :, 0][buf[k, :, 0] > 0].min()) for k in range(3))
print(f"{cfg} flow mask {os.environ.get('SDE_FLOW_MASK', 'default')}")
spans = []
for k in range(3):
    n = int((buf[k, :, 0] > 0).sum())
    st = (buf[k, :n, 0].astype(np.int64) - t0) / 1e3
    en = (buf[k, :n, 1].astype(np.int64) - t0) / 1e3
    wt = (buf[k, :n, 3].astype(np.int64) - t0) / 1e3 if k > 0 else st
    spans.append((st, en))
    print(f"{names[k]:5s} {n:5d} CTAs  first start {st.min():7.1f}  last start {st.max():7.1f}  last end {en.max():7.1f} us  "
          f"CTA life mean {np.mean(en - st):6.1f} p10 {np.percentile(en - st, 10):6.1f} p90 {np.percentile(en - st, 90):6.1f}  "
          f"start->past wait mean {np.mean(wt - st):5.2f} max {np.max(wt - st):6.1f}")
# refill gaps: on every SM, the time between a CTA's exit and the start of the CTA that takes its slot
for k in range(3):
    n = int((buf[k, :, 0] > 0).sum())
    st = (buf[k, :n, 0].astype(np.int64) - t0) / 1e3
    en = (buf[k, :n, 1].astype(np.int64) - t0) / 1e3
    sm = buf[k, :n, 2].astype(np.int64)
    gaps = []
    for m in np.unique(sm):
        idx = np.where(sm == m)[0]
        ev = sorted([(st[i], 1) for i in idx] + [(en[i], -1) for i in idx])
        free_since = []
        for t, kind in ev:
            if kind == -1:
                free_since.append(t)
            elif free_since:
                gaps.append(t - free_since.pop(0))
    gaps = np.array(gaps) if gaps else np.zeros(1)
    print(f"{names[k]:5s} refill gap per CTA: mean {gaps.mean():5.2f} us  p50 {np.percentile(gaps, 50):5.2f}  p90 {np.percentile(gaps, 90):5.2f}  "
          f"sum {gaps.sum() / 1e3:6.2f} ms over {len(gaps)} refills = {gaps.sum() / max(len(np.unique(sm)), 1):6.1f} us per SM")
end = max(en.max() for _, en in spans)
print("resident CTAs per 5 us bin (warp / fwd / bwd):")
for lo in np.arange(0, end, 5.0):
    mid = lo + 2.5
    print(f"  {lo:6.0f} us  " + "  ".join(f"{int(((st <= mid) & (en > mid)).sum()):5d}" for st, en in spans))
