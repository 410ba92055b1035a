#!/usr/bin/env python
"""Benchmark of the fused view-synthesis loss (fwd+bwd) -- prints ONE JSON line.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

ours:       a "step" is one forward + one backward pass of the fused MonoDepth2 loss (three launches:
            warp kernel, loss forward, loss backward) over one batch of synthetic KITTI-shaped input
            (BASELINE.json configs[1]: 640x192, batch 12 per GPU, 4 scales, 2 sources, automask +
            smoothness).  `value` = warped Mpix/s with inputs resident in HBM (three input sets are
            rotated so that every step reads from HBM, not L2); `e2e` = the same through host buffers:
            the full-resolution frames, depth pyramid, intrinsics and poses are copied from pinned host
            memory every step, the image pyramid is built on the device, losses and gradients are
            copied back -- all inside the timed region, copies pipelined against the kernels.  N>1: one process per GPU (torchrun), each rank
            its own batch of 12 (weak scaling; at N=8 the global batch is configs[4]'s 96); the only
            cross-GPU traffic is one all-reduce of the two loss scalars per step.
reference:  the reference's CPU implementation of the same path (oracle/port.py, the same ATen op
            sequence as the reference; /root/reference itself is not present on the GPU box), all
            host threads, on a bounded sample of the workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "fused view-synthesis loss fwd+bwd warped Mpix/s"
UNIT = "Mpix/s"
H, W, B_PER_GPU, SCALES, S = 192, 640, 12, 4, 2
BYTES_FWD_PER_TARGET_PX = 16 + 12 * S        # depth 4 + target 12 + sources 12*S   
BYTES_BWD_PER_TARGET_PX = 16 + 12 * S + 4    # same reads + grad-depth write


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples SM clocks / throttle reasons with nvidia-smi while the timed region runs."""

    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for n, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------- CPU arm
def cpu_loss_step(inp, pyr, threads):
    """One fwd+bwd of the oracle port (the reference's ATen op sequence) on the host."""
    from oracle import port
    from simpledepthestimation_b200.synthetic import euler_pose
    torch.set_num_threads(threads)
    depth = [d.clone().requires_grad_() for d in inp["depth"]]
    pose = [euler_pose(v).requires_grad_() for v in inp["pose_vec"]]
    out = port.mono_loss(inp["img"], None, inp["K"], depth, pose, pyramid=pyr)
    (out["rec_loss"] + out["smooth_loss"]).backward()
    return float(out["rec_loss"].detach())


def cpu_baseline(steps, warmup, batch=1):
    from oracle import port
    from simpledepthestimation_b200.synthetic import mono_inputs
    threads = os.cpu_count() or 1
    inp = mono_inputs(batch, H, W, SCALES, S, seed=0)
    pyr = [(port.resize_bilinear(inp["img"], d.shape[-2:]), [port.resize_bilinear(c, d.shape[-2:]) for c in inp["ctx"]])
           for d in inp["depth"]]
    for _ in range(warmup):
        cpu_loss_step(inp, pyr, threads)
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_loss_step(inp, pyr, threads)
    dt = (time.perf_counter() - t0) / steps
    warped = S * sum(batch * (H >> i) * (W >> i) for i in range(SCALES))
    return {"value": warped / dt / 1e6, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{steps} steps of 640x192 batch {batch} (4 scales, 2 sources) fwd+bwd, oracle/port.py, "
                      f"{threads} threads, {dt * 1e3:.1f} ms/step"}, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 10))
    cb, dt = cpu_baseline(steps, max(1, min(args.warmup, 2)), batch=1)
    line = {
        "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": max(1, min(args.warmup, 2)), "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "MonoDepth2 loss 640x192, 4 scales, 2 sources, automask+smoothness, fwd+bwd; "
                               "bounded sample: batch 1 per step on the host CPU"},
        "cpu_baseline": cb,
        "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=args.out)


# ------------------------------------------------------------------------------------------- GPU arm
def make_sets(dev, rank, nsets):
    from simpledepthestimation_b200.geometry.camera import resize_img
    from simpledepthestimation_b200.synthetic import euler_pose, mono_inputs
    sets, host = [], []
    for k in range(nsets):
        inp = mono_inputs(B_PER_GPU, H, W, SCALES, S, seed=1000 * rank + k)
        sizes = [tuple(d.shape[-2:]) for d in inp["depth"]]
        tgt = [resize_img(inp["img"], s).contiguous() for s in sizes]
        src = [[resize_img(c, s).contiguous() for c in inp["ctx"]] for s in sizes]
        pose = [euler_pose(v).contiguous() for v in inp["pose_vec"]]
        depth = [d.contiguous() for d in inp["depth"]]
        # host side of the e2e leg: what the data loader and the networks hand over (full-resolution frames)
        host.append((inp["img"].contiguous(), [c.contiguous() for c in inp["ctx"]], depth, inp["K"].contiguous(), pose))
        mv = lambda t: t.to(dev)  # noqa: E731
        sets.append(([mv(t) for t in tgt], [[mv(x) for x in row] for row in src], [mv(d) for d in depth], mv(inp["K"]),
                     [mv(p) for p in pose]))
    return sets, host


def other_configs(dev):
    """Secondary measurements (not the bench line): BASELINE.json configs[2] (MonoDepth2 1024x320, batch 8) and
    configs[3] (MotionLearning 1920x1280, batch 4, both directions, translation field), fwd+bwd, HBM-resident."""
    from simpledepthestimation_b200.functional import MonoLossPlan, MotionLossPlan
    from simpledepthestimation_b200.geometry.camera import resize_img
    from simpledepthestimation_b200.synthetic import euler_pose, mono_inputs, motion_inputs

    def timeit(fn, iters=5):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters

    out = {}
    B3, H3, W3 = 8, 320, 1024
    inp = mono_inputs(B3, H3, W3, SCALES, S, seed=3)
    sizes = [tuple(d.shape[-2:]) for d in inp["depth"]]
    mv = lambda t: t.to(dev).contiguous()  # noqa: E731
    tgt = [mv(resize_img(inp["img"], s)) for s in sizes]
    src = [[mv(resize_img(c, s)) for c in inp["ctx"]] for s in sizes]
    depth, K, pose = [mv(d) for d in inp["depth"]], mv(inp["K"]), [mv(euler_pose(v)) for v in inp["pose_vec"]]
    plan = MonoLossPlan(B3, sizes, S, (H3, W3), dev)
    losses, ones, warped = torch.empty(2, device=dev), torch.ones(2, device=dev), plan.new_warped()
    _, argm = plan.forward(tgt, src, depth, K, pose, out=losses, warped=warped)
    gd, gp = [torch.empty_like(d) for d in depth], [torch.empty_like(p) for p in pose]

    def mono_step():
        plan.forward(tgt, src, depth, K, pose, out=losses, argmin_out=argm, warped=warped)
        plan.backward(tgt, src, depth, K, pose, argm, ones, gd, gp, warped=warped)
    ms = timeit(mono_step)
    px = S * sum(B3 * h * w for h, w in sizes)
    out["cfg3_mono_1024x320_b8"] = {"ms_per_step": ms, "warped_mpix_s": px / (ms * 1e-3) / 1e6,
                                    "frac_of_hbm_roofline": (px / S * 84.0 / (ms * 1e-3) / 1e9) / peaks()[0]}
    del tgt, src, depth, warped, gd, argm, plan

    B4, H4, W4 = 4, 1280, 1920
    mi = motion_inputs(B4, H4, W4, seed=0)
    f1, f2, d1, d2, K4 = mv(mi["img1"]), mv(mi["img2"]), mv(mi["depth1"]), mv(mi["depth2"]), mv(mi["K"])
    pose4, mo = mv(euler_pose(mi["pose_vec"])), mv(mi["motion"])
    mplan = MotionLossPlan(B4, (H4, W4), dev, 2, with_field=True)
    args4 = ([f1, f2], [f2, f1], [d1, d2], [d2, d1], K4, [pose4[:B4].contiguous(), pose4[B4:].contiguous()],
             [mo[:B4].contiguous(), mo[B4:].contiguous()])
    ml, gl = torch.empty(2, 4, device=dev), torch.ones(2, 4, device=dev)
    mgd, mgp = [torch.empty_like(d1), torch.empty_like(d2)], [torch.empty(B4, 4, 4, device=dev) for _ in range(2)]
    mgf = [torch.empty_like(args4[6][0]) for _ in range(2)]
    mwarped = mplan.new_warped()

    def motion_step():
        mplan.forward(*args4, want_maps=False, out=ml, warped=mwarped)
        mplan.backward(*args4, gl, mgd, mgp, mgf, warped=mwarped)
    ms = timeit(motion_step, iters=3)
    px = 2 * B4 * H4 * W4
    out["cfg4_motion_1920x1280_b4"] = {"ms_per_step": ms, "warped_mpix_s": px / (ms * 1e-3) / 1e6,
                                       "frac_of_hbm_roofline": (px * 104.0 / (ms * 1e-3) / 1e9) / peaks()[0]}
    return out


def run_ours(args):
    import torch.distributed as dist
    from simpledepthestimation_b200 import build
    from simpledepthestimation_b200.functional import HostLossRunner, MonoLossPlan

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    build.build()

    nsets = 3
    sets, host = make_sets(dev, rank, nsets)
    sizes = [(H >> i, W >> i) for i in range(SCALES)]
    plan = MonoLossPlan(B_PER_GPU, sizes, S, (H, W), dev)
    ones = torch.ones(2, device=dev)
    losses = [torch.empty(2, device=dev) for _ in range(nsets)]
    gd = [[torch.empty_like(d) for d in s[2]] for s in sets]
    gp = [[torch.empty_like(p) for p in s[4]] for s in sets]
    warped = [plan.new_warped() for _ in sets]
    argm = [plan.forward(*s, out=losses[k], warped=warped[k])[1] for k, s in enumerate(sets)]
    loss_sum = torch.zeros(2, device=dev)
    side = torch.cuda.Stream(device=dev) if world > 1 else None

    def step(i):
        k = i % nsets
        plan.forward(*sets[k], out=losses[k], argmin_out=argm[k], warped=warped[k])
        plan.backward(*sets[k], argm[k], ones, gd[k], gp[k], warped=warped[k])
        if world > 1:
            # the loss scalars are logging-only in the reference (comm.reduce_dict, train.py:95): reduce them
            # asynchronously on a side stream so the 8-byte collective never stalls the compute stream
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                loss_sum.copy_(losses[k])
                dist.all_reduce(loss_sum)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for i in range(warmup):
            fn(i)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        if side is not None:
            torch.cuda.current_stream().wait_stream(side)
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms / steps

    target_px = sum(B_PER_GPU * h * w for h, w in sizes)
    warped_px = S * target_px
    with ClockSampler(local) as clk:
        ms_step = timed(step, args.steps, args.warmup)
        # keep the GPU busy a little longer if the run was too short for nvidia-smi to sample it
        if len(clk.rows) < 3:
            t_end = time.time() + 0.5
            while time.time() < t_end:
                step(0)
            torch.cuda.synchronize()
    clocks = clk.summary()
    value = world * warped_px / (ms_step * 1e-3) / 1e6

    # per-kernel durations (CUDA events on the launching stream) for the roofline of the dominant kernel
    ms_fwd = timed(lambda i: plan.forward(*sets[i % nsets], out=losses[i % nsets], argmin_out=argm[i % nsets],
                                          warped=warped[i % nsets]), args.steps, 3)
    ms_bwd = timed(lambda i: plan.backward(*sets[i % nsets], argm[i % nsets], ones, gd[i % nsets], gp[i % nsets],
                                           warped=warped[i % nsets]), args.steps, 3)
    peak, peak_src = peaks()
    dom = "mono_bwd_kernel" if ms_bwd >= ms_fwd else "mono_fwd_kernel"
    dom_bytes = target_px * (BYTES_BWD_PER_TARGET_PX if dom == "mono_bwd_kernel" else BYTES_FWD_PER_TARGET_PX)
    dom_ms = max(ms_bwd, ms_fwd)
    achieved = dom_bytes / (dom_ms * 1e-3) / 1e9
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "dram_traffic.json")) as fh:
            traffic = json.load(fh).get(dom)
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": dom_bytes, "kernel_ms": dom_ms,
                "fwd_ms": ms_fwd, "bwd_ms": ms_bwd,
                "step_frac_of_hbm_roofline": (target_px * 84.0 / (ms_step * 1e-3) / 1e9) / peak}

    # end to end through host buffers
    runner = HostLossRunner(plan, dev)
    pinned = [runner.pin(h) for h in host]
    e2e_ms = timed(lambda i: runner.step(pinned[i % nsets]), max(3, min(args.steps, 20)), 3)
    runner.finish()
    e2e = {"value": world * warped_px / (e2e_ms * 1e-3) / 1e6, "unit": UNIT,
           "h2d_bytes_per_step": runner.h2d_bytes, "d2h_bytes_per_step": runner.d2h_bytes, "ms_per_step": e2e_ms,
           "gpu_launches_per_step": runner.launches_per_step,
           "path": "pinned host frames/depth/K/pose -> H2D -> device pyramid -> warp + loss fwd + loss bwd -> D2H losses+grads"}

    extra = None
    if rank == 0 and world == 1 and not args.no_extra:
        try:
            extra = other_configs(dev)
        except Exception as exc:  # secondary numbers must never break the bench line
            extra = {"error": repr(exc)[:200]}

    cb = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cb, _ = cpu_baseline(6, 1, batch=1)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": f"MonoDepth2 loss {W}x{H}, batch {B_PER_GPU} per GPU, {SCALES} scales, {S} sources, "
                                   "automask + smoothness, fwd+bwd (BASELINE.json configs[1]; global batch 96 at 8 GPUs = configs[4])",
                       "global_batch": B_PER_GPU * world, "parallelism": f"dp{world}",
                       "l2": f"{nsets} input sets rotated ({nsets * 86} MB > 126 MB L2)"},
            "roofline": roofline, "cpu_baseline": cb, "e2e": e2e,
            "gpu_launches": (3 if plan.save_warped else 2) * args.steps, "clocks": clocks, "other_configs": extra,
        }
        print(json.dumps(line), file=args.out)
    if world > 1:
        dist.destroy_process_group()


def _quiet_stdout():
    """Everything libraries print to fd 1 while the bench runs (NCCL's version banner ...) goes to stderr; returns the
    stream that writes to the real stdout, which receives exactly one line: the JSON result."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return real


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    ap.add_argument("--no-extra", action="store_true", help="skip the secondary configs (cfg3, cfg4)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    args.out = _quiet_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
    args.out.flush()


if __name__ == "__main__":
    main()
