#!/usr/bin/env python
"""Benchmark of the fused view-synthesis loss (fwd+bwd) -- prints ONE JSON line.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

ours:       a "step" is one forward + one backward pass of the fused MonoDepth2 loss (MonoLossPlan.forward_backward ->
            sde_mono_loss_step: three launches -- warp kernel, loss forward, loss backward -- per sub-batch; the plan's
            default runs the batch as two sub-batches of 6 on two streams, `config.scheduling` says what ran) over one
            batch of synthetic KITTI-shaped input (BASELINE.json configs[1]: 640x192, batch 12 per GPU, 4
            scales, 2 sources, automask + smoothness).  The launches of a step are captured once per input set in a
            CUDA graph, so the host enqueues one graph launch per step.
            `value` = warped Mpix/s with inputs resident in HBM (three input sets are rotated so that every step reads
            from HBM, not L2).  Timing: W warm-up steps, then BLOCKS of exactly K steps, each block bracketed by a
            barrier + torch.cuda.synchronize() on both sides and timed with CUDA events on the launching stream; a
            block's time is the MAX over ranks; blocks repeat until >= 0.5 s have been timed and `ms_per_step` is the
            MEDIAN block (`per_rank_ms` shows every rank's median, `block_ms` every block).
            `e2e` = the same through host buffers: the full-resolution frames, depth pyramid, intrinsics and poses are
            copied from pinned host memory every step, the image pyramid is built on the device, losses and gradients
            are copied back -- all inside the timed region, copies pipelined against the kernels.
            N>1: one process per GPU (torchrun), each rank its own batch of 12 (weak scaling; at N=8 the global batch is
            configs[4]'s 96); the only cross-GPU traffic is one all-reduce of the two loss scalars per step.
reference:  the reference's own CPU implementation of the same path -- the unmodified reference files staged under
            oracle/_ref by oracle/make_ref.sh ("kind": "reference"), else the restatement oracle/port.py ("port") --
            all host threads, on a bounded sample of the workload.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import statistics
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
MIN_TIMED_SECONDS = 0.5


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


# ------------------------------------------------------------------------------------------- reference arms
def _reference_available():
    try:
        from oracle import ref_import
        return ref_import.available()
    except Exception:
        return False


def reference_loss_step(inp, device, dtype=torch.float32):
    """One fwd+bwd of the reference's OWN MonoDepth2Model.forward (oracle/_ref, unmodified) on `device`."""
    from oracle import ref_import
    ns = ref_import.load()
    model = reference_loss_step.models.get(str(device))
    if model is None:
        model = ns.MonoDepth2.MonoDepth2Model(ref_import.mono_cfg()).train().to(device=device, dtype=dtype)
        reference_loss_step.models[str(device)] = model
    depth = [d.detach().clone().requires_grad_() for d in inp["depth"]]
    vecs = [v.detach().clone().requires_grad_() for v in inp["pose_vec"]]
    model.depth_net.payload = {"depth_pred": depth}
    model.pose_net.payload = {"pose_pred": [ns.pose_utils.pose_vec2mat(v) for v in vecs]}
    out = model({"img": inp["img"], "ctx_img": inp["ctx"], "img_orig": inp["img"], "ctx_img_orig": inp["ctx"],
                 "intrinsics": inp["K"]})
    (out["rec_loss"] + out["smooth_loss"]).backward()
    return out["rec_loss"].detach()


reference_loss_step.models = {}


def port_loss_step(inp, pyr):
    """One fwd+bwd of the restatement oracle/port.py (the reference's ATen op sequence) on the host."""
    from oracle import port
    from simpledepthestimation_b200.synthetic import euler_pose
    depth = [d.clone().requires_grad_() for d in inp["depth"]]
    pose = [euler_pose(v).requires_grad_() for v in inp["pose_vec"]]
    out = port.mono_loss(inp["img"], None, inp["K"], depth, pose, pyramid=pyr)
    (out["rec_loss"] + out["smooth_loss"]).backward()
    return out["rec_loss"].detach()


def cpu_baseline(steps, warmup, batch=1):
    """The reference's CPU path on all host threads over `steps` steps of a batch-`batch` sample of the workload."""
    from simpledepthestimation_b200.synthetic import mono_inputs
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    inp = mono_inputs(batch, H, W, SCALES, S, seed=0)
    if _reference_available():
        kind, what = "reference", "the reference's MonoDepth2Model.forward + backward (oracle/_ref, unmodified files)"
        fn = lambda: reference_loss_step(inp, torch.device("cpu"))  # noqa: E731
    else:
        from oracle import port
        kind, what = "port", "oracle/port.py (restatement; oracle/_ref not staged)"
        pyr = [(port.resize_bilinear(inp["img"], d.shape[-2:]), [port.resize_bilinear(c, d.shape[-2:]) for c in inp["ctx"]])
               for d in inp["depth"]]
        fn = lambda: port_loss_step(inp, pyr)  # noqa: E731
    for _ in range(warmup):
        fn()
    t0 = time.perf_counter()
    for _ in range(steps):
        fn()
    dt = (time.perf_counter() - t0) / steps
    warped = S * sum(batch * (H >> i) * (W >> i) for i in range(SCALES))
    return {"value": warped / dt / 1e6, "unit": UNIT, "cores": threads, "kind": kind,
            "sample": f"{steps} steps of 640x192 batch {batch} (4 scales, 2 sources) fwd+bwd, {what}, "
                      f"{threads} threads, {dt * 1e3:.1f} ms/step"}, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, min(args.steps, 6)), max(1, min(args.warmup, 2))
    cb, dt = cpu_baseline(steps, warmup, batch=B_PER_GPU)
    line = {
        "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"MonoDepth2 loss {W}x{H}, batch {B_PER_GPU}, {SCALES} scales, {S} sources, automask + "
                               "smoothness, fwd+bwd (BASELINE.json configs[1]) on the host CPU; bounded sample: "
                               f"{steps} steps"},
        "cpu_baseline": cb,
        "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=args.out)


# ------------------------------------------------------------------------------------------- GPU arm
def make_sets(dev, rank, nsets, batch=B_PER_GPU):
    from simpledepthestimation_b200.geometry.camera import resize_img
    from simpledepthestimation_b200.synthetic import euler_pose, mono_inputs
    sets, host, raw = [], [], []
    for k in range(nsets):
        inp = mono_inputs(batch, H, W, SCALES, S, seed=1000 * rank + k)
        sizes = [tuple(d.shape[-2:]) for d in inp["depth"]]
        tgt = [resize_img(inp["img"], s).contiguous() for s in sizes]
        src = [[resize_img(c, s).contiguous() for c in inp["ctx"]] for s in sizes]
        pose = [euler_pose(v).contiguous() for v in inp["pose_vec"]]
        depth = [d.contiguous() for d in inp["depth"]]
        # host side of the e2e leg: what the data loader and the networks hand over (full-resolution frames)
        host.append((inp["img"].contiguous(), [c.contiguous() for c in inp["ctx"]], depth, inp["K"].contiguous(), pose))
        raw.append(inp)
        mv = lambda t: t.to(dev)  # noqa: E731
        sets.append(([mv(t) for t in tgt], [[mv(x) for x in row] for row in src], [mv(d) for d in depth], mv(inp["K"]),
                     [mv(p) for p in pose]))
    return sets, host, raw


def event_time(fn, iters, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def other_configs(dev, raw_cfg2):
    """Secondary measurements on one GPU (not the bench line's value): BASELINE.json configs[2] (MonoDepth2 1024x320,
    batch 8), configs[3] (MotionLearning 1920x1280, batch 4, both directions, translation field), configs[4] at N=1
    (global batch 96 on one GPU), the drop-in model step (MonoDepth2Model.forward + .backward() through
    torch.autograd.Function) and the reference in eager CUDA on the same B200."""
    from simpledepthestimation_b200.functional import MonoLossPlan, MotionLossPlan
    from simpledepthestimation_b200.geometry.camera import resize_img
    from simpledepthestimation_b200.synthetic import euler_pose, mono_inputs, motion_inputs

    out = {}
    peak = peaks()[0]
    mv = lambda t: t.to(dev).contiguous()  # noqa: E731

    def mono_cfg(name, B, Hc, Wc, seed, iters):
        inp = mono_inputs(B, Hc, Wc, SCALES, S, seed=seed)
        sizes = [tuple(d.shape[-2:]) for d in inp["depth"]]
        tgt = [mv(resize_img(inp["img"], s)) for s in sizes]
        src = [[mv(resize_img(c, s)) for c in inp["ctx"]] for s in sizes]
        depth, K, pose = [mv(d) for d in inp["depth"]], mv(inp["K"]), [mv(euler_pose(v)) for v in inp["pose_vec"]]
        plan = MonoLossPlan(B, sizes, S, (Hc, Wc), dev)
        losses, ones, warped = torch.empty(2, device=dev), torch.ones(2, device=dev), plan.new_warped()
        _, argm = plan.forward(tgt, src, depth, K, pose, out=losses, warped=warped)
        gd, gp = [torch.empty_like(d) for d in depth], [torch.empty_like(p) for p in pose]

        def mono_step():
            plan.forward_backward(tgt, src, depth, K, pose, ones, out=losses, argmin_out=argm, grad_depth=gd, grad_pose=gp,
                                  warped=warped)
        ms = event_time(mono_step, iters)
        px = S * sum(B * h * w for h, w in sizes)
        out[name] = {"ms_per_step": ms, "warped_mpix_s": px / (ms * 1e-3) / 1e6, "iters": iters,
                     "frac_of_hbm_roofline": (px / S * 84.0 / (ms * 1e-3) / 1e9) / peak}

    mono_cfg("cfg3_mono_1024x320_b8", 8, 320, 1024, 3, 20)
    torch.cuda.empty_cache()
    mono_cfg("cfg5_n1_b96_mono_640x192", 96, H, W, 5, 10)
    torch.cuda.empty_cache()

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
    ms = event_time(motion_step, 12)
    px = 2 * B4 * H4 * W4
    out["cfg4_motion_1920x1280_b4"] = {"ms_per_step": ms, "warped_mpix_s": px / (ms * 1e-3) / 1e6, "iters": 12,
                                       "frac_of_hbm_roofline": (px * 104.0 / (ms * 1e-3) / 1e9) / peak}
    del f1, f2, d1, d2, mo, mwarped, mgd, mgf, args4, mplan
    torch.cuda.empty_cache()

    # the API north_star says stays unchanged: MonoDepth2Model(cfg).forward(batch) -> dict of losses, .backward()
    try:
        out["model_step_cfg2"] = model_step(dev, raw_cfg2)
    except Exception as exc:
        out["model_step_cfg2"] = {"error": repr(exc)[:300]}
    torch.cuda.empty_cache()
    try:
        out["model_step_cfg4"] = motion_model_step(dev)
    except Exception as exc:
        out["model_step_cfg4"] = {"error": repr(exc)[:300]}
    torch.cuda.empty_cache()
    # the comparator SURVEY.md 2.3 / 8d names: the reference itself, eager PyTorch on the same B200
    try:
        out["eager_cuda_reference_cfg2"] = eager_cuda_reference(dev, raw_cfg2)
    except Exception as exc:
        out["eager_cuda_reference_cfg2"] = {"error": repr(exc)[:300]}
    torch.cuda.empty_cache()
    return out


class _AttrDict(dict):
    __getattr__ = dict.__getitem__


class _Inject(torch.nn.Module):
    def __init__(self, cfg=None):
        super().__init__()
        self.payload = {}

    def forward(self, batch):
        batch.update(self.payload)
        return batch


def model_step(dev, inp):
    """MonoDepth2Model.forward(batch) + backward of rec_loss + smooth_loss (depth / pose predictions injected in place
    of the networks), cfg2, CUDA events.  Includes the image pyramid, the pose matrices, the
    autograd.Function plumbing and every allocation the model path makes."""
    from simpledepthestimation_b200.geometry.pose_utils import pose_vec2mat
    from simpledepthestimation_b200.modeling import DEPTH_NET_REGISTRY, POSE_NET_REGISTRY, build_model
    if "BenchInjectDepth" not in DEPTH_NET_REGISTRY:
        DEPTH_NET_REGISTRY._do_register("BenchInjectDepth", _Inject)
        POSE_NET_REGISTRY._do_register("BenchInjectPose", _Inject)
    loss = _AttrDict(SSIM_WEIGHT=0.85, C1=1e-4, C2=9e-4, CLIP=0.0, AUTOMASK=True, SMOOTHNESS_WEIGHT=1e-3,
                     PHOTOMETRIC_REDUCE="min", SUPERVISED_WEIGHT=0.0, VARIANCE_FOCUS=0.85, VAR_LOSS_WEIGHT=0.0)
    cfg = _AttrDict(LOSS=loss, MODEL=_AttrDict(META_ARCHITECTURE="MonoDepth2Model", DEVICE=str(dev),
                                               PIXEL_MEAN=[0.45, 0.45, 0.45], PIXEL_STD=[0.225, 0.225, 0.225],
                                               DEPTH_NET=_AttrDict(NAME="BenchInjectDepth"),
                                               POSE_NET=_AttrDict(NAME="BenchInjectPose")))
    model = build_model(cfg).train()
    g = lambda t: t.to(dev).contiguous()  # noqa: E731
    img, ctx, K = g(inp["img"]), [g(c) for c in inp["ctx"]], g(inp["K"])
    depth = [g(d).requires_grad_() for d in inp["depth"]]
    vecs = [g(v).requires_grad_() for v in inp["pose_vec"]]

    def step():
        for t in depth + vecs:
            t.grad = None
        model.depth_net.payload = {"depth_pred": depth}
        model.pose_net.payload = {"pose_pred": [pose_vec2mat(v) for v in vecs]}   # PoseNet.py:63
        out = model({"img": img, "ctx_img": ctx, "img_orig": img, "ctx_img_orig": ctx, "intrinsics": K})
        (out["rec_loss"] + out["smooth_loss"]).backward()
    ms = event_time(step, 30, warm=3)
    px = S * sum(B_PER_GPU * (H >> i) * (W >> i) for i in range(SCALES))
    return {"ms_per_step": ms, "warped_mpix_s": px / (ms * 1e-3) / 1e6, "iters": 30,
            "path": "MonoDepth2Model.forward(batch) -> rec_loss + smooth_loss -> .backward(); depth / pose injected"}


def motion_model_step(dev):
    """MotionLearningModel.forward(batch) + backward of every *loss key (rgbd consistency both directions, smoothness,
    rotation / translation cycle, field smoothness / sparsity), cfg4: 1920x1280, batch 4, residual translation field;
    depth / pose / motion predictions injected in place of the networks."""
    from simpledepthestimation_b200.geometry.pose_utils import pose_vec2mat
    from simpledepthestimation_b200.modeling import DEPTH_NET_REGISTRY, POSE_NET_REGISTRY, build_model
    from simpledepthestimation_b200.synthetic import motion_inputs
    if "BenchInjectDepth" not in DEPTH_NET_REGISTRY:
        DEPTH_NET_REGISTRY._do_register("BenchInjectDepth", _Inject)
        POSE_NET_REGISTRY._do_register("BenchInjectPose", _Inject)
    loss = _AttrDict(NUM_SCALES=1, SSIM_WEIGHT=3.0, C1="inf", C2=9e-6, CLIP=0.0, DEPTH_L1_WEIGHT=0.0, SMOOTHNESS_WEIGHT=1e-3,
                     SUPERVISED_WEIGHT=0.0, VARIANCE_FOCUS=0.85, VAR_LOSS_WEIGHT=0.0, MOTION_SMOOTHNESS_WEIGHT=1.0,
                     MOTION_SPARSITY_WEIGHT=0.2, ROT_CYCLE_WEIGHT=1e-3, TRANS_CYCLE_WEIGHT=5e-2, SCALE_NORMALIZE=False)
    cfg = _AttrDict(LOSS=loss, MODEL=_AttrDict(META_ARCHITECTURE="MotionLearningModel", DEVICE=str(dev),
                                               PIXEL_MEAN=[0.45, 0.45, 0.45], PIXEL_STD=[0.225, 0.225, 0.225],
                                               DEPTH_NET=_AttrDict(NAME="BenchInjectDepth"),
                                               POSE_NET=_AttrDict(NAME="BenchInjectPose", USE_DEPTH=True)))
    model = build_model(cfg).train()
    B4, H4, W4 = 4, 1280, 1920
    inp = motion_inputs(B4, H4, W4, seed=0)
    g = lambda t: t.to(dev).contiguous()  # noqa: E731
    d = g(torch.cat([inp["depth1"], inp["depth2"]], 0)).requires_grad_()
    vec, mo = g(inp["pose_vec"]).requires_grad_(), g(inp["motion"]).requires_grad_()
    feed = {"img": g(inp["img1"]), "ctx_img": [g(inp["img2"])], "intrinsics": g(inp["K"])}

    def step():
        for t in (d, vec, mo):
            t.grad = None
        model.depth_net.payload = {"depth_pred": [d]}
        model.pose_net.payload = {"pose_pred": pose_vec2mat(vec), "motion_pred": mo}
        out = model(dict(feed))
        sum(v for k, v in out.items() if "loss" in k).backward()
    ms = event_time(step, 8, warm=2)
    px = 2 * B4 * H4 * W4
    return {"ms_per_step": ms, "warped_mpix_s": px / (ms * 1e-3) / 1e6, "iters": 8,
            "path": "MotionLearningModel.forward(batch) -> all *loss keys -> .backward(); depth / pose / motion injected"}


def eager_cuda_reference(dev, inp):
    """The reference's own MonoDepth2Model (oracle/_ref, unmodified) in eager PyTorch on this GPU, cfg2, fp32."""
    px = S * sum(B_PER_GPU * (H >> i) * (W >> i) for i in range(SCALES))
    gi = {k: ([t.to(dev) for t in v] if isinstance(v, list) else v.to(dev)) for k, v in inp.items()}
    if _reference_available():
        kind = "reference (oracle/_ref, unmodified files), eager CUDA"
        fn = lambda: reference_loss_step(gi, dev)  # noqa: E731
    else:
        from oracle import port
        kind = "oracle/port.py (restatement), eager CUDA"
        pyr = [(port.resize_bilinear(gi["img"], d.shape[-2:]), [port.resize_bilinear(c, d.shape[-2:]) for c in gi["ctx"]])
               for d in gi["depth"]]
        fn = lambda: port_loss_step(gi, pyr)  # noqa: E731
    ms = event_time(fn, 5, warm=2)
    return {"ms_per_step": ms, "warped_mpix_s": px / (ms * 1e-3) / 1e6, "iters": 5, "kind": kind}


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
    sets, host, raw = make_sets(dev, rank, nsets)
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

    def launch_step(k):
        # losses and gradients in one call (upstream gradients = ones, as in the trainer that sums the loss keys): the same
        # launches as plan.forward() + plan.backward(), without the stream join between the two passes
        plan.forward_backward(*sets[k], ones, out=losses[k], argmin_out=argm[k], grad_depth=gd[k], grad_pose=gp[k],
                              warped=warped[k])

    # the three launches of a step as one CUDA graph per input set: the host then enqueues one graph launch per step
    graphs = None
    if not args.no_graph:
        try:
            for k in range(nsets):
                launch_step(k)
            torch.cuda.synchronize()
            graphs = []
            for k in range(nsets):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    launch_step(k)
                graphs.append(g)
            graphs[0].replay()
            torch.cuda.synchronize()
        except Exception as exc:   # never lose the bench line to a capture problem: plain launches instead
            sys.stderr.write(f"bench.py: CUDA graph capture failed ({exc!r}); timing plain launches\n")
            graphs = None
            torch.cuda.synchronize()

    def step(i):
        k = i % nsets
        if graphs is not None:
            graphs[k].replay()
        else:
            launch_step(k)
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

    def timed(fn, steps, warmup, min_seconds=MIN_TIMED_SECONDS, max_blocks=400, after=None):
        """Blocks of exactly `steps` steps, barrier + synchronize on both sides of every block, CUDA events on the
        launching stream; a block's time is the MAX over ranks.  Returns per-step ms: median block, every block, and
        every rank's own median."""
        for i in range(warmup):
            fn(i)
        blocks, mine, total, it = [], [], 0.0, warmup
        while len(blocks) < 3 or (total < min_seconds * 1e3 and len(blocks) < max_blocks):
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(steps):
                fn(it + i)
            if after is not None:
                after()          # e.g. order the timed stream after the last device->host copy
            e1.record()
            if side is not None:
                torch.cuda.current_stream().wait_stream(side)
            barrier()
            it += steps
            ms = e0.elapsed_time(e1)
            mine.append(ms / steps)
            if world > 1:
                # every rank must take the same number of blocks: the loop condition uses the reduced time
                t = torch.tensor([ms], device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = float(t)
            blocks.append(ms / steps)
            total += ms
        per_rank = [statistics.median(mine)]
        if world > 1:
            t = torch.zeros(world, device=dev)
            t[rank] = per_rank[0]
            dist.all_reduce(t)
            per_rank = [float(x) for x in t]
        return {"ms": statistics.median(blocks), "blocks": blocks, "per_rank": per_rank, "seconds": total * 1e-3}

    target_px = sum(B_PER_GPU * h * w for h, w in sizes)
    warped_px = S * target_px
    with ClockSampler(local) as clk:
        main = timed(step, args.steps, args.warmup)
        # keep the GPU busy a little longer if the run was too short for nvidia-smi to sample it
        if len(clk.rows) < 3:
            t_end = time.time() + 0.5
            while time.time() < t_end:
                step(0)
            torch.cuda.synchronize()
    clocks = clk.summary()
    ms_step = main["ms"]
    value = world * warped_px / (ms_step * 1e-3) / 1e6

    # per-kernel durations (CUDA events on the launching stream) for the roofline of the dominant kernel: every kernel as
    # ONE launch over the whole batch (a plan with streams=1), timed alone
    k_steps = max(20, args.steps)
    plan1 = MonoLossPlan(B_PER_GPU, sizes, S, (H, W), dev, streams=1)
    ms_fwd = timed(lambda i: plan1.forward(*sets[i % nsets], out=losses[i % nsets], argmin_out=argm[i % nsets],
                                           warped=warped[i % nsets]), k_steps, 3, 0.1)["ms"]
    ms_bwd = timed(lambda i: plan1.backward(*sets[i % nsets], argm[i % nsets], ones, gd[i % nsets], gp[i % nsets],
                                            warped=warped[i % nsets]), k_steps, 3, 0.1)["ms"]
    peak, peak_src = peaks()
    bwd_name = "mono_bwd_pair_kernel" if (S % 2 == 0 and plan1.save_warped and os.environ.get("SDE_BWD_PAIR", "1") != "0") else "mono_bwd_kernel"
    # the forward CALL is two kernels of similar length (warp 61 us + loss forward 65 us at cfg2, profiles/r2_launches.csv),
    # the backward call one: the backward kernel is the dominant KERNEL unless the forward call takes twice as long
    dom = bwd_name if 2.0 * ms_bwd >= ms_fwd or not plan1.save_warped else "mono_warp_kernel + mono_fwd_kernel"
    dom_bytes = target_px * (BYTES_BWD_PER_TARGET_PX if dom == bwd_name else BYTES_FWD_PER_TARGET_PX)
    dom_ms = ms_bwd if dom == bwd_name else ms_fwd
    achieved = dom_bytes / (dom_ms * 1e-3) / 1e9
    traffic, traffic_src = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "dram_traffic.json")) as fh:
            tr = json.load(fh)
        traffic, traffic_src = tr.get(dom), tr.get("source", "profiles/dram_traffic.json (static file from an ncu --set full capture)")
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": dom_bytes, "kernel_ms": dom_ms,
                "fwd_ms": ms_fwd, "bwd_ms": ms_bwd, "kernel_timing": "one launch over the whole batch, timed alone",
                "step_frac_of_hbm_roofline": (target_px * 84.0 / (ms_step * 1e-3) / 1e9) / peak}

    # End to end through host buffers.  Headline leg: everything a step consumes travels from pinned host memory --
    # the three frames as the DECODED uint8 images (converted by byte / 255 = torchvision's ToTensor inside the pyramid
    # kernel), the predicted depth pyramid, intrinsics and poses as fp32 -- and the losses and all gradients travel back.
    e_steps = max(3, min(args.steps, 20))

    def e2e_leg(**kw):
        r = HostLossRunner(plan, dev, **kw)
        pins = [r.pin(h) for h in host]
        if kw.get("frames_only"):
            r.set_resident(sets[0][2], sets[0][3], sets[0][4])
        t = timed(lambda i: r.step(pins[i % nsets]), e_steps, 3, 0.3, after=r.join)
        r.finish()
        return {"value": world * warped_px / (t["ms"] * 1e-3) / 1e6, "unit": UNIT, "h2d_bytes_per_step": r.h2d_bytes,
                "d2h_bytes_per_step": r.d2h_bytes, "ms_per_step": t["ms"], "per_rank_ms": t["per_rank"],
                "gpu_launches_per_step": r.launches_per_step, "what": r.describe()}

    e2e = e2e_leg(u8_frames=True)
    e2e["path"] = ("pinned host: uint8 frames + fp32 depth pyramid / K / poses -> one H2D copy -> device pyramid (byte / 255, "
                   "resize) -> warp + loss fwd + loss bwd -> one D2H copy of losses + gradients; two pipelined slots")
    # further legs of the same path (not the headline): fp32 frames as in round 1; depth / K / poses device-resident as
    # in the trainer, where they are network outputs
    e2e_variants = {}
    for name, kw in (("e2e_f32_frames", dict()), ("e2e_f32_frames_only", dict(frames_only=True)),
                     ("e2e_u8_frames_only", dict(u8_frames=True, frames_only=True))):
        try:
            e2e_variants[name] = e2e_leg(**kw)
        except Exception as exc:
            e2e_variants[name] = {"error": repr(exc)[:200]}
    e2e["variants"] = e2e_variants

    extra = None
    if rank == 0 and world == 1 and not args.no_extra:
        try:
            extra = other_configs(dev, raw[0])
        except Exception as exc:  # secondary numbers must never break the bench line
            extra = {"error": repr(exc)[:200]}

    cb = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cb, _ = cpu_baseline(4, 1, batch=B_PER_GPU)

    if rank == 0:
        n_blocks = len(main["blocks"])
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": f"MonoDepth2 loss {W}x{H}, batch {B_PER_GPU} per GPU, {SCALES} scales, {S} sources, "
                                   "automask + smoothness, fwd+bwd (BASELINE.json configs[1]; global batch 96 at 8 GPUs = configs[4])",
                       "global_batch": B_PER_GPU * world, "parallelism": f"dp{world}",
                       "l2": f"{nsets} input sets rotated ({nsets * 86} MB > 126 MB L2)",
                       "timing": f"{n_blocks} blocks of {args.steps} steps (>= {MIN_TIMED_SECONDS} s), barrier + synchronize "
                                 "around every block, CUDA events, max over ranks per block, median block",
                       "cuda_graph": graphs is not None,
                       "scheduling": ("one stream, three launches chained with tile-level dependencies (sde_mono_loss_step)"
                                      if plan.parts == 1 else
                                      f"{plan.parts} sub-batches of {plan.sub_batch} samples on {plan.parts} streams "
                                      "(MonoLossPlan(streams=...): their kernels overlap each other's start-up and drain)")},
            "per_rank_ms": main["per_rank"], "block_ms": [round(b, 5) for b in main["blocks"][:64]],
            "timed_seconds": main["seconds"],
            "roofline": roofline, "cpu_baseline": cb, "e2e": e2e,
            "gpu_launches": (3 if plan.save_warped else 2) * plan.parts * args.steps * n_blocks,
            "gpu_launches_per_step": (3 if plan.save_warped else 2) * plan.parts, "clocks": clocks, "other_configs": extra,
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
    ap.add_argument("--no-extra", action="store_true", help="skip the secondary configs (cfg3, cfg4, cfg5 at N=1, model step, eager reference)")
    ap.add_argument("--no-graph", action="store_true", help="plain launches instead of one CUDA graph per step")
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
