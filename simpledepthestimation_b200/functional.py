"""torch.autograd.Function wrappers over the C ABI (libsde_loss.so).

PyTorch is plumbing here: it owns device memory and the stream; the arithmetic is the
hand-written sm_100a kernels.  There is no CPU / eager fallback -- calling these on a CPU
tensor or without the built library raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Sequence

import torch

from . import _lib

__all__ = ["MonoLossPlan", "HostLossRunner", "mono_photometric_smoothness_loss", "MotionLossPlan",
           "motion_rgbd_smoothness_loss"]


def _require_cuda(t: torch.Tensor, name: str):
    if not t.is_cuda:
        raise _lib.SdeError(f"{name} must be a CUDA tensor: the view-synthesis loss has no CPU path")
    if t.dtype != torch.float32:
        raise _lib.SdeError(f"{name} must be float32, got {t.dtype}")


def _contig(t: torch.Tensor) -> torch.Tensor:
    return t if t.is_contiguous() else t.contiguous()


class MonoSaved:
    """Device buffers a MonoDepth2 loss step keeps between its forward and backward pass (MonoLossPlan.new_warped)."""

    __slots__ = ("warped", "smooth_g")

    def __init__(self, warped, smooth_g):
        self.warped, self.smooth_g = warped, smooth_g


class MonoLossPlan:
    """Shape-specialised launcher of the fused MonoDepth2 loss (forward + backward).

    Holds the descriptor, the zero-initialised workspace and the output buffers for one
    (B, scales, S, sizes, loss options) signature so that a training step costs three kernel
    launches per sub-batch and no allocation.  A plan (its workspace: ticket counters and partial-sum
    slots) belongs to ONE stream at a time -- use one plan per stream; the workspace is re-zeroed if
    a call reports an error.  Mirrors the options read by the reference's
    MonoDepth2Model.__init__ (detectron2/modeling/meta_arch/MonoDepth2.py:26-46).

    Scheduling of a step's three kernels (warp, loss forward, loss backward); measured at 640x192 x 12 with the
    two-sources-per-pass backward kernel (profiles/r2_notes.md), where the launches of a batch of 12 alone would spend
    ~17 % of their time starting up and draining:
      streams=1   one stream, the launches chained with TILE-level dependencies (sde_mono_loss_step /
                  sde_mono_loss_forward, include/sde_loss.h: a forward tile waits for the chunks of the warp kernel that
                  hold its rows, a backward tile for the forward tiles of its image), so each kernel fills the SMs its
                  predecessor leaves idle.  forward_backward() 246 us, forward() + backward() 251 us.
      streams=K   the batch as K contiguous sub-batches, each on its own side stream of the calling stream with
                  grid-level dependencies (sde_mono_desc.norm_batch makes their losses and gradients add up to the whole
                  batch's; a sub-batch's kept planes stay in L2 between its kernels).  K = 2: forward_backward() 244 us,
                  forward() + backward() 259 us (each must rejoin the calling stream before it returns).
      None        (default) each call in its faster form: forward_backward() on two sub-batch streams when the batch is
                  even and >= 4, forward() / backward() through a single-stream plan (split_plan()).
    The calling stream waits for the side streams before a call returns, so callers see ordinary stream semantics.
    """

    def __init__(self, batch: int, sizes: Sequence[Sequence[int]], n_sources: int, full_size: Sequence[int],
                 device, ssim_weight=0.85, c1=1e-4, c2=9e-4, smooth_weight=1e-3, automask=True, reduce="min",
                 save_warped=True, depth_mode="depth", min_depth=0.1, max_depth=80.0, streams=None):
        if reduce not in ("min", "mean"):
            raise NotImplementedError(reduce)  # same as MonoDepth2.py:120-121
        if depth_mode not in _lib.DEPTH_MODES:
            raise _lib.SdeError(f"depth_mode must be one of {sorted(_lib.DEPTH_MODES)}")
        if len(sizes) > _lib.MAX_SCALES or n_sources > _lib.MAX_SOURCES:
            raise _lib.SdeError("too many scales / sources")
        self.lib = _lib.load()
        self.device = torch.device(device)
        self.batch, self.sizes, self.n_sources = batch, [tuple(s) for s in sizes], n_sources
        self.full_size = tuple(full_size)
        auto = streams is None and not os.environ.get("SDE_MONO_STREAMS")
        if streams is None:
            streams = int(os.environ.get("SDE_MONO_STREAMS", "0")) or (2 if batch >= 4 and batch % 2 == 0 else 1)
        if streams < 1 or batch % streams != 0:
            raise _lib.SdeError(f"streams ({streams}) must divide the batch ({batch})")
        self.parts, self.sub_batch = int(streams), batch // int(streams)
        # keep the warped sources from forward to backward (default: the backward kernel then stages them with
        # coalesced loads instead of re-projecting and re-gathering, 16 % faster per step at 640x192x12 for 24 B
        # per pixel and source of extra HBM traffic, which this issue-bound path has to spare) or recompute them
        # (save_warped=False: nothing but the argmin bytes and O(B) scalars live between the passes)
        self.save_warped = bool(save_warped)
        d = _lib.MonoDesc()
        d.batch, d.n_scales, d.n_sources = self.sub_batch, len(sizes), n_sources
        d.norm_batch = batch
        for i, (h, w) in enumerate(self.sizes):
            d.height[i], d.width[i] = h, w
        d.full_height, d.full_width = self.full_size
        d.ssim_weight, d.c1, d.c2, d.smooth_weight = ssim_weight, c1, c2, smooth_weight
        d.flags = (_lib.FLAG_AUTOMASK if automask else 0) | (_lib.FLAG_REDUCE_MEAN if reduce == "mean" else 0) | \
                  (_lib.FLAG_NO_TMA if _lib.tma_disabled() else 0) | (_lib.FLAG_NO_FLOW if self.parts > 1 else 0)
        self.lib.sde_reload_env()   # developer switches are read when a plan is built, not per call
        # depth_mode "disp" / "logit": the `depth` tensors hold the decoder's disparity / pre-softplus output and the
        # kernels apply disp_to_depth(., min_depth, max_depth) (depth_decoder.py:9-18,108) themselves; the gradient comes
        # back w.r.t. that tensor
        self.depth_mode = depth_mode
        d.depth_mode, d.min_depth, d.max_depth = _lib.DEPTH_MODES[depth_mode], float(min_depth), float(max_depth)
        self.desc = d
        nbytes = self.lib.sde_mono_workspace_bytes(C.byref(d))
        if nbytes == 0:
            raise _lib.SdeError("invalid loss descriptor (sizes must be >= 2, 1..6 scales, 1..4 sources)")
        self._ws_bytes, self._ws_stride = nbytes, (nbytes + 255) // 256 * 256
        self.workspace = torch.zeros(self.parts * self._ws_stride, dtype=torch.uint8, device=self.device)
        # per sub-batch a block of [n_scales * sub_batch * 2] floats: the same total as for the whole batch
        self.stats = torch.empty(len(sizes) * batch * 2, dtype=torch.float32, device=self.device)
        self._partial = torch.empty(self.parts, 2, dtype=torch.float32, device=self.device) if self.parts > 1 else None
        self._side = None
        # streams=None: forward_backward() runs the sub-batches on two streams, separate forward() / backward() calls (the
        # autograd path) go to a single-stream plan with tile-level dependencies -- each the faster form for its call
        self._split = None
        if auto and self.parts > 1:
            self._split = MonoLossPlan(batch, sizes, n_sources, full_size, device, ssim_weight, c1, c2, smooth_weight, automask,
                                       reduce, save_warped, depth_mode, min_depth, max_depth, streams=1)

    def split_plan(self) -> "MonoLossPlan":
        """The plan that serves separate forward() / backward() calls (its `stats` travel between them)."""
        return self._split if self._split is not None else self

    def _check_status(self, st, what):
        if st != 0:
            self.workspace.zero_()   # a failed call may leave tickets behind: restore the zero-filled state
        _lib.check(st, what)

    def _side_streams(self):
        if self._side is None:
            self._side = [torch.cuda.Stream(device=self.device) for _ in range(self.parts)]
        return self._side

    # ---------------------------------------------------------------------------------
    def _buffers(self, q, target, source, depth, K, pose) -> _lib.MonoBuffers:
        """The buffer block of sub-batch q: every per-sample tensor enters as base pointer + q * sub_batch samples."""
        b = _lib.MonoBuffers()
        n0 = q * self.sub_batch
        for i, (h, w) in enumerate(self.sizes):
            b.target[i] = target[i].data_ptr() + n0 * 3 * h * w * 4
            b.depth[i] = depth[i].data_ptr() + n0 * h * w * 4
            for j in range(self.n_sources):
                b.source[i][j] = source[i][j].data_ptr() + n0 * 3 * h * w * 4
        for j in range(self.n_sources):
            b.pose[j] = pose[j].data_ptr() + n0 * 64
        b.intrinsics = K.data_ptr() + n0 * 36
        b.saved_stats = self.stats.data_ptr() + q * len(self.sizes) * self.sub_batch * 2 * 4
        b.workspace = self.workspace.data_ptr() + q * self._ws_stride
        return b

    def _check(self, target, source, depth, K, pose):
        B = self.batch
        for i, (h, w) in enumerate(self.sizes):
            _require_cuda(target[i], "target"); _require_cuda(depth[i], "depth")
            if tuple(target[i].shape) != (B, 3, h, w) or tuple(depth[i].shape) != (B, 1, h, w):
                raise _lib.SdeError(f"scale {i}: expected target [B,3,{h},{w}] and depth [B,1,{h},{w}]")
            if not (target[i].is_contiguous() and depth[i].is_contiguous()):
                raise _lib.SdeError(f"scale {i}: target and depth must be contiguous")
            for j in range(self.n_sources):
                _require_cuda(source[i][j], "source")
                if tuple(source[i][j].shape) != (B, 3, h, w) or not source[i][j].is_contiguous():
                    raise _lib.SdeError(f"scale {i} source {j}: expected contiguous [B,3,{h},{w}]")
        _require_cuda(K, "intrinsics")
        if tuple(K.shape) != (B, 3, 3) or not K.is_contiguous():
            raise _lib.SdeError("intrinsics must be contiguous [B,3,3]")
        for j in range(self.n_sources):
            _require_cuda(pose[j], "pose")
            if tuple(pose[j].shape) != (B, 4, 4) or not pose[j].is_contiguous():
                raise _lib.SdeError("pose must be contiguous [B,4,4]")

    def new_warped(self):
        """What one step keeps from the forward to the backward pass (sde_mono_buffers.warped / .smooth_g), or None:
        per (scale, source) a [B,9,h,w] buffer -- the warped source and its derivatives w.r.t. the sample coordinate
        (divided by the projective denominator) -- and per scale the [B,1,h,w] local smoothness gradient."""
        if not self.save_warped:
            return None
        new = lambda *shape: torch.empty(*shape, dtype=torch.float32, device=self.device)  # noqa: E731
        return MonoSaved([[new(self.batch, _lib.MONO_SAVED_PLANES, h, w) for _ in range(self.n_sources)] for h, w in self.sizes],
                         [new(self.batch, 1, h, w) for h, w in self.sizes])

    def _set_warped(self, b, q, saved):
        if saved is not None:
            n0 = q * self.sub_batch
            for i, (h, w) in enumerate(self.sizes):
                b.smooth_g[i] = saved.smooth_g[i].data_ptr() + n0 * h * w * 4
                for j in range(self.n_sources):
                    b.warped[i][j] = saved.warped[i][j].data_ptr() + n0 * _lib.MONO_SAVED_PLANES * h * w * 4

    def _launch(self, fn, what, fill):
        """Calls entry point `fn` once per sub-batch: on the current stream, or each on its own side stream, forked from
        and joined back into the current stream."""
        cur = torch.cuda.current_stream()
        if self.parts == 1:
            self._check_status(fn(C.byref(self.desc), C.byref(fill(0)), cur.cuda_stream), what)
            return
        side = self._side_streams()
        for q in range(self.parts):
            side[q].wait_stream(cur)
            self._check_status(fn(C.byref(self.desc), C.byref(fill(q)), side[q].cuda_stream), what)
        for q in range(self.parts):
            cur.wait_stream(side[q])

    def forward_backward(self, target, source, depth, K, pose, grad_losses, out=None, argmin_out=None, grad_depth=None,
                         grad_pose=None, warped=None):
        """Losses AND gradients in one call (value-and-grad), for callers that know the upstream gradients of
        (rec_loss, smooth_loss) when they ask for the losses -- a trainer that sums the loss keys knows them to be ones
        (projects/MonoDepth2/train.py:91-101).  Same kernels and results as forward() followed by backward(); the
        difference is scheduling: every sub-batch runs warp -> loss forward -> loss backward on its own side stream
        without rejoining the calling stream in between, so one sub-batch's backward kernel overlaps the other's
        forward tail (forward() must join before it returns, because its outputs are handed to the caller).
        Returns (losses[2], argmin list, grad_depth list, grad_pose list)."""
        self._check(target, source, depth, K, pose)
        losses = out if out is not None else torch.empty(2, dtype=torch.float32, device=self.device)
        argmin = argmin_out if argmin_out is not None else \
            [torch.empty(self.batch, h, w, dtype=torch.uint8, device=self.device) for h, w in self.sizes]
        if grad_depth is None:
            grad_depth = [torch.empty_like(d) for d in depth]
        if grad_pose is None:
            grad_pose = [torch.empty_like(p) for p in pose]
        if warped is None:
            warped = self.new_warped()

        def fill(q):
            b = self._buffers(q, target, source, depth, K, pose)
            self._set_warped(b, q, warped)
            b.losses = losses.data_ptr() if self.parts == 1 else self._partial.data_ptr() + q * 8
            b.grad_losses = grad_losses.data_ptr()
            for i, (h, w) in enumerate(self.sizes):
                b.argmin[i] = argmin[i].data_ptr() + q * self.sub_batch * h * w
                b.grad_depth[i] = grad_depth[i].data_ptr() + q * self.sub_batch * h * w * 4
            for j in range(self.n_sources):
                b.grad_pose[j] = grad_pose[j].data_ptr() + q * self.sub_batch * 64
            return b

        self._launch(self.lib.sde_mono_loss_step, "sde_mono_loss_step", fill)
        if self.parts > 1:
            torch.sum(self._partial, 0, out=losses)
        return losses, argmin, grad_depth, grad_pose

    def forward(self, target, source, depth, K, pose, want_argmin=True, out=None, argmin_out=None, warped=None):
        """Runs the forward kernel.  Returns (losses[2], argmin list).  All inputs contiguous fp32 CUDA.
        `warped` (from new_warped()) receives the warped sources for the backward pass."""
        if self._split is not None:
            return self._split.forward(target, source, depth, K, pose, want_argmin, out, argmin_out, warped)
        self._check(target, source, depth, K, pose)
        losses = out if out is not None else torch.empty(2, dtype=torch.float32, device=self.device)
        argmin = []
        if argmin_out is not None:
            argmin = argmin_out
        elif want_argmin:
            argmin = [torch.empty(self.batch, h, w, dtype=torch.uint8, device=self.device) for h, w in self.sizes]

        def fill(q):
            b = self._buffers(q, target, source, depth, K, pose)
            self._set_warped(b, q, warped)
            b.losses = losses.data_ptr() if self.parts == 1 else self._partial.data_ptr() + q * 8
            for i, (h, w) in enumerate(self.sizes):
                if argmin:
                    b.argmin[i] = argmin[i].data_ptr() + q * self.sub_batch * h * w
            return b
        self._launch(self.lib.sde_mono_loss_forward, "sde_mono_loss_forward", fill)
        if self.parts > 1:
            torch.sum(self._partial, 0, out=losses)   # the sub-batches' shares of the batch means
        return losses, argmin

    def backward(self, target, source, depth, K, pose, argmin, grad_losses, grad_depth=None, grad_pose=None,
                 warped=None):
        """Runs the backward kernel (reads `warped` if given, else recomputes the warp).
        Returns (grad_depth list, grad_pose list)."""
        if self._split is not None:
            return self._split.backward(target, source, depth, K, pose, argmin, grad_losses, grad_depth, grad_pose, warped)
        if grad_depth is None:
            grad_depth = [torch.empty_like(d) for d in depth]
        if grad_pose is None:
            grad_pose = [torch.empty_like(p) for p in pose]

        def fill(q):
            b = self._buffers(q, target, source, depth, K, pose)
            self._set_warped(b, q, warped)
            b.grad_losses = grad_losses.data_ptr()
            for i, (h, w) in enumerate(self.sizes):
                b.grad_depth[i] = grad_depth[i].data_ptr() + q * self.sub_batch * h * w * 4
                if argmin:
                    b.argmin[i] = argmin[i].data_ptr() + q * self.sub_batch * h * w
            for j in range(self.n_sources):
                b.grad_pose[j] = grad_pose[j].data_ptr() + q * self.sub_batch * 64
            return b
        self._launch(self.lib.sde_mono_loss_backward, "sde_mono_loss_backward", fill)
        return grad_depth, grad_pose


class _MonoLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, plan: MonoLossPlan, K, n_scales, n_sources, *tensors):
        depth = [_contig(t) for t in tensors[:n_scales]]
        pose = [_contig(t) for t in tensors[n_scales:n_scales + n_sources]]
        rest = tensors[n_scales + n_sources:]
        target = [_contig(t) for t in rest[:n_scales]]
        source = [[_contig(rest[n_scales + i * n_sources + j]) for j in range(n_sources)] for i in range(n_scales)]
        K = _contig(K)
        warped = plan.new_warped()
        losses, argmin = plan.forward(target, source, depth, K, pose, warped=warped)
        ctx.plan, ctx.n_scales, ctx.n_sources = plan, n_scales, n_sources
        ctx.warped = warped
        ctx.save_for_backward(K, *depth, *pose, *target, *[s for row in source for s in row], *argmin)
        ctx.stats = plan.split_plan().stats.clone()  # the plan may be reused before backward runs
        for a in argmin:
            ctx.mark_non_differentiable(a)
        return (losses[0], losses[1], *argmin)

    @staticmethod
    def backward(ctx, g_rec, g_smooth, *_):
        n, S, plan = ctx.n_scales, ctx.n_sources, ctx.plan
        saved = ctx.saved_tensors
        K = saved[0]
        depth = list(saved[1:1 + n])
        pose = list(saved[1 + n:1 + n + S])
        target = list(saved[1 + n + S:1 + 2 * n + S])
        flat = saved[1 + 2 * n + S:1 + 2 * n + S + n * S]
        source = [[flat[i * S + j] for j in range(S)] for i in range(n)]
        argmin = list(saved[1 + 2 * n + S + n * S:])
        zero = torch.zeros((), dtype=torch.float32, device=K.device)
        g = torch.stack([g_rec if g_rec is not None else zero, g_smooth if g_smooth is not None else zero]).float()
        plan.split_plan().stats.copy_(ctx.stats)
        grad_depth, grad_pose = plan.backward(target, source, depth, K, pose, argmin, g.contiguous(),
                                              warped=ctx.warped)
        return (None, None, None, None, *grad_depth, *grad_pose, *([None] * (n + n * S)))


def mono_photometric_smoothness_loss(plan: MonoLossPlan, target: List[torch.Tensor],
                                     source: List[List[torch.Tensor]], depth: List[torch.Tensor],
                                     K: torch.Tensor, pose: List[torch.Tensor]):
    """Differentiable fused loss: returns (rec_loss, smooth_loss, argmin_maps).

    target[i] [B,3,h_i,w_i], source[i][j] likewise, depth[i] [B,1,h_i,w_i] (requires_grad),
    K [B,3,3] at full resolution, pose[j] [B,4,4] (requires_grad).  Gradients flow to depth
    and pose only (images and intrinsics are data, as in the reference)."""
    n, S = len(depth), len(pose)
    flat_src = [source[i][j] for i in range(n) for j in range(S)]
    out = _MonoLossFn.apply(plan, K, n, S, *depth, *pose, *target, *flat_src)
    return out[0], out[1], list(out[2:])


class HostLossRunner:
    """End-to-end step through HOST buffers, the call a host-side integration makes: the full-resolution
    frames, the predicted depth pyramid, intrinsics and poses start in pinned host memory; every step copies
    them to the device, builds the image pyramid there (resize_img, MonoDepth2.py:82,88), runs the fused
    forward + backward and copies the losses and gradients back to pinned host memory.

    Copies and compute are software-pipelined over two input slots and two result slots: the H2D copy of step i+1
    (copy stream) and the D2H copy of step i-1 (its own stream) overlap the kernels of step i (calling stream).  Every step still performs its own H2D
    and D2H; `h2d_bytes` / `d2h_bytes` count exactly the tensors copied per step.  step() is asynchronous;
    finish() waits for the last step and returns its host results.

    Options (fewer bytes on the wire; the default copies everything as fp32):
      u8_frames    the frames travel as the decoded uint8 images and become fp32 inside the pyramid kernel
                   (byte / 255 = torchvision's ToTensor, kitti_v2.py:207-208): a quarter of the frame bytes;
      frames_only  the depth pyramid, intrinsics and poses are device-resident (set_resident()): in the trainer
                   they are network outputs that never leave the device, only the frames come from the host."""

    def __init__(self, plan: MonoLossPlan, device, slots=2, u8_frames=False, frames_only=False):
        from .ops import resize_pyramid, resize_pyramid_u8
        self._pyramid, self._pyramid_u8 = resize_pyramid, resize_pyramid_u8
        self.plan, self.device = plan, torch.device(device)
        self.u8_frames, self.frames_only = bool(u8_frames), bool(frames_only)
        B, S = plan.batch, plan.n_sources
        H, W = plan.full_size
        # One arena per direction: the inputs of a step travel as ONE host->device copy and the results as ONE
        # device->host copy (ten / seven separate copies cost ~8 us of launch gap each on a PCIe-bound step).
        # Every tensor is a 16-byte aligned view into the byte arena (TMA needs aligned plane bases).
        fdt = torch.uint8 if self.u8_frames else torch.float32
        self.in_shapes = [("img", (B, 3, H, W), fdt)] + [(f"ctx{j}", (B, 3, H, W), fdt) for j in range(S)]
        if not self.frames_only:
            self.in_shapes += ([(f"depth{i}", (B, 1, h, w), torch.float32) for i, (h, w) in enumerate(plan.sizes)] +
                               [("K", (B, 3, 3), torch.float32)] + [(f"pose{j}", (B, 4, 4), torch.float32) for j in range(S)])
        self.out_shapes = ([("losses", (2,), torch.float32)] +
                           [(f"grad_depth{i}", (B, 1, h, w), torch.float32) for i, (h, w) in enumerate(plan.sizes)] +
                           [(f"grad_pose{j}", (B, 4, 4), torch.float32) for j in range(S)])
        self.in_bytes, self.out_bytes = self._arena_bytes(self.in_shapes), self._arena_bytes(self.out_shapes)
        new = lambda *shape, dt=torch.float32: torch.empty(*shape, dtype=dt, device=self.device)  # noqa: E731
        self.slots = []
        for _ in range(slots):
            arena = new(self.in_bytes, dt=torch.uint8)
            v = self._views(arena, self.in_shapes)
            sl = dict(arena=arena, img=v["img"], ctx=[v[f"ctx{j}"] for j in range(S)], ready=torch.cuda.Event(),
                      free=torch.cuda.Event())
            if not self.frames_only:
                sl.update(depth=[v[f"depth{i}"] for i in range(len(plan.sizes))], K=v["K"],
                          pose=[v[f"pose{j}"] for j in range(S)])
            sl["free"].record()
            self.slots.append(sl)
        self.resident = None
        # results: two device arenas and two pinned host arenas, so that the device->host copy of step i (its own stream)
        # overlaps the kernels of step i + 1 -- PCIe is full duplex, the copies of the two directions overlap as well
        self.outs = []
        for _ in range(2):
            arena = new(self.out_bytes, dt=torch.uint8)
            ov = self._views(arena, self.out_shapes)
            h = torch.empty(self.out_bytes, dtype=torch.uint8, pin_memory=True)
            hv = self._views(h, self.out_shapes)
            o = dict(arena=arena, losses=ov["losses"], grad_depth=[ov[f"grad_depth{i}"] for i in range(len(plan.sizes))],
                     grad_pose=[ov[f"grad_pose{j}"] for j in range(S)], host=h, h_losses=hv["losses"],
                     h_grad_depth=[hv[f"grad_depth{i}"] for i in range(len(plan.sizes))],
                     h_grad_pose=[hv[f"grad_pose{j}"] for j in range(S)], ready=torch.cuda.Event(), free=torch.cuda.Event())
            o["free"].record()
            self.outs.append(o)
        self._last = self.outs[0]
        self.argmin = [new(B, h, w, dt=torch.uint8) for h, w in plan.sizes]
        self.ones = torch.ones(2, device=self.device)
        self.warped = plan.new_warped()
        self.d2h_stream = torch.cuda.Stream(device=self.device)
        self.copy_stream = torch.cuda.Stream(device=self.device)
        # bytes per step, counted from the tensors (the arenas add at most 15 bytes of padding per tensor)
        count = lambda shapes: sum(int(torch.Size(shape).numel()) * torch.empty((), dtype=dt).element_size()  # noqa: E731
                                   for _, shape, dt in shapes)
        self.h2d_bytes, self.d2h_bytes = count(self.in_shapes), count(self.out_shapes)
        self._i = 0
        # kernels per step: pyramid (all frames and scales in one launch); per sub-batch warp + loss forward, backward
        pyr = 1 if (self.u8_frames or len(plan.sizes) > 1) else 0
        self.launches_per_step = pyr + ((2 if plan.save_warped else 1) + 1) * plan.parts

    def describe(self):
        frames = "uint8 frames (byte / 255 in the pyramid kernel)" if self.u8_frames else "fp32 frames"
        rest = "depth pyramid / K / poses device-resident" if self.frames_only else "depth pyramid, K, poses copied too"
        return f"{frames}; {rest}; losses + gradients copied back"

    @staticmethod
    def _arena_bytes(shapes):
        return sum((int(torch.Size(shape).numel()) * torch.empty((), dtype=dt).element_size() + 15) // 16 * 16
                   for _, shape, dt in shapes)

    @staticmethod
    def _views(arena, shapes):
        out, off = {}, 0
        for name, shape, dt in shapes:
            n = int(torch.Size(shape).numel()) * torch.empty((), dtype=dt).element_size()
            out[name] = arena[off:off + n].view(dt).view(*shape)
            off += (n + 15) // 16 * 16
        return out

    def set_resident(self, depth, K, pose):
        """frames_only: the device tensors that stand for the network outputs (depth pyramid, poses) and intrinsics."""
        self.resident = dict(depth=list(depth), K=K, pose=list(pose))

    def pin(self, host_set):
        """Packs an (img, ctx list, depth list, K, pose list) tuple of CPU tensors into one pinned arena (the layout of
        the device slots) -- what a data loader that writes into pinned memory hands over.  With u8_frames the fp32
        frames are quantised to the uint8 images a decoder delivers (round(x * 255))."""
        img, ctx, depth, K, pose = host_set
        arena = torch.empty(self.in_bytes, dtype=torch.uint8, pin_memory=True)
        v = self._views(arena, self.in_shapes)
        q = (lambda t: (t * 255.0).round().clamp(0, 255).to(torch.uint8)) if self.u8_frames else (lambda t: t)
        v["img"].copy_(q(img))
        for j, c in enumerate(ctx):
            v[f"ctx{j}"].copy_(q(c))
        if not self.frames_only:
            for i, d in enumerate(depth):
                v[f"depth{i}"].copy_(d)
            v["K"].copy_(K)
            for j, x in enumerate(pose):
                v[f"pose{j}"].copy_(x)
        return arena

    def step(self, host_arena):
        sl = self.slots[self._i % len(self.slots)]
        self._i += 1
        main = torch.cuda.current_stream()
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(sl["free"])       # the kernels that last read this slot are done
            sl["arena"].copy_(host_arena, non_blocking=True)
            sl["ready"].record()
        main.wait_event(sl["ready"])
        sizes = self.plan.sizes
        frames = [sl["img"]] + sl["ctx"]
        pyr = self._pyramid_u8(frames, sizes) if self.u8_frames else self._pyramid(frames, sizes)   # [frame][scale]
        target = pyr[0]
        source = [[pyr[1 + j][i] for j in range(len(sl["ctx"]))] for i in range(len(sizes))]
        rest = self.resident if self.frames_only else sl
        if rest is None:
            raise _lib.SdeError("HostLossRunner(frames_only=True): call set_resident(depth, K, pose) first")
        o = self.outs[(self._i - 1) % 2]
        main.wait_event(o["free"])                        # the copy that last read this result arena is done
        self.plan.forward_backward(target, source, rest["depth"], rest["K"], rest["pose"], self.ones, out=o["losses"],
                                   argmin_out=self.argmin, grad_depth=o["grad_depth"], grad_pose=o["grad_pose"],
                                   warped=self.warped)
        sl["free"].record()
        o["ready"].record()
        with torch.cuda.stream(self.d2h_stream):
            self.d2h_stream.wait_event(o["ready"])
            o["host"].copy_(o["arena"], non_blocking=True)
            o["free"].record()
        self._last = o

    def join(self):
        """Orders the calling stream after the device->host copies issued so far (for event timing on that stream)."""
        torch.cuda.current_stream().wait_stream(self.d2h_stream)

    def finish(self):
        """Waits for the last step (kernels and its device->host copy) and returns its host results."""
        self.d2h_stream.synchronize()
        o = self._last
        return o["h_losses"], o["h_grad_depth"], o["h_grad_pose"]


# =================================================================================================
# MotionLearning two-frame loss
# =================================================================================================
class MotionLossPlan:
    """Shape-specialised launcher of the fused MotionLearning loss of one scale (both directions).

    Mirrors the options MotionLearningModel.__init__ reads for this path
    (detectron2/modeling/meta_arch/MotionLearning.py:36-42): LOSS.SSIM_WEIGHT, LOSS.C1 ('inf'
    allowed), LOSS.C2.  `scale` is the scale_intrinsics factor of the scale (MotionLearning.py:131-132).
    """

    def __init__(self, batch: int, size: Sequence[int], device, n_dirs=2, ssim_weight=3.0, c1=float("inf"), c2=9e-6,
                 scale=(1.0, 1.0), with_field=True, save_warped=True):
        self.lib = _lib.load()
        # keep the warped planes (rgb, depth error, valid/occlusion: 20 B per pixel and direction) from the statistics
        # pass for the forward and backward loss kernels (TMA-staged, no re-projection / re-gather), or recompute
        self.save_warped = bool(save_warped)
        self.device = torch.device(device)
        self.batch, self.size, self.n_dirs, self.with_field = batch, tuple(size), n_dirs, bool(with_field)
        d = _lib.MotionDesc()
        d.batch, d.n_dirs, d.height, d.width = batch, n_dirs, self.size[0], self.size[1]
        d.scale_x, d.scale_y = float(scale[0]), float(scale[1])
        d.ssim_weight, d.c1, d.c2 = float(ssim_weight), float(c1), float(c2)
        d.flags = (_lib.MOTION_FLAG_FIELD if with_field else 0) | (_lib.MOTION_FLAG_NO_TMA if _lib.tma_disabled() else 0)
        self.lib.sde_reload_env()
        self.desc = d
        nbytes = self.lib.sde_motion_workspace_bytes(C.byref(d))
        if nbytes == 0:
            raise _lib.SdeError("invalid motion-loss descriptor (size >= 2, 1..2 directions, C1 and C2 not both inf)")
        self.workspace = torch.zeros(nbytes, dtype=torch.uint8, device=self.device)
        self.stats = torch.empty(n_dirs * batch * 4, dtype=torch.float32, device=self.device)

    def _check(self, frame_a, frame_b, depth_a, depth_b, K, pose, field):
        B, (h, w) = self.batch, self.size
        for k in range(self.n_dirs):
            for t, name, c in ((frame_a[k], "frame_a", 3), (frame_b[k], "frame_b", 3), (depth_a[k], "depth_a", 1),
                               (depth_b[k], "depth_b", 1)):
                _require_cuda(t, name)
                if tuple(t.shape) != (B, c, h, w) or not t.is_contiguous():
                    raise _lib.SdeError(f"{name}[{k}] must be contiguous [B,{c},{h},{w}]")
            _require_cuda(pose[k], "pose")
            if tuple(pose[k].shape) != (B, 4, 4) or not pose[k].is_contiguous():
                raise _lib.SdeError("pose must be contiguous [B,4,4]")
            if self.with_field:
                _require_cuda(field[k], "field")
                if tuple(field[k].shape) != (B, 3, h, w) or not field[k].is_contiguous():
                    raise _lib.SdeError(f"field[{k}] must be contiguous [B,3,{h},{w}]")
        _require_cuda(K, "intrinsics")
        if tuple(K.shape) != (B, 3, 3) or not K.is_contiguous():
            raise _lib.SdeError("intrinsics must be contiguous [B,3,3]")

    def new_warped(self):
        """Buffers for what one step keeps from the forward to the backward pass (per direction [B,16,h,w]: warped
        rgb, depth error, valid/occlusion, derivatives of the warp, local smoothness gradient, and four planes of
        scratch for the interleaved copy of frame B + depth B that the gather reads), or None."""
        if not self.save_warped:
            return None
        h, w = self.size
        return [torch.empty(self.batch, _lib.MOTION_SAVED_PLANES, h, w, dtype=torch.float32, device=self.device)
                for _ in range(self.n_dirs)]

    def _buffers(self, frame_a, frame_b, depth_a, depth_b, K, pose, field, warped=None):
        b = _lib.MotionBuffers()
        if warped is not None:
            for k in range(self.n_dirs):
                b.warped[k] = warped[k].data_ptr()
        for k in range(self.n_dirs):
            b.frame_a[k], b.frame_b[k] = frame_a[k].data_ptr(), frame_b[k].data_ptr()
            b.depth_a[k], b.depth_b[k] = depth_a[k].data_ptr(), depth_b[k].data_ptr()
            b.pose[k] = pose[k].data_ptr()
            if self.with_field:
                b.field[k] = field[k].data_ptr()
        b.intrinsics = K.data_ptr()
        b.saved_stats = self.stats.data_ptr()
        b.workspace = self.workspace.data_ptr()
        return b

    def forward(self, frame_a, frame_b, depth_a, depth_b, K, pose, field=None, want_maps=True, out=None, warped=None):
        """Statistics / warp pass + fused loss.  Returns (losses [n_dirs,4], maps) where maps is a list per
        direction of dict(occlusion_mask [B,1,h,w], depth_proximity_weight [B,1,h,w], coords_A_in_B [B,h,w,2]).
        `warped` (from new_warped()) receives the warped planes for the loss kernels and the backward pass."""
        self._check(frame_a, frame_b, depth_a, depth_b, K, pose, field)
        b = self._buffers(frame_a, frame_b, depth_a, depth_b, K, pose, field, warped)
        losses = out if out is not None else torch.empty(self.n_dirs, _lib.MOTION_N_LOSSES, dtype=torch.float32,
                                                         device=self.device)
        b.losses = losses.data_ptr()
        maps = []
        if want_maps:
            B, (h, w) = self.batch, self.size
            for k in range(self.n_dirs):
                m = dict(occlusion_mask=torch.empty(B, 1, h, w, dtype=torch.float32, device=self.device),
                         depth_proximity_weight=torch.empty(B, 1, h, w, dtype=torch.float32, device=self.device),
                         coords_A_in_B=torch.empty(B, h, w, 2, dtype=torch.float32, device=self.device))
                b.occlusion[k] = m["occlusion_mask"].data_ptr()
                b.weight[k] = m["depth_proximity_weight"].data_ptr()
                b.coords[k] = m["coords_A_in_B"].data_ptr()
                maps.append(m)
        st = self.lib.sde_motion_loss_forward(C.byref(self.desc), C.byref(b), torch.cuda.current_stream().cuda_stream)
        if st != 0:
            self.workspace.zero_()   # a failed call may leave tickets behind
        _lib.check(st, "sde_motion_loss_forward")
        return losses, maps

    def backward(self, frame_a, frame_b, depth_a, depth_b, K, pose, field, grad_losses, grad_depth=None,
                 grad_pose=None, grad_field=None, warped=None):
        b = self._buffers(frame_a, frame_b, depth_a, depth_b, K, pose, field, warped)
        b.grad_losses = grad_losses.data_ptr()
        if grad_depth is None:
            grad_depth = [torch.empty_like(d) for d in depth_a]
        if grad_pose is None:
            grad_pose = [torch.empty_like(p) for p in pose]
        if grad_field is None and self.with_field:
            grad_field = [torch.empty_like(f) for f in field]
        for k in range(self.n_dirs):
            b.grad_depth_a[k] = grad_depth[k].data_ptr()
            b.grad_pose[k] = grad_pose[k].data_ptr()
            if self.with_field:
                b.grad_field[k] = grad_field[k].data_ptr()
        st = self.lib.sde_motion_loss_backward(C.byref(self.desc), C.byref(b), torch.cuda.current_stream().cuda_stream)
        if st != 0:
            self.workspace.zero_()   # a failed call may leave tickets behind
        _lib.check(st, "sde_motion_loss_backward")
        return grad_depth, grad_pose, grad_field


class _MotionLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, plan: MotionLossPlan, K, want_maps, *tensors):
        n = plan.n_dirs
        depth_a = [_contig(t) for t in tensors[0:n]]
        pose = [_contig(t) for t in tensors[n:2 * n]]
        field = [_contig(t) for t in tensors[2 * n:3 * n]] if plan.with_field else None
        rest = tensors[3 * n:] if plan.with_field else tensors[2 * n:]
        frame_a = [_contig(t) for t in rest[0:n]]
        frame_b = [_contig(t) for t in rest[n:2 * n]]
        depth_b = [_contig(t) for t in rest[2 * n:3 * n]]
        K = _contig(K)
        warped = plan.new_warped()
        losses, maps = plan.forward(frame_a, frame_b, depth_a, depth_b, K, pose, field, want_maps=want_maps, warped=warped)
        ctx.plan, ctx.warped = plan, warped
        ctx.save_for_backward(K, *depth_a, *pose, *(field or []), *frame_a, *frame_b, *depth_b)
        ctx.stats = plan.stats.clone()
        flat = []
        for m in maps:
            flat += [m["occlusion_mask"], m["depth_proximity_weight"], m["coords_A_in_B"]]
        for t in flat:
            ctx.mark_non_differentiable(t)
        return (losses, *flat)

    @staticmethod
    def backward(ctx, g_losses, *_):
        plan = ctx.plan
        n = plan.n_dirs
        sv = ctx.saved_tensors
        K = sv[0]
        i = 1
        depth_a = list(sv[i:i + n]); i += n
        pose = list(sv[i:i + n]); i += n
        field = None
        if plan.with_field:
            field = list(sv[i:i + n]); i += n
        frame_a = list(sv[i:i + n]); i += n
        frame_b = list(sv[i:i + n]); i += n
        depth_b = list(sv[i:i + n]); i += n
        plan.stats.copy_(ctx.stats)
        gd, gp, gf = plan.backward(frame_a, frame_b, depth_a, depth_b, K, pose, field, g_losses.contiguous().float(),
                                   warped=ctx.warped)
        grads = [*gd, *gp] + (list(gf) if plan.with_field else [])
        return (None, None, None, *grads, *([None] * (3 * n)))


def motion_rgbd_smoothness_loss(plan: MotionLossPlan, frame_a, frame_b, depth_a, depth_b, K, pose, field=None,
                                want_maps=True):
    """Differentiable fused MotionLearning loss of one scale.

    Per direction k: rgbd_consistency_loss(frame_a[k], frame_b[k], depth_a[k], depth_b[k], K, R_k, t_k + field[k])
    (MotionLearning.py:248-291) and smoothness_loss(depth_a[k], frame_a[k]).  Returns
    (losses [n_dirs,4] = rgb_l1_loss, ssim_loss, smooth_loss, 0; maps per direction).  Gradients reach
    depth_a, pose and field; depth_b receives none (occlusion is a comparison, the proximity weight is
    detached and DEPTH_L1_WEIGHT is 0 in the reference's configs)."""
    if plan.with_field and field is None:
        raise _lib.SdeError("plan was built with_field=True but no field was given")
    ts = [*depth_a, *pose] + (list(field) if plan.with_field else []) + [*frame_a, *frame_b, *depth_b]
    out = _MotionLossFn.apply(plan, K, want_maps, *ts)
    losses, flat = out[0], out[1:]
    maps = [dict(occlusion_mask=flat[3 * k], depth_proximity_weight=flat[3 * k + 1], coords_A_in_B=flat[3 * k + 2])
            for k in range(len(flat) // 3)]
    return losses, maps
