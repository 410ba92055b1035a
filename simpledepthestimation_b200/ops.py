"""torch.autograd.Function wrappers of the stand-alone CUDA operators (csrc/ops.cu): view_synthesis,
SSIM / WeightedSSIM, smoothness_loss, resize_img.  The reference-named entry points live in
geometry/camera.py and modeling/losses/*.py; this module is the plumbing (device memory, stream,
workspaces).  No CPU path: CPU tensors raise."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib


def _cuda_f32(t: torch.Tensor, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise _lib.SdeError(f"{name} must be a CUDA tensor: the view-synthesis loss path has no CPU implementation")
    if t.dtype != torch.float32:
        raise _lib.SdeError(f"{name} must be float32, got {t.dtype}")
    return t if t.is_contiguous() else t.contiguous()


def _stream():
    return torch.cuda.current_stream().cuda_stream


_WS = {}


def _zero_workspace(kind, key, nbytes, device):
    """Zero-initialised workspaces are re-zeroed by the kernels, so they are cached per shape and stream."""
    k = (kind, key, str(device), _stream())
    ws = _WS.get(k)
    if ws is None or ws.numel() < nbytes:
        ws = torch.zeros(max(nbytes, 16), dtype=torch.uint8, device=device)
        _WS[k] = ws
    return ws


# ------------------------------------------------------------------------------------------------ view_synthesis
class _ViewSynthesisFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, image_B, depth_A, K, R, t):
        lib = _lib.load()
        image_B, depth_A = _cuda_f32(image_B, "image_B"), _cuda_f32(depth_A, "depth_A")
        K, R, t = _cuda_f32(K, "intrinsics"), _cuda_f32(R, "R_A_to_B"), _cuda_f32(t, "t_A_to_B")
        B, Cc, h, w = image_B.shape
        if tuple(depth_A.shape) != (B, 1, h, w) or tuple(K.shape) != (B, 3, 3) or tuple(R.shape) != (B, 3, 3):
            raise _lib.SdeError("view_synthesis: expected image_B [B,C,H,W], depth_A [B,1,H,W], intrinsics / R [B,3,3]")
        per_pixel = tuple(t.shape) == (B, 3, h, w) and (h, w) != (1, 1)
        if not per_pixel and t.numel() != B * 3:
            raise _lib.SdeError("view_synthesis: t_A_to_B must be [B,3,1,1] or [B,3,H,W]")
        d = _lib.VsDesc(B, Cc, h, w, _lib_flag(per_pixel))
        b = _lib.VsBuffers()
        dev = image_B.device
        sampled = torch.empty_like(image_B)
        z = torch.empty(B, 1, h, w, device=dev)
        coords = torch.empty(B, h, w, 2, device=dev)
        valid = torch.empty(B, 1, h, w, dtype=torch.uint8, device=dev)
        b.image_b, b.depth_a, b.intrinsics, b.rotation, b.translation = (x.data_ptr() for x in (image_B, depth_A, K, R, t))
        b.sampled, b.depth_in_b, b.coords, b.valid = sampled.data_ptr(), z.data_ptr(), coords.data_ptr(), valid.data_ptr()
        _lib.check(lib.sde_view_synthesis_forward(C.byref(d), C.byref(b), _stream()), "sde_view_synthesis_forward")
        ctx.save_for_backward(image_B, depth_A, K, R, t)
        ctx.per_pixel, ctx.t_shape = per_pixel, tuple(t.shape)
        valid = valid.bool()
        ctx.mark_non_differentiable(valid)
        return sampled, z, coords, valid

    @staticmethod
    def backward(ctx, g_sampled, g_z, g_coords, _g_valid):
        lib = _lib.load()
        image_B, depth_A, K, R, t = ctx.saved_tensors
        B, Cc, h, w = image_B.shape
        dev = image_B.device
        d = _lib.VsDesc(B, Cc, h, w, _lib_flag(ctx.per_pixel))
        nbytes = lib.sde_view_synthesis_workspace_bytes(C.byref(d))
        ws = _zero_workspace("vs", (B, Cc, h, w), nbytes, dev)
        b = _lib.VsBuffers()
        b.image_b, b.depth_a, b.intrinsics, b.rotation, b.translation = (x.data_ptr() for x in (image_B, depth_A, K, R, t))
        keep = []
        gs = g_sampled.contiguous().float() if g_sampled is not None else torch.zeros_like(image_B)
        keep.append(gs)
        b.grad_sampled = gs.data_ptr()
        if g_z is not None:
            g_z = g_z.contiguous().float(); keep.append(g_z); b.grad_depth_in_b = g_z.data_ptr()
        if g_coords is not None:
            g_coords = g_coords.contiguous().float(); keep.append(g_coords); b.grad_coords = g_coords.data_ptr()
        g_depth = torch.empty_like(depth_A)
        g_R = torch.empty_like(R)
        g_t = torch.empty(B, 3, h, w, device=dev) if ctx.per_pixel else torch.empty(B, 3, device=dev)
        b.grad_depth_a, b.grad_rotation, b.grad_translation = g_depth.data_ptr(), g_R.data_ptr(), g_t.data_ptr()
        g_img = None
        if ctx.needs_input_grad[0]:
            g_img = torch.empty_like(image_B)
            b.grad_image_b = g_img.data_ptr()
        b.workspace = ws.data_ptr()
        _lib.check(lib.sde_view_synthesis_backward(C.byref(d), C.byref(b), _stream()), "sde_view_synthesis_backward")
        return g_img, g_depth, None, g_R, g_t.view(ctx.t_shape)


def _lib_flag(per_pixel):
    return 1 if per_pixel else 0


def view_synthesis(image_B, depth_A, intrinsics, R_A_to_B, t_A_to_B):
    return _ViewSynthesisFn.apply(image_B, depth_A, intrinsics, R_A_to_B, t_A_to_B)


# ------------------------------------------------------------------------------------------------ SSIM
class _SsimFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, y, w, c1, c2):
        lib = _lib.load()
        x, y = _cuda_f32(x, "x"), _cuda_f32(y, "y")
        if x.shape != y.shape or x.dim() != 4:
            raise _lib.SdeError("SSIM: x and y must be [B,C,H,W] of the same shape")
        B, Cc, h, w_ = x.shape
        d = _lib.SsimDesc(B, Cc, h, w_, float(c1), float(c2))
        b = _lib.SsimBuffers()
        out = torch.empty_like(x)
        b.x, b.y, b.out = x.data_ptr(), y.data_ptr(), out.data_ptr()
        avg_w = None
        if w is not None:
            w = _cuda_f32(w, "w")
            if tuple(w.shape) != (B, 1, h, w_):
                raise _lib.SdeError("WeightedSSIM: w must be [B,1,H,W]")
            avg_w = torch.empty_like(w)
            b.weight, b.avg_w = w.data_ptr(), avg_w.data_ptr()
        if lib.sde_ssim_workspace_bytes(C.byref(d)) == 0:
            raise _lib.SdeError("SSIM: invalid sizes (H, W >= 2) or C1 and C2 both infinite")
        _lib.check(lib.sde_ssim_forward(C.byref(d), C.byref(b), _stream()), "sde_ssim_forward")
        ctx.save_for_backward(x, y, w if w is not None else x.new_empty(0))
        ctx.consts = (float(c1), float(c2), w is not None)
        if avg_w is None:
            return out
        ctx.mark_non_differentiable(avg_w)
        return out, avg_w

    @staticmethod
    def backward(ctx, g_out, *_):
        lib = _lib.load()
        x, y, w = ctx.saved_tensors
        c1, c2, weighted = ctx.consts
        B, Cc, h, w_ = x.shape
        d = _lib.SsimDesc(B, Cc, h, w_, c1, c2)
        b = _lib.SsimBuffers()
        g_out = g_out.contiguous().float()
        ws = torch.empty(lib.sde_ssim_workspace_bytes(C.byref(d)), dtype=torch.uint8, device=x.device)
        gx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        gy = torch.empty_like(y) if ctx.needs_input_grad[1] else None
        if gx is None and gy is None:
            return None, None, None, None, None
        b.x, b.y, b.grad_out, b.workspace = x.data_ptr(), y.data_ptr(), g_out.data_ptr(), ws.data_ptr()
        if weighted:
            b.weight = w.data_ptr()
        if gx is not None:
            b.grad_x = gx.data_ptr()
        if gy is not None:
            b.grad_y = gy.data_ptr()
        _lib.check(lib.sde_ssim_backward(C.byref(d), C.byref(b), _stream()), "sde_ssim_backward")
        return gx, gy, None, None, None


def ssim(x, y, c1, c2):
    return _SsimFn.apply(x, y, None, c1, c2)


def weighted_ssim(x, y, w, c1, c2):
    return _SsimFn.apply(x, y, w.detach(), c1, c2)


# ------------------------------------------------------------------------------------------------ smoothness
class _SmoothnessFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, depth, image):
        lib = _lib.load()
        depth, image = _cuda_f32(depth, "depth"), _cuda_f32(image, "image")
        B, Cc, h, w = image.shape
        if tuple(depth.shape) != (B, 1, h, w):
            raise _lib.SdeError("smoothness_loss: expected depth [B,1,H,W] and image [B,C,H,W]")
        d = _lib.SmoothDesc(B, Cc, h, w)
        nbytes = lib.sde_smoothness_workspace_bytes(C.byref(d))
        if nbytes == 0:
            raise _lib.SdeError("smoothness_loss: H and W must be >= 2")
        ws = _zero_workspace("smooth", (B, Cc, h, w), nbytes, depth.device)
        loss = torch.empty(1, device=depth.device)
        stats = torch.empty(B * 2, device=depth.device)
        b = _lib.SmoothBuffers()
        b.depth, b.image, b.loss, b.saved_stats, b.workspace = (depth.data_ptr(), image.data_ptr(), loss.data_ptr(),
                                                               stats.data_ptr(), ws.data_ptr())
        _lib.check(lib.sde_smoothness_forward(C.byref(d), C.byref(b), _stream()), "sde_smoothness_forward")
        ctx.save_for_backward(depth, image, stats)
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        depth, image, stats = ctx.saved_tensors
        B, Cc, h, w = image.shape
        d = _lib.SmoothDesc(B, Cc, h, w)
        g = g.reshape(1).contiguous().float()
        gd = torch.empty_like(depth)
        b = _lib.SmoothBuffers()
        b.depth, b.image, b.saved_stats, b.grad_loss, b.grad_depth = (depth.data_ptr(), image.data_ptr(), stats.data_ptr(),
                                                                      g.data_ptr(), gd.data_ptr())
        _lib.check(lib.sde_smoothness_backward(C.byref(d), C.byref(b), _stream()), "sde_smoothness_backward")
        return gd, None


def smoothness(depth, image):
    return _SmoothnessFn.apply(depth, image.detach())


# ------------------------------------------------------------------------------------------------ resize
def resize_bilinear(image: torch.Tensor, dst_size) -> torch.Tensor:
    """F.interpolate(image, dst_size, mode='bilinear', align_corners=True) for a CUDA fp32 [..., H, W] tensor."""
    lib = _lib.load()
    image = _cuda_f32(image, "image")
    sh, sw = image.shape[-2:]
    dh, dw = int(dst_size[-2]), int(dst_size[-1])
    planes = image.numel() // (sh * sw)
    out = torch.empty(*image.shape[:-2], dh, dw, device=image.device)
    _lib.check(lib.sde_resize_bilinear(image.data_ptr(), out.data_ptr(), planes, sh, sw, dh, dw, _stream()),
               "sde_resize_bilinear")
    return out


class _AvgPoolFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, image, dh, dw):
        lib = _lib.load()
        image = _cuda_f32(image, "image")
        sh, sw = image.shape[-2:]
        planes = image.numel() // (sh * sw)
        out = torch.empty(*image.shape[:-2], dh, dw, device=image.device)
        _lib.check(lib.sde_resize_avgpool_forward(image.data_ptr(), out.data_ptr(), planes, sh, sw, dh, dw, _stream()),
                   "sde_resize_avgpool_forward")
        ctx.geom = (tuple(image.shape), planes, sh, sw, dh, dw)
        return out

    @staticmethod
    def backward(ctx, g):
        shape, planes, sh, sw, dh, dw = ctx.geom
        g = _cuda_f32(g, "grad")
        out = torch.empty(shape, device=g.device)
        _lib.check(_lib.load().sde_resize_avgpool_backward(g.data_ptr(), out.data_ptr(), planes, sh, sw, dh, dw, _stream()),
                   "sde_resize_avgpool_backward")
        return out, None, None


def resize_avgpool(image: torch.Tensor, dst_size) -> torch.Tensor:
    """F.adaptive_avg_pool2d(image, dst_size) for a CUDA fp32 [..., H, W] tensor (camera.py:49-54), differentiable."""
    return _AvgPoolFn.apply(image, int(dst_size[-2]), int(dst_size[-1]))


def resize_pyramid(frames, sizes):
    """resize_img of every frame in `frames` (CUDA fp32 [..., H, W], all the same shape) to every size in `sizes`, in
    one launch.  Returns out[f][l]; a size equal to the source size returns the frame itself (camera.py:41-42)."""
    lib = _lib.load()
    frames = [_cuda_f32(f, "frame") for f in frames]
    sh, sw = frames[0].shape[-2:]
    if any(f.shape != frames[0].shape for f in frames):
        raise _lib.SdeError("resize_pyramid: frames must share one shape")
    levels = [(int(s[-2]), int(s[-1])) for s in sizes if (int(s[-2]), int(s[-1])) != (sh, sw)]
    if len(frames) > _lib.MAX_SOURCES + 1 or len(levels) > _lib.MAX_SCALES:
        raise _lib.SdeError("resize_pyramid: too many frames / levels")
    out = {}
    if levels:
        planes = frames[0].numel() // (sh * sw)
        b = _lib.PyramidBuffers()
        for f, fr in enumerate(frames):
            b.src[f] = fr.data_ptr()
            for l, (dh, dw) in enumerate(levels):
                out[f, (dh, dw)] = torch.empty(*fr.shape[:-2], dh, dw, device=fr.device)
                b.dst[f][l] = out[f, (dh, dw)].data_ptr()
        dh = (C.c_int32 * len(levels))(*[s[0] for s in levels])
        dw = (C.c_int32 * len(levels))(*[s[1] for s in levels])
        _lib.check(lib.sde_resize_pyramid(len(frames), planes, sh, sw, len(levels), dh, dw, C.byref(b), _stream()),
                   "sde_resize_pyramid")
    return [[fr if (int(s[-2]), int(s[-1])) == (sh, sw) else out[f, (int(s[-2]), int(s[-1]))] for s in sizes]
            for f, fr in enumerate(frames)]


def resize_pyramid_u8(frames, sizes):
    """The pyramid from DECODED frames: `frames` are CUDA uint8 [..., H, W] tensors of one shape; returns out[f][l] fp32
    with the bits of resize_img(frame / 255, sizes[l]) (a size equal to the source size is the conversion itself,
    torchvision ToTensor).  One launch."""
    lib = _lib.load()
    for f in frames:
        if not f.is_cuda or f.dtype != torch.uint8:
            raise _lib.SdeError("resize_pyramid_u8: frames must be CUDA uint8 tensors")
    frames = [f.contiguous() for f in frames]
    sh, sw = frames[0].shape[-2:]
    if any(f.shape != frames[0].shape for f in frames):
        raise _lib.SdeError("resize_pyramid_u8: frames must share one shape")
    levels = [(int(s[-2]), int(s[-1])) for s in sizes]
    if len(frames) > _lib.MAX_SOURCES + 1 or not 1 <= len(levels) <= _lib.MAX_SCALES:
        raise _lib.SdeError("resize_pyramid_u8: too many frames / levels")
    planes = frames[0].numel() // (sh * sw)
    b = _lib.PyramidBuffers()
    out = []
    for f, fr in enumerate(frames):
        b.src[f] = fr.data_ptr()
        row = []
        for l, (dh, dw) in enumerate(levels):
            row.append(torch.empty(*fr.shape[:-2], dh, dw, dtype=torch.float32, device=fr.device))
            b.dst[f][l] = row[-1].data_ptr()
        out.append(row)
    dh = (C.c_int32 * len(levels))(*[s[0] for s in levels])
    dw = (C.c_int32 * len(levels))(*[s[1] for s in levels])
    _lib.check(lib.sde_resize_pyramid_u8(len(frames), planes, sh, sw, len(levels), dh, dw, C.byref(b), _stream()),
               "sde_resize_pyramid_u8")
    return out


# ------------------------------------------------------------------------------------------------ motion regularisers
class _MotionConsistencyFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, coords, mask, R, t_ab, t_ba):
        lib = _lib.load()
        coords, mask, R = _cuda_f32(coords.detach(), "coords_A_in_B"), _cuda_f32(mask.detach(), "mask"), _cuda_f32(R, "R_A2B")
        t_ab, t_ba = _cuda_f32(t_ab, "t_A2B"), _cuda_f32(t_ba, "t_B2A")
        B, _, h, w = t_ab.shape
        if tuple(coords.shape) != (B, h, w, 2) or tuple(mask.shape) != (B, 1, h, w) or tuple(R.shape) != (B, 3, 3) \
                or tuple(t_ab.shape) != (B, 3, h, w) or tuple(t_ba.shape) != (B, 3, h, w):
            raise _lib.SdeError("motion_consistency_loss: expected coords [B,H,W,2], mask [B,1,H,W], R [B,3,3], t [B,3,H,W]")
        d = _lib.McDesc(B, h, w)
        nbytes = lib.sde_motion_consistency_workspace_bytes(C.byref(d))
        if nbytes == 0:
            raise _lib.SdeError("motion_consistency_loss: H and W must be >= 2")
        ws = _zero_workspace("mcons", (B, h, w), nbytes, t_ab.device)
        loss = torch.empty(1, device=t_ab.device)
        b = _lib.McBuffers()
        b.coords, b.mask, b.rotation, b.t_ab, b.t_ba = (x.data_ptr() for x in (coords, mask, R, t_ab, t_ba))
        b.loss, b.workspace = loss.data_ptr(), ws.data_ptr()
        _lib.check(lib.sde_motion_consistency_forward(C.byref(d), C.byref(b), _stream()), "sde_motion_consistency_forward")
        ctx.save_for_backward(coords, mask, R, t_ab, t_ba)
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        coords, mask, R, t_ab, t_ba = ctx.saved_tensors
        B, _, h, w = t_ab.shape
        d = _lib.McDesc(B, h, w)
        ws = _zero_workspace("mcons", (B, h, w), lib.sde_motion_consistency_workspace_bytes(C.byref(d)), t_ab.device)
        g = g.reshape(1).contiguous().float()
        g_ab, g_ba, g_R = torch.empty_like(t_ab), torch.empty_like(t_ba), torch.empty_like(R)
        b = _lib.McBuffers()
        b.coords, b.mask, b.rotation, b.t_ab, b.t_ba = (x.data_ptr() for x in (coords, mask, R, t_ab, t_ba))
        b.grad_loss, b.grad_t_ab, b.grad_t_ba, b.grad_rotation = g.data_ptr(), g_ab.data_ptr(), g_ba.data_ptr(), g_R.data_ptr()
        b.workspace = ws.data_ptr()
        _lib.check(lib.sde_motion_consistency_backward(C.byref(d), C.byref(b), _stream()), "sde_motion_consistency_backward")
        return None, None, g_R, g_ab, g_ba


def motion_translation_consistency(coords, mask, R_A2B, t_A2B, t_B2A):
    return _MotionConsistencyFn.apply(coords, mask, R_A2B, t_A2B, t_B2A)


class _MotionConsistencySplitFn(torch.autograd.Function):
    """The translation cycle term on (pose, residual field) pairs: t = pose[:, :3, 3] + field is formed per pixel inside
    the kernels (sde_mcons_buffers.pose_ab / pose_ba), the [B,3,H,W] overall fields are never materialised."""

    @staticmethod
    def forward(ctx, coords, mask, R, pose_ab, field_ab, pose_ba, field_ba):
        lib = _lib.load()
        coords, mask, R = _cuda_f32(coords.detach(), "coords_A_in_B"), _cuda_f32(mask.detach(), "mask"), _cuda_f32(R, "R_A2B")
        pose_ab, pose_ba = _cuda_f32(pose_ab, "pose_A2B"), _cuda_f32(pose_ba, "pose_B2A")
        field_ab = None if field_ab is None else _cuda_f32(field_ab, "field_A2B")
        field_ba = None if field_ba is None else _cuda_f32(field_ba, "field_B2A")
        B, h, w, _ = coords.shape
        for f in (field_ab, field_ba):
            if f is not None and tuple(f.shape) != (B, 3, h, w):
                raise _lib.SdeError("motion_consistency_loss: the residual fields must be [B,3,H,W]")
        if tuple(mask.shape) != (B, 1, h, w) or tuple(R.shape) != (B, 3, 3) or tuple(pose_ab.shape) != (B, 4, 4) \
                or tuple(pose_ba.shape) != (B, 4, 4):
            raise _lib.SdeError("motion_consistency_loss: expected coords [B,H,W,2], mask [B,1,H,W], R [B,3,3], pose [B,4,4]")
        d = _lib.McDesc(B, h, w)
        nbytes = lib.sde_motion_consistency_workspace_bytes(C.byref(d))
        if nbytes == 0:
            raise _lib.SdeError("motion_consistency_loss: H and W must be >= 2")
        ws = _zero_workspace("mcons", (B, h, w), nbytes, coords.device)
        loss = torch.empty(1, device=coords.device)
        b = _lib.McBuffers()
        b.coords, b.mask, b.rotation, b.pose_ab, b.pose_ba = (x.data_ptr() for x in (coords, mask, R, pose_ab, pose_ba))
        b.t_ab = None if field_ab is None else field_ab.data_ptr()
        b.t_ba = None if field_ba is None else field_ba.data_ptr()
        b.loss, b.workspace = loss.data_ptr(), ws.data_ptr()
        _lib.check(lib.sde_motion_consistency_forward(C.byref(d), C.byref(b), _stream()), "sde_motion_consistency_forward")
        ctx.has = (field_ab is not None, field_ba is not None)
        ctx.save_for_backward(coords, mask, R, pose_ab, pose_ba, *[f for f in (field_ab, field_ba) if f is not None])
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        coords, mask, R, pose_ab, pose_ba, *fields = ctx.saved_tensors
        field_ab = fields.pop(0) if ctx.has[0] else None
        field_ba = fields.pop(0) if ctx.has[1] else None
        B, h, w, _ = coords.shape
        d = _lib.McDesc(B, h, w)
        ws = _zero_workspace("mcons", (B, h, w), lib.sde_motion_consistency_workspace_bytes(C.byref(d)), coords.device)
        g = g.reshape(1).contiguous().float()
        g_R = torch.empty_like(R)
        g_pab, g_pba = torch.empty(B, 3, device=R.device), torch.empty(B, 3, device=R.device)
        g_fab = None if field_ab is None else torch.empty_like(field_ab)
        g_fba = None if field_ba is None else torch.empty_like(field_ba)
        b = _lib.McBuffers()
        b.coords, b.mask, b.rotation, b.pose_ab, b.pose_ba = (x.data_ptr() for x in (coords, mask, R, pose_ab, pose_ba))
        b.grad_loss, b.grad_rotation, b.workspace = g.data_ptr(), g_R.data_ptr(), ws.data_ptr()
        b.grad_pose_t_ab, b.grad_pose_t_ba = g_pab.data_ptr(), g_pba.data_ptr()
        if field_ab is not None:
            b.t_ab, b.grad_t_ab = field_ab.data_ptr(), g_fab.data_ptr()
        if field_ba is not None:
            b.t_ba, b.grad_t_ba = field_ba.data_ptr(), g_fba.data_ptr()
        _lib.check(lib.sde_motion_consistency_backward(C.byref(d), C.byref(b), _stream()), "sde_motion_consistency_backward")

        def pose_grad(gt):   # [B,3] -> [B,4,4] with the translation column filled
            out = torch.zeros(B, 4, 4, device=gt.device)
            out[:, :3, 3] = gt
            return out
        return None, None, g_R, pose_grad(g_pab), g_fab, pose_grad(g_pba), g_fba


def motion_translation_consistency_split(coords, mask, R_A2B, pose_A2B, field_A2B, pose_B2A, field_B2A):
    return _MotionConsistencySplitFn.apply(coords, mask, R_A2B, pose_A2B, field_A2B, pose_B2A, field_B2A)


class _MotionFieldRegFn(torch.autograd.Function):
    """motion_smoothness_loss_fn + motion_sparsity_loss_fn of the normalised field m / sqrt(3 mean(t^2) + 1e-12),
    t = pose[:, :3, 3] + m (MotionLearning.py:203-220), fused: sde_motion_field_reg_forward / _backward."""

    @staticmethod
    def forward(ctx, pose, field):
        lib = _lib.load()
        pose, field = _cuda_f32(pose, "pose"), _cuda_f32(field, "motion_field")
        if field.dim() != 4 or field.shape[1] != 3 or tuple(pose.shape) != (field.shape[0], 4, 4):
            raise _lib.SdeError("motion field regularisers: expected pose [B,4,4] and a [B,3,H,W] field")
        B, _, h, w = field.shape
        d = _lib.MregDesc(B, 3, h, w)
        nbytes = lib.sde_motion_field_reg_workspace_bytes(C.byref(d))
        if nbytes == 0:
            raise _lib.SdeError("motion field regularisers: H and W must be >= 2")
        ws = _zero_workspace("mfield", (B, h, w), nbytes, field.device)
        losses = torch.empty(2, device=field.device)
        stats = torch.empty(B * 12, device=field.device)
        b = _lib.MfieldBuffers()
        b.pose, b.field, b.losses, b.saved_stats, b.workspace = (x.data_ptr() for x in (pose, field, losses, stats, ws))
        _lib.check(lib.sde_motion_field_reg_forward(C.byref(d), C.byref(b), _stream()), "sde_motion_field_reg_forward")
        ctx.save_for_backward(pose, field, stats)
        return losses[0], losses[1]

    @staticmethod
    def backward(ctx, g_sm, g_sp):
        lib = _lib.load()
        pose, field, stats = ctx.saved_tensors
        B, _, h, w = field.shape
        d = _lib.MregDesc(B, 3, h, w)
        zero = torch.zeros((), device=field.device)
        g = torch.stack([g_sm if g_sm is not None else zero, g_sp if g_sp is not None else zero]).float().contiguous()
        gf, gt = torch.empty_like(field), torch.empty(B, 3, device=field.device)
        b = _lib.MfieldBuffers()
        b.pose, b.field, b.saved_stats = pose.data_ptr(), field.data_ptr(), stats.data_ptr()
        b.grad_losses, b.grad_field, b.grad_pose_t = g.data_ptr(), gf.data_ptr(), gt.data_ptr()
        _lib.check(lib.sde_motion_field_reg_backward(C.byref(d), C.byref(b), _stream()), "sde_motion_field_reg_backward")
        gp = torch.zeros(B, 4, 4, device=field.device)
        gp[:, :3, 3] = gt
        return gp, gf


def motion_field_regularizers(pose, field):
    """(motion_smoothness_loss_fn(mn), motion_sparsity_loss_fn(mn)) with mn = field / sqrt(3 mean((pose_t + field)^2) + 1e-12)."""
    return _MotionFieldRegFn.apply(pose, field)


class _MotionRegFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, field, kind):
        lib = _lib.load()
        field = _cuda_f32(field, "motion_field")
        if field.dim() != 4:
            raise _lib.SdeError("motion regulariser: expected a [B,C,H,W] field")
        B, Cc, h, w = field.shape
        d = _lib.MregDesc(B, Cc, h, w)
        nbytes = lib.sde_motion_reg_workspace_bytes(C.byref(d))
        if nbytes == 0:
            raise _lib.SdeError("motion regulariser: H and W must be >= 2")
        ws = _zero_workspace("mreg", (B, Cc, h, w), nbytes, field.device)
        loss = torch.empty(1, device=field.device)
        stats = torch.empty(B * Cc, device=field.device)
        b = _lib.MregBuffers()
        b.field, b.loss, b.saved_stats, b.workspace = field.data_ptr(), loss.data_ptr(), stats.data_ptr(), ws.data_ptr()
        fn = lib.sde_motion_smoothness_forward if kind == "smoothness" else lib.sde_motion_sparsity_forward
        _lib.check(fn(C.byref(d), C.byref(b), _stream()), f"sde_motion_{kind}_forward")
        ctx.save_for_backward(field, stats)
        ctx.kind = kind
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        field, stats = ctx.saved_tensors
        B, Cc, h, w = field.shape
        d = _lib.MregDesc(B, Cc, h, w)
        g = g.reshape(1).contiguous().float()
        gf = torch.empty_like(field)
        b = _lib.MregBuffers()
        b.field, b.saved_stats, b.grad_loss, b.grad_field = field.data_ptr(), stats.data_ptr(), g.data_ptr(), gf.data_ptr()
        fn = lib.sde_motion_smoothness_backward if ctx.kind == "smoothness" else lib.sde_motion_sparsity_backward
        _lib.check(fn(C.byref(d), C.byref(b), _stream()), f"sde_motion_{ctx.kind}_backward")
        return gf, None


def motion_smoothness(field):
    return _MotionRegFn.apply(field, "smoothness")


def motion_sparsity(field):
    return _MotionRegFn.apply(field, "sparsity")


# ------------------------------------------------------------------------------------------------ variance_loss
class _VarianceFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, depth):
        lib = _lib.load()
        depth = _cuda_f32(depth, "depth")
        n = depth.numel()
        ws = _zero_workspace("var", n, lib.sde_variance_workspace_bytes(n), depth.device)
        loss, stats = torch.empty(1, device=depth.device), torch.empty(2, device=depth.device)
        b = _lib.VarBuffers()
        b.depth, b.loss, b.saved_stats, b.workspace = depth.data_ptr(), loss.data_ptr(), stats.data_ptr(), ws.data_ptr()
        _lib.check(lib.sde_variance_loss_forward(n, C.byref(b), _stream()), "sde_variance_loss_forward")
        ctx.save_for_backward(depth, stats)
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        depth, stats = ctx.saved_tensors
        g = g.reshape(1).contiguous().float()
        gd = torch.empty_like(depth)
        b = _lib.VarBuffers()
        b.depth, b.saved_stats, b.grad_loss, b.grad_depth = depth.data_ptr(), stats.data_ptr(), g.data_ptr(), gd.data_ptr()
        _lib.check(lib.sde_variance_loss_backward(depth.numel(), C.byref(b), _stream()), "sde_variance_loss_backward")
        return gd


def variance(depth):
    return _VarianceFn.apply(depth)


# ------------------------------------------------------------------------------------------------ silog_loss
class _SilogFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, depth_est, depth_gt, variance_focus):
        lib = _lib.load()
        est, gt = _cuda_f32(depth_est, "depth_est"), _cuda_f32(depth_gt, "depth_gt")
        if est.shape != gt.shape:
            raise _lib.SdeError(f"silog_loss: depth_est {tuple(est.shape)} and depth_gt {tuple(gt.shape)} differ")
        n = est.numel()
        ws = _zero_workspace("silog", n, lib.sde_silog_workspace_bytes(n), est.device)
        loss, stats = torch.empty(1, device=est.device), torch.empty(3, device=est.device)
        b = _lib.SilogBuffers()
        b.depth_est, b.depth_gt, b.loss, b.saved_stats, b.workspace = (est.data_ptr(), gt.data_ptr(), loss.data_ptr(),
                                                                       stats.data_ptr(), ws.data_ptr())
        _lib.check(lib.sde_silog_loss_forward(n, float(variance_focus), C.byref(b), _stream()), "sde_silog_loss_forward")
        ctx.save_for_backward(est, gt, stats)
        ctx.vf = float(variance_focus)
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        est, gt, stats = ctx.saved_tensors
        g = g.reshape(1).contiguous().float()
        ge = torch.empty_like(est)
        b = _lib.SilogBuffers()
        b.depth_est, b.depth_gt, b.saved_stats, b.grad_loss, b.grad_depth_est = (est.data_ptr(), gt.data_ptr(),
                                                                                stats.data_ptr(), g.data_ptr(), ge.data_ptr())
        _lib.check(lib.sde_silog_loss_backward(est.numel(), ctx.vf, C.byref(b), _stream()), "sde_silog_loss_backward")
        return ge, None, None


def silog(depth_est, depth_gt, variance_focus):
    return _SilogFn.apply(depth_est, depth_gt, variance_focus)


# ------------------------------------------------------------------------------------------------ disp_to_depth
class _DispToDepthFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, disp, min_depth, max_depth):
        lib = _lib.load()
        disp = _cuda_f32(disp, "disp")
        scaled, depth = torch.empty_like(disp), torch.empty_like(disp)
        b = _lib.DispBuffers()
        b.disp, b.scaled_disp, b.depth = disp.data_ptr(), scaled.data_ptr(), depth.data_ptr()
        _lib.check(lib.sde_disp_to_depth_forward(disp.numel(), float(min_depth), float(max_depth), C.byref(b), _stream()),
                   "sde_disp_to_depth_forward")
        ctx.save_for_backward(disp)
        ctx.range = (float(min_depth), float(max_depth))
        return scaled, depth

    @staticmethod
    def backward(ctx, g_scaled, g_depth):
        lib = _lib.load()
        (disp,) = ctx.saved_tensors
        gd = torch.empty_like(disp)
        b = _lib.DispBuffers()
        b.disp, b.grad_disp = disp.data_ptr(), gd.data_ptr()
        keep = []
        if g_scaled is not None:
            keep.append(g_scaled.contiguous().float())
            b.grad_scaled_disp = keep[-1].data_ptr()
        if g_depth is not None:
            keep.append(g_depth.contiguous().float())
            b.grad_depth = keep[-1].data_ptr()
        if not keep:
            return torch.zeros_like(disp), None, None
        _lib.check(lib.sde_disp_to_depth_backward(disp.numel(), ctx.range[0], ctx.range[1], C.byref(b), _stream()),
                   "sde_disp_to_depth_backward")
        return gd, None, None


def disp_to_depth(disp, min_depth, max_depth):
    """(scaled_disp, depth) as detectron2/layers/depth_decoder.py:9-18."""
    return _DispToDepthFn.apply(disp, min_depth, max_depth)


# ------------------------------------------------------------------------------------------------ pose_vec2mat
class _PoseVec2MatFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, vec):
        lib = _lib.load()
        vec = _cuda_f32(vec, "vec")
        if vec.dim() != 2 or vec.shape[1] != 6:
            raise _lib.SdeError(f"pose_vec2mat: expected [B,6], got {tuple(vec.shape)}")
        pose = torch.empty(vec.shape[0], 4, 4, device=vec.device)
        b = _lib.PoseVecBuffers()
        b.vec, b.pose = vec.data_ptr(), pose.data_ptr()
        _lib.check(lib.sde_pose_vec2mat_forward(vec.shape[0], C.byref(b), _stream()), "sde_pose_vec2mat_forward")
        ctx.save_for_backward(vec)
        return pose

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        (vec,) = ctx.saved_tensors
        g = g.contiguous().float()
        gv = torch.empty_like(vec)
        b = _lib.PoseVecBuffers()
        b.vec, b.grad_pose, b.grad_vec = vec.data_ptr(), g.data_ptr(), gv.data_ptr()
        _lib.check(lib.sde_pose_vec2mat_backward(vec.shape[0], C.byref(b), _stream()), "sde_pose_vec2mat_backward")
        return gv


def pose_vec2mat(vec):
    """[B,6] (tx,ty,tz,rx,ry,rz) -> [B,4,4] as detectron2/geometry/pose_utils.py:130-137."""
    return _PoseVec2MatFn.apply(vec)
