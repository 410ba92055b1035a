"""ctypes binding of libsde_loss.so (include/sde_loss.h).  No fallback: if the library is
missing or a call fails, an exception is raised."""
from __future__ import annotations

import ctypes as C
import os

ABI_VERSION = 5   # SDE_ABI_VERSION of include/sde_loss.h this binding was written against
MAX_SCALES = 6
MAX_SOURCES = 4
MONO_SAVED_PLANES = 9   # SDE_MONO_SAVED_PLANES: planes per sample of a `warped` buffer
MOTION_SAVED_PLANES = 16  # SDE_MOTION_SAVED_PLANES
FLAG_AUTOMASK = 1
FLAG_REDUCE_MEAN = 2
FLAG_NO_TMA = 4            # SDE_MONO_NO_TMA
FLAG_NO_FLOW = 8
MOTION_FLAG_NO_TMA = 2     # SDE_MOTION_NO_TMA
DEPTH_MODES = {"depth": 0, "disp": 1, "logit": 2}   # SDE_DEPTH_IS_*
MAX_DIRS = 2
MOTION_FLAG_FIELD = 1
MOTION_N_LOSSES = 4

_f32p = C.c_void_p  # device pointers travel as integers


class MonoDesc(C.Structure):
    _fields_ = [
        ("batch", C.c_int32), ("n_scales", C.c_int32), ("n_sources", C.c_int32),
        ("height", C.c_int32 * MAX_SCALES), ("width", C.c_int32 * MAX_SCALES),
        ("full_height", C.c_int32), ("full_width", C.c_int32),
        ("ssim_weight", C.c_float), ("c1", C.c_float), ("c2", C.c_float), ("smooth_weight", C.c_float),
        ("flags", C.c_uint32),
        ("depth_mode", C.c_int32), ("min_depth", C.c_float), ("max_depth", C.c_float), ("norm_batch", C.c_int32),
    ]


class MonoBuffers(C.Structure):
    _fields_ = [
        ("target", _f32p * MAX_SCALES),
        ("source", (_f32p * MAX_SOURCES) * MAX_SCALES),
        ("depth", _f32p * MAX_SCALES),
        ("intrinsics", _f32p),
        ("pose", _f32p * MAX_SOURCES),
        ("losses", _f32p),
        ("argmin", _f32p * MAX_SCALES),
        ("saved_stats", _f32p),
        ("warped", (_f32p * MAX_SOURCES) * MAX_SCALES),
        ("smooth_g", _f32p * MAX_SCALES),
        ("grad_losses", _f32p),
        ("grad_depth", _f32p * MAX_SCALES),
        ("grad_pose", _f32p * MAX_SOURCES),
        ("workspace", _f32p),
    ]


class MotionDesc(C.Structure):
    _fields_ = [
        ("batch", C.c_int32), ("n_dirs", C.c_int32), ("height", C.c_int32), ("width", C.c_int32),
        ("scale_x", C.c_float), ("scale_y", C.c_float), ("ssim_weight", C.c_float),
        ("c1", C.c_float), ("c2", C.c_float), ("flags", C.c_uint32),
    ]


class MotionBuffers(C.Structure):
    _fields_ = [
        ("frame_a", _f32p * MAX_DIRS), ("frame_b", _f32p * MAX_DIRS),
        ("depth_a", _f32p * MAX_DIRS), ("depth_b", _f32p * MAX_DIRS),
        ("pose", _f32p * MAX_DIRS), ("field", _f32p * MAX_DIRS),
        ("intrinsics", _f32p),
        ("losses", _f32p), ("saved_stats", _f32p),
        ("occlusion", _f32p * MAX_DIRS), ("weight", _f32p * MAX_DIRS), ("coords", _f32p * MAX_DIRS),
        ("warped", _f32p * MAX_DIRS),
        ("grad_losses", _f32p),
        ("grad_depth_a", _f32p * MAX_DIRS), ("grad_pose", _f32p * MAX_DIRS), ("grad_field", _f32p * MAX_DIRS),
        ("workspace", _f32p),
    ]


class VsDesc(C.Structure):
    _fields_ = [("batch", C.c_int32), ("channels", C.c_int32), ("height", C.c_int32), ("width", C.c_int32),
                ("flags", C.c_uint32)]


class VsBuffers(C.Structure):
    _fields_ = [(n, _f32p) for n in (
        "image_b", "depth_a", "intrinsics", "rotation", "translation", "sampled", "depth_in_b", "coords", "valid",
        "grad_sampled", "grad_depth_in_b", "grad_coords", "grad_depth_a", "grad_rotation", "grad_translation",
        "grad_image_b", "workspace")]


class SsimDesc(C.Structure):
    _fields_ = [("batch", C.c_int32), ("channels", C.c_int32), ("height", C.c_int32), ("width", C.c_int32),
                ("c1", C.c_float), ("c2", C.c_float)]


class SsimBuffers(C.Structure):
    _fields_ = [(n, _f32p) for n in ("x", "y", "weight", "out", "avg_w", "grad_out", "grad_x", "grad_y", "workspace")]


class SmoothDesc(C.Structure):
    _fields_ = [("batch", C.c_int32), ("channels", C.c_int32), ("height", C.c_int32), ("width", C.c_int32)]


class SmoothBuffers(C.Structure):
    _fields_ = [(n, _f32p) for n in ("depth", "image", "loss", "saved_stats", "grad_loss", "grad_depth", "workspace")]


class McDesc(C.Structure):
    _fields_ = [("batch", C.c_int32), ("height", C.c_int32), ("width", C.c_int32)]


class McBuffers(C.Structure):
    _fields_ = [(n, _f32p) for n in ("coords", "mask", "rotation", "t_ab", "t_ba", "loss", "grad_loss", "grad_t_ab",
                                     "grad_t_ba", "grad_rotation", "workspace", "pose_ab", "pose_ba", "grad_pose_t_ab",
                                     "grad_pose_t_ba")]


class MfieldBuffers(C.Structure):
    _fields_ = [(n, _f32p) for n in ("pose", "field", "losses", "saved_stats", "grad_losses", "grad_field", "grad_pose_t",
                                     "workspace")]


class MregDesc(C.Structure):
    _fields_ = [("batch", C.c_int32), ("channels", C.c_int32), ("height", C.c_int32), ("width", C.c_int32)]


class MregBuffers(C.Structure):
    _fields_ = [(n, _f32p) for n in ("field", "loss", "saved_stats", "grad_loss", "grad_field", "workspace")]


class VarBuffers(C.Structure):
    _fields_ = [(n, _f32p) for n in ("depth", "loss", "saved_stats", "grad_loss", "grad_depth", "workspace")]


class SilogBuffers(C.Structure):
    _fields_ = [(n, _f32p) for n in ("depth_est", "depth_gt", "loss", "saved_stats", "grad_loss", "grad_depth_est",
                                     "workspace")]


class DispBuffers(C.Structure):
    _fields_ = [(n, _f32p) for n in ("disp", "scaled_disp", "depth", "grad_scaled_disp", "grad_depth", "grad_disp")]


class PoseVecBuffers(C.Structure):
    _fields_ = [(n, _f32p) for n in ("vec", "pose", "grad_pose", "grad_vec")]


class PyramidBuffers(C.Structure):
    _fields_ = [("src", _f32p * (MAX_SOURCES + 1)), ("dst", (_f32p * MAX_SCALES) * (MAX_SOURCES + 1))]


class SdeError(RuntimeError):
    pass


_lib = None


def lib_path() -> str:
    # SDE_LIB_PATH: developer switch for A/B timing of two builds in one GPU session (scratch/ab.sh)
    if os.environ.get("SDE_LIB_PATH"):
        return os.environ["SDE_LIB_PATH"]
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "libsde_loss.so")


def load():
    """Loads the shared library once.  Raises if it has not been built
    (python -m simpledepthestimation_b200.build)."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise SdeError(f"{path} not found: build it with `python -m simpledepthestimation_b200.build` "
                       "(there is no CPU or PyTorch fallback)")
    lib = C.CDLL(path)
    lib.sde_version.restype = C.c_int
    if lib.sde_version() != ABI_VERSION:
        raise SdeError(f"libsde_loss.so has ABI version {lib.sde_version()}, the binding expects {ABI_VERSION}: rebuild it")
    lib.sde_strerror.restype = C.c_char_p
    lib.sde_strerror.argtypes = [C.c_int]
    lib.sde_last_cuda_error.restype = C.c_char_p
    lib.sde_reload_env.restype = None
    lib.sde_mono_workspace_bytes.restype = C.c_size_t
    lib.sde_mono_workspace_bytes.argtypes = [C.POINTER(MonoDesc)]
    for name in ("sde_mono_loss_forward", "sde_mono_loss_backward", "sde_mono_loss_step"):
        fn = getattr(lib, name)
        fn.restype = C.c_int
        fn.argtypes = [C.POINTER(MonoDesc), C.POINTER(MonoBuffers), C.c_void_p]
    lib.sde_motion_workspace_bytes.restype = C.c_size_t
    lib.sde_motion_workspace_bytes.argtypes = [C.POINTER(MotionDesc)]
    for name in ("sde_motion_loss_forward", "sde_motion_loss_backward"):
        fn = getattr(lib, name)
        fn.restype = C.c_int
        fn.argtypes = [C.POINTER(MotionDesc), C.POINTER(MotionBuffers), C.c_void_p]
    for prefix, D, B in (("sde_view_synthesis", VsDesc, VsBuffers), ("sde_ssim", SsimDesc, SsimBuffers),
                         ("sde_smoothness", SmoothDesc, SmoothBuffers)):
        ws = getattr(lib, prefix + "_workspace_bytes")
        ws.restype, ws.argtypes = C.c_size_t, [C.POINTER(D)]
        for suffix in ("_forward", "_backward"):
            fn = getattr(lib, prefix + suffix)
            fn.restype, fn.argtypes = C.c_int, [C.POINTER(D), C.POINTER(B), C.c_void_p]
    lib.sde_motion_consistency_workspace_bytes.restype = C.c_size_t
    lib.sde_motion_consistency_workspace_bytes.argtypes = [C.POINTER(McDesc)]
    for suffix in ("_forward", "_backward"):
        fn = getattr(lib, "sde_motion_consistency" + suffix)
        fn.restype, fn.argtypes = C.c_int, [C.POINTER(McDesc), C.POINTER(McBuffers), C.c_void_p]
    lib.sde_motion_reg_workspace_bytes.restype = C.c_size_t
    lib.sde_motion_reg_workspace_bytes.argtypes = [C.POINTER(MregDesc)]
    for name in ("smoothness", "sparsity"):
        for suffix in ("_forward", "_backward"):
            fn = getattr(lib, f"sde_motion_{name}{suffix}")
            fn.restype, fn.argtypes = C.c_int, [C.POINTER(MregDesc), C.POINTER(MregBuffers), C.c_void_p]
    lib.sde_motion_field_reg_workspace_bytes.restype = C.c_size_t
    lib.sde_motion_field_reg_workspace_bytes.argtypes = [C.POINTER(MregDesc)]
    for suffix in ("_forward", "_backward"):
        fn = getattr(lib, "sde_motion_field_reg" + suffix)
        fn.restype, fn.argtypes = C.c_int, [C.POINTER(MregDesc), C.POINTER(MfieldBuffers), C.c_void_p]
    lib.sde_variance_workspace_bytes.restype, lib.sde_variance_workspace_bytes.argtypes = C.c_size_t, [C.c_int64]
    for suffix in ("_forward", "_backward"):
        fn = getattr(lib, "sde_variance_loss" + suffix)
        fn.restype, fn.argtypes = C.c_int, [C.c_int64, C.POINTER(VarBuffers), C.c_void_p]
    lib.sde_silog_workspace_bytes.restype, lib.sde_silog_workspace_bytes.argtypes = C.c_size_t, [C.c_int64]
    for suffix in ("_forward", "_backward"):
        fn = getattr(lib, "sde_silog_loss" + suffix)
        fn.restype, fn.argtypes = C.c_int, [C.c_int64, C.c_float, C.POINTER(SilogBuffers), C.c_void_p]
        fn = getattr(lib, "sde_disp_to_depth" + suffix)
        fn.restype, fn.argtypes = C.c_int, [C.c_int64, C.c_float, C.c_float, C.POINTER(DispBuffers), C.c_void_p]
        fn = getattr(lib, "sde_pose_vec2mat" + suffix)
        fn.restype, fn.argtypes = C.c_int, [C.c_int32, C.POINTER(PoseVecBuffers), C.c_void_p]
    lib.sde_resize_bilinear.restype = C.c_int
    lib.sde_resize_bilinear.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                        C.c_void_p]
    lib.sde_resize_pyramid.restype = C.c_int
    lib.sde_resize_pyramid.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_int32),
                                       C.POINTER(C.c_int32), C.POINTER(PyramidBuffers), C.c_void_p]
    for name in ("sde_resize_avgpool_forward", "sde_resize_avgpool_backward"):
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = C.c_int, lib.sde_resize_bilinear.argtypes
    lib.sde_resize_pyramid_u8.restype = C.c_int
    lib.sde_resize_pyramid_u8.argtypes = lib.sde_resize_pyramid.argtypes
    _lib = lib
    return lib


def tma_disabled() -> bool:
    """SDE_DISABLE_TMA=1 (testing): plans built while it is set stage their tile planes with plain loads; the choice is
    recorded in the plan's descriptor, so its forward and backward calls always agree."""
    return os.environ.get("SDE_DISABLE_TMA", "") == "1"


def check(status: int, what: str):
    if status != 0:
        lib = load()
        msg = lib.sde_strerror(status).decode()
        if status == -3:
            msg += ": " + lib.sde_last_cuda_error().decode()
        raise SdeError(f"{what} failed: {msg} (status {status})")
