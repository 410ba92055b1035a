"""Import the real reference (read-only) for oracle validation and for the CPU arm of bench.py.

TEST / BENCH INFRASTRUCTURE ONLY.  In the build container the reference is /root/reference; on the GPU box
(where /root/reference does not exist) it is the unmodified copy that oracle/make_ref.sh stages under the
git-ignored oracle/_ref/ -- available() tells whether either is present.

``import detectron2`` fails in this image (fvcore/iopath/... absent), so the
package ``__init__`` files are bypassed with empty module stubs whose
``__path__`` points into the reference tree; the arithmetic files
(geometry/camera.py, modeling/losses/*.py, modeling/meta_arch/MonoDepth2.py,
MotionLearning.py) only need torch and import cleanly.

One semantic fix is applied, documented in SURVEY.md section 0: MonoDepth2Model passes
the translation as [B,3,1,1] (MonoDepth2.py:94) but view_synthesis takes the
output size from it (camera.py:169), giving NaN at HEAD; the wrapper broadcasts
``t`` to [B,3,h,w], which is the only meaningful reading.
"""
from __future__ import annotations

import os
import sys
import types

import torch
import torch.nn as nn

def _ref_root():
    """/root/reference in the build container; on the GPU box the copy staged by oracle/make_ref.sh (git-ignored)."""
    staged = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
    for root in (os.environ.get("SDE_REFERENCE_ROOT"), "/root/reference", staged):
        if root and os.path.isdir(os.path.join(root, "detectron2", "modeling", "meta_arch")):
            return root
    return "/root/reference"


REF_ROOT = _ref_root()
_REF = os.path.join(REF_ROOT, "detectron2")


def available() -> bool:
    return os.path.isdir(_REF)


class _Inject(nn.Module):
    """Stand-in for depth_net / pose_net: merges a payload into the batch."""

    def __init__(self):
        super().__init__()
        self.payload = {}

    def forward(self, batch):
        batch.update(self.payload)
        return batch


class AttrDict(dict):
    __getattr__ = dict.__getitem__

    def __setattr__(self, k, v):
        self[k] = v


def _stub(name, path=None):
    m = types.ModuleType(name)
    m.__package__ = name
    if path is not None:
        m.__path__ = [path]
    sys.modules[name] = m
    return m


_loaded = None


def load():
    """Returns a namespace with the reference's modules (camera, ssim_loss, ...)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    for name in [n for n in sys.modules if n == "detectron2" or n.startswith("detectron2.")]:
        del sys.modules[name]
    _stub("detectron2", _REF)
    _stub("detectron2.modeling", _REF + "/modeling")
    _stub("detectron2.utils", _REF + "/utils")
    _stub("detectron2.modeling.meta_arch", _REF + "/modeling/meta_arch")
    reg = type("Registry", (), {"register": lambda self: (lambda cls: cls)})()
    _stub("detectron2.modeling.meta_arch.build").META_ARCH_REGISTRY = reg
    _stub("detectron2.modeling.depth_net").build_depth_net = lambda cfg: _Inject()
    _stub("detectron2.modeling.pose_net").build_pose_net = lambda cfg: _Inject()

    import importlib
    import warnings

    warnings.filterwarnings("ignore", message=".*meshgrid.*")
    ns = types.SimpleNamespace()
    ns.camera = importlib.import_module("detectron2.geometry.camera")
    ns.pose_utils = importlib.import_module("detectron2.geometry.pose_utils")
    ns.ssim_loss = importlib.import_module("detectron2.modeling.losses.ssim_loss")
    ns.smoothness_loss = importlib.import_module("detectron2.modeling.losses.smoothness_loss")
    ns.motion_loss = importlib.import_module("detectron2.modeling.losses.motion_loss")
    ns.losses = importlib.import_module("detectron2.modeling.losses.losses")
    ns.MonoDepth2 = importlib.import_module("detectron2.modeling.meta_arch.MonoDepth2")
    ns.MotionLearning = importlib.import_module("detectron2.modeling.meta_arch.MotionLearning")
    # the t-shape fix (see module docstring)
    vs = ns.camera.view_synthesis
    ns.MonoDepth2.view_synthesis = lambda img, d, K, R, t: vs(img, d, K, R, t.expand(-1, -1, *d.shape[-2:]))
    _loaded = ns
    return ns


def mono_cfg(**over):
    loss = AttrDict(SSIM_WEIGHT=0.85, C1=1e-4, C2=9e-4, CLIP=0.0, AUTOMASK=True, SMOOTHNESS_WEIGHT=1e-3,
                    PHOTOMETRIC_REDUCE="min", SUPERVISED_WEIGHT=0.0, VARIANCE_FOCUS=0.85, VAR_LOSS_WEIGHT=0.0)
    loss.update(over)
    return AttrDict(LOSS=loss, MODEL=AttrDict(PIXEL_MEAN=[0.45, 0.45, 0.45], PIXEL_STD=[0.225, 0.225, 0.225]))


def motion_cfg(**over):
    loss = AttrDict(NUM_SCALES=1, SSIM_WEIGHT=3.0, C1="inf", C2=9e-6, CLIP=0.0, DEPTH_L1_WEIGHT=0.0,
                    SMOOTHNESS_WEIGHT=1e-3, SUPERVISED_WEIGHT=0.0, VARIANCE_FOCUS=0.85, VAR_LOSS_WEIGHT=0.0,
                    MOTION_SMOOTHNESS_WEIGHT=1.0, MOTION_SPARSITY_WEIGHT=0.2, ROT_CYCLE_WEIGHT=1e-3,
                    TRANS_CYCLE_WEIGHT=5e-2, SCALE_NORMALIZE=False)
    loss.update(over)
    model = AttrDict(PIXEL_MEAN=[0.45, 0.45, 0.45], PIXEL_STD=[0.225, 0.225, 0.225],
                     POSE_NET=AttrDict(USE_DEPTH=True))
    return AttrDict(LOSS=loss, MODEL=model)


def run_mono(inp, dtype=torch.float64, grads=True, **cfg_over):
    """Runs the reference's own MonoDepth2Model.forward (+backward) on `inp`
    (a dict from simpledepthestimation_b200.synthetic.mono_inputs).  Returns dict
    with rec_loss, smooth_loss, grad_depth (list), grad_pose_vec (list)."""
    ns = load()
    model = ns.MonoDepth2.MonoDepth2Model(mono_cfg(**cfg_over)).train().to(dtype)
    img = inp["img"].to(dtype)
    ctx = [c.to(dtype) for c in inp["ctx"]]
    depth = [d.to(dtype).clone().requires_grad_(grads) for d in inp["depth"]]
    vecs = [v.to(dtype).clone().requires_grad_(grads) for v in inp["pose_vec"]]
    model.depth_net.payload = {"depth_pred": depth}
    model.pose_net.payload = {"pose_pred": [ns.pose_utils.pose_vec2mat(v) for v in vecs]}
    out = model({"img": img, "ctx_img": ctx, "img_orig": img, "ctx_img_orig": ctx,
                 "intrinsics": inp["K"].to(dtype)})
    res = {k: v.detach() for k, v in out.items() if torch.is_tensor(v) and v.dim() == 0}
    if grads:
        total = sum(v for k, v in out.items() if "loss" in k)
        total.backward()
        res["grad_depth"] = [d.grad for d in depth]
        res["grad_pose_vec"] = [v.grad for v in vecs]
    return res


def run_motion(inp, dtype=torch.float64, grads=True, with_motion=True, model_over=None, extra_batch=None, **cfg_over):
    """Runs the reference's MotionLearningModel.forward (+backward of summed *loss* keys).  `model_over` updates
    cfg.MODEL (WITH_MASK, MASK_DILATION), `extra_batch` adds batch entries (mask, ctx_mask)."""
    ns = load()
    cfg = motion_cfg(**cfg_over)
    cfg.MODEL.update(model_over or {})
    model = ns.MotionLearning.MotionLearningModel(cfg).train().to(dtype)
    d1 = inp["depth1"].to(dtype).clone().requires_grad_(grads)
    d2 = inp["depth2"].to(dtype).clone().requires_grad_(grads)
    vec = inp["pose_vec"].to(dtype).clone().requires_grad_(grads)
    mo = inp["motion"].to(dtype).clone().requires_grad_(grads)
    model.depth_net.payload = {"depth_pred": [torch.cat([d1, d2], 0)]}
    payload = {"pose_pred": ns.pose_utils.pose_vec2mat(vec)}
    if with_motion:
        payload["motion_pred"] = mo
    model.pose_net.payload = payload
    feed = {"img": inp["img1"].to(dtype), "ctx_img": [inp["img2"].to(dtype)], "intrinsics": inp["K"].to(dtype)}
    feed.update(extra_batch or {})
    batch = model(feed)
    res = {k: v.detach() for k, v in batch.items() if "loss" in k and torch.is_tensor(v)}
    res["depth_proximity_weight"] = [tuple(t.detach() for t in pair) for pair in batch["depth_proximity_weight"]]
    res["overall_motion"] = [tuple(t.detach() for t in pair) for pair in batch["overall_motion"]]
    if grads:
        total = sum(v for k, v in batch.items() if "loss" in k)
        total.backward()
        res["grad_depth1"], res["grad_depth2"] = d1.grad, d2.grad
        res["grad_pose_vec"] = vec.grad
        res["grad_motion"] = mo.grad if with_motion else None
    return res
