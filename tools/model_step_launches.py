"""One training step of the drop-in models with injected depth / pose predictions, for a kernel launch
list under ncu:  ncu --metrics gpu__time_duration.sum --csv python tools/model_step_launches.py [motion|mono] [B H W]
Prints the losses; the launch list shows which kernels a model step issues (library kernels = torch eager ops)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from simpledepthestimation_b200.geometry.pose_utils import pose_vec2mat  # noqa: E402
from simpledepthestimation_b200.modeling import DEPTH_NET_REGISTRY, POSE_NET_REGISTRY, build_model  # noqa: E402
from simpledepthestimation_b200.synthetic import mono_inputs, motion_inputs  # noqa: E402


class AttrDict(dict):
    __getattr__ = dict.__getitem__


class Inject(torch.nn.Module):
    def __init__(self, cfg=None):
        super().__init__()
        self.payload = {}

    def forward(self, batch):
        batch.update(self.payload)
        return batch


DEPTH_NET_REGISTRY._do_register("ToolInjectDepth", Inject)
POSE_NET_REGISTRY._do_register("ToolInjectPose", Inject)
kind = sys.argv[1] if len(sys.argv) > 1 else "motion"
B, H, W = (int(v) for v in sys.argv[2:5]) if len(sys.argv) > 4 else ((4, 1280, 1920) if kind == "motion" else (12, 192, 640))
model_cfg = dict(DEVICE="cuda:0", PIXEL_MEAN=[0.45] * 3, PIXEL_STD=[0.225] * 3, DEPTH_NET=AttrDict(NAME="ToolInjectDepth"),
                 POSE_NET=AttrDict(NAME="ToolInjectPose", USE_DEPTH=True))
if kind == "motion":
    loss = AttrDict(NUM_SCALES=1, SSIM_WEIGHT=3.0, C1="inf", C2=9e-6, CLIP=0.0, DEPTH_L1_WEIGHT=0.0, SMOOTHNESS_WEIGHT=1e-3,
                    SUPERVISED_WEIGHT=0.0, VARIANCE_FOCUS=0.85, VAR_LOSS_WEIGHT=0.0, MOTION_SMOOTHNESS_WEIGHT=1.0,
                    MOTION_SPARSITY_WEIGHT=0.2, ROT_CYCLE_WEIGHT=1e-3, TRANS_CYCLE_WEIGHT=5e-2, SCALE_NORMALIZE=False)
    model = build_model(AttrDict(LOSS=loss, MODEL=AttrDict(META_ARCHITECTURE="MotionLearningModel", **model_cfg))).train()
    inp = motion_inputs(B, H, W, seed=0)
    dev = model.device
    g = lambda t: t.to(dev).contiguous()  # noqa: E731
    d = g(torch.cat([inp["depth1"], inp["depth2"]], 0)).requires_grad_()
    vec, mo = g(inp["pose_vec"]).requires_grad_(), g(inp["motion"]).requires_grad_()
    feed = {"img": g(inp["img1"]), "ctx_img": [g(inp["img2"])], "intrinsics": g(inp["K"])}

    def step():
        model.depth_net.payload = {"depth_pred": [d]}
        model.pose_net.payload = {"pose_pred": pose_vec2mat(vec), "motion_pred": mo}
        out = model(dict(feed))
        total = sum(v for k, v in out.items() if "loss" in k)
        total.backward()
        return {k: float(v) for k, v in out.items() if "loss" in k}
else:
    loss = AttrDict(SSIM_WEIGHT=0.85, C1=1e-4, C2=9e-4, CLIP=0.0, AUTOMASK=True, SMOOTHNESS_WEIGHT=1e-3,
                    PHOTOMETRIC_REDUCE="min", SUPERVISED_WEIGHT=0.0, VARIANCE_FOCUS=0.85, VAR_LOSS_WEIGHT=0.0)
    model = build_model(AttrDict(LOSS=loss, MODEL=AttrDict(META_ARCHITECTURE="MonoDepth2Model", **model_cfg))).train()
    inp = mono_inputs(B, H, W, seed=0)
    dev = model.device
    g = lambda t: t.to(dev).contiguous()  # noqa: E731
    depth = [g(x).requires_grad_() for x in inp["depth"]]
    vecs = [g(v).requires_grad_() for v in inp["pose_vec"]]
    img, ctx, K = g(inp["img"]), [g(c) for c in inp["ctx"]], g(inp["K"])

    def step():
        model.depth_net.payload = {"depth_pred": depth}
        model.pose_net.payload = {"pose_pred": [pose_vec2mat(v) for v in vecs]}
        out = model({"img": img, "ctx_img": ctx, "img_orig": img, "ctx_img_orig": ctx, "intrinsics": K})
        (out["rec_loss"] + out["smooth_loss"]).backward()
        return {k: float(v) for k, v in out.items() if "loss" in k}

step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
print(step())
torch.cuda.synchronize()
torch.cuda.profiler.stop()
if os.environ.get("SDE_CPROFILE"):
    import cProfile
    import pstats
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(5):
        step()
    torch.cuda.synchronize()
    pr.disable()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(45)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    step()
e1.record()
torch.cuda.synchronize()
print(f"{kind} model step {B}x{H}x{W}: {e0.elapsed_time(e1) / 10:.3f} ms")
