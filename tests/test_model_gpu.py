"""MonoDepth2Model drop-in: same constructor, batch dict and output keys as the reference
meta-architecture; numbers against the reference's golden outputs."""
import pytest
import torch
import torch.nn as nn

from helpers import golden_mono_inputs, load_golden, rel_err
from simpledepthestimation_b200.synthetic import euler_pose

pytestmark = pytest.mark.gpu


class AttrDict(dict):
    __getattr__ = dict.__getitem__


class _Inject(nn.Module):
    def __init__(self, cfg=None):
        super().__init__()
        self.payload = {}

    def forward(self, batch):
        batch.update(self.payload)
        return batch


def make_cfg(**loss):
    L = AttrDict(SSIM_WEIGHT=0.85, C1=1e-4, C2=9e-4, CLIP=0.0, AUTOMASK=True, SMOOTHNESS_WEIGHT=1e-3,
                 PHOTOMETRIC_REDUCE="min", SUPERVISED_WEIGHT=0.0, VARIANCE_FOCUS=0.85, VAR_LOSS_WEIGHT=0.0)
    L.update(loss)
    return AttrDict(LOSS=L, MODEL=AttrDict(META_ARCHITECTURE="MonoDepth2Model", DEVICE="cuda:0",
                                           PIXEL_MEAN=[0.45, 0.45, 0.45], PIXEL_STD=[0.225, 0.225, 0.225],
                                           DEPTH_NET=AttrDict(NAME="InjectDepth"), POSE_NET=AttrDict(NAME="InjectPose")))


@pytest.fixture(scope="module")
def registered(sde_lib):
    from simpledepthestimation_b200.modeling import DEPTH_NET_REGISTRY, POSE_NET_REGISTRY

    if "InjectDepth" not in DEPTH_NET_REGISTRY:
        DEPTH_NET_REGISTRY._do_register("InjectDepth", _Inject)
        POSE_NET_REGISTRY._do_register("InjectPose", _Inject)
    return True


def test_registry_build_and_training_forward_matches_reference(registered):
    from simpledepthestimation_b200.modeling import build_model

    g = load_golden("mono_2x32x64")
    inp = golden_mono_inputs(g)
    model = build_model(make_cfg()).train()
    dev = model.device
    depth = [d.to(dev).requires_grad_() for d in inp["depth"]]
    vecs = [v.to(dev).requires_grad_() for v in inp["pose_vec"]]
    model.depth_net.payload = {"depth_pred": depth}
    model.pose_net.payload = {"pose_pred": [euler_pose(v) for v in vecs]}
    # host-side batch, exactly what the reference's loader hands to forward()
    batch = {"img": inp["img"], "ctx_img": list(inp["ctx"]), "img_orig": inp["img"], "ctx_img_orig": list(inp["ctx"]),
             "intrinsics": inp["K"]}
    out = model(batch)
    assert set(out) == {"rec_loss", "smooth_loss"}
    losses = sum(v for k, v in out.items() if "loss" in k)  # the trainer's reduction (train.py:91-92)
    assert torch.isfinite(losses)
    losses.backward()
    # the pyramid is resized on the GPU here (library bilinear), so allow its ulp-level differences
    assert rel_err(out["rec_loss"].detach(), g["rec_loss_f64"]) < 1e-5
    assert rel_err(out["smooth_loss"].detach(), g["smooth_loss_f64"]) < 1e-5
    for i, d in enumerate(depth):
        err = (d.grad.cpu().double() - torch.from_numpy(g[f"grad_depth{i}_f64"])).abs()
        assert float(torch.quantile((err / err.new_tensor(g[f"grad_depth{i}_f64"]).abs().max()).flatten(), 0.99)) < 1e-4
    for j, v in enumerate(vecs):
        assert rel_err(v.grad, g[f"grad_pose_vec{j}_f64"]) < 1e-3


def test_eval_mode_returns_depth_pred(registered):
    from simpledepthestimation_b200.modeling import build_model

    g = load_golden("mono_2x32x64")
    inp = golden_mono_inputs(g)
    model = build_model(make_cfg()).eval()
    model.depth_net.payload = {"depth_pred": [d.to(model.device) for d in inp["depth"]]}
    out = model({"img": inp["img"]})
    assert set(out) == {"depth_pred"} and out["depth_pred"].shape == inp["depth"][0].shape


def test_unsupported_options_fail_loudly(registered):
    from simpledepthestimation_b200.modeling import build_model

    with pytest.raises(NotImplementedError):
        build_model(make_cfg(PHOTOMETRIC_REDUCE="median"))


@pytest.mark.parametrize("over", [dict(CLIP=0.5), dict(CLIP=1.0, AUTOMASK=False), dict(CLIP=0.5, PHOTOMETRIC_REDUCE="mean")])
def test_clip_configuration_runs_the_unfused_operator_loop(registered, over):
    """LOSS.CLIP > 0 (MonoDepth2.py:146-149: every photometric map capped at its mean + CLIP * std) goes through the
    reference's loop over the stand-alone CUDA operators; losses and gradients against the oracle."""
    from helpers import port_mono_from_vec
    from simpledepthestimation_b200.modeling import build_model
    from simpledepthestimation_b200.synthetic import mono_inputs

    inp = mono_inputs(2, 48, 80, seed=9)
    model = build_model(make_cfg(**over)).train()
    dev = model.device
    depth = [d.to(dev).requires_grad_() for d in inp["depth"]]
    vecs = [v.to(dev).requires_grad_() for v in inp["pose_vec"]]
    model.depth_net.payload = {"depth_pred": depth}
    model.pose_net.payload = {"pose_pred": [euler_pose(v) for v in vecs]}
    out = model({"img": inp["img"], "ctx_img": list(inp["ctx"]), "img_orig": inp["img"], "ctx_img_orig": list(inp["ctx"]),
                 "intrinsics": inp["K"]})
    (out["rec_loss"] + out["smooth_loss"]).backward()
    kw = dict(clip=over["CLIP"], automask=over.get("AUTOMASK", True), reduce=over.get("PHOTOMETRIC_REDUCE", "min"))
    ref = port_mono_from_vec(inp, torch.float64, **kw)
    assert rel_err(out["rec_loss"].detach(), ref["rec_loss"].detach()) < 1e-5
    assert rel_err(out["smooth_loss"].detach(), ref["smooth_loss"].detach()) < 1e-5
    for d, r in zip(depth, ref["grad_depth"]):
        err = (d.grad.double().cpu() - r).abs() / r.abs().max()
        assert float(torch.quantile(err.flatten(), 0.99)) < 1e-4 and float(err.max()) < 5e-3
    for v, r in zip(vecs, ref["grad_pose_vec"]):
        assert rel_err(v.grad, r) < 2e-3
