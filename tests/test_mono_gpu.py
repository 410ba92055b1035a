"""Parity of the fused CUDA loss (through the C ABI) with the oracle / the reference's golden
outputs.  Tolerances (BASELINE.json north_star): loss 1e-5 relative; argmin maps exact except
where the two best candidates are within 1e-5; gradients 1e-4 relative (to the largest gradient
of the tensor) on decision-stable pixels against the fp64 reference."""
import numpy as np
import pytest
import torch

from oracle import port
from helpers import (build_pyramid, golden_mono_inputs, gpu_mono_from_pose, gpu_mono_from_vec, load_golden, oracle_mono,
                     port_mono_from_vec, rel_err, stable_mask)
from simpledepthestimation_b200.synthetic import euler_pose, mono_inputs

pytestmark = pytest.mark.gpu
LOSS_TOL, GRAD_TOL, TIE_TOL = 1e-5, 1e-4, 1e-5

VARIANTS = {"": ({}, {}), "_noauto": (dict(automask=False), dict(automask=False)),
            "_mean": (dict(reduce="mean"), dict(reduce="mean")), "_l1": (dict(ssim_weight=0.0), dict(ssim_w=0.0))}


@pytest.fixture(scope="module")
def dev(sde_lib):
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda", 0)


def check_against(out, ref_rec, ref_smooth, ref_gd, ref_gv, masks=None, ref_gv32=None, tag=""):
    from parity_log import record
    achieved = {"rec_loss_rel": rel_err(out["rec_loss"], ref_rec), "smooth_loss_rel": rel_err(out["smooth_loss"], ref_smooth)}
    for i, (g, r) in enumerate(zip(out["grad_depth"], ref_gd)):
        r = torch.as_tensor(r, dtype=torch.float64)
        e = ((g.double() - r).abs() / r.abs().max())[:, 0]
        achieved[f"grad_depth_s{i}_max" + ("_stable" if masks is not None else "")] = float((e[masks[i]] if masks is not None else e).max())
    for j, (g, r) in enumerate(zip(out["grad_pose_vec"], ref_gv)):
        achieved[f"grad_pose_vec{j}_rel"] = rel_err(g, r)
        if ref_gv32 is not None:
            achieved[f"grad_pose_vec{j}_rel_reference_fp32"] = rel_err(ref_gv32[j], r)
    record(**{("variant" + tag if tag else "variant_default"): achieved})
    assert rel_err(out["rec_loss"], ref_rec) < LOSS_TOL
    assert rel_err(out["smooth_loss"], ref_smooth) < LOSS_TOL
    for i, (g, r) in enumerate(zip(out["grad_depth"], ref_gd)):
        r = torch.as_tensor(r, dtype=torch.float64)
        err = (g.double() - r).abs() / r.abs().max()
        if masks is not None:
            frac_unstable = 1.0 - float(masks[i].double().mean())
            assert frac_unstable < 0.05, f"scale {i}: {frac_unstable:.3f} of pixels decision-unstable"
            err = err[:, 0][masks[i]]
        assert float(err.max()) < GRAD_TOL, f"grad_depth[{i}] {float(err.max()):.2e}"
    for j, (g, r) in enumerate(zip(out["grad_pose_vec"], ref_gv)):
        # A pose gradient is a sum over all pixels: one decision flip (bilinear cell, argmin, |.| sign)
        # at a coarse-scale pixel moves it by ~1e-3 in any fp32 implementation.  Where the reference's
        # own fp32 run deviates from its fp64 run by more than the tolerance, bound ours by 3x that.
        tol = GRAD_TOL if ref_gv32 is None else max(GRAD_TOL, 3 * rel_err(ref_gv32[j], r))
        assert rel_err(g, r) < tol, f"grad_pose_vec[{j}] {rel_err(g, r):.2e} (tol {tol:.2e})"


@pytest.mark.parametrize("name", ["mono_2x32x64", "mono_1x50x70", "mono_1x48x160_bigpose"])
def test_golden_reference_outputs(dev, name):
    """CUDA path vs outputs of the real reference (fp64 run) on the committed inputs."""
    g = load_golden(name)
    inp = golden_mono_inputs(g)
    tgt, src = build_pyramid(inp)
    for tag, (plan_kw, _) in VARIANTS.items():
        if f"rec_loss{tag}_f64" not in g:
            continue
        out = gpu_mono_from_vec(inp, dev, **plan_kw)
        n = len(inp["depth"])
        masks = None
        if tag in ("", "_noauto"):
            masks = [stable_mask(inp, tgt, src, i, automask=(tag == "")) for i in range(n)]
        check_against(out, g[f"rec_loss{tag}_f64"], g[f"smooth_loss{tag}_f64"],
                      [g[f"grad_depth{i}{tag}_f64"] for i in range(n)],
                      [g[f"grad_pose_vec{j}{tag}_f64"] for j in range(len(inp["pose_vec"]))], masks,
                      [g[f"grad_pose_vec{j}{tag}_f32"] for j in range(len(inp["pose_vec"]))], tag=tag)
        if tag == "":
            for i, a in enumerate(out["argmin"]):
                mism = a.numpy() != g[f"argmin{i}"]
                assert not (mism & (g[f"argmin_gap{i}"] > TIE_TOL)).any(), f"argmin scale {i}"


def test_cfg1_kitti_shape_against_live_oracle(dev):
    """BASELINE.json configs[0]: 640x192, batch 1, 4 scales, 2 sources."""
    inp = mono_inputs(1, 192, 640)
    tgt, src = build_pyramid(inp)
    ref = port_mono_from_vec(inp, torch.float64)
    out = gpu_mono_from_vec(inp, dev)
    assert rel_err(out["rec_loss"], ref["rec_loss"].detach()) < LOSS_TOL
    assert rel_err(out["smooth_loss"], ref["smooth_loss"].detach()) < LOSS_TOL
    g = load_golden("mono_cfg1_1x192x640")
    if golden_mono_inputs(g) is not None:  # same bits as in the build container -> compare with the real reference
        assert rel_err(out["rec_loss"], g["rec_loss_f64"]) < LOSS_TOL
        assert rel_err(out["smooth_loss"], g["smooth_loss_f64"]) < LOSS_TOL
    for i, a in enumerate(out["argmin"]):
        top2 = ref["cand"][i].topk(2, dim=1, largest=False)[0]
        gap = top2[:, 1] - top2[:, 0]
        mism = a.long() != ref["argmin"][i]
        assert not (mism & (gap > TIE_TOL)).any()
        assert float(mism.double().mean()) < 1e-3
    ref32 = port_mono_from_vec(inp, torch.float32)
    for i, (gd, r) in enumerate(zip(out["grad_depth"], ref["grad_depth"])):
        m = stable_mask(inp, tgt, src, i)
        assert float(m.double().mean()) > 0.95
        err64 = ((gd.double() - r).abs() / r.abs().max())[:, 0]
        err32 = ((gd.double() - ref32["grad_depth"][i].double()).abs() / r.abs().max())[:, 0]
        # On decision-stable pixels the gradient must agree with the reference at 1e-4 of the largest
        # gradient.  At 640 px wide an fp32 pixel coordinate has 3e-5..6e-5 px resolution and the SSIM
        # gradient of low-variance windows amplifies that ~1e-3-fold, so a few stable pixels of the
        # reference's OWN fp32 run sit 3e-4 from its fp64 run (measured: 10 of 122880 at scale 0); there
        # the kernel must match the fp32 reference instead, or -- two fp32 evaluations of such an
        # ill-conditioned pixel differ by their rounding order alone -- stay within 3x the reference's own
        # fp32 deviation at that pixel.
        dev32 = ((ref32["grad_depth"][i].double() - r).abs() / r.abs().max())[:, 0]
        ok = (err64 < GRAD_TOL) | (err32 < GRAD_TOL) | (err64 < 3.0 * dev32)
        assert bool(ok[m].all())
        assert float(torch.quantile(err64.flatten(), 0.999)) < GRAD_TOL
        assert int((err64[m] > GRAD_TOL).sum()) <= 20
    from parity_log import record
    ach = {"rec_loss_rel": rel_err(out["rec_loss"], ref["rec_loss"].detach()),
           "smooth_loss_rel": rel_err(out["smooth_loss"], ref["smooth_loss"].detach())}
    for j, (gv, r) in enumerate(zip(out["grad_pose_vec"], ref["grad_pose_vec"])):
        # a handful of decision-flip pixels at the coarse scales move a pose gradient by ~1e-3 in ANY
        # fp32 implementation; bound ours by the reference's own fp32 deviation
        own = rel_err(gv, r)
        ref_dev = rel_err(ref32["grad_pose_vec"][j], r)
        ach[f"grad_pose_vec{j}_rel"], ach[f"grad_pose_vec{j}_rel_reference_fp32"] = own, ref_dev
        assert own < max(GRAD_TOL, 5 * ref_dev), (own, ref_dev)
    record(config="cfg1 1x192x640", **ach)


def test_bit_identical_across_runs(dev):
    inp = mono_inputs(3, 96, 320, seed=11)
    outs = [gpu_mono_from_vec(inp, dev) for _ in range(4)]
    for o in outs[1:]:
        assert torch.equal(o["rec_loss"], outs[0]["rec_loss"]) and torch.equal(o["smooth_loss"], outs[0]["smooth_loss"])
        for a, b in zip(o["grad_depth"] + o["grad_pose_vec"] + o["argmin"],
                        outs[0]["grad_depth"] + outs[0]["grad_pose_vec"] + outs[0]["argmin"]):
            assert torch.equal(a, b)


@pytest.mark.parametrize("B,H,W", [(12, 192, 640), (8, 320, 1024)])
def test_batch_linearity_at_full_size(dev, B, H, W):
    """BASELINE.json configs[1] (640x192, batch 12) and configs[2] (1024x320, batch 8): the batch loss is the
    mean of the per-sample losses and a sample's gradients are 1/B of its stand-alone gradients
    (size-independent check that needs no CPU oracle at this size)."""
    inp = mono_inputs(B, H, W, seed=2)
    # pyramid and pose matrices are built once on the CPU and sliced, so the batched and the
    # per-sample runs see identical input bits (library resize / bmm kernels vary with batch size)
    tgt, src = build_pyramid(inp)
    pose = [euler_pose(v) for v in inp["pose_vec"]]
    full = gpu_mono_from_pose(inp["depth"], inp["K"], pose, tgt, src, (H, W), dev)
    rec, sm = [], []
    for b in range(B):
        sl = slice(b, b + 1)
        o = gpu_mono_from_pose([d[sl] for d in inp["depth"]], inp["K"][sl], [p[sl] for p in pose],
                               [t[sl] for t in tgt], [[x[sl] for x in row] for row in src], (H, W), dev)
        rec.append(float(o["rec_loss"]))
        sm.append(float(o["smooth_loss"]))
        for i in range(4):
            assert torch.equal(o["argmin"][i][0], full["argmin"][i][b])
            assert rel_err(full["grad_depth"][i][b] * B, o["grad_depth"][i][0]) < 1e-5
        for j in range(2):
            assert rel_err(full["grad_pose"][j][b] * B, o["grad_pose"][j][0]) < 1e-4
    assert abs(np.mean(rec) - float(full["rec_loss"])) < 2e-6 * float(full["rec_loss"])
    assert abs(np.mean(sm) - float(full["smooth_loss"])) < 2e-6 * float(full["smooth_loss"])


# ------------------------------------------------------------------------------------------- edge cases
def _against_oracle(inp, dev, loss_tol=LOSS_TOL, **kw):
    plan_kw = {k: v for k, v in kw.items()}
    kw = {k: v for k, v in kw.items() if k != "save_warped"}
    port_kw = dict(automask=kw.get("automask", True), reduce=kw.get("reduce", "min"),
                   ssim_w=kw.get("ssim_weight", 0.85), smooth_w=kw.get("smooth_weight", 1e-3))
    out = gpu_mono_from_vec(inp, dev, **plan_kw)
    tgt, src = build_pyramid(inp)
    ref = oracle_mono(inp, torch.float64, tgt, src, **port_kw)
    assert rel_err(out["rec_loss"], ref["rec_loss"].detach()) < loss_tol
    if port_kw["smooth_w"] > 0:
        assert rel_err(out["smooth_loss"], ref["smooth_loss"].detach()) < loss_tol
    return out, ref


def test_source_equal_to_target_gives_zero_loss_and_gradients(dev):
    """(i): identity candidates are exactly 0, the automask wins everywhere, no gradient flows."""
    inp = mono_inputs(2, 32, 64, seed=4)
    inp["ctx"] = [inp["img"].clone(), inp["img"].clone()]
    out = gpu_mono_from_vec(inp, dev, smooth_weight=0.0)
    assert float(out["rec_loss"]) == 0.0
    for a in out["argmin"]:
        assert bool(((a == 1) | (a == 0)).all())  # first exact zero wins; the warp can only tie
    for g in out["grad_depth"] + out["grad_pose_vec"]:
        assert float(g.abs().max()) == 0.0


def test_identical_sources_break_ties_towards_the_lowest_index(dev):
    """(ii): torch.min keeps the first index among equal candidates."""
    inp = mono_inputs(2, 32, 64, seed=5)
    inp["ctx"][1] = inp["ctx"][0].clone()
    inp["pose_vec"][1] = inp["pose_vec"][0].clone()
    out = gpu_mono_from_vec(inp, dev)
    for a in out["argmin"]:
        assert int(a.max()) <= 1


def test_identity_pose(dev):
    """(iii): R = I, t = 0."""
    inp = mono_inputs(2, 32, 64, seed=6)
    inp["pose_vec"] = [torch.zeros_like(v) for v in inp["pose_vec"]]
    _against_oracle(inp, dev)


def test_points_behind_the_camera_and_out_of_view(dev):
    """(iv)/(v): large motion sends samples behind the camera or far outside the image; the
    clamped border sampling and zero coordinate gradients must match the reference."""
    inp = mono_inputs(2, 48, 80, seed=7)
    inp["pose_vec"][0] = torch.tensor([[0.5, 0.1, -60.0, 0.0, 0.3, 0.0], [30.0, -20.0, 1.0, 0.2, 0.0, 0.1]])
    inp["pose_vec"][1] = torch.tensor([[0.0, 0.0, -1.0, 0.0, 3.0, 0.0], [-200.0, 0.0, 0.0, 0.0, 0.0, 1.0]])
    out, ref = _against_oracle(inp, dev)
    for g in out["grad_depth"] + out["grad_pose_vec"]:
        assert bool(torch.isfinite(g).all())
    for i, (g, r) in enumerate(zip(out["grad_depth"], ref["grad_depth"])):
        err = (g.double() - r).abs() / r.abs().max()
        assert float(torch.quantile(err.flatten(), 0.99)) < GRAD_TOL


def test_translation_broadcast_shapes_agree(dev):
    """(vi): the kernel takes t from the 4x4 pose; the reference's two spellings ([B,3,1,1] and the
    expanded [B,3,h,w]) are the same numbers, so one oracle run covers both."""
    inp = mono_inputs(1, 24, 80, seed=8)
    _against_oracle(inp, dev)


def test_depth_at_the_clamps(dev):
    """(vii): depth below the 1e-6 clamp of the smoothness term and near-zero projected depth."""
    inp = mono_inputs(1, 32, 64, seed=9)
    for d in inp["depth"]:
        d[:, :, ::5, ::7] = 5e-7
        d[:, :, 1::6, 2::9] = 2e-6
    out, ref = _against_oracle(inp, dev, loss_tol=5e-5)
    for i, (g, r) in enumerate(zip(out["grad_depth"], ref["grad_depth"])):
        assert bool(torch.isfinite(g).all())
        tiny = inp["depth"][i] < 1e-6
        # below the clamp the smoothness term passes no gradient (clamp(min=1e-6), smoothness_loss.py:62);
        # what is left is the photometric part, which the oracle has too
        err = (g.double() - r).abs() / r.abs().max()
        assert float(torch.quantile(err.flatten(), 0.98)) < GRAD_TOL
        assert float(err[tiny].max()) < 1e-2


def test_shallow_depths_down_to_5cm(dev):
    inp = mono_inputs(2, 32, 64, seed=10, min_depth=0.05)
    _against_oracle(inp, dev)


@pytest.mark.parametrize("shape", [(1, 24, 80), (1, 37, 53), (2, 16, 130), (1, 6, 8)])
def test_odd_sizes_single_scale(dev, shape):
    """(viii): sizes that are not multiples of the tile, one scale only."""
    B, H, W = shape
    inp = mono_inputs(B, H, W, scales=1, seed=12)
    out, ref = _against_oracle(inp, dev)
    g, r = out["grad_depth"][0], ref["grad_depth"][0]
    err = (g.double() - r).abs() / r.abs().max()
    assert float(torch.quantile(err.flatten(), 0.99)) < GRAD_TOL


def test_recompute_and_saved_warp_backward_agree(dev):
    """The backward kernel either reads the warped sources kept by the forward pass (default) or
    recomputes the warp (`save_warped=False`, nothing but argmin bytes and O(B) scalars kept): losses and argmin maps
    must give the same bits; the gradients agree to rounding (the derivative of the warp and the smoothness gradient
    are evaluated by different kernels in the two modes, with different fused-multiply-add contractions and
    exp / reciprocal roundings)."""
    inp = mono_inputs(2, 48, 160, seed=21)
    a = gpu_mono_from_vec(inp, dev, save_warped=True)
    b = gpu_mono_from_vec(inp, dev, save_warped=False)
    assert torch.equal(a["rec_loss"], b["rec_loss"]) and torch.equal(a["smooth_loss"], b["smooth_loss"])
    for x, y in zip(a["argmin"], b["argmin"]):
        assert torch.equal(x, y)
    for x, y in zip(a["grad_depth"] + a["grad_pose_vec"], b["grad_depth"] + b["grad_pose_vec"]):
        assert float((x - y).abs().max()) <= 2e-6 * float(y.abs().max())
    _against_oracle(inp, dev, save_warped=False)


def test_single_source_single_sample(dev):
    """(x): B = 1, S = 1."""
    inp = mono_inputs(1, 32, 64, S=1, seed=13)
    _against_oracle(inp, dev)
    _against_oracle(inp, dev, automask=False)


def test_three_sources(dev):
    inp = mono_inputs(1, 32, 64, S=3, seed=14)
    _against_oracle(inp, dev)


def test_nan_depth_propagates_to_the_loss(dev):
    """The trainer asserts torch.isfinite(losses) (projects/MonoDepth2/train.py:93): a NaN in the
    predictions must surface, not be hidden."""
    inp = mono_inputs(1, 32, 64, seed=15)
    inp["depth"][0][0, 0, 3, 4] = float("nan")
    out = gpu_mono_from_vec(inp, dev)
    assert not bool(torch.isfinite(out["smooth_loss"]))


def test_upstream_gradient_scaling(dev):
    """Backward honours the upstream gradients of the two scalars independently."""
    from simpledepthestimation_b200.functional import MonoLossPlan, mono_photometric_smoothness_loss
    from oracle import port

    inp = mono_inputs(1, 32, 64, seed=16)
    g = lambda t: t.to(dev).contiguous()  # noqa: E731
    sizes = [tuple(d.shape[-2:]) for d in inp["depth"]]
    tgt = [g(port.resize_bilinear(inp["img"], s)) for s in sizes]
    src = [[g(port.resize_bilinear(c, s)) for c in inp["ctx"]] for s in sizes]
    plan = MonoLossPlan(1, sizes, 2, (32, 64), dev)
    grads = []
    for wr, ws in ((1.0, 1.0), (3.0, 0.0), (0.0, 2.0)):
        depth = [g(d).requires_grad_() for d in inp["depth"]]
        pose = [g(euler_pose(v)).requires_grad_() for v in inp["pose_vec"]]
        rec, sm, _ = mono_photometric_smoothness_loss(plan, tgt, src, depth, g(inp["K"]), pose)
        (wr * rec + ws * sm).backward()
        grads.append(([d.grad.clone() for d in depth], [p.grad.clone() for p in pose]))
    for i in range(4):
        combo = grads[1][0][i] / 3.0 + grads[2][0][i] / 2.0
        assert rel_err(combo, grads[0][0][i]) < 1e-5
    for j in range(2):
        assert rel_err(grads[1][1][j] / 3.0, grads[0][1][j]) < 1e-5   # smoothness does not touch the pose
        assert float(grads[2][1][j].abs().max()) == 0.0


def test_shape_errors_raise(dev):
    from simpledepthestimation_b200 import _lib
    from simpledepthestimation_b200.functional import MonoLossPlan

    plan = MonoLossPlan(1, [(16, 32)], 1, (16, 32), dev)
    z = lambda *s: torch.zeros(*s, device=dev)  # noqa: E731
    with pytest.raises(_lib.SdeError):
        plan.forward([z(1, 3, 16, 33)], [[z(1, 3, 16, 32)]], [z(1, 1, 16, 32)], z(1, 3, 3), [z(1, 4, 4)])
    with pytest.raises(_lib.SdeError):
        plan.forward([z(1, 3, 16, 32).cpu()], [[z(1, 3, 16, 32)]], [z(1, 1, 16, 32)], z(1, 3, 3), [z(1, 4, 4)])
    with pytest.raises(NotImplementedError):
        MonoLossPlan(1, [(16, 32)], 1, (16, 32), dev, reduce="median")


def test_tma_and_thread_staged_planes_give_the_same_bits(dev, monkeypatch):
    """SDE_DISABLE_TMA=1 forces the thread-staged planes on a shape that takes the TMA path (row pitch a multiple of
    16 bytes): same warp kernel, same arithmetic, so losses, argmin maps and gradients must be bit-identical; the
    thread-staged path is also checked against the oracle."""
    inp = mono_inputs(2, 48, 160, seed=23)
    monkeypatch.delenv("SDE_DISABLE_TMA", raising=False)
    a = gpu_mono_from_vec(inp, dev)
    monkeypatch.setenv("SDE_DISABLE_TMA", "1")
    b = gpu_mono_from_vec(inp, dev)
    assert torch.equal(a["rec_loss"], b["rec_loss"]) and torch.equal(a["smooth_loss"], b["smooth_loss"])
    for x, y in zip(a["argmin"] + a["grad_depth"] + a["grad_pose_vec"], b["argmin"] + b["grad_depth"] + b["grad_pose_vec"]):
        assert torch.equal(x, y)
    _against_oracle(inp, dev)


def test_chained_and_plain_launches_give_the_same_bits(dev, monkeypatch):
    """Programmatic dependent launch (forward kernel chained behind the warp kernel, backward kernel behind its
    predecessor) only moves launch latency: every output must be bit-identical to plain launches (SDE_DISABLE_PDL=1)
    and to the all-chained configuration (SDE_PDL_MASK=7), over repeated back-to-back steps."""
    inp = mono_inputs(2, 96, 320, seed=29)
    runs = []
    for env in ({"SDE_DISABLE_PDL": "1"}, {}, {"SDE_PDL_MASK": "7"}):
        for k in ("SDE_DISABLE_PDL", "SDE_PDL_MASK"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        for _ in range(3):
            runs.append(gpu_mono_from_vec(inp, dev))
    a = runs[0]
    for b in runs[1:]:
        assert torch.equal(a["rec_loss"], b["rec_loss"]) and torch.equal(a["smooth_loss"], b["smooth_loss"])
        for x, y in zip(a["argmin"] + a["grad_depth"] + a["grad_pose_vec"], b["argmin"] + b["grad_depth"] + b["grad_pose_vec"]):
            assert torch.equal(x, y)


def test_host_runner_end_to_end_matches_device_path(dev):
    """HostLossRunner (pinned host arena -> one H2D copy -> device pyramid -> fused fwd + bwd -> one D2H copy, two
    pipelined slots) returns the bits of the device-resident path, step after step."""
    from simpledepthestimation_b200.functional import HostLossRunner, MonoLossPlan
    from simpledepthestimation_b200.geometry.camera import resize_img

    B, H, W = 2, 48, 160
    sets = [mono_inputs(B, H, W, seed=40 + k) for k in range(3)]
    sizes = [tuple(d.shape[-2:]) for d in sets[0]["depth"]]
    plan = MonoLossPlan(B, sizes, 2, (H, W), dev)
    runner = HostLossRunner(plan, dev)
    host = [runner.pin((s["img"], list(s["ctx"]), list(s["depth"]), s["K"], [euler_pose(v) for v in s["pose_vec"]]))
            for s in sets]
    g = lambda t: t.to(dev).contiguous()  # noqa: E731
    for step in range(5):
        k = step % 3
        runner.step(host[k])
        losses, gd, gp = runner.finish()
        s = sets[k]
        tgt = [resize_img(g(s["img"]), sz) for sz in sizes]
        src = [[resize_img(g(c), sz) for c in s["ctx"]] for sz in sizes]
        depth, K, pose = [g(d) for d in s["depth"]], g(s["K"]), [g(euler_pose(v)) for v in s["pose_vec"]]
        saved = plan.new_warped()
        ref_l, argm = plan.forward(tgt, src, depth, K, pose, warped=saved)
        ref_gd, ref_gp = plan.backward(tgt, src, depth, K, pose, argm, torch.ones(2, device=dev), warped=saved)
        torch.cuda.synchronize()
        assert torch.equal(losses, ref_l.cpu())
        for a, b in zip(gd + gp, ref_gd + ref_gp):
            assert torch.equal(a, b.cpu())


@pytest.mark.parametrize("u8,frames_only", [(True, False), (False, True), (True, True)])
def test_host_runner_options_match_device_path(dev, u8, frames_only):
    """HostLossRunner(u8_frames / frames_only): uint8 frames converted inside the pyramid kernel, and depth / K / poses
    kept device-resident, give the bits of the device-resident path on the same (quantised) frames."""
    from simpledepthestimation_b200.functional import HostLossRunner, MonoLossPlan
    from simpledepthestimation_b200.geometry.camera import resize_img

    B, H, W = 2, 48, 160
    s = mono_inputs(B, H, W, seed=44)
    sizes = [tuple(d.shape[-2:]) for d in s["depth"]]
    plan = MonoLossPlan(B, sizes, 2, (H, W), dev)
    g = lambda t: t.to(dev).contiguous()  # noqa: E731
    depth, K, pose = [g(d) for d in s["depth"]], g(s["K"]), [g(euler_pose(v)) for v in s["pose_vec"]]
    runner = HostLossRunner(plan, dev, u8_frames=u8, frames_only=frames_only)
    if frames_only:
        runner.set_resident(depth, K, pose)
    arena = runner.pin((s["img"], list(s["ctx"]), list(s["depth"]), s["K"], [euler_pose(v) for v in s["pose_vec"]]))
    frame_bytes = 3 * B * 3 * H * W * (1 if u8 else 4)
    rest_bytes = 0 if frames_only else 4 * (sum(B * h * w for h, w in sizes) + B * 9 + 2 * B * 16)
    assert runner.h2d_bytes == frame_bytes + rest_bytes
    for _ in range(3):
        runner.step(arena)
    losses, gd, gp = runner.finish()
    quant = (lambda t: (t * 255.0).round().clamp(0, 255) / 255.0) if u8 else (lambda t: t)
    img, ctx = g(quant(s["img"])), [g(quant(c)) for c in s["ctx"]]
    tgt = [resize_img(img, sz) for sz in sizes]
    src = [[resize_img(c, sz) for c in ctx] for sz in sizes]
    saved = plan.new_warped()
    ref_l, argm = plan.forward(tgt, src, depth, K, pose, warped=saved)
    ref_gd, ref_gp = plan.backward(tgt, src, depth, K, pose, argm, torch.ones(2, device=dev), warped=saved)
    torch.cuda.synchronize()
    assert torch.equal(losses, ref_l.cpu())
    for a, b in zip(gd + gp, ref_gd + ref_gp):
        assert torch.equal(a, b.cpu())


@pytest.mark.parametrize("mode", ["disp", "logit"])
@pytest.mark.parametrize("B,H,W,save", [(2, 32, 64, True), (1, 50, 70, True), (2, 32, 64, False)])
def test_depth_decoder_tail_folded_into_the_loss(dev, mode, B, H, W, save):
    """the loss kernels take the decoder's disparity (or its pre-softplus logits) and apply
    disp_to_depth (layers/depth_decoder.py:9-18; softplus :108; DepthResNet.py:41,57) themselves; the gradient comes back
    w.r.t. that tensor.  Against (a) the oracle in fp64 and (b) the same kernels fed the depth torch decodes, with
    autograd through the decode."""
    from parity_log import record
    from simpledepthestimation_b200.functional import MonoLossPlan, mono_photometric_smoothness_loss

    inp = mono_inputs(B, H, W, seed=21)
    min_d, max_d = 0.1, 80.0
    gen = torch.Generator().manual_seed(6)
    # what the network hands over: disparities in (0, 1) or logits; the same smooth fields as the depth generator
    raw = []
    for d in inp["depth"]:
        disp = ((1.0 / d.double() - 1.0 / max_d) / (1.0 / min_d - 1.0 / max_d)).clamp(1e-4, 0.999)
        raw.append((disp if mode == "disp" else torch.log(torch.expm1(disp))).float())
    tgt, src = build_pyramid(inp)
    g = lambda t: t.to(dev).contiguous()  # noqa: E731
    sizes = [tuple(d.shape[-2:]) for d in inp["depth"]]
    pose = [g(euler_pose(v)) for v in inp["pose_vec"]]
    T, S_, K = [g(t) for t in tgt], [[g(x) for x in row] for row in src], g(inp["K"])

    def decode(x):   # depth_decoder.py:9-18 (+ nn.Softplus, :93,108)
        disp = torch.nn.functional.softplus(x) if mode == "logit" else x
        return 1 / (1 / max_d + (1 / min_d - 1 / max_d) * disp)

    # (b) torch decodes, the kernels see depth
    x1 = [g(r).requires_grad_() for r in raw]
    p1 = [p.clone().requires_grad_() for p in pose]
    plan1 = MonoLossPlan(B, sizes, 2, (H, W), dev, save_warped=save)
    rec1, sm1, arg1 = mono_photometric_smoothness_loss(plan1, T, S_, [decode(x) for x in x1], K, p1)
    (rec1 + sm1).backward()
    # folded: the kernels see the raw tensor
    x2 = [g(r).requires_grad_() for r in raw]
    p2 = [p.clone().requires_grad_() for p in pose]
    plan2 = MonoLossPlan(B, sizes, 2, (H, W), dev, save_warped=save, depth_mode=mode, min_depth=min_d, max_depth=max_d)
    rec2, sm2, arg2 = mono_photometric_smoothness_loss(plan2, T, S_, x2, K, p2)
    (rec2 + sm2).backward()
    torch.cuda.synchronize()
    e = dict(rec=rel_err(rec2.detach(), rec1.detach()), smooth=rel_err(sm2.detach(), sm1.detach()),
             grad_raw=max(rel_err(a.grad, b.grad) for a, b in zip(x2, x1)),
             grad_pose=max(rel_err(a.grad, b.grad) for a, b in zip(p2, p1)))
    for a, b in zip(arg2, arg1):
        assert float((a != b).double().mean()) < 1e-3
    assert e["rec"] < 2e-6 and e["smooth"] < 2e-6 and e["grad_raw"] < 2e-5 and e["grad_pose"] < 2e-5, e
    # (a) the oracle in fp64 on the same raw tensors
    xd = [r.double().requires_grad_() for r in raw]
    pd_ = [euler_pose(v.float()).double().requires_grad_() for v in inp["pose_vec"]]
    pyr = [(t.double(), [s_.double() for s_ in ss]) for t, ss in zip(tgt, src)]
    out = port.mono_loss(inp["img"].double(), None, inp["K"].double(), [decode(x) for x in xd], pd_, pyramid=pyr)
    (out["rec_loss"] + out["smooth_loss"]).backward()
    e["rec_vs_oracle"] = rel_err(rec2.detach(), out["rec_loss"].detach())
    e["smooth_vs_oracle"] = rel_err(sm2.detach(), out["smooth_loss"].detach())
    e["grad_raw_vs_oracle_q999"] = max(
        float(torch.quantile(((a.grad.cpu().double() - b.grad).abs() / b.grad.abs().max()).flatten(), 0.999))
        for a, b in zip(x2, xd))
    record(mode=mode, shape=[B, H, W], save_warped=save, **e)
    assert e["rec_vs_oracle"] < LOSS_TOL and e["smooth_vs_oracle"] < LOSS_TOL and e["grad_raw_vs_oracle_q999"] < GRAD_TOL, e


def test_sub_batch_streams_reproduce_the_single_launch(dev):
    """MonoLossPlan(streams=K): the batch as K sub-batches on K streams (sde_mono_desc.norm_batch = the whole batch)
    gives the single-launch losses up to the summation order and the same argmin maps and gradients bit for bit
    (every gradient element belongs to one sample), for K = 2 and 3; odd batches refuse to split."""
    from simpledepthestimation_b200 import _lib
    from simpledepthestimation_b200.functional import MonoLossPlan

    inp = mono_inputs(6, 48, 160, seed=31)
    ref = gpu_mono_from_vec(inp, dev, streams=1)
    for k in (2, 3):
        out = gpu_mono_from_vec(inp, dev, streams=k)
        assert rel_err(out["rec_loss"], ref["rec_loss"]) < 1e-6 and rel_err(out["smooth_loss"], ref["smooth_loss"]) < 1e-6
        for a, b in zip(out["argmin"] + out["grad_depth"], ref["argmin"] + ref["grad_depth"]):
            assert torch.equal(a, b)
        for a, b in zip(out["grad_pose_vec"], ref["grad_pose_vec"]):
            assert rel_err(a, b) < 1e-6
    with pytest.raises(_lib.SdeError):
        MonoLossPlan(6, [(48, 160)], 2, (48, 160), dev, streams=4)
    auto = MonoLossPlan(12, [(48, 160)], 2, (48, 160), dev)
    assert auto.parts == 2 and auto.split_plan().parts == 1 and MonoLossPlan(3, [(48, 160)], 2, (48, 160), dev).parts == 1
    explicit = MonoLossPlan(12, [(48, 160)], 2, (48, 160), dev, streams=2)
    assert explicit.parts == 2 and explicit.split_plan() is explicit


def test_forward_backward_in_one_call_matches_the_two_calls(dev):
    """MonoLossPlan.forward_backward (losses and gradients in one call, sub-batch streams not rejoined between the
    passes) returns the bits of forward() followed by backward()."""
    from simpledepthestimation_b200.functional import MonoLossPlan

    B, H, W = 4, 48, 160
    inp = mono_inputs(B, H, W, seed=33)
    tgt, src = build_pyramid(inp)
    g = lambda t: t.to(dev).contiguous()  # noqa: E731
    sizes = [tuple(d.shape[-2:]) for d in inp["depth"]]
    args = ([g(t) for t in tgt], [[g(x) for x in row] for row in src], [g(d) for d in inp["depth"]], g(inp["K"]),
            [g(euler_pose(v)) for v in inp["pose_vec"]])
    gl = torch.tensor([0.7, 1.3], device=dev)
    for streams in (1, 2):
        plan = MonoLossPlan(B, sizes, 2, (H, W), dev, streams=streams)
        saved = plan.new_warped()
        l1, a1 = plan.forward(*args, warped=saved)
        gd1, gp1 = plan.backward(*args, a1, gl, warped=saved)
        l2, a2, gd2, gp2 = plan.forward_backward(*args, gl)
        torch.cuda.synchronize()
        assert torch.equal(l1, l2)
        for x, y in zip(a1 + gd1 + gp1, a2 + gd2 + gp2):
            assert torch.equal(x, y)


@pytest.mark.parametrize("B,H,W", [(6, 96, 320), (3, 50, 70)])
def test_tile_level_dependencies_give_the_same_bits(dev, monkeypatch, B, H, W):
    """sde_mono_loss_step chains warp -> forward -> backward with tile-level dependencies (SDE_FLOW_MASK bit 0: a
    forward tile waits for the chunk flags of its rows, bit 1: a backward tile for its image's flag) instead of
    whole-grid ones.  Every output must carry the bits of the grid-level chain (mask 0) over repeated back-to-back
    steps on alternating inputs -- one plan, one set of kept planes, so a flag left set by one step would let the next
    one read the previous step's planes -- with and without sub-batch streams, on a TMA shape and on an odd width
    (thread-staged planes); the flag words (the tail of the workspace) must be clear again after every step."""
    from simpledepthestimation_b200.functional import MonoLossPlan

    g = lambda t: t.to(dev).contiguous()  # noqa: E731
    sets = []
    for seed in (41, 42):
        inp = mono_inputs(B, H, W, seed=seed)
        tgt, src = build_pyramid(inp)
        sets.append(([g(t) for t in tgt], [[g(x) for x in row] for row in src], [g(d) for d in inp["depth"]], g(inp["K"]),
                     [g(euler_pose(v)) for v in inp["pose_vec"]]))
    sizes = [tuple(d.shape[-2:]) for d in inp["depth"]]
    gl = torch.tensor([0.9, 1.1], device=dev)
    ref = None
    for mask in ("0", "1", "2", "3"):
        monkeypatch.setenv("SDE_FLOW_MASK", mask)
        for streams in (1, 3):
            plan = MonoLossPlan(B, sizes, 2, (H, W), dev, streams=streams)
            saved = plan.new_warped()
            sub = B // streams
            outs = []
            for it in range(6):
                l, a, gd, gp = plan.forward_backward(*sets[it % 2], gl, warped=saved)
                outs.append([t.clone() for t in [l] + a + gd + gp])
            torch.cuda.synchronize()
            # the flag words are the last two regions of a sub-batch's workspace (sde_api.cu: mono_layout): one word per
            # block of the warp kernel (2048 pixels), one per (scale, image), each region padded to 16 bytes
            a16 = lambda n: (n + 15) // 16 * 16  # noqa: E731
            flag_bytes = a16(4 * sum(sub * (-(-h * w // 2048)) for h, w in sizes)) + a16(4 * len(sizes) * sub)
            for q in range(streams):
                end = q * plan._ws_stride + plan._ws_bytes
                assert int(plan.workspace[end - flag_bytes:end].count_nonzero()) == 0, f"mask {mask} streams {streams}: flags left set"
            for it in range(2, 6):
                for x, y in zip(outs[it], outs[it - 2]):
                    assert torch.equal(x, y), f"mask {mask} streams {streams}: step {it} differs from step {it - 2}"
            if ref is None:
                ref = outs[:2]
            else:
                for k in range(2):
                    assert torch.allclose(outs[k][0], ref[k][0], rtol=1e-6, atol=0)   # losses: summation order of the sub-batches
                    for x, y in zip(outs[k][1:], ref[k][1:]):
                        assert torch.equal(x, y), f"mask {mask} streams {streams}"


def test_persistent_work_queues_give_the_same_bits(dev, monkeypatch):
    """SDE_PERSIST selects which kernels run one CTA per resident slot over a work queue in the workspace (bit 0 warp,
    1 forward, 2 backward; default 4) instead of one CTA per chunk / tile.  Same tiles, same arithmetic: every output
    must carry the bits of the plain launches, over repeated steps on alternating inputs (a queue that was not cleared
    would leave the next step without work), with and without tile-level dependencies, for both calling forms."""
    from simpledepthestimation_b200.functional import MonoLossPlan

    B, H, W = 4, 96, 320
    g = lambda t: t.to(dev).contiguous()  # noqa: E731
    sets = []
    for seed in (51, 52):
        inp = mono_inputs(B, H, W, seed=seed)
        tgt, src = build_pyramid(inp)
        sets.append(([g(t) for t in tgt], [[g(x) for x in row] for row in src], [g(d) for d in inp["depth"]], g(inp["K"]),
                     [g(euler_pose(v)) for v in inp["pose_vec"]]))
    sizes = [tuple(d.shape[-2:]) for d in inp["depth"]]
    gl = torch.tensor([1.0, 1.0], device=dev)
    ref = None
    for flow in ("3", "0"):
        monkeypatch.setenv("SDE_FLOW_MASK", flow)
        for persist in ("0", "4", "7"):
            monkeypatch.setenv("SDE_PERSIST", persist)
            plan = MonoLossPlan(B, sizes, 2, (H, W), dev, streams=1)
            saved = plan.new_warped()
            outs = []
            for it in range(4):
                if it < 2:
                    l, a, gd, gp = plan.forward_backward(*sets[it % 2], gl, warped=saved)
                else:
                    l, a = plan.forward(*sets[it % 2], warped=saved)
                    gd, gp = plan.backward(*sets[it % 2], a, gl, warped=saved)
                outs.append([t.clone() for t in [l] + a + gd + gp])
            torch.cuda.synchronize()
            if ref is None:
                ref = outs
            for it in range(4):
                for x, y in zip(outs[it], ref[it % 2]):
                    assert torch.equal(x, y), f"flow {flow} persist {persist} step {it}"


@pytest.mark.parametrize("shape", [(2, 48, 160), (1, 37, 53)])
def test_four_sources_two_pairs_per_tile(dev, monkeypatch, shape):
    """Four sources: the two-sources-per-pass backward kernel reloads the warped planes of the second pair in the
    middle of a tile (TMA shape) or stages them with plain loads (odd width).  Losses against the fp64 oracle,
    gradients against the oracle's and against the one-source-per-pass kernel (SDE_BWD_PAIR=0: same math, other order)."""
    B, H, W = shape
    inp = mono_inputs(B, H, W, S=4, seed=61)
    monkeypatch.delenv("SDE_BWD_PAIR", raising=False)
    out, ref = _against_oracle(inp, dev)
    monkeypatch.setenv("SDE_BWD_PAIR", "0")
    single = gpu_mono_from_vec(inp, dev)
    for a, b in zip(out["argmin"], single["argmin"]):
        assert torch.equal(a, b)
    # measured 1e-6 of the largest element (the pair kernel forms the second source's box sums as a difference); a pose
    # gradient is a cancelling sum over all pixels: 9e-6 on the 37 x 53 image
    for a, b in zip(out["grad_depth"], single["grad_depth"]):
        assert float((a - b).abs().max()) <= 4e-6 * float(b.abs().max()), float((a - b).abs().max() / b.abs().max())
    for a, b in zip(out["grad_pose_vec"], single["grad_pose_vec"]):
        assert float((a - b).abs().max()) <= 4e-5 * float(b.abs().max()), float((a - b).abs().max() / b.abs().max())
    tgt, src = build_pyramid(inp)
    masks = [stable_mask(inp, tgt, src, i) for i in range(len(inp["depth"]))]
    for i, (g, r) in enumerate(zip(out["grad_depth"], ref["grad_depth"])):
        r = torch.as_tensor(r, dtype=torch.float64)
        err = ((g.double() - r).abs() / r.abs().max())[:, 0][masks[i]]
        assert float(err.max()) < GRAD_TOL, f"grad_depth[{i}] {float(err.max()):.2e}"
