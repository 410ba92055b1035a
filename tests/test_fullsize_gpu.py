"""Oracle parity at the FULL size of BASELINE.json's GPU configurations (round-1 review: only configs[0] met the
oracle at its own size).  The CUDA path runs the whole batch in one call through the C ABI; the oracle
(oracle/port.py = the reference's ATen op sequence, fp64 and fp32) is evaluated sample by sample to bound host
memory -- exact, because the reference's losses are batch means of per-sample terms (MonoDepth2.py:119 `.mean()`
over [B,1,h,w]; smoothness_loss.py:62-80 normalises per image), so a sample's share of the batch loss is 1/B of
its stand-alone loss and likewise for its gradients.

Protocol: loss 1e-5 relative; argmin exact except where the two best candidates are within 1e-5
(or within 3x the largest gap at which the reference's own fp32 run deviates from its fp64 run -- measured 1.4e-5 at
cfg2); depth gradients 1e-4 of the largest gradient on decision-stable pixels, held to the reference's own fp32
behaviour where that exceeds the tolerance (worst error <= 3x the reference's, no more outliers than 3x the reference's
+ 5), 99.9 % of ALL pixels within 1e-4; pose gradients 1e-4 or 5x the reference's own fp32 deviation.  Every achieved
error is written next to the reference's own fp32-vs-fp64 figure to profiles/parity_r2.json."""
import gc

import pytest
import torch

from helpers import build_pyramid, gpu_mono_from_vec, port_mono_from_vec, rel_err, stable_mask
from parity_log import record
from simpledepthestimation_b200.synthetic import mono_inputs, motion_inputs

pytestmark = pytest.mark.gpu
LOSS_TOL, GRAD_TOL, TIE_TOL = 1e-5, 1e-4, 1e-5


def _slice_inputs(inp, b):
    sl = slice(b, b + 1)
    return dict(img=inp["img"][sl], ctx=[c[sl] for c in inp["ctx"]], K=inp["K"][sl], depth=[d[sl] for d in inp["depth"]],
                pose_vec=[v[sl] for v in inp["pose_vec"]])


def _oracle_per_sample(inp, dtype):
    """Batch result of the oracle assembled from per-sample runs (see module docstring)."""
    B = inp["img"].shape[0]
    n, S = len(inp["depth"]), len(inp["pose_vec"])
    rec = smooth = 0.0
    gd = [[] for _ in range(n)]
    gv = [[] for _ in range(S)]
    argmin = [[] for _ in range(n)]
    gap = [[] for _ in range(n)]
    for b in range(B):
        o = port_mono_from_vec(_slice_inputs(inp, b), dtype)
        rec += float(o["rec_loss"].detach().double()) / B
        smooth += float(o["smooth_loss"].detach().double()) / B
        for i in range(n):
            gd[i].append(o["grad_depth"][i].double() / B)
            argmin[i].append(o["argmin"][i])
            top2 = o["cand"][i].detach().topk(2, dim=1, largest=False)[0]
            gap[i].append((top2[:, 1] - top2[:, 0]).double())
        for j in range(S):
            gv[j].append(o["grad_pose_vec"][j].double() / B)
        del o
        gc.collect()
    return dict(rec_loss=rec, smooth_loss=smooth, grad_depth=[torch.cat(x) for x in gd],
                grad_pose_vec=[torch.cat(x) for x in gv], argmin=[torch.cat(x) for x in argmin],
                gap=[torch.cat(x) for x in gap])


def _check_mono(inp, dev, tag):
    out = gpu_mono_from_vec(inp, dev)
    r64 = _oracle_per_sample(inp, torch.float64)
    r32 = _oracle_per_sample(inp, torch.float32)
    tgt, src = build_pyramid(inp)
    m = dict(config=tag, loss_tol=LOSS_TOL, grad_tol=GRAD_TOL)
    m["rec_loss_rel"] = abs(float(out["rec_loss"]) - r64["rec_loss"]) / abs(r64["rec_loss"])
    m["smooth_loss_rel"] = abs(float(out["smooth_loss"]) - r64["smooth_loss"]) / abs(r64["smooth_loss"])
    m["rec_loss_rel_reference_fp32"] = abs(r32["rec_loss"] - r64["rec_loss"]) / abs(r64["rec_loss"])
    fails = []
    if m["rec_loss_rel"] >= LOSS_TOL or m["smooth_loss_rel"] >= LOSS_TOL:
        fails.append("loss")
    for i, a in enumerate(out["argmin"]):
        a64 = r64["argmin"][i].reshape(a.shape)
        gap = r64["gap"][i].reshape(a.shape)
        mism = a.long() != a64
        mism32 = r32["argmin"][i].reshape(a.shape) != a64          # the reference's own fp32 run against its fp64 run
        worst = float(gap[mism].max()) if bool(mism.any()) else 0.0
        worst32 = float(gap[mism32].max()) if bool(mism32.any()) else 0.0
        m[f"argmin_s{i}"] = dict(mismatch_frac=float(mism.double().mean()), mismatches=int(mism.sum()),
                                 off_ties=int((mism & (gap > TIE_TOL)).sum()), largest_gap_at_mismatch=worst,
                                 reference_fp32_mismatches=int(mism32.sum()),
                                 reference_fp32_off_ties=int((mism32 & (gap > TIE_TOL)).sum()),
                                 reference_fp32_largest_gap_at_mismatch=worst32)
        # exact except at ties: a mismatch needs a candidate gap within the tie tolerance -- or, at two million pixels,
        # within 3x the largest gap at which the reference's own fp32 run picks another candidate than its fp64 run
        if worst > max(TIE_TOL, 3.0 * worst32) or m[f"argmin_s{i}"]["mismatch_frac"] >= 1e-4:
            fails.append(f"argmin{i}")
    for i, (gd, r) in enumerate(zip(out["grad_depth"], r64["grad_depth"])):
        mask = stable_mask(inp, tgt, src, i)
        scale = r.abs().max()
        err64 = ((gd.double() - r).abs() / scale)[:, 0]
        err32 = ((gd.double() - r32["grad_depth"][i]).abs() / scale)[:, 0]
        dev32 = ((r32["grad_depth"][i] - r).abs() / scale)[:, 0]
        ok = (err64 < GRAD_TOL) | (err32 < GRAD_TOL) | (err64 < 3.0 * dev32)
        g = dict(
            stable_frac=float(mask.double().mean()), max_err_stable=float(err64[mask].max()),
            max_err_all=float(err64.max()), q999_all=float(torch.quantile(err64.flatten()[:4_000_000], 0.999)),
            stable_over_tol=int((err64[mask] > GRAD_TOL).sum()), stable_unexplained=int((~ok[mask]).sum()),
            reference_fp32_max_err_stable=float(dev32[mask].max()), reference_fp32_max_err_all=float(dev32.max()),
            reference_fp32_stable_over_tol=int((dev32[mask] > GRAD_TOL).sum()))
        m[f"grad_depth_s{i}"] = g
        # 1e-4 of the largest gradient on decision-stable pixels; where the reference's OWN fp32 run exceeds that against
        # its fp64 run (ill-conditioned SSIM windows, fp32 pixel coordinates: SURVEY.md App. C) the kernel is held to the
        # reference's fp32 behaviour instead: worst error within 3x the reference's worst, no more pixels over the
        # tolerance than 3x the reference's count (+5), and 99.9 % of ALL pixels within the plain tolerance
        if g["stable_frac"] <= 0.95 or g["q999_all"] >= GRAD_TOL or \
                g["max_err_stable"] > max(GRAD_TOL, 3.0 * g["reference_fp32_max_err_stable"]) or \
                g["stable_over_tol"] > 5 + 3 * g["reference_fp32_stable_over_tol"]:
            fails.append(f"grad_depth{i}")
    for j, (gv, r) in enumerate(zip(out["grad_pose_vec"], r64["grad_pose_vec"])):
        own, ref_dev = rel_err(gv, r), rel_err(r32["grad_pose_vec"][j], r)
        m[f"grad_pose_vec{j}_rel"], m[f"grad_pose_vec{j}_rel_reference_fp32"] = own, ref_dev
        if own >= max(GRAD_TOL, 5 * ref_dev):
            fails.append(f"grad_pose{j}")
    record(**m)
    assert not fails, (fails, m)


@pytest.fixture(scope="module")
def dev(sde_lib):
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda", 0)


def test_cfg2_full_batch_against_oracle(dev):
    """BASELINE.json configs[1]: 640x192, batch 12, 4 scales, 2 sources, automask + smoothness."""
    _check_mono(mono_inputs(12, 192, 640, seed=0), dev, "cfg2 12x192x640")


def test_cfg3_full_batch_against_oracle(dev):
    """BASELINE.json configs[2]: 1024x320, batch 8 (the widest MonoDepth2 shape: fp32 pixel coordinates are coarsest here)."""
    _check_mono(mono_inputs(8, 320, 1024, seed=3), dev, "cfg3 8x320x1024")


def test_cfg5_rank_slices_reassemble_the_global_batch(dev):
    """BASELINE.json configs[4]: the global batch of 96 is sharded into contiguous slices; two 12-sample slices of one
    24-sample batch, run separately, must reproduce the 24-sample result (loss = mean of the slice losses, gradients
    = slice gradients / 2) -- the arithmetic the 8-rank run relies on."""
    inp = mono_inputs(24, 192, 640, seed=5)
    full = gpu_mono_from_vec(inp, dev)
    parts = []
    for r in range(2):
        sl = slice(12 * r, 12 * r + 12)
        parts.append(gpu_mono_from_vec(dict(img=inp["img"][sl], ctx=[c[sl] for c in inp["ctx"]], K=inp["K"][sl],
                                            depth=[d[sl] for d in inp["depth"]], pose_vec=[v[sl] for v in inp["pose_vec"]]), dev))
    e_loss = rel_err(full["rec_loss"], (parts[0]["rec_loss"] + parts[1]["rec_loss"]) / 2)
    e_gd = max(rel_err(full["grad_depth"][i], torch.cat([p["grad_depth"][i] for p in parts]) / 2) for i in range(4))
    e_gv = max(rel_err(full["grad_pose_vec"][j], torch.cat([p["grad_pose_vec"][j] for p in parts]) / 2) for j in range(2))
    record(config="cfg5 slices 2x12 of 24", loss_rel=e_loss, grad_depth_rel=e_gd, grad_pose_rel=e_gv)
    assert e_loss < 2e-6 and e_gd < 1e-6 and e_gv < 1e-5
    for i in range(4):
        assert torch.equal(full["argmin"][i], torch.cat([p["argmin"][i] for p in parts]))


def test_cfg4_full_size_against_oracle(sde_lib):
    """BASELINE.json configs[3]: MotionLearning 1920x1280 with the residual translation field, both directions
    (batch 1: the per-sample statistic depth_err_2nd_mom and every mean are per sample / linear in the batch)."""
    from test_motion_gpu import check_grad, gpu_motion, oracle_motion

    inp = motion_inputs(1, 1280, 1920, seed=5)
    rg = gpu_motion(inp, True)
    rg.pop("raw")
    torch.cuda.empty_cache()
    r64 = oracle_motion(inp, torch.float64, True)
    r64m = [{k: (v.detach() if torch.is_tensor(v) else v) for k, v in m.items()} for m in r64.pop("maps")]
    gc.collect()
    r32 = oracle_motion(inp, torch.float32, True)
    r32.pop("maps")
    gc.collect()
    m = dict(config="cfg4 1x1280x1920 field, both directions", loss_tol=1e-5, grad_tol=1e-4)
    names = ["rgb_l1", "ssim", "smooth"]
    fails = []
    for d in range(2):
        for k in range(3):
            ref = float(r64["losses"][d, k])
            own = abs(float(rg["losses"][d, k]) - ref) / abs(ref)
            dev32 = abs(float(r32["losses"][d, k]) - ref) / abs(ref)
            m[f"loss_{names[k]}_dir{d}_rel"], m[f"loss_{names[k]}_dir{d}_rel_reference_fp32"] = own, dev32
            if own > max(1e-5, 3.0 * dev32):
                fails.append(f"loss {names[k]} dir {d}")
    for k in ("gd1", "gd2", "gpose", "gmo"):
        scale = float(r64[k].abs().max())
        err = (rg[k].double() - r64[k]).abs() / scale
        m[f"{k}_max"] = float(err.max())
        m[f"{k}_q99"] = float(torch.quantile(err.flatten()[:1_000_000], 0.99)) if err.numel() >= 1000 else None
        m[f"{k}_max_reference_fp32"] = float((r32[k].double() - r64[k]).abs().max() / scale)
    for d in range(2):
        mism = rg["maps"][d]["occlusion_mask"].double() != r64m[d]["occlusion_mask"]
        m[f"occlusion_mismatch_dir{d}"] = int(mism.sum())
        m[f"coords_max_abs_dir{d}"] = float((rg["maps"][d]["coords_A_in_B"].double() - r64m[d]["coords_A_in_B"]).abs().max())
        if int(mism.sum()) > max(2, int(1e-4 * mism.numel())) or m[f"coords_max_abs_dir{d}"] >= 1e-4:
            fails.append(f"maps dir {d}")
    record(**m)
    assert not fails, (fails, m)
    for k in ("gd1", "gd2", "gpose", "gmo"):
        check_grad(k, rg[k], r64[k], r32[k])
