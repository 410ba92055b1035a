"""Writes tests/golden/*.npz from the REAL reference (build container only).

TEST INFRASTRUCTURE ONLY.  Usage: python -m oracle.make_golden
Every fixture holds the inputs (fp32) and the outputs of the unmodified reference code
(imported read-only from /root/reference by oracle/ref_import.py) in fp64 and fp32.  For the
large KITTI-shaped case only the seed, a checksum of the generated inputs and the scalar
outputs are stored.
"""
import hashlib
import os

import numpy as np
import torch

from oracle import ref_import
from simpledepthestimation_b200.synthetic import mono_inputs, motion_inputs

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def checksum(tensors):
    h = hashlib.sha256()
    for t in tensors:
        h.update(np.ascontiguousarray(t.numpy()).tobytes())
    return h.hexdigest()


def flat_inputs(inp):
    return [inp["img"], *inp["ctx"], inp["K"], *inp["depth"], *inp["pose_vec"]]


def ref_argmin(inp, dtype, automask=True):
    """argmin maps from the reference's own rgb_consistency_loss / resize_img / scale_intrinsics
    (the loop of MonoDepth2.py:78-119 re-run to expose the per-pixel candidate stack)."""
    ns = ref_import.load()
    model = ns.MonoDepth2.MonoDepth2Model(ref_import.mono_cfg(AUTOMASK=automask)).train().to(dtype)
    img = inp["img"].to(dtype)
    ctx = [c.to(dtype) for c in inp["ctx"]]
    K = inp["K"].to(dtype)
    poses = [ns.pose_utils.pose_vec2mat(v.to(dtype)) for v in inp["pose_vec"]]
    maps, gaps = [], []
    for d in inp["depth"]:
        d = d.to(dtype)
        A = ns.camera.resize_img(img, d.shape[-2:])
        Ki = ns.camera.scale_intrinsics(K.clone(), d.shape[-1] / img.shape[-1], d.shape[-2] / img.shape[-2])
        cand = []
        for c, T in zip(ctx, poses):
            Bi = ns.camera.resize_img(c, d.shape[-2:])
            cand.append(model.rgb_consistency_loss(A, Bi, d, Ki, T[:, :3, :3], T[:, :3, [3], None]))
            if automask:
                cand.append(model.rgb_consistency_loss(A, Bi, d, Ki, None, None))
        stack = torch.cat(cand, 1)
        maps.append(stack.min(1)[1].to(torch.uint8).numpy())
        top2 = stack.topk(2, dim=1, largest=False)[0]
        gaps.append((top2[:, 1] - top2[:, 0]).numpy())
    return maps, gaps


def mono_case(name, B, H, W, store_inputs=True, seed=0, pose_scale=1.0, variants=(), **gen):
    inp = mono_inputs(B, H, W, seed=seed, pose_scale=pose_scale, **gen)
    data = {"meta": np.array([B, H, W, seed]), "pose_scale": np.array(pose_scale),
            "input_sha256": np.array(checksum(flat_inputs(inp)))}
    if store_inputs:
        data["img"] = inp["img"].numpy()
        for j, c in enumerate(inp["ctx"]):
            data[f"ctx{j}"] = c.numpy()
        data["K"] = inp["K"].numpy()
        for i, d in enumerate(inp["depth"]):
            data[f"depth{i}"] = d.numpy()
        for j, v in enumerate(inp["pose_vec"]):
            data[f"pose_vec{j}"] = v.numpy()
    for tag, over in (("", {}),) + tuple(variants):
        for dt, dn in ((torch.float64, "f64"), (torch.float32, "f32")):
            r = ref_import.run_mono(inp, dt, **over)
            data[f"rec_loss{tag}_{dn}"] = r["rec_loss"].numpy()
            data[f"smooth_loss{tag}_{dn}"] = r["smooth_loss"].numpy()
            if dn == "f64" or store_inputs:
                for j, g in enumerate(r["grad_pose_vec"]):
                    data[f"grad_pose_vec{j}{tag}_{dn}"] = g.numpy()
            if store_inputs:
                for i, g in enumerate(r["grad_depth"]):
                    data[f"grad_depth{i}{tag}_{dn}"] = g.numpy()
    if store_inputs:
        maps, gaps = ref_argmin(inp, torch.float64)
        for i, (m, g) in enumerate(zip(maps, gaps)):
            data[f"argmin{i}"] = m
            data[f"argmin_gap{i}"] = g.astype(np.float32)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **data)
    print(name, {k: float(v) for k, v in data.items() if k.startswith(("rec_loss", "smooth_loss"))})


def motion_case(name, B, H, W, seed=0):
    inp = motion_inputs(B, H, W, seed=seed)
    data = {"meta": np.array([B, H, W, seed])}
    for k, v in inp.items():
        data[k] = v.numpy()
    for wm, tag in ((True, ""), (False, "_rigid")):
        for dt, dn in ((torch.float64, "f64"), (torch.float32, "f32")):
            r = ref_import.run_motion(inp, dt, with_motion=wm)
            for k, v in r.items():
                if torch.is_tensor(v):
                    data[f"{k}{tag}_{dn}"] = v.numpy()
            if dn == "f64":
                w12, w21 = r["depth_proximity_weight"][0]
                data[f"weight12{tag}"] = w12.numpy().astype(np.float32)
                data[f"weight21{tag}"] = w21.numpy().astype(np.float32)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **data)
    print(name, {k: float(v) for k, v in data.items() if "loss" in k and k.endswith("f64")})


MOTION_VARIANTS = {
    # tag: (LOSS overrides, MODEL overrides, needs masks)      -- MotionLearning.py:108-117,126-161
    "scales2": (dict(NUM_SCALES=2), {}, False),
    "scalenorm": (dict(SCALE_NORMALIZE=True), {}, False),
    "scales2_scalenorm": (dict(NUM_SCALES=2, SCALE_NORMALIZE=True), {}, False),
    "mask": ({}, dict(WITH_MASK=True, MASK_DILATION=2), True),
}


def motion_masks(B, H, W):
    """Instance-mask stand-ins (batch['mask'], batch['ctx_mask'][0]): a few rectangles, int64 ids."""
    gen = torch.Generator().manual_seed(17)
    masks = []
    for _ in range(2):
        m = torch.zeros(B, 1, H, W, dtype=torch.int64)
        for b in range(B):
            for k in range(2):
                y0, x0 = int(torch.randint(0, H - 8, (1,), generator=gen)), int(torch.randint(0, W - 12, (1,), generator=gen))
                m[b, 0, y0:y0 + 8, x0:x0 + 12] = k + 1
        masks.append(m)
    return masks


def motion_variants_case(name, base="motion_2x32x64"):
    """MotionLearningModel variants the shipped configs leave off -- NUM_SCALES = 2 (resize_img_avgpool,
    camera.py:49-54), SCALE_NORMALIZE, WITH_MASK -- run by the real reference on the inputs of `base`."""
    g = np.load(os.path.join(OUT, base + ".npz"))
    inp = {k: torch.from_numpy(g[k]) for k in ("img1", "img2", "depth1", "depth2", "K", "pose_vec", "motion")}
    B, _, H, W = inp["img1"].shape
    m1, m2 = motion_masks(B, H, W)
    data = {"mask1": m1.numpy().astype(np.uint8), "mask2": m2.numpy().astype(np.uint8)}
    for tag, (loss_over, model_over, need_mask) in MOTION_VARIANTS.items():
        extra = {"mask": m1, "ctx_mask": [m2]} if need_mask else None
        for dt, dn in ((torch.float64, "f64"), (torch.float32, "f32")):
            r = ref_import.run_motion(inp, dt, model_over=model_over, extra_batch=extra, **loss_over)
            for k, v in r.items():
                if torch.is_tensor(v):
                    data[f"{tag}.{k}_{dn}"] = v.numpy()
            if dn == "f64":
                for i, (a, b) in enumerate(r["overall_motion"]):
                    data[f"{tag}.overall_motion{i}_12"] = a.numpy().astype(np.float32)
                    data[f"{tag}.overall_motion{i}_21"] = b.numpy().astype(np.float32)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **data)
    print(name, {k: float(v) for k, v in data.items() if "loss" in k and k.endswith("f64")})


def main():
    import sys
    os.makedirs(OUT, exist_ok=True)
    only = set(sys.argv[1:])
    want = lambda n: not only or n in only  # noqa: E731
    variants = (("_noauto", dict(AUTOMASK=False)), ("_mean", dict(PHOTOMETRIC_REDUCE="mean")),
                ("_l1", dict(SSIM_WEIGHT=0.0)))
    if want("mono_2x32x64"):
        mono_case("mono_2x32x64", 2, 32, 64, variants=variants)
    if want("mono_1x50x70"):
        mono_case("mono_1x50x70", 1, 50, 70, seed=3, pose_scale=2.0)
    if want("mono_1x48x160_bigpose"):
        mono_case("mono_1x48x160_bigpose", 1, 48, 160, seed=5, pose_scale=8.0)
    if want("mono_cfg1_1x192x640"):
        mono_case("mono_cfg1_1x192x640", 1, 192, 640, store_inputs=False)
    if want("mono_cfg3_1x320x1024"):
        mono_case("mono_cfg3_1x320x1024", 1, 320, 1024, store_inputs=False, seed=3)
    if want("motion_2x32x64"):
        motion_case("motion_2x32x64", 2, 32, 64)
    if want("motion_variants_2x32x64"):
        motion_variants_case("motion_variants_2x32x64")


if __name__ == "__main__":
    main()
