"""MotionLearningModel: drop-in for the reference meta-architecture
(detectron2/modeling/meta_arch/MotionLearning.py:27-246).  Same constructor (cfg), batch-dict
contract and output keys.  Per scale, both rgbd_consistency_loss calls (MotionLearning.py:166-176)
and both smoothness_loss calls (:231-235) are ONE fused CUDA forward (+ one backward) launch; the
motion regularisers (motion_loss.py) consume the coords / occlusion maps the fused
kernel emits."""
from __future__ import annotations

from collections import defaultdict

import torch
import torch.nn as nn
import torch.nn.functional as F

from ...functional import MotionLossPlan, motion_rgbd_smoothness_loss
from ...geometry.camera import resize_img, resize_img_avgpool, scale_intrinsics, view_synthesis
from ...utils.memory import to_cuda
from ..losses.motion_loss import motion_consistency_loss_split, motion_field_regularizer_losses
from ..losses.losses import silog_loss, variance_loss
from ..losses.ssim_loss import WeightedSSIM
from ..nets import build_depth_net, build_pose_net
from .build import META_ARCH_REGISTRY


def _scaled_translation(pose, s):
    """[B,4,4] with the translation column divided by s (MotionLearning.py:160-161)."""
    out = pose.clone()
    out[:, :3, 3] = pose[:, :3, 3] / s
    return out


class OverallMotion:
    """batch['overall_motion'] entry (MotionLearning.py:143-154): the translation field t = pose[:, :3, [3], None] +
    residual field, [B,3,H,W].  The losses never need it as a tensor (the kernels add the two parts per pixel), and its
    only reader is the periodic image logging of projects/MotionLearning/train.py:150-153, so it is formed on first
    use: indexing, attributes and methods are those of the tensor."""

    def __init__(self, pose, field, size):
        self._pose, self._field, self._size, self._t = pose, field, size, None

    def tensor(self):
        if self._t is None:
            t = self._pose[:, :3, 3][:, :, None, None]
            self._t = t + self._field if self._field is not None else t.expand(-1, -1, *self._size)
        return self._t

    def __getitem__(self, i):
        return self.tensor()[i]

    def __getattr__(self, name):
        return getattr(self.tensor(), name)


@META_ARCH_REGISTRY.register()
class MotionLearningModel(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.depth_net = build_depth_net(cfg)
        self.pose_net = build_pose_net(cfg)

        self.num_scales = cfg.LOSS.NUM_SCALES
        self.depth_l1_loss_w = cfg.LOSS.DEPTH_L1_WEIGHT
        self.ssim_loss_w = cfg.LOSS.SSIM_WEIGHT
        self.c1, self.c2 = float(cfg.LOSS.C1), float(cfg.LOSS.C2)
        self.smooth_loss_w = cfg.LOSS.SMOOTHNESS_WEIGHT
        self.sup_loss_w = cfg.LOSS.SUPERVISED_WEIGHT
        self.var_loss_w = cfg.LOSS.VAR_LOSS_WEIGHT
        self.motion_smooth_loss_w = cfg.LOSS.MOTION_SMOOTHNESS_WEIGHT
        self.motion_sparsity_loss_w = cfg.LOSS.MOTION_SPARSITY_WEIGHT
        self.rot_cycle_loss_w = cfg.LOSS.ROT_CYCLE_WEIGHT
        self.trans_cycle_loss_w = cfg.LOSS.TRANS_CYCLE_WEIGHT
        self.scale_normalize = cfg.LOSS.SCALE_NORMALIZE
        self.pose_use_depth = cfg.MODEL.POSE_NET.USE_DEPTH
        self.with_mask = cfg.MODEL.get("WITH_MASK", False)
        self.mask_dilation = cfg.MODEL.get("MASK_DILATION", 8)
        self.return_loss = cfg.MODEL.get("RETURN_LOSS", False)
        # DEPTH_L1_WEIGHT / SUPERVISED_WEIGHT are 0 in every shipped config (projects/MotionLearning/configs/Base.yaml:11-26);
        # when set, their terms are added next to the fused loss from the stand-alone CUDA operators
        self.supervise_loss = silog_loss(cfg.LOSS.VARIANCE_FOCUS)

        self.register_buffer("pixel_mean", torch.Tensor(cfg.MODEL.PIXEL_MEAN).view(1, -1, 1, 1))
        self.register_buffer("pixel_std", torch.Tensor(cfg.MODEL.PIXEL_STD).view(1, -1, 1, 1))
        self._plans = {}
        self.ssim = WeightedSSIM(self.c1, self.c2)

    @property
    def device(self):
        return self.pixel_mean.device

    def _plan(self, batch, size, scale_w, with_field):
        key = (batch, tuple(size), scale_w, with_field, str(self.device))
        plan = self._plans.get(key)
        if plan is None:
            plan = MotionLossPlan(batch, size, self.device, 2, self.ssim_loss_w, self.c1, self.c2,
                                  scale=(scale_w, scale_w), with_field=with_field)
            self._plans[key] = plan
        return plan

    def forward(self, batch):
        batch = to_cuda(batch, self.device)
        if not (self.training or self.return_loss):
            batch["depth_net_input"] = (batch["img"] - self.pixel_mean) / self.pixel_std
            return self.depth_net(batch)

        frame1, frame2 = batch["img"], batch["ctx_img"][0]
        batch["depth_net_input"] = torch.cat([(frame1 - self.pixel_mean) / self.pixel_std,
                                              (frame2 - self.pixel_mean) / self.pixel_std], 0)
        batch = self.depth_net(batch)
        depth1, depth2 = zip(*[torch.chunk(d, 2, dim=0) for d in batch["depth_pred"]])

        in1, in2 = frame1, frame2
        if self.pose_use_depth:
            in1, in2 = torch.cat([in1, depth1[0]], 1), torch.cat([in2, depth2[0]], 1)
        batch["pose_net_input"] = torch.cat([torch.cat([in1, in2], 1), torch.cat([in2, in1], 1)], 0)
        batch = self.pose_net(batch)

        pose_1to2, pose_2to1 = torch.chunk(batch["pose_pred"], 2, dim=0)
        motion_1to2 = motion_2to1 = None
        if "motion_pred" in batch:
            motion_1to2, motion_2to1 = torch.chunk(batch["motion_pred"], 2, dim=0)
            if self.with_mask:
                mask1, mask2 = (batch["mask"] > 0).float(), (batch["ctx_mask"][0] > 0).float()
                if self.mask_dilation > 0:
                    k = self.mask_dilation * 2 + 1
                    mask1 = F.max_pool2d(mask1, k, stride=1, padding=self.mask_dilation)
                    mask2 = F.max_pool2d(mask2, k, stride=1, padding=self.mask_dilation)
                motion_1to2, motion_2to1 = motion_1to2 * mask1, motion_2to1 * mask2

        batch["depth_proximity_weight"], batch["overall_motion"] = [], []
        losses = defaultdict(lambda: 0)
        B = frame1.shape[0]
        K = batch["intrinsics"].float().contiguous()
        for i in reversed(range(self.num_scales)):
            scale_w = 1.0 / 2 ** i
            H, W = int(depth1[0].shape[-2] * scale_w), int(depth1[0].shape[-1] * scale_w)
            f1, f2 = resize_img_avgpool(frame1, (H, W)).contiguous(), resize_img_avgpool(frame2, (H, W)).contiguous()
            d1, d2 = resize_img_avgpool(depth1[0], (H, W)), resize_img_avgpool(depth2[0], (H, W))
            P12, P21 = pose_1to2, pose_2to1
            with_field = motion_1to2 is not None
            m12 = m21 = None
            if with_field:
                m12, m21 = resize_img_avgpool(motion_1to2, (H, W)), resize_img_avgpool(motion_2to1, (H, W))
            # overall translation fields (MotionLearning.py:143-155): the output dict carries them un-normalised (:154 comes
            # before the SCALE_NORMALIZE block); formed lazily, the losses take pose and residual field separately
            batch["overall_motion"].append((OverallMotion(P12, m12, (H, W)), OverallMotion(P21, m21, (H, W))))
            if self.scale_normalize:
                depth_mean = torch.mean(torch.cat([d1, d2], 0))
                d1, d2 = d1 / depth_mean, d2 / depth_mean
                P12, P21 = _scaled_translation(P12, depth_mean), _scaled_translation(P21, depth_mean)
                if with_field:
                    m12, m21 = m12 / depth_mean, m21 / depth_mean

            plan = self._plan(B, (H, W), scale_w, with_field)
            out, maps = motion_rgbd_smoothness_loss(
                plan, [f1, f2], [f2, f1], [d1, d2], [d2, d1], K, [P12, P21],
                [m12, m21] if with_field else None)
            # merge_loss(losses, output, scale_w) for both directions (MotionLearning.py:170,176)
            losses["rgb_l1_loss"] += (out[0, 0] + out[1, 0]) * scale_w
            if self.ssim_loss_w > 0.0:
                losses["ssim_loss"] += (out[0, 1] + out[1, 1]) * scale_w
            if self.smooth_loss_w > 0.0:
                losses["smooth_loss"] += (out[0, 2] + out[1, 2]) * (scale_w * self.smooth_loss_w)
            batch["depth_proximity_weight"].append((maps[0]["depth_proximity_weight"], maps[1]["depth_proximity_weight"]))

            if self.rot_cycle_loss_w > 0 or self.trans_cycle_loss_w > 0:   # MotionLearning.py:188-201
                for m, Pa, Pb, ma, mb in ((maps[0], P12, P21, m12, m21), (maps[1], P21, P12, m21, m12)):
                    rot, tr = motion_consistency_loss_split(m["coords_A_in_B"], m["occlusion_mask"], Pa, Pb, ma, mb)
                    losses["rot_loss"] += rot * scale_w * self.rot_cycle_loss_w
                    losses["trans_loss"] += tr * scale_w * self.trans_cycle_loss_w

            if with_field and (self.motion_smooth_loss_w > 0.0 or self.motion_sparsity_loss_w > 0.0):
                # MotionLearning.py:203-220: both regularisers of m / sqrt(3 mean(t^2) + 1e-12), fused per direction
                for P, m in ((P12, m12), (P21, m21)):
                    sm, sp = motion_field_regularizer_losses(P, m)
                    if self.motion_smooth_loss_w > 0.0:
                        losses["motion_smooth_loss"] += sm * scale_w * self.motion_smooth_loss_w
                    if self.motion_sparsity_loss_w > 0.0:
                        losses["motion_sparsity_loss"] += sp * scale_w * self.motion_sparsity_loss_w

            if self.depth_l1_loss_w > 0:   # MotionLearning.py:264-267, both directions (:166-176)
                Ki = scale_intrinsics(K.clone(), scale_w, scale_w)
                t12, t21 = OverallMotion(P12, m12, (H, W)).tensor(), OverallMotion(P21, m21, (H, W)).tensor()
                for fb, da, db, R, t in ((f2, d1, d2, P12[:, :3, :3], t12), (f1, d2, d1, P21[:, :3, :3], t21)):
                    losses["depth_l1_loss"] += self._depth_l1_loss(fb, da, db, Ki, R.contiguous(), t.contiguous()) * scale_w
            if self.sup_loss_w > 0.0:   # MotionLearning.py:222-229 (on the un-normalised resized depths)
                r1, r2 = resize_img_avgpool(depth1[0], (H, W)), resize_img_avgpool(depth2[0], (H, W))
                g1 = resize_img(batch["depth"], (H, W), mode="nearest")
                g2 = resize_img(batch["ctx_depth"][0], (H, W), mode="nearest")
                losses["sup_loss"] += (self.supervise_loss(r1, g1) + self.supervise_loss(r2, g2)) * scale_w * self.sup_loss_w
            if self.var_loss_w > 0.0:   # MotionLearning.py:237-239 (on the un-normalised resized depths)
                r1, r2 = resize_img_avgpool(depth1[0], (H, W)), resize_img_avgpool(depth2[0], (H, W))
                losses["var_loss"] += (variance_loss(r1) + variance_loss(r2)) * scale_w * self.var_loss_w

        batch.update(losses)
        return batch

    def _depth_l1_loss(self, frame_B, depth_A, depth_B, intrinsics, R_A2B, t_A2B):
        """depth_l1_loss of one direction (MotionLearning.py:264-267) from the view_synthesis operator."""
        sampled, depth_in_B, _, valid = view_synthesis(torch.cat([frame_B, depth_B], 1), depth_A, intrinsics, R_A2B, t_A2B)
        sampled_depth_B = sampled[:, 3:4]
        occ = (depth_in_B < sampled_depth_B).float() * valid.float()
        l1 = (sampled_depth_B.detach() - depth_in_B).abs() * occ
        return (l1.sum([1, 2, 3]) / (occ.sum([1, 2, 3]) + 1)).mean() * self.depth_l1_loss_w

    def rgbd_consistency_loss(self, frame_A, frame_B, depth_A, depth_B, intrinsics, R_A2B, t_A2B):
        """One direction of the rgb-d consistency loss as a dict (MotionLearning.py:248-291): the un-fused form
        of what forward() computes in the fused kernel -- view_synthesis + WeightedSSIM as CUDA operators, the
        per-sample statistics in torch."""
        out = {}
        sampled, depth_in_B, coords, valid = view_synthesis(torch.cat([frame_B, depth_B], 1), depth_A, intrinsics,
                                                            R_A2B, t_A2B)
        out["coords_A_in_B"] = coords
        sampled_frame_B, sampled_depth_B = torch.split(sampled, [3, 1], dim=1)
        occ = (depth_in_B < sampled_depth_B).float() * valid.float()
        out["occlusion_mask"] = occ
        normalizer = occ.sum([1, 2, 3]) + 1
        if self.depth_l1_loss_w > 0:
            l1 = (sampled_depth_B.detach() - depth_in_B).abs() * occ
            out["depth_l1_loss"] = (l1.sum([1, 2, 3]) / normalizer).mean() * self.depth_l1_loss_w
        out["rgb_l1_loss"] = ((sampled_frame_B - frame_A).abs() * occ).mean()
        if self.ssim_loss_w > 0.0:
            err = (depth_in_B - sampled_depth_B) ** 2
            m2 = ((err * occ).sum([1, 2, 3]) / normalizer + 1e-4).view(-1, 1, 1, 1)
            weight = ((m2 / (err + m2)) * valid.float()).detach()
            ssim_map, avg_w = self.ssim(sampled_frame_B, frame_A, weight)
            out["depth_proximity_weight"] = weight
            out["ssim_loss"] = (ssim_map * avg_w).mean() * self.ssim_loss_w * 0.5
        return out
