"""MonoDepth2Model: drop-in for the reference meta-architecture
(detectron2/modeling/meta_arch/MonoDepth2.py:20-128).  Same constructor (cfg), same batch-dict
contract, same output keys (`rec_loss`, `smooth_loss`, eval: `depth_pred`), same LOSS.* config
keys -- but the multi-scale loss loop (MonoDepth2.py:78-124) is two fused CUDA launches."""
from __future__ import annotations

import torch
import torch.nn as nn

from ...functional import MonoLossPlan, mono_photometric_smoothness_loss
from ...geometry.camera import resize_img, scale_intrinsics, view_synthesis
from ...ops import resize_pyramid
from ...utils.memory import to_cuda
from ..losses.losses import silog_loss, variance_loss
from ..losses.smoothness_loss import smoothness_loss
from ..losses.ssim_loss import SSIM
from ..nets import build_depth_net, build_pose_net
from .build import META_ARCH_REGISTRY


@META_ARCH_REGISTRY.register()
class MonoDepth2Model(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.depth_net = build_depth_net(cfg)
        self.pose_net = build_pose_net(cfg)

        self.ssim_loss_weight = cfg.LOSS.SSIM_WEIGHT
        self.c1, self.c2 = cfg.LOSS.C1, cfg.LOSS.C2
        self.photometric_reduce = cfg.LOSS.PHOTOMETRIC_REDUCE
        self.use_automask = cfg.LOSS.AUTOMASK
        self.clip_loss = cfg.LOSS.CLIP
        self.var_loss_w = cfg.LOSS.VAR_LOSS_WEIGHT
        self.sup_loss_w = cfg.LOSS.SUPERVISED_WEIGHT
        self.smooth_loss_w = cfg.LOSS.SMOOTHNESS_WEIGHT
        if self.photometric_reduce not in ("min", "mean"):
            raise NotImplementedError(self.photometric_reduce)
        # LOSS.CLIP > 0 (off in every shipped config, Base.yaml:8) caps every photometric map at its own mean + CLIP * std
        # (MonoDepth2.py:146-149), a global statistic per map with a host read: that configuration runs the reference's
        # loop over the stand-alone CUDA operators (_unfused_losses) instead of the fused kernels
        self.supervise_loss = silog_loss(cfg.LOSS.VARIANCE_FOCUS)

        self.register_buffer("pixel_mean", torch.Tensor(cfg.MODEL.PIXEL_MEAN).view(1, -1, 1, 1))
        self.register_buffer("pixel_std", torch.Tensor(cfg.MODEL.PIXEL_STD).view(1, -1, 1, 1))
        self._plans = {}
        self.ssim = SSIM(self.c1, self.c2)
        # disp_to_depth range of the depth nets (DepthResNet.py:41: min_depth=0.1, max_depth=cfg.MODEL.MAX_DEPTH): used when
        # a depth net hands over `disp_pred` / `disp_logit_pred` instead of `depth_pred`
        self.min_depth = float(cfg.MODEL.get("MIN_DEPTH", 0.1)) if hasattr(cfg.MODEL, "get") else 0.1
        self.max_depth = float(cfg.MODEL.get("MAX_DEPTH", 80.0)) if hasattr(cfg.MODEL, "get") else 80.0

    @property
    def device(self):
        return self.pixel_mean.device

    def _plan(self, batch, sizes, n_sources, full_size, depth_mode="depth"):
        key = (batch, tuple(sizes), n_sources, tuple(full_size), str(self.device), depth_mode)
        plan = self._plans.get(key)
        if plan is None:
            plan = MonoLossPlan(batch, sizes, n_sources, full_size, self.device,
                                ssim_weight=self.ssim_loss_weight, c1=self.c1, c2=self.c2,
                                smooth_weight=self.smooth_loss_w, automask=self.use_automask,
                                reduce=self.photometric_reduce, depth_mode=depth_mode,
                                min_depth=self.min_depth, max_depth=self.max_depth)
            self._plans[key] = plan
        return plan

    def forward(self, batch):
        batch = to_cuda(batch, self.device)
        output = {}
        batch["depth_net_input"] = (batch["img"] - self.pixel_mean) / self.pixel_std
        batch = self.depth_net(batch)

        if self.training:
            batch["pose_net_input"] = torch.cat([batch["img"]] + batch["ctx_img"], 1)
            batch = self.pose_net(batch)  # num_ctx * [B, 4, 4]

            image = batch["img_orig"]
            contexts = batch["ctx_img_orig"]
            # A depth net may skip its tail (softplus + disp_to_depth, depth_decoder.py:9-18,108; DepthResNet.py:57) and
            # hand over the disparities (`disp_pred`) or the last convolution's output (`disp_logit_pred`): the fused
            # loss decodes them in-kernel and returns the gradient w.r.t. that tensor
            depth_mode = "depth"
            if "depth_pred" not in batch:
                depth_mode = "disp" if "disp_pred" in batch else "logit"
                batch["depth_pred"] = batch["disp_pred"] if depth_mode == "disp" else batch["disp_logit_pred"]
                if self.clip_loss > 0.0 or self.sup_loss_w > 0.0 or self.var_loss_w > 0.0:
                    raise NotImplementedError("disp_pred / disp_logit_pred need the fused loss (CLIP, SUPERVISED_WEIGHT and "
                                              "VAR_LOSS_WEIGHT off): hand over depth_pred for those terms")
            depth_pred = batch["depth_pred"]
            pose_pred = list(batch["pose_pred"])
            sizes = [tuple(d.shape[-2:]) for d in depth_pred]

            # image pyramid (resize_img, MonoDepth2.py:82,88): every frame to every scale in one launch
            if image.is_cuda and image.dtype == torch.float32 and len(contexts) + 1 <= 5 and \
                    all(c.shape == image.shape for c in contexts):
                pyr = resize_pyramid([image] + list(contexts), sizes)
                target = pyr[0]
                source = [[pyr[1 + j][i] for j in range(len(contexts))] for i in range(len(sizes))]
            else:
                target = [resize_img(image, s) for s in sizes]
                source = [[resize_img(c, s) for c in contexts] for s in sizes]

            if self.clip_loss > 0.0:
                rec, smooth = self._unfused_losses(target, source, list(depth_pred), batch["intrinsics"].float(),
                                                   pose_pred, tuple(image.shape[-2:]))
            else:
                plan = self._plan(image.shape[0], sizes, len(contexts), tuple(image.shape[-2:]), depth_mode)
                rec, smooth, _ = mono_photometric_smoothness_loss(plan, target, source, list(depth_pred),
                                                                  batch["intrinsics"].float(), pose_pred)
            output["rec_loss"] = rec
            if self.smooth_loss_w > 0.0:
                output["smooth_loss"] = smooth
            if self.sup_loss_w > 0.0:   # MonoDepth2.py:107-110 (weighted by SMOOTHNESS_WEIGHT there, kept as is)
                n = len(depth_pred)
                output["sup_loss"] = sum(
                    self.supervise_loss(d, resize_img(batch["depth"], d.shape[-2:], mode="nearest"))
                    * (1.0 / 2 ** (n - i - 1)) * self.smooth_loss_w / n for i, d in enumerate(depth_pred))
            if self.var_loss_w > 0.0:   # PackNet (packnet_1a.yaml:12); MonoDepth2.py:112-113
                n = len(depth_pred)
                output["var_loss"] = sum(variance_loss(d) * (1.0 / 2 ** (n - i - 1)) * self.var_loss_w / n
                                         for i, d in enumerate(depth_pred))
        else:
            output["depth_pred"] = batch["depth_pred"][0]
        return output

    def _unfused_losses(self, target, source, depth_pred, intrinsics, pose_pred, full_size):
        """The loss loop of MonoDepth2.py:78-121 over the stand-alone CUDA operators (view_synthesis, SSIM,
        smoothness_loss); torch only concatenates, reduces and clips.  Used when LOSS.CLIP > 0."""
        n = len(depth_pred)
        H, W = full_size
        rec, smooth = 0.0, 0.0
        for i, d in enumerate(depth_pred):
            h, w = d.shape[-2:]
            K = scale_intrinsics(intrinsics.clone(), w / W, h / H)
            cand = []
            for src, T in zip(source[i], pose_pred):
                cand.append(self.rgb_consistency_loss(target[i], src, d, K, T[:, :3, :3].contiguous(), T[:, :3, 3:, None]))
                if self.use_automask:
                    cand.append(self.rgb_consistency_loss(target[i], src, d, K, None, None))
            if self.photometric_reduce == "mean":
                rec = rec + sum(c.mean() for c in cand) / len(cand) / n
            else:
                rec = rec + torch.cat(cand, 1).min(1, True)[0].mean() / n
            if self.smooth_loss_w > 0.0:
                smooth = smooth + smoothness_loss(d, target[i]) * (1.0 / 2 ** (n - i - 1)) * self.smooth_loss_w / n
        return rec, smooth

    def rgb_consistency_loss(self, frame_A, frame_B, depth_A, intrinsics, R_A2B=None, t_A2B=None):
        """Photometric error of one (scale, source) pair, [B,1,H,W] (MonoDepth2.py:130-151): the un-fused
        form of what forward() computes inside the fused kernels -- view_synthesis + SSIM as two CUDA
        operators.  With R_A2B / t_A2B omitted frame_B is compared unwarped (identity / automask term)."""
        if R_A2B is not None and t_A2B is not None:
            if t_A2B.dim() == 2:
                t_A2B = t_A2B[:, :, None, None]
            sampled, _, _, _ = view_synthesis(frame_B, depth_A, intrinsics, R_A2B, t_A2B)
        else:
            sampled = frame_B
        loss = (sampled - frame_A).abs().mean(1, True)
        if self.ssim_loss_weight > 0.0:
            loss = self.ssim(sampled, frame_A).mean(1, True) * self.ssim_loss_weight + loss * (1 - self.ssim_loss_weight)
        if self.clip_loss > 0.0:   # MonoDepth2.py:146-149 (the float() is the reference's host read)
            loss = torch.clamp(loss, max=float((loss.mean() + self.clip_loss * loss.std()).detach()))
        return loss
