"""Plug-in points for the depth / pose networks.

The CNNs are outside this package's scope (they are cuDNN convolutions in the reference,
detectron2/modeling/depth_net, pose_net); a trainer registers its own builders here under the
names its config uses (`cfg.MODEL.DEPTH_NET.NAME`, `cfg.MODEL.POSE_NET.NAME`), exactly like the
reference's DEPTH_NET_REGISTRY / POSE_NET_REGISTRY.  A network is an nn.Module whose
forward(batch) returns the batch with `depth_pred` (list, finest first) / `pose_pred` added.
"""
from ..utils.registry import Registry

DEPTH_NET_REGISTRY = Registry("DEPTH_NET")
POSE_NET_REGISTRY = Registry("POSE_NET")


def build_depth_net(cfg):
    return DEPTH_NET_REGISTRY.get(cfg.MODEL.DEPTH_NET.NAME)(cfg)


def build_pose_net(cfg):
    return POSE_NET_REGISTRY.get(cfg.MODEL.POSE_NET.NAME)(cfg)
