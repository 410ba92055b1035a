from .meta_arch import META_ARCH_REGISTRY, build_model  # noqa: F401
from .nets import DEPTH_NET_REGISTRY, POSE_NET_REGISTRY, build_depth_net, build_pose_net  # noqa: F401
