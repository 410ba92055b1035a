from .camera import inv_intrinsics, resize_img, resize_img_avgpool, scale_intrinsics  # noqa: F401
from .pose_utils import pose_vec2mat  # noqa: F401
