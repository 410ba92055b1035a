from .camera import inv_intrinsics, resize_img, resize_img_avgpool, scale_intrinsics  # noqa: F401
