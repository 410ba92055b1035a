from .motion_loss import motion_consistency_loss, motion_smoothness_loss_fn, motion_sparsity_loss_fn  # noqa: F401
from .smoothness_loss import smoothness_loss  # noqa: F401
from .ssim_loss import SSIM, WeightedSSIM  # noqa: F401
from .losses import silog_loss, variance_loss  # noqa: F401
