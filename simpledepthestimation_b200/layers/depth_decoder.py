"""disp_to_depth with the reference's name (detectron2/layers/depth_decoder.py:9-18), backed by the
sde_disp_to_depth_* CUDA entry points.  The decoder itself (convolutions) is out of scope."""
from __future__ import annotations

from ..ops import disp_to_depth  # noqa: F401
