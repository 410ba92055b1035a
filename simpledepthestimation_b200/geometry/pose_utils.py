"""pose_vec2mat with the reference's name (detectron2/geometry/pose_utils.py:98-137), backed by the
sde_pose_vec2mat_* CUDA entry points (forward and backward)."""
from __future__ import annotations

from ..ops import pose_vec2mat  # noqa: F401
