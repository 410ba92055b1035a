from .build import META_ARCH_REGISTRY, build_model  # noqa: F401
from .MonoDepth2 import MonoDepth2Model  # noqa: F401
from .MotionLearning import MotionLearningModel  # noqa: F401
