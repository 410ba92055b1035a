from .depth_decoder import disp_to_depth  # noqa: F401
