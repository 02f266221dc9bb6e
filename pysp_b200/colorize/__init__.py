from .transform import cam_to_lin_srgb, cam_to_rgb_norm, clip_rgb, lin_srgb_to_srgb  # noqa: F401
from .rgb_space import ArbitraryRgbColorspace, LinRgbColorspace  # noqa: F401
