"""Reference: colorize/rgb_space.py:9-56 (host-side 3x3 set-up, float64 NumPy)."""
from ..colour import ArbitraryRgbColorspace, LinRgbColorspace  # noqa: F401
