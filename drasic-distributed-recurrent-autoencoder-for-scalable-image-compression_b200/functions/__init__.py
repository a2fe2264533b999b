"""The reference's `functions` package surface: Quantize (autograd function) and the functional pixel
(un)shuffle."""
from .quantize import Quantize  # noqa: F401
from .shuffle import pixel_shuffle, pixel_unshuffle  # noqa: F401
