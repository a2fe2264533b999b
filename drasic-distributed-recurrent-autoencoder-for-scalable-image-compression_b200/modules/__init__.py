"""The reference's `modules` package surface (src/modules/__init__.py): cell classes, the quantizer module and
the pixel (un)shuffle modules, all backed by the sm_100a kernels of this repo."""
from .cell import *  # noqa: F401,F403
from .quantization import Quantization  # noqa: F401
from .shuffle import PixelShuffle, PixelUnShuffle  # noqa: F401
