"""The reference's `models` package surface: the zero-argument factories train_model.py:58 evaluates
(`models.dist_codec()`, `models.sep_codec()`) and their Model classes."""
from . import dist_codec as _dist, sep_codec as _sep
from .dist_codec import *  # noqa: F401,F403
from .sep_codec import *  # noqa: F401,F403
