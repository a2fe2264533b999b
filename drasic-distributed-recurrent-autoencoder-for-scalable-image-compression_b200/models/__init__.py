from .dist_codec import *  # noqa: F401,F403
from .sep_codec import *   # noqa: F401,F403
