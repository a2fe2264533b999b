from .shuffle import *   # noqa: F401,F403
from .quantize import *  # noqa: F401,F403
