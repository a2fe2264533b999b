from .cell import *          # noqa: F401,F403
from .quantization import *  # noqa: F401,F403
from .shuffle import *       # noqa: F401,F403
