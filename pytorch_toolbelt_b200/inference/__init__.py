from . import functional, tiles, tiles_3d, tta  # noqa: F401
from .tiles import *  # noqa: F401,F403
from .tta import *  # noqa: F401,F403
from .engine import *  # noqa: F401,F403
from .tiles_3d import *  # noqa: F401,F403
