from . import functional, tiles, tta  # noqa: F401
from .tiles import *  # noqa: F401,F403
from .tta import *  # noqa: F401,F403
from .engine import *  # noqa: F401,F403
