"""Segmentation losses on fused B200 kernels; drop-in for the hot-path subset of ``pytorch_toolbelt.losses``
(focal, dice, jaccard, lovasz, wing) plus the element-wise losses of SURVEY 8f rank 4 (soft BCE, balanced BCE, quality focal,
soft F1). The other reference losses are out of scope."""
from . import functional  # noqa: F401
from .dice import *  # noqa: F401,F403
from .elementwise import *  # noqa: F401,F403
from .focal import *  # noqa: F401,F403
from .functional import *  # noqa: F401,F403
from .jaccard import *  # noqa: F401,F403
from .joint import *  # noqa: F401,F403
from .lovasz import *  # noqa: F401,F403
from .wing_loss import *  # noqa: F401,F403
