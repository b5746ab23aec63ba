"""pytorch_toolbelt_b200: a B200 (sm_100a) native engine for the tiled-inference / TTA / segmentation-loss
hot path of BloodAxe/pytorch-toolbelt, behind the reference's own Python API.

    from pytorch_toolbelt_b200.inference.tiles import ImageSlicer, TileMerger, CudaTileMerger
    from pytorch_toolbelt_b200.inference import tta
    from pytorch_toolbelt_b200 import losses as L

All pixel-touching work runs in hand-written CUDA kernels reached through the C-ABI of include/tb200.h
(libtb200.so next to this file). There is no CPU / eager fallback.
"""
__version__ = "0.1.0"

from . import _native  # noqa: F401  (does not dlopen until first use)
