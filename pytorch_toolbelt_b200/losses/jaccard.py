"""``JaccardLoss`` (IoU loss): the Dice machinery of ``dice.py`` with the Jaccard score
``(I + smooth) / max(S - I + smooth, eps)`` selected in the kernels. Constructor and attributes follow
``pytorch_toolbelt/losses/jaccard.py:18-103`` (no ``ignore_index`` there, hence none here)."""
from typing import List

from .dice import BINARY_MODE, MULTICLASS_MODE, MULTILABEL_MODE, _RegionLossModule

__all__ = ["JaccardLoss", "BINARY_MODE", "MULTICLASS_MODE", "MULTILABEL_MODE"]


class JaccardLoss(_RegionLossModule):
    """Jaccard loss for image segmentation: binary, multi-class and multi-label targets."""

    _jaccard = True

    def __init__(self, mode: str, classes: List[int] = None, log_loss=False, from_logits=True, smooth=0, eps=1e-7):
        super().__init__(mode, classes, log_loss, from_logits, smooth, eps, ignore_index=None)
