"""JaccardLoss: drop-in for ``pytorch_toolbelt/losses/jaccard.py`` (shares the fused kernels with DiceLoss)."""
from typing import List

import torch
from torch import Tensor
from torch.nn.modules.loss import _Loss

from .dice import BINARY_MODE, MULTICLASS_MODE, MULTILABEL_MODE, _RegionLoss

__all__ = ["JaccardLoss", "BINARY_MODE", "MULTICLASS_MODE", "MULTILABEL_MODE"]


class JaccardLoss(_Loss):
    """Jaccard (IoU) loss for image segmentation: binary, multi-class and multi-label (reference jaccard.py:18-103)."""

    def __init__(self, mode: str, classes: List[int] = None, log_loss=False, from_logits=True, smooth=0, eps=1e-7):
        assert mode in {BINARY_MODE, MULTILABEL_MODE, MULTICLASS_MODE}
        super(JaccardLoss, self).__init__()
        self.mode = mode
        if classes is not None:
            assert mode != BINARY_MODE, "Masking classes is not supported with mode=binary"
            classes = [int(c) for c in (classes.tolist() if torch.is_tensor(classes) else classes)]
        self.classes = classes
        self.from_logits = from_logits
        self.smooth = smooth
        self.eps = eps
        self.log_loss = log_loss

    def forward(self, y_pred: Tensor, y_true: Tensor) -> Tensor:
        assert y_true.size(0) == y_pred.size(0)
        cfg = dict(jaccard=True, mode=self.mode, classes=self.classes, log_loss=self.log_loss,
                   from_logits=self.from_logits, smooth=self.smooth, eps=self.eps, ignore_index=None)
        cfg.update(self._sync_cfg())
        return _RegionLoss.apply(y_pred, y_true, cfg)

    def _sync_cfg(self) -> dict:
        return {}
