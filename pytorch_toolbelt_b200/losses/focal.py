"""BinaryFocalLoss / CrossEntropyFocalLoss / FocalLoss: drop-in for ``pytorch_toolbelt/losses/focal.py``."""
from typing import Optional

import torch
from torch import Tensor, nn

from ..utils.support import pytorch_toolbelt_deprecated
from .functional import _focal_from_labels, focal_loss_with_logits, softmax_focal_loss_with_logits

__all__ = ["CrossEntropyFocalLoss", "BinaryFocalLoss", "FocalLoss"]


class BinaryFocalLoss(nn.Module):
    __constants__ = ["alpha", "gamma", "reduction", "ignore_index", "normalized", "reduced_threshold", "activation"]

    def __init__(
        self,
        alpha: Optional[float] = None,
        gamma: float = 2.0,
        ignore_index: Optional[int] = None,
        reduction: str = "mean",
        normalized: bool = False,
        reduced_threshold: Optional[float] = None,
        activation: str = "sigmoid",
        softmax_dim: Optional[int] = None,
        class_weights: Optional[Tensor] = None,
    ):
        """
        :param alpha: Prior probability of having positive value in target.
        :param gamma: Power factor for dampening weight (focal strength).
        :param ignore_index: targets equal to ignore_index do not contribute to the loss.
        :param reduced_threshold: switch to reduced focal loss (use reduction="sum" with it, as upstream advises).
        :param activation: "sigmoid" or "softmax" (then softmax_dim must be 1).
        """
        super().__init__()
        self.alpha = alpha
        self.gamma = gamma
        self.ignore_index = ignore_index
        self.reduction = reduction
        self.normalized = normalized
        self.reduced_threshold = reduced_threshold
        self.activation = activation
        self.softmax_dim = softmax_dim
        if class_weights is not None and not torch.is_tensor(class_weights):
            class_weights = torch.tensor(list(class_weights), dtype=torch.float32)
        self.register_buffer("class_weights", class_weights, persistent=False)

    def __repr__(self):
        cw = None if self.class_weights is None else self.class_weights.tolist()
        return (
            f"{self.__class__.__name__}(alpha={self.alpha}, gamma={self.gamma}, ignore_index={self.ignore_index}, "
            f"reduction={self.reduction}, normalized={self.normalized}, reduced_threshold={self.reduced_threshold}, "
            f"activation={self.activation}, softmax_dim={self.softmax_dim}, class_weights={cw})"
        )

    def forward(self, inputs: Tensor, targets: Tensor) -> Tensor:
        """
        :param inputs: [B,C,H,W] logits
        :param targets: [B,C,H,W] (same shape) or [B,H,W] long class indices, one-hot encoded w.r.t. ignore_index
        """
        kw = dict(alpha=self.alpha, gamma=self.gamma, reduced_threshold=self.reduced_threshold, reduction=self.reduction,
                  normalized=self.normalized, ignore_index=self.ignore_index, activation=self.activation, eps=1e-6)
        if len(targets.shape) + 1 == len(inputs.shape):
            if self.activation != "sigmoid" and self.softmax_dim != 1:
                raise NotImplementedError("activation='softmax' is implemented for softmax_dim=1")
            return _focal_from_labels(inputs, targets, self.class_weights, **kw)
        kw.pop("eps")
        return focal_loss_with_logits(inputs, targets, softmax_dim=self.softmax_dim, class_weights=self.class_weights, **kw)


class CrossEntropyFocalLoss(nn.Module):
    """Focal loss for multi-class problems (softmax focal term); targets have one dim less, like nn.CrossEntropyLoss."""

    def __init__(
        self,
        gamma: float = 2.0,
        reduction: str = "mean",
        normalized: bool = False,
        reduced_threshold: Optional[float] = None,
        ignore_index: int = -100,
        class_weights: Optional[Tensor] = None,
    ):
        super().__init__()
        self.gamma = gamma
        self.reduction = reduction
        self.reduced_threshold = reduced_threshold
        self.normalized = normalized
        self.ignore_index = ignore_index
        self.register_buffer("class_weights", class_weights, persistent=False)

    def forward(self, inputs: Tensor, targets: Tensor) -> Tensor:
        return softmax_focal_loss_with_logits(
            inputs,
            targets,
            gamma=self.gamma,
            reduction=self.reduction,
            normalized=self.normalized,
            reduced_threshold=self.reduced_threshold,
            ignore_index=self.ignore_index,
            class_weights=self.class_weights,
        )


@pytorch_toolbelt_deprecated("Class FocalLoss is deprecated. Please use CrossEntropyFocalLoss instead.")
def FocalLoss(*input, **kwargs):
    return CrossEntropyFocalLoss(*input, **kwargs)
