"""Focal-loss modules over the fused kernels: ``BinaryFocalLoss`` (sigmoid or channel-softmax focal term on dense or
label targets), ``CrossEntropyFocalLoss`` (softmax focal term, ``nn.CrossEntropyLoss``-style targets) and the
deprecated ``FocalLoss`` alias.

Constructor arguments, defaults, public attributes and ``forward`` conventions are those of
``pytorch_toolbelt/losses/focal.py:13-166``; nothing is computed here -- ``forward`` hands the options to
``losses.functional`` which launches ``tb_focal_fwd_bwd`` / ``tb_softmax_focal_fwd_bwd`` (loss and gradient in one
pass; the one-hot expansion of label targets, focal.py:94-105, happens inside the kernel).
"""
from typing import Optional

import torch
from torch import Tensor, nn

from ..utils.support import pytorch_toolbelt_deprecated
from .functional import _focal_from_labels, focal_loss_with_logits, softmax_focal_loss_with_logits

__all__ = ["CrossEntropyFocalLoss", "BinaryFocalLoss", "FocalLoss"]

_BINARY_OPTIONS = ("alpha", "gamma", "ignore_index", "reduction", "normalized", "reduced_threshold", "activation",
                   "softmax_dim")
_MULTICLASS_OPTIONS = ("gamma", "reduction", "normalized", "reduced_threshold", "ignore_index")


def _as_weight_buffer(class_weights) -> Optional[Tensor]:
    if class_weights is None or torch.is_tensor(class_weights):
        return class_weights
    return torch.tensor(list(class_weights), dtype=torch.float32)


class BinaryFocalLoss(nn.Module):
    """Focal loss for binary / multi-label problems.

    ``alpha``: prior of the positive class (None = no class balancing); ``gamma``: focusing power; ``ignore_index``:
    target value whose elements contribute nothing; ``reduction``: "mean" | "sum" | "batchwise_mean" | "none";
    ``normalized``: divide by the sum of the focal terms; ``reduced_threshold``: reduced focal loss (upstream advises
    ``reduction="sum"`` with it); ``activation``: "sigmoid", or "softmax" with ``softmax_dim=1``; ``class_weights``:
    per-channel multipliers ``[C]``.
    """

    __constants__ = list(_BINARY_OPTIONS)

    def __init__(self, alpha: Optional[float] = None, gamma: float = 2.0, ignore_index: Optional[int] = None,
                 reduction: str = "mean", normalized: bool = False, reduced_threshold: Optional[float] = None,
                 activation: str = "sigmoid", softmax_dim: Optional[int] = None,
                 class_weights: Optional[Tensor] = None):
        super().__init__()
        given = dict(alpha=alpha, gamma=gamma, ignore_index=ignore_index, reduction=reduction, normalized=normalized,
                     reduced_threshold=reduced_threshold, activation=activation, softmax_dim=softmax_dim)
        for name in _BINARY_OPTIONS:
            setattr(self, name, given[name])
        self.register_buffer("class_weights", _as_weight_buffer(class_weights), persistent=False)

    def _options(self) -> dict:
        return {name: getattr(self, name) for name in _BINARY_OPTIONS}

    def __repr__(self):
        shown = ", ".join(f"{name}={value}" for name, value in self._options().items())
        weights = None if self.class_weights is None else self.class_weights.tolist()
        return f"{type(self).__name__}({shown}, class_weights={weights})"

    def forward(self, inputs: Tensor, targets: Tensor) -> Tensor:
        """``inputs`` [B,C,*] logits; ``targets`` of the same shape, or [B,*] integer labels that are one-hot encoded
        with respect to ``ignore_index`` (inside the kernel)."""
        options = self._options()
        label_targets = targets.dim() + 1 == inputs.dim()
        if not label_targets:
            return focal_loss_with_logits(inputs, targets, class_weights=self.class_weights, **options)
        if self.activation != "sigmoid" and self.softmax_dim != 1:
            raise NotImplementedError("activation='softmax' is implemented for softmax_dim=1")
        options.pop("softmax_dim")
        return _focal_from_labels(inputs, targets, self.class_weights, eps=1e-6, **options)


class CrossEntropyFocalLoss(nn.Module):
    """Multi-class focal loss whose focal term comes from the channel softmax; ``targets`` are class indices with one
    dimension less than ``inputs``; ``ignore_index`` (default -100) marks pixels that contribute nothing."""

    def __init__(self, gamma: float = 2.0, reduction: str = "mean", normalized: bool = False,
                 reduced_threshold: Optional[float] = None, ignore_index: int = -100,
                 class_weights: Optional[Tensor] = None):
        super().__init__()
        given = dict(gamma=gamma, reduction=reduction, normalized=normalized, reduced_threshold=reduced_threshold,
                     ignore_index=ignore_index)
        for name in _MULTICLASS_OPTIONS:
            setattr(self, name, given[name])
        self.register_buffer("class_weights", class_weights, persistent=False)

    def forward(self, inputs: Tensor, targets: Tensor) -> Tensor:
        options = {name: getattr(self, name) for name in _MULTICLASS_OPTIONS}
        return softmax_focal_loss_with_logits(inputs, targets, class_weights=self.class_weights, **options)


@pytorch_toolbelt_deprecated("Class FocalLoss is deprecated. Please use CrossEntropyFocalLoss instead.")
def FocalLoss(*input, **kwargs):
    return CrossEntropyFocalLoss(*input, **kwargs)
