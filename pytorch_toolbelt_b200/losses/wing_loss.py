"""``WingLoss`` module over the fused wing kernel (``tb_wing_fwd_bwd``).

Same constructor, attributes and call convention as ``pytorch_toolbelt/losses/wing_loss.py:8-15``; the arithmetic
(``|t - x| < width ? width * ln(1 + d / curvature) : d - C``) lives in ``losses.functional.wing_loss``.
"""
import math

from torch import Tensor
from torch.nn.modules.loss import _Loss

from .functional import wing_loss

__all__ = ["WingLoss"]


class WingLoss(_Loss):
    """Wing loss for landmark regression (Feng et al., CVPR 2018): logarithmic near zero, L1 beyond ``width``."""

    def __init__(self, width: float = 5, curvature: float = 0.5, reduction: str = "mean") -> None:
        super().__init__(reduction=reduction)
        self.width, self.curvature = width, curvature

    @property
    def linear_offset(self) -> float:
        """The constant C that makes the two branches meet at ``|t - x| == width``."""
        return self.width - self.width * math.log(1.0 + self.width / self.curvature)

    def extra_repr(self) -> str:
        return f"width={self.width}, curvature={self.curvature}, reduction={self.reduction!r}"

    def forward(self, prediction: Tensor, target: Tensor) -> Tensor:
        return wing_loss(prediction, target, width=self.width, curvature=self.curvature, reduction=self.reduction)
