"""SoftBCEWithLogitsLoss, BalancedBCEWithLogitsLoss, QualityFocalLoss, BinarySoftF1Loss, SoftF1Loss.

Drop-ins for ``pytorch_toolbelt/losses/{soft_bce,balanced_bce,quality_focal_loss,soft_f1}.py`` on the fused
forward+gradient kernels of ``csrc/losses_extra.cu`` (``tb_elem_loss_fwd_bwd``, ``tb_soft_f1_fwd_bwd``): one streaming pass
over (logits, target) yields the value and d loss / d logits (two passes where the loss is a ratio of global sums), where
the reference runs 6-10 element-wise ATen kernels each way through autograd. Constructor signatures, defaults and
reductions are the reference's; so are its quirks (BalancedBCE raises its weights to ``gamma`` twice, balanced_bce.py:29-33;
SoftF1Loss compares the one-hot VALUES with ``ignore_index``, soft_f1.py:97-101).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch
from torch import Tensor, nn

from .. import _native as N
from .._native import TB
from . import _kernels as K
from .functional import _StoredGradLoss

__all__ = ["SoftBCEWithLogitsLoss", "BalancedBCEWithLogitsLoss", "balanced_binary_cross_entropy_with_logits",
           "QualityFocalLoss", "soft_micro_f1", "BinarySoftF1Loss", "SoftF1Loss"]

_REDUCTION = {"mean": TB.LOSS_MEAN, "sum": TB.LOSS_SUM, "normalized": TB.LOSS_NORMALIZED}


def _channel_vector(v: Optional[Tensor], c: int, device, what: str):
    """`weight` / `pos_weight` of BCE-with-logits broadcast against [B,C,*]: a scalar or one entry per channel."""
    if v is None:
        return None, 0
    w = v.detach().to(device=device, dtype=torch.float32).reshape(-1).contiguous()
    if w.numel() not in (1, c):
        raise NotImplementedError(f"{what} must hold 1 or C={c} values (shapes like [C,1,1]); got {tuple(v.shape)}")
    return w, w.numel()


class _ElemLoss(_StoredGradLoss):
    @staticmethod
    def forward(ctx, output, target, weight, pos_weight, cfg):
        x, origin = K.to_cuda_f32(output, cfg["what"])
        device = x.device
        if tuple(target.shape) != tuple(output.shape):
            raise ValueError(f"target shape {tuple(target.shape)} must match the logits {tuple(output.shape)}")
        t = target.detach().to(device=device, dtype=torch.float32).contiguous()
        b, c, inner = K.bci(x)
        red = _REDUCTION.get(cfg["reduction"], TB.LOSS_NONE)
        w, wn = _channel_vector(weight, c, device, "weight")
        pw, pn = _channel_vector(pos_weight, c, device, "pos_weight")
        p = N.ElemLossParams(cfg["kind"], red, float(cfg.get("smooth") or 0.0), 0 if cfg.get("smooth") is None else 1,
                             float(cfg["ignore"]) if cfg.get("ignore") is not None else 0.0,
                             0 if cfg.get("ignore") is None else 1, wn, pn, float(cfg.get("gamma", 1.0)), float(cfg.get("beta", 2.0)))
        want_grad = ctx.needs_input_grad[0]
        elementwise = red == TB.LOSS_NONE
        grad = torch.empty((b, c, inner), dtype=torch.float32, device=device) if want_grad else None
        out = torch.empty(x.shape if elementwise else (), dtype=torch.float32, device=device)
        ws = K.workspace(b, c, inner, device)
        with torch.cuda.device(device):
            rc = N.lib().tb_elem_loss_fwd_bwd(x.data_ptr(), t.data_ptr(), b, c, inner, C.byref(p), N.ptr(w), N.ptr(pw), out.data_ptr(),
                                              N.ptr(grad), ws.data_ptr(), ws.numel(), N.stream_ptr(device))
        N.check(rc, "tb_elem_loss_fwd_bwd")
        ctx.stored_grad, ctx.origin, ctx.x_shape, ctx.x_dtype = grad, origin, output.shape, output.dtype
        ctx.mode, ctx.elem_shape, ctx.n_extra = ("elementwise" if elementwise else "scalar"), tuple(x.shape), 4
        if grad is not None and elementwise:
            ctx.stored_grad = grad.reshape(x.shape)
        return out if origin.type == "cuda" else out.to(origin)


class SoftBCEWithLogitsLoss(nn.Module):
    """``nn.BCEWithLogitsLoss`` with ``ignore_index`` and label smoothing (reference soft_bce.py:9-50)."""

    __constants__ = ["weight", "pos_weight", "reduction", "ignore_index", "smooth_factor"]

    def __init__(self, weight=None, ignore_index: Optional[int] = -100, reduction="mean", smooth_factor=None, pos_weight=None):
        super().__init__()
        self.ignore_index = ignore_index
        self.reduction = reduction
        self.smooth_factor = smooth_factor
        self.register_buffer("weight", weight)
        self.register_buffer("pos_weight", pos_weight)

    def forward(self, input: Tensor, target: Tensor) -> Tensor:
        cfg = dict(kind=TB.EL_SOFT_BCE, what="SoftBCEWithLogitsLoss", reduction=self.reduction, smooth=self.smooth_factor,
                   ignore=self.ignore_index)
        return _ElemLoss.apply(input, target, self.weight, self.pos_weight, cfg)


def balanced_binary_cross_entropy_with_logits(logits: Tensor, targets: Tensor, gamma: float = 1.0,
                                              ignore_index: Optional[int] = None, reduction: str = "mean") -> Tensor:
    """Class-count balanced BCE (reference balanced_bce.py:10-49); ``targets`` are hard 0/1 labels of the logits' shape."""
    cfg = dict(kind=TB.EL_BALANCED_BCE, what="balanced_binary_cross_entropy_with_logits", reduction=reduction, gamma=gamma,
               ignore=ignore_index)
    return _ElemLoss.apply(logits, targets, None, None, cfg)


class BalancedBCEWithLogitsLoss(nn.Module):
    """Balanced binary cross-entropy, https://arxiv.org/pdf/1504.06375.pdf formula 2 (reference balanced_bce.py:51-79)."""

    __constants__ = ["gamma", "reduction", "ignore_index"]

    def __init__(self, gamma: float = 1.0, reduction="mean", ignore_index: Optional[int] = None):
        super().__init__()
        self.gamma, self.reduction, self.ignore_index = gamma, reduction, ignore_index

    def forward(self, output: Tensor, target: Tensor) -> Tensor:
        return balanced_binary_cross_entropy_with_logits(output, target, gamma=self.gamma, ignore_index=self.ignore_index,
                                                         reduction=self.reduction)


class QualityFocalLoss(nn.Module):
    """Quality Focal Loss, https://arxiv.org/abs/2006.04388 (reference quality_focal_loss.py:5-53):
    ``|sigmoid(x) - t|**beta * BCE(x, t)``; ``reduction`` mean | sum | normalized (sum of losses / sum of focal terms)."""

    __constants__ = ["beta", "reduction"]

    def __init__(self, beta: float = 2, reduction="mean"):
        super().__init__()
        self.beta, self.reduction = beta, reduction

    def forward(self, predictions: Tensor, targets: Tensor) -> Tensor:
        cfg = dict(kind=TB.EL_QFL, what="QualityFocalLoss", reduction=self.reduction, beta=self.beta)
        return _ElemLoss.apply(predictions, targets, None, None, cfg)


# --------------------------------------------------------------------------------------------------------------
# soft F1
# --------------------------------------------------------------------------------------------------------------
class _SoftF1(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, targets, mode, ignore, eps):
        x, origin = K.to_cuda_f32(logits, "soft F1 loss")
        device = x.device
        if mode == 0:
            x2 = x.reshape(-1, 1)
            t = targets.detach().to(device=device, dtype=torch.float32).reshape(-1).contiguous()
        else:
            if x.dim() != 2:
                raise ValueError("SoftF1Loss takes [N, C] logits and [N] class labels (as the reference's one_hot implies)")
            x2 = x
            t = targets.detach().to(device=device, dtype=torch.int64).reshape(-1).contiguous()
        x2 = x2.contiguous()
        n, c = x2.shape
        if t.numel() != n:
            raise ValueError(f"{t.numel()} targets for {n} samples")
        out = torch.zeros((), dtype=torch.float32, device=device)
        want_grad = ctx.needs_input_grad[0]
        grad = torch.zeros_like(x2) if want_grad else None
        if n > 0:
            lib = N.lib()
            ws = torch.empty(lib.tb_soft_f1_workspace_bytes(n, c), dtype=torch.uint8, device=device)
            with torch.cuda.device(device):
                rc = lib.tb_soft_f1_fwd_bwd(x2.data_ptr(), t.data_ptr(), n, c, mode, 0 if ignore is None else 1,
                                            0.0 if ignore is None else float(ignore), float(eps), out.data_ptr(), N.ptr(grad),
                                            ws.data_ptr(), ws.numel(), N.stream_ptr(device))
            N.check(rc, "tb_soft_f1_fwd_bwd")
        ctx.stored_grad, ctx.origin, ctx.x_shape, ctx.x_dtype = grad, origin, logits.shape, logits.dtype
        return out if origin.type == "cuda" else out.to(origin)

    @staticmethod
    def backward(ctx, upstream):
        grad = ctx.stored_grad
        if grad is None:
            raise RuntimeError("backward called twice on a fused loss; recompute the loss instead of retaining the graph")
        ctx.stored_grad = None
        grad = K.scale_stored_grad(grad, upstream).reshape(ctx.x_shape).to(ctx.x_dtype)
        if ctx.origin.type != "cuda":
            grad = grad.to(ctx.origin)
        return grad, None, None, None, None


def soft_micro_f1(preds: Tensor, targets: Tensor, eps=1e-6) -> Tensor:
    """Mean over labels of (1 - soft F1) on PROBABILITIES ``preds`` [N, L] (reference soft_f1.py:8-29). Host-side formula on
    the caller's tensors (three column sums of [N, L]): the fused kernels below fold it into the losses."""
    tp = torch.sum(preds * targets, dim=0)
    fp = torch.sum(preds * (1 - targets), dim=0)
    fn = torch.sum((1 - preds) * targets, dim=0)
    return (1 - 2 * tp / (2 * tp + fn + fp + eps)).mean()


class BinarySoftF1Loss(nn.Module):
    """1 - soft F1 of ``sigmoid(preds)`` against binary targets, any shape (reference soft_f1.py:57-77)."""

    def __init__(self, ignore_index: Optional[int] = None, eps=1e-6):
        super().__init__()
        self.ignore_index, self.eps = ignore_index, eps

    def forward(self, preds: Tensor, targets: Tensor) -> Tensor:
        return _SoftF1.apply(preds, targets, 0, self.ignore_index, self.eps)


class SoftF1Loss(nn.Module):
    """Mean over classes of 1 - soft F1 of ``softmax(preds, 1)`` [N, C] against class labels [N] (reference soft_f1.py:80-107)."""

    def __init__(self, ignore_index: Optional[int] = None, eps=1e-6):
        super().__init__()
        self.ignore_index, self.eps = ignore_index, eps

    def forward(self, preds: Tensor, targets: Tensor) -> Tensor:
        return _SoftF1.apply(preds, targets, 1 if self.ignore_index is None else 2, self.ignore_index, self.eps)
