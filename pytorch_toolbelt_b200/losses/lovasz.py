"""Lovasz hinge / Lovasz-softmax losses (Berman et al. 2018): drop-in for ``pytorch_toolbelt/losses/lovasz.py``.

The reference evaluates these with Python loops over images and classes, a boolean-mask gather for ignored pixels, one
``torch.sort`` per (image, class) and float cumsums (lovasz.py:37-160). Here every (image, class) pair is a SEGMENT of one
kernel sequence (``tb_lovasz_fwd_bwd``, csrc/lovasz.cu): error + sort-key construction, a hand-written segmented radix
sort that carries the ground-truth bit and the pixel index as payload, the positives scan, the Jaccard increments, the
dot product, the "present classes" / per-image means and the gradient scattered back through the permutation -- forward
value and d loss / d input from one call, no host loop, no host synchronisation.

Behaviour kept from the reference: ``LovaszLoss`` takes class PROBABILITIES and applies no softmax; ``BinaryLovaszLoss``
takes logits; both run in float32 only (lovasz.py:29-31,71); a 3-D ``[B,H,W]`` input to ``LovaszLoss`` is the "sigmoid
output" case with foreground = (label == 0) (lovasz.py:129-132,145-148); only classes present in the (valid) labels are
averaged. Deviations, deliberate: positives are counted in exact integers (the reference's float32 cumsum stops counting
at 2**24); an image / batch without a single valid pixel contributes 0 (the reference returns an empty tensor there).
"""
from __future__ import annotations

from typing import Optional, Union

import torch
from torch.nn.modules.loss import _Loss

from .. import _native as N
from .._native import TB
from . import _kernels as K

__all__ = ["BinaryLovaszLoss", "LovaszLoss", "lovasz_hinge_loss", "lovasz_softmax_loss"]

_HINGE, _SOFTMAX = 0, 1


class _LovaszFused(torch.autograd.Function):
    """value + stored gradient from one tb_lovasz_fwd_bwd call; backward only scales by the upstream scalar."""

    @staticmethod
    def forward(ctx, scores: torch.Tensor, labels: torch.Tensor, mode: int, per_image: bool, ignore):
        N.require_cuda("Lovasz loss")
        if scores.dtype != torch.float32:
            raise TypeError(f"Lovasz losses run in float32 like the reference (lovasz.py:29-31,71), got {scores.dtype}")
        origin = scores.device
        x = scores.detach()
        x = (x if x.is_cuda else x.cuda()).contiguous()
        device = x.device
        if mode == _HINGE:
            if x.dim() >= 2 and labels.dim() == x.dim() - 1 and x.shape[1] == 1:
                x = x.squeeze(1)  # [B,1,H,W] logits against [B,H,W] labels
            batch, classes = (x.shape[0] if x.dim() >= 1 and x.numel() else 0), 1
            inner = x.numel() // batch if batch else 0
        else:
            if x.dim() == 3:  # "sigmoid output": one probability plane (lovasz.py:145-148)
                batch, classes = x.shape[0], 1
            else:
                batch, classes = x.shape[0], x.shape[1]
            inner = x.numel() // (batch * classes) if batch * classes else 0
        if labels.numel() != batch * inner:
            raise ValueError(f"labels with {labels.numel()} elements do not match scores of shape {tuple(scores.shape)}")
        y = labels.detach()
        y = y if y.is_cuda else y.to(device)
        if y.dtype == torch.float32:
            t_kind = TB.T_DENSE_F32
        else:
            y, t_kind = y.to(torch.int64), TB.T_LABELS_I64
        y = y.contiguous()
        out = torch.zeros((), dtype=torch.float32, device=device)
        want_grad = ctx.needs_input_grad[0]
        grad = torch.empty_like(x) if want_grad else None
        if batch * inner > 0:
            lib = N.lib()
            ws = torch.empty(lib.tb_lovasz_fused_workspace_bytes(batch, classes, inner, int(per_image)), dtype=torch.uint8,
                             device=device)
            with torch.cuda.device(device):
                rc = lib.tb_lovasz_fwd_bwd(x.data_ptr(), y.data_ptr(), t_kind, batch, classes, inner, mode, int(per_image),
                                           0 if ignore is None else 1, 0.0 if ignore is None else float(ignore),
                                           out.data_ptr(), N.ptr(grad), ws.data_ptr(), ws.numel(), N.stream_ptr(device))
            N.check(rc, "tb_lovasz_fwd_bwd")
        elif grad is not None:
            grad.zero_()
        ctx.stored_grad, ctx.origin, ctx.x_shape = grad, origin, scores.shape
        return out if origin.type == "cuda" else out.to(origin)

    @staticmethod
    def backward(ctx, upstream):
        grad = ctx.stored_grad
        if grad is None:
            raise RuntimeError("backward called twice on a fused loss; recompute the loss instead of retaining the graph")
        ctx.stored_grad = None
        grad = K.scale_stored_grad(grad, upstream).reshape(ctx.x_shape)
        if ctx.origin.type != "cuda":
            grad = grad.to(ctx.origin)
        return grad, None, None, None, None


def lovasz_hinge_loss(logits: torch.Tensor, labels: torch.Tensor, per_image: bool = False,
                      ignore_index: Optional[Union[int, float]] = None) -> torch.Tensor:
    """Binary Lovasz hinge: ``logits`` [B,H,W] (or [B,1,H,W]), ``labels`` [B,H,W] in {0,1} (lovasz.py:37-86)."""
    return _LovaszFused.apply(logits, labels, _HINGE, bool(per_image), ignore_index)


def lovasz_softmax_loss(probas: torch.Tensor, labels: torch.Tensor, per_image: bool = False,
                        ignore_index: Optional[Union[int, float]] = None) -> torch.Tensor:
    """Multi-class Lovasz-softmax over the classes present in ``labels``: ``probas`` [B,C,H,W] class probabilities (or
    [B,H,W] for a single sigmoid output), ``labels`` [B,H,W] (lovasz.py:92-160)."""
    return _LovaszFused.apply(probas, labels, _SOFTMAX, bool(per_image), ignore_index)


class BinaryLovaszLoss(_Loss):
    def __init__(self, per_image: bool = False, ignore_index: Optional[Union[int, float]] = None):
        super().__init__()
        self.ignore_index = ignore_index
        self.per_image = per_image

    def forward(self, logits, target):
        return lovasz_hinge_loss(logits, target, per_image=self.per_image, ignore_index=self.ignore_index)


class LovaszLoss(_Loss):
    def __init__(self, per_image=False, ignore=None):
        super().__init__()
        self.ignore = ignore
        self.per_image = per_image

    def forward(self, logits, target):
        return lovasz_softmax_loss(logits, target, per_image=self.per_image, ignore_index=self.ignore)
