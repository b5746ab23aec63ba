"""Functional segmentation losses; drop-in for the hot-path subset of ``pytorch_toolbelt/losses/functional.py``:
focal_loss_with_logits (:19-107), softmax_focal_loss_with_logits (:110-173), soft_jaccard_score (:188-218),
soft_dice_score (:221-247), wing_loss (:250-277).

Each loss is one fused CUDA pass that produces the value and d(loss)/d(logits) together (tb_*_fwd_bwd in
include/tb200.h); autograd just hands the stored gradient back (scaled by the upstream gradient on the
device). label_smoothed_nll_loss / log_cosh_loss are out of scope.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from .. import _native as N
from .._native import TB
from ..utils.support import pytorch_toolbelt_deprecated
from . import _kernels as K

__all__ = [
    "focal_loss_with_logits",
    "softmax_focal_loss_with_logits",
    "sigmoid_focal_loss",
    "soft_jaccard_score",
    "soft_dice_score",
    "wing_loss",
]


def _needs_grad(*tensors) -> bool:
    return torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors)


def _finish(out: torch.Tensor, origin: torch.device) -> torch.Tensor:
    return out if origin.type == "cuda" else out.to(origin)


class _StoredGradLoss(torch.autograd.Function):
    """Shared backward: the forward kernel already produced d(loss)/dx (for the reduced scalar) or the
    per-element derivative (for reduction none / batchwise)."""

    @staticmethod
    def backward(ctx, upstream):
        grad = ctx.stored_grad
        if grad is None:
            raise RuntimeError("backward called twice on a fused loss; pass retain_graph=False or recompute the loss")
        ctx.stored_grad = None
        if ctx.mode == "scalar":
            grad = K.scale_stored_grad(grad, upstream)
        elif ctx.mode == "elementwise":
            grad = grad * upstream.to(grad.device).reshape(ctx.elem_shape)
        else:  # batchwise: loss[c,pos] = sum_b elem[b,c,pos]
            grad = grad * upstream.to(grad.device).reshape((1,) + tuple(ctx.elem_shape[1:]))
        grad = grad.reshape(ctx.x_shape).to(ctx.x_dtype)
        if ctx.origin.type != "cuda":
            grad = grad.to(ctx.origin)
        return (grad,) + (None,) * ctx.n_extra


# --------------------------------------------------------------------------------------------------------------
# sigmoid focal
# --------------------------------------------------------------------------------------------------------------
class _FocalLoss(_StoredGradLoss):
    @staticmethod
    def forward(ctx, output, target, class_weights, cfg):
        x, origin = K.to_cuda_f32(output, "focal_loss_with_logits")
        device = x.device
        b, c, inner = K.bci(x)
        labels = cfg["labels"]
        if labels:
            t = target.detach().to(device=device, dtype=torch.int64).contiguous()
            t_kind = TB.T_LABELS_I64
        else:
            t = target.detach().to(device=device, dtype=torch.float32).contiguous()
            t_kind = TB.T_DENSE_F32
        red = K.reduction_code(cfg["reduction"])
        want_grad = ctx.needs_input_grad[0]  # grad mode is off inside Function.forward: ask the ctx
        if cfg["normalized"] and red in (TB.LOSS_NONE, TB.LOSS_BATCHWISE) and want_grad:
            raise NotImplementedError("normalized focal loss with reduction none/batchwise_mean has no fused backward")
        p = N.FocalParams(
            gamma=float(cfg["gamma"]),
            alpha=float(cfg["alpha"]) if cfg["alpha"] is not None else 0.0,
            has_alpha=int(cfg["alpha"] is not None),
            reduced_threshold=float(cfg["reduced_threshold"]) if cfg["reduced_threshold"] is not None else 0.0,
            has_reduced_threshold=int(cfg["reduced_threshold"] is not None),
            ignore_index=float(cfg["ignore_index"]) if cfg["ignore_index"] is not None else 0.0,
            has_ignore=int(cfg["ignore_index"] is not None),
            normalized=int(bool(cfg["normalized"])),
            eps=float(cfg["eps"]),
            reduction=red,
            softmax=int(cfg["softmax"]),
        )
        elem_shape = (b, c, inner)
        if red == TB.LOSS_NONE:
            out = torch.empty(elem_shape, dtype=torch.float32, device=device)
        elif red == TB.LOSS_BATCHWISE:
            out = torch.empty((b + 1) * c * inner, dtype=torch.float32, device=device)
        else:
            out = torch.empty((), dtype=torch.float32, device=device)
        grad = torch.empty(elem_shape, dtype=torch.float32, device=device) if want_grad else None
        ws = K.workspace(b, c, inner, device)
        cw_keep, cw_ptr = K.class_weights_ptr(class_weights, c, device)
        with torch.cuda.device(device):
            rc = N.lib().tb_focal_fwd_bwd(x.data_ptr(), t.data_ptr(), t_kind, b, c, inner, C.byref(p), cw_ptr,
                                          out.data_ptr(), N.ptr(grad), ws.data_ptr(), ws.numel(), N.stream_ptr(device))
        N.check(rc, "tb_focal_fwd_bwd")
        ctx.stored_grad, ctx.origin, ctx.x_shape, ctx.x_dtype = grad, origin, output.shape, output.dtype
        ctx.elem_shape, ctx.n_extra = elem_shape, 3
        ctx.mode = "scalar" if red in (TB.LOSS_MEAN, TB.LOSS_SUM) else ("elementwise" if red == TB.LOSS_NONE else "batchwise")
        if red == TB.LOSS_NONE:
            out = out.reshape(output.shape)
        elif red == TB.LOSS_BATCHWISE:
            out = out[: c * inner].reshape(output.shape[1:])
        return _finish(out, origin)


def focal_loss_with_logits(
    output: torch.Tensor,
    target: torch.Tensor,
    gamma: float = 2.0,
    alpha: Optional[float] = 0.25,
    reduction: str = "mean",
    normalized: bool = False,
    reduced_threshold: Optional[float] = None,
    eps: float = 1e-6,
    ignore_index=None,
    activation: str = "sigmoid",
    softmax_dim: Optional[int] = None,
    class_weights: Optional[torch.Tensor] = None,
) -> torch.Tensor:
    """Binary focal loss between target and output logits (same arguments as the reference, functional.py:20-33).

    ``target`` has the shape of ``output``. ``activation="softmax"`` needs ``softmax_dim=1``.
    """
    if activation != "sigmoid" and softmax_dim not in (1, 1 - output.dim()):
        raise NotImplementedError("activation='softmax' is implemented for softmax_dim=1 (the channel dim)")
    if target.shape != output.shape:
        raise ValueError(f"target shape {tuple(target.shape)} must match output shape {tuple(output.shape)}")
    cfg = dict(gamma=gamma, alpha=alpha, reduction=reduction, normalized=normalized, reduced_threshold=reduced_threshold,
               eps=eps, ignore_index=ignore_index, softmax=activation != "sigmoid", labels=False)
    return _FocalLoss.apply(output, target, class_weights, cfg)


def _focal_from_labels(output, labels, class_weights, **kw):
    """BinaryFocalLoss with [B,*] long targets: the one-hot expansion (focal.py:88-105) happens inside the kernel."""
    cfg = dict(kw)
    cfg["labels"] = True
    cfg["softmax"] = kw["activation"] != "sigmoid"
    cfg.pop("activation")
    return _FocalLoss.apply(output, labels, class_weights, cfg)


# --------------------------------------------------------------------------------------------------------------
# softmax focal
# --------------------------------------------------------------------------------------------------------------
class _SoftmaxFocalLoss(_StoredGradLoss):
    @staticmethod
    def forward(ctx, output, target, class_weights, cfg):
        x, origin = K.to_cuda_f32(output, "softmax_focal_loss_with_logits")
        device = x.device
        b, c, inner = K.bci(x)
        t = target.detach().to(device=device, dtype=torch.int64).contiguous()
        if t.numel() != b * inner:
            raise ValueError(f"target must have shape [B,*] matching output [B,C,*]; got {tuple(target.shape)}")
        red = K.reduction_code(cfg["reduction"])
        want_grad = ctx.needs_input_grad[0]  # grad mode is off inside Function.forward: ask the ctx
        if cfg["normalized"] and red in (TB.LOSS_NONE, TB.LOSS_BATCHWISE) and want_grad:
            raise NotImplementedError("normalized focal loss with reduction none/batchwise_mean has no fused backward")
        if red == TB.LOSS_NONE:
            out = torch.empty((b, inner), dtype=torch.float32, device=device)
        elif red == TB.LOSS_BATCHWISE:
            out = torch.empty((b + 1) * inner, dtype=torch.float32, device=device)
        else:
            out = torch.empty((), dtype=torch.float32, device=device)
        grad = torch.empty((b, c, inner), dtype=torch.float32, device=device) if want_grad else None
        ws = K.workspace(b, c, inner, device)
        cw_keep, cw_ptr = K.class_weights_ptr(class_weights, c, device)
        rt = cfg["reduced_threshold"]
        with torch.cuda.device(device):
            rc = N.lib().tb_softmax_focal_fwd_bwd(
                x.data_ptr(), t.data_ptr(), b, c, inner, float(cfg["gamma"]), float(rt) if rt is not None else 0.0,
                int(rt is not None), int(cfg["ignore_index"]), int(bool(cfg["normalized"])), float(cfg["eps"]), red, cw_ptr,
                out.data_ptr(), N.ptr(grad), ws.data_ptr(), ws.numel(), N.stream_ptr(device))
        N.check(rc, "tb_softmax_focal_fwd_bwd")
        ctx.origin, ctx.x_shape, ctx.x_dtype, ctx.n_extra = origin, output.shape, output.dtype, 3
        if red in (TB.LOSS_MEAN, TB.LOSS_SUM):
            ctx.mode, ctx.stored_grad, ctx.elem_shape = "scalar", grad, (b, c, inner)
        else:
            # per-position upstream gradients broadcast over the channel dim
            ctx.mode = "elementwise" if red == TB.LOSS_NONE else "batchwise"
            ctx.stored_grad, ctx.elem_shape = grad, (b, 1, inner)
        if red == TB.LOSS_NONE:
            out = out.reshape(target.shape)
        elif red == TB.LOSS_BATCHWISE:
            out = out[:inner].reshape(target.shape[1:])
        return _finish(out, origin)


def softmax_focal_loss_with_logits(
    output: torch.Tensor,
    target: torch.Tensor,
    class_weights: Optional[torch.Tensor] = None,
    gamma: float = 2.0,
    reduction: str = "mean",
    normalized: bool = False,
    reduced_threshold: Optional[float] = None,
    eps: float = 1e-6,
    ignore_index: int = -100,
) -> torch.Tensor:
    """Softmax focal loss: output [B,C,*] logits, target [B,*] class indices (functional.py:110-173)."""
    cfg = dict(gamma=gamma, reduction=reduction, normalized=normalized, reduced_threshold=reduced_threshold, eps=eps,
               ignore_index=ignore_index)
    return _SoftmaxFocalLoss.apply(output, target, class_weights, cfg)


@pytorch_toolbelt_deprecated("Function sigmoid_focal_loss is deprecated. Please use focal_loss_with_logits instead.")
def sigmoid_focal_loss(*input, **kwargs):
    return focal_loss_with_logits(*input, **kwargs)


# --------------------------------------------------------------------------------------------------------------
# dice / jaccard scores on probabilities (metric-style helpers: tiny outputs, plain reductions)
# --------------------------------------------------------------------------------------------------------------
def soft_jaccard_score(output: torch.Tensor, target: torch.Tensor, smooth: float = 0.0, eps: float = 1e-7, dims=None):
    """(I + smooth) / max(S - I + smooth, eps) with I = sum(output*target), S = sum(output+target) over ``dims``."""
    assert output.size() == target.size()
    intersection, cardinality = _region_sums(output, target, dims)
    return (intersection + smooth) / (cardinality - intersection + smooth).clamp_min(eps)


def soft_dice_score(output: torch.Tensor, target: torch.Tensor, smooth: float = 0.0, eps: float = 1e-7, dims=None):
    """(2 I + smooth) / max(S + smooth, eps)."""
    assert output.size() == target.size()
    intersection, cardinality = _region_sums(output, target, dims)
    return (2.0 * intersection + smooth) / (cardinality + smooth).clamp_min(eps)


def _region_sums(output, target, dims):
    """I and S. The per-class case dims=(0,2) on [B,C,P] float32 CUDA tensors without autograd runs through the
    fused statistics kernel; every other call pattern (arbitrary dims, autograd through probabilities, CPU metric
    evaluation on a handful of numbers as in the reference's tests) is plain tensor algebra on tiny outputs."""
    fused = (
        output.is_cuda and output.dim() == 3 and dims is not None and tuple(dims) == (0, 2)
        and output.dtype == torch.float32 and not _needs_grad(output, target)
    )
    if fused:
        from .dice import region_stats

        stats = region_stats(output, target.to(torch.float32), mode=2, from_logits=False, ignore_index=None)
        return stats[:, 0].to(output.dtype), stats[:, 1].to(output.dtype)
    if dims is not None:
        return torch.sum(output * target, dim=dims), torch.sum(output + target, dim=dims)
    return torch.sum(output * target), torch.sum(output + target)


# --------------------------------------------------------------------------------------------------------------
# wing
# --------------------------------------------------------------------------------------------------------------
class _WingLoss(_StoredGradLoss):
    @staticmethod
    def forward(ctx, output, target, width, curvature, reduction):
        x, origin = K.to_cuda_f32(output, "wing_loss")
        device = x.device
        t = target.detach().to(device=device, dtype=torch.float32).contiguous()
        if t.shape != x.shape:
            t = t.expand_as(x).contiguous()
        red = K.reduction_code(reduction) if reduction in ("mean", "sum") else TB.LOSS_NONE
        want_grad = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        n = x.numel()
        out = torch.empty(x.shape if red == TB.LOSS_NONE else (), dtype=torch.float32, device=device)
        grad = torch.empty(x.shape, dtype=torch.float32, device=device) if want_grad else None
        ws = K.workspace(1, 1, n, device)
        with torch.cuda.device(device):
            rc = N.lib().tb_wing_fwd_bwd(x.data_ptr(), t.data_ptr(), n, float(width), float(curvature), red, out.data_ptr(),
                                         N.ptr(grad), ws.data_ptr(), ws.numel(), N.stream_ptr(device))
        N.check(rc, "tb_wing_fwd_bwd")
        ctx.stored_grad, ctx.origin, ctx.x_shape, ctx.x_dtype = grad, origin, output.shape, output.dtype
        ctx.elem_shape, ctx.n_extra = tuple(x.shape), 0
        ctx.mode = "scalar" if red != TB.LOSS_NONE else "elementwise"
        ctx.target_needs_grad = target.requires_grad
        return _finish(out, origin)

    @staticmethod
    def backward(ctx, upstream):
        (gx,) = _StoredGradLoss.backward(ctx, upstream)
        # loss depends on (target - output): d/dtarget = -d/doutput
        return gx, (-gx if ctx.target_needs_grad else None), None, None, None


def wing_loss(output: torch.Tensor, target: torch.Tensor, width=5, curvature=0.5, reduction="mean"):
    """Wing loss, https://arxiv.org/pdf/1711.06753.pdf (functional.py:250-277)."""
    return _WingLoss.apply(output, target, width, curvature, reduction)
