"""BinaryFocalLoss + DiceLoss / JaccardLoss as ONE fused forward+gradient (BASELINE config 3).

New relative to the reference, where the two losses are separate modules summed by the caller
(``losses/focal.py:13-105`` + ``losses/dice.py:18-131``): two passes over (logits, target) and one gradient write
instead of two gradient buffers and an add. When the pair cannot be fused (softmax / normalised focal, multiclass
region loss, label targets, ...) the two fused single-loss kernels run one after the other -- same value.

Validated on a B200 in round 2 (tests/test_joint_labels_gpu.py, compute-sanitizer clean).
"""
import ctypes as C
from typing import Union

import torch
from torch import Tensor, nn

from .. import _native as N
from .._native import TB
from . import _kernels as K
from .dice import _MODE_CODE, BINARY_MODE, MULTICLASS_MODE, DiceLoss, _params
from .focal import BinaryFocalLoss
from .functional import _finish, _StoredGradLoss
from .jaccard import JaccardLoss

__all__ = ["BinaryFocalDiceLoss"]


class _JointFocalRegion(_StoredGradLoss):
    @staticmethod
    def forward(ctx, y_pred, y_true, class_weights, cfg):
        x, origin = K.to_cuda_f32(y_pred, "BinaryFocalDiceLoss")
        device = x.device
        bs = x.size(0)
        if cfg["mode"] == BINARY_MODE:
            x3 = x.reshape(bs, 1, -1)
        else:
            x3 = x.reshape(bs, x.size(1), -1)
        x3 = x3.contiguous()
        b, c, inner = x3.shape
        t = y_true.detach().to(device=device, dtype=torch.float32).reshape(b, c, inner).contiguous()
        fp = N.FocalParams(
            gamma=float(cfg["gamma"]),
            alpha=float(cfg["alpha"]) if cfg["alpha"] is not None else 0.0,
            has_alpha=int(cfg["alpha"] is not None),
            reduced_threshold=float(cfg["reduced_threshold"]) if cfg["reduced_threshold"] is not None else 0.0,
            has_reduced_threshold=int(cfg["reduced_threshold"] is not None),
            ignore_index=float(cfg["focal_ignore"]) if cfg["focal_ignore"] is not None else 0.0,
            has_ignore=int(cfg["focal_ignore"] is not None),
            normalized=0, eps=1e-6, reduction=K.reduction_code(cfg["reduction"]), softmax=0,
        )
        dp = _params(cfg["jaccard"], _MODE_CODE[cfg["mode"]], True, cfg["log_loss"], cfg["smooth"], cfg["eps"],
                     cfg["region_ignore"])
        classes = cfg["classes"]
        cls_dev = None
        if classes is not None:
            cls_dev = torch.as_tensor(list(classes), dtype=torch.int32).to(device)
            if cls_dev.numel() == 0 or int(cls_dev.min()) < 0 or int(cls_dev.max()) >= c:
                raise IndexError(f"classes {list(classes)} out of range for {c} channels")
        want_grad = ctx.needs_input_grad[0]
        grad = torch.empty((b, c, inner), dtype=torch.float32, device=device) if want_grad else None
        out = torch.empty(3, dtype=torch.float32, device=device)
        ws = K.workspace(b, c, inner, device)
        cw_keep, cw_ptr = K.class_weights_ptr(class_weights, c, device)
        with torch.cuda.device(device):
            rc = N.lib().tb_focal_dice_fwd_bwd(
                x3.data_ptr(), t.data_ptr(), b, c, inner, C.byref(fp), cw_ptr, C.byref(dp), N.ptr(cls_dev),
                0 if cls_dev is None else cls_dev.numel(), float(cfg["w_focal"]), float(cfg["w_region"]), out.data_ptr(),
                N.ptr(grad), ws.data_ptr(), ws.numel(), N.stream_ptr(device))
        N.check(rc, "tb_focal_dice_fwd_bwd")
        ctx.stored_grad, ctx.origin, ctx.x_shape, ctx.x_dtype = grad, origin, y_pred.shape, y_pred.dtype
        ctx.elem_shape, ctx.n_extra, ctx.mode = (b, c, inner), 3, "scalar"
        total, terms = _finish(out[0], origin), _finish(out[1:], origin)
        ctx.mark_non_differentiable(terms)  # the two terms are reported, only the total carries the gradient
        return total, terms

    @staticmethod
    def backward(ctx, upstream, _terms):
        return _StoredGradLoss.backward(ctx, upstream)


class BinaryFocalDiceLoss(nn.Module):
    """``focal_weight * focal(y_pred, y_true) + region_weight * region(y_pred, y_true)`` with ``focal`` a
    ``BinaryFocalLoss`` and ``region`` a ``DiceLoss`` or ``JaccardLoss`` of this package, fused when possible.
    ``last_terms`` holds the two unweighted terms of the last call (device tensor, detached)."""

    def __init__(self, focal: BinaryFocalLoss, region: Union[DiceLoss, JaccardLoss], focal_weight: float = 1.0,
                 region_weight: float = 1.0):
        super().__init__()
        if not isinstance(focal, BinaryFocalLoss) or not isinstance(region, (DiceLoss, JaccardLoss)):
            raise TypeError("BinaryFocalDiceLoss(focal=BinaryFocalLoss, region=DiceLoss | JaccardLoss)")
        self.focal, self.region = focal, region
        self.focal_weight, self.region_weight = float(focal_weight), float(region_weight)
        self.last_terms = None

    def fusable(self, y_pred: Tensor, y_true: Tensor) -> bool:
        f, r = self.focal, self.region
        if f.activation != "sigmoid" or f.normalized or f.reduction not in ("mean", "sum"):
            return False
        if r.mode == MULTICLASS_MODE or not r.from_logits or y_true.shape != y_pred.shape:
            return False
        if y_pred.dim() < 2 or y_pred.size(0) * (1 if r.mode == BINARY_MODE else y_pred.size(1)) > 65535:
            return False
        if r.mode == BINARY_MODE and y_pred.size(1) != 1 and f.class_weights is not None:
            return False  # the region term flattens the channels, the focal class weights need them
        return not hasattr(r, "process_group")  # data-parallel statistics (Sync*Loss) go through the two-kernel path

    def forward(self, y_pred: Tensor, y_true: Tensor) -> Tensor:
        if not self.fusable(y_pred, y_true):
            a, b = self.focal(y_pred, y_true), self.region(y_pred, y_true)
            self.last_terms = torch.stack([a.detach(), b.detach()])
            return self.focal_weight * a + self.region_weight * b
        f, r = self.focal, self.region
        cfg = dict(gamma=f.gamma, alpha=f.alpha, reduced_threshold=f.reduced_threshold, focal_ignore=f.ignore_index,
                   reduction=f.reduction, jaccard=isinstance(r, JaccardLoss), mode=r.mode, classes=r.classes,
                   log_loss=r.log_loss, smooth=r.smooth, eps=r.eps, region_ignore=getattr(r, "ignore_index", None),
                   w_focal=self.focal_weight, w_region=self.region_weight)
        total, terms = _JointFocalRegion.apply(y_pred, y_true, f.class_weights, cfg)
        self.last_terms = terms
        return total
