"""DiceLoss / JaccardLoss core: drop-in for ``pytorch_toolbelt/losses/dice.py`` and ``jaccard.py``.

Two fused passes: (1) per-class statistics I = sum p*t, S = sum (p+t), T = sum t over batch and space with the
activation (sigmoid / channel softmax) and the ignore mask computed on the fly; (2) loss value and the
closed-form gradient w.r.t. the logits. Nothing of size [B,C,H,W] is materialised besides
the gradient itself.
"""
import ctypes as C
from typing import List, Optional

import torch
from torch import Tensor
from torch.nn.modules.loss import _Loss

from .. import _native as N
from .._native import TB
from . import _kernels as K

__all__ = ["DiceLoss", "region_stats"]

BINARY_MODE = "binary"
MULTICLASS_MODE = "multiclass"
MULTILABEL_MODE = "multilabel"
_MODE_CODE = {BINARY_MODE: 0, MULTICLASS_MODE: 1, MULTILABEL_MODE: 2}


def _params(jaccard, mode, from_logits, log_loss, smooth, eps, ignore_index) -> N.DiceParams:
    return N.DiceParams(
        jaccard=int(jaccard), mode=int(mode), from_logits=int(bool(from_logits)), log_loss=int(bool(log_loss)),
        smooth=float(smooth), eps=float(eps), has_ignore=int(ignore_index is not None),
        ignore_index=float(ignore_index) if ignore_index is not None else 0.0,
    )


def region_stats(x: Tensor, t: Tensor, mode: int, from_logits: bool, ignore_index) -> Tensor:
    """float64 [C,3] = (I, S, T) per class. x: [B,C,P] float32 CUDA; t: dense float32 [B,C,P] or int64 labels [B,P]."""
    b, c, inner = x.shape
    p = _params(False, mode, from_logits, False, 0.0, 1e-7, ignore_index)
    stats = torch.empty((c, 3), dtype=torch.float64, device=x.device)
    ws = K.workspace(b, c, inner, x.device)
    t_kind = TB.T_LABELS_I64 if t.dtype == torch.int64 else TB.T_DENSE_F32
    with torch.cuda.device(x.device):
        rc = N.lib().tb_dice_stats(x.data_ptr(), t.data_ptr(), t_kind, b, c, inner, C.byref(p), stats.data_ptr(),
                                   ws.data_ptr(), ws.numel(), N.stream_ptr(x.device))
    N.check(rc, "tb_dice_stats")
    return stats


class _RegionLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y_pred, y_true, cfg):
        x, origin = K.to_cuda_f32(y_pred, "DiceLoss/JaccardLoss")
        device = x.device
        mode = _MODE_CODE[cfg["mode"]]
        bs = y_true.size(0)
        if mode == 0:
            x3 = x.reshape(bs, 1, -1)
            t = y_true.detach().to(device=device, dtype=torch.float32).reshape(bs, 1, -1).contiguous()
        elif mode == 1:
            x3 = x.reshape(bs, x.size(1), -1)
            t = y_true.detach().to(device=device, dtype=torch.int64).reshape(bs, -1).contiguous()
        else:
            x3 = x.reshape(bs, x.size(1), -1)
            t = y_true.detach().to(device=device, dtype=torch.float32).reshape(bs, x.size(1), -1).contiguous()
        x3 = x3.contiguous()
        b, c, inner = x3.shape
        t_kind = TB.T_LABELS_I64 if mode == 1 else TB.T_DENSE_F32
        p = _params(cfg["jaccard"], mode, cfg["from_logits"], cfg["log_loss"], cfg["smooth"], cfg["eps"], cfg["ignore_index"])
        stats = torch.empty((c, 3), dtype=torch.float64, device=device)
        ws = K.workspace(b, c, inner, device)
        want_grad = ctx.needs_input_grad[0]
        grad = torch.empty((b, c, inner), dtype=torch.float32, device=device) if want_grad else None
        out = torch.empty((), dtype=torch.float32, device=device)
        classes = cfg["classes"]
        cls_dev = None
        if classes is not None:
            cls_dev = torch.as_tensor(list(classes), dtype=torch.int32).to(device)
            if cls_dev.numel() == 0 or int(cls_dev.min()) < 0 or int(cls_dev.max()) >= c:
                raise IndexError(f"classes {list(classes)} out of range for {c} channels")
        with torch.cuda.device(device):
            stream = N.stream_ptr(device)
            rc = N.lib().tb_dice_stats(x3.data_ptr(), t.data_ptr(), t_kind, b, c, inner, C.byref(p), stats.data_ptr(),
                                       ws.data_ptr(), ws.numel(), stream)
            N.check(rc, "tb_dice_stats")
            if cfg.get("sync"):
                # data-parallel over the batch: the per-class sums are the only thing that crosses ranks.
                # The loss value becomes the global-batch loss on every rank and the gradient kernel below, fed the
                # global sums, yields d(global loss)/d(local logits) -- no further reduction is needed.
                import torch.distributed as dist

                dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=cfg.get("group"))
            rc = N.lib().tb_dice_loss_grad(x3.data_ptr(), t.data_ptr(), t_kind, b, c, inner, C.byref(p), stats.data_ptr(),
                                           N.ptr(cls_dev), 0 if cls_dev is None else cls_dev.numel(), out.data_ptr(),
                                           N.ptr(grad), stream)
            N.check(rc, "tb_dice_loss_grad")
        ctx.stored_grad, ctx.origin, ctx.x_shape, ctx.x_dtype = grad, origin, y_pred.shape, y_pred.dtype
        ctx.grad_mul = float(cfg.get("grad_mul", 1.0))
        return out if origin.type == "cuda" else out.to(origin)

    @staticmethod
    def backward(ctx, upstream):
        grad = ctx.stored_grad
        if grad is None:
            raise RuntimeError("backward called twice on a fused loss; recompute the loss instead of retaining the graph")
        ctx.stored_grad = None
        if ctx.grad_mul != 1.0:  # data-parallel variants under DDP gradient AVERAGING (distributed/losses.py)
            upstream = upstream * ctx.grad_mul
        grad = K.scale_stored_grad(grad, upstream).reshape(ctx.x_shape).to(ctx.x_dtype)
        if ctx.origin.type != "cuda":
            grad = grad.to(ctx.origin)
        return grad, None, None


class _RegionLossModule(_Loss):
    """Shared module body of DiceLoss / JaccardLoss: holds the options, launches the two fused kernels."""

    _jaccard = False

    def __init__(self, mode, classes, log_loss, from_logits, smooth, eps, ignore_index):
        if mode not in _MODE_CODE:
            raise AssertionError(f"mode must be one of {sorted(_MODE_CODE)}, got {mode!r}")
        super().__init__()
        if classes is not None:
            assert mode != BINARY_MODE, "Masking classes is not supported with mode=binary"
            classes = [int(c) for c in (classes.tolist() if torch.is_tensor(classes) else classes)]
        self.mode, self.classes = mode, classes
        self.log_loss, self.from_logits = log_loss, from_logits
        self.smooth, self.eps, self.ignore_index = smooth, eps, ignore_index

    def _sync_cfg(self) -> dict:
        """Overridden by the data-parallel variants (distributed/losses.py) to request the statistics all-reduce."""
        return {}

    def forward(self, y_pred: Tensor, y_true: Tensor) -> Tensor:
        """``y_pred`` [N,C,*]; ``y_true`` [N,*] labels (multiclass) or [N,C,*] / [N,1,*] dense targets."""
        assert y_true.size(0) == y_pred.size(0)
        cfg = dict(jaccard=self._jaccard, mode=self.mode, classes=self.classes, log_loss=self.log_loss,
                   from_logits=self.from_logits, smooth=self.smooth, eps=self.eps, ignore_index=self.ignore_index)
        cfg.update(self._sync_cfg())
        return _RegionLoss.apply(y_pred, y_true, cfg)


class DiceLoss(_RegionLossModule):
    """Dice loss for image segmentation: binary, multiclass and multilabel (reference dice.py:18-131).

    ``mode``: 'binary' | 'multiclass' | 'multilabel'; ``classes``: optional list of channels that contribute (the
    documented intent -- the reference's own handling of this argument is broken); ``log_loss``:
    ``-log(dice)`` instead of ``1 - dice``; ``from_logits``: the input is raw logits; ``ignore_index``: label of
    ignored pixels.
    """

    def __init__(self, mode: str, classes: List[int] = None, log_loss=False, from_logits=True, smooth: float = 0.0,
                 ignore_index=None, eps=1e-7):
        super().__init__(mode, classes, log_loss, from_logits, smooth, eps, ignore_index)
