"""Lovasz hinge / Lovasz-softmax: drop-in for ``pytorch_toolbelt/losses/lovasz.py`` (Berman et al. 2018).

Sort-bound. The descending key sort is ``torch.sort`` (a library radix sort, like cuBLAS for a GEMM); everything
after it -- positives count scan, Jaccard increments (lovasz.py:23-34), the dot product and the gradient w.r.t. the
sorted errors -- is one fused CUDA pipeline (tb_lovasz_sorted), and the gradient is scattered back through the
permutation. Counts are carried as exact integers (the reference's float32 cumsum loses integers above 2^24).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Union

import torch
from torch.nn.modules.loss import _Loss

from .. import _native as N

__all__ = ["BinaryLovaszLoss", "LovaszLoss"]


class _LovaszSortedDot(torch.autograd.Function):
    """loss = dot(relu?(errors_sorted), lovasz_grad(gt_sorted)) for one flat segment."""

    @staticmethod
    def forward(ctx, errors: torch.Tensor, fg: torch.Tensor, relu: bool):
        N.require_cuda("Lovasz loss")
        if errors.dtype != torch.float32:
            raise TypeError("Lovasz losses run in float32 (like the reference, lovasz.py:29-31,71)")
        err_sorted, perm = torch.sort(errors.detach(), dim=0, descending=True)
        gt_sorted = fg.detach().to(torch.float32)[perm].contiguous()
        n = err_sorted.numel()
        device = errors.device
        loss = torch.empty(1, dtype=torch.float32, device=device)
        want_grad = ctx.needs_input_grad[0]
        grad_sorted = torch.empty(n, dtype=torch.float32, device=device) if want_grad else None
        ws = torch.empty(N.lib().tb_lovasz_workspace_bytes(n, 1), dtype=torch.uint8, device=device)
        offsets = (C.c_int64 * 2)(0, n)
        with torch.cuda.device(device):
            rc = N.lib().tb_lovasz_sorted(err_sorted.data_ptr(), gt_sorted.data_ptr(), offsets, 1, int(relu), loss.data_ptr(),
                                          N.ptr(grad_sorted), ws.data_ptr(), ws.numel(), N.stream_ptr(device))
        N.check(rc, "tb_lovasz_sorted")
        ctx.perm, ctx.grad_sorted, ctx.n = perm, grad_sorted, n
        return loss[0]

    @staticmethod
    def backward(ctx, upstream):
        grad = torch.empty(ctx.n, dtype=torch.float32, device=ctx.grad_sorted.device)
        grad[ctx.perm] = ctx.grad_sorted * upstream
        return grad, None, None


def _lovasz_grad(gt_sorted):
    """Gradient of the Lovasz extension w.r.t. sorted errors (Alg. 1 of the paper); host-visible helper kept for
    API parity with the reference (lovasz.py:23-34); the losses below use the fused kernel instead."""
    p = len(gt_sorted)
    gts = gt_sorted.sum()
    intersection = gts - gt_sorted.float().cumsum(0)
    union = gts + (1 - gt_sorted).float().cumsum(0)
    jaccard = 1.0 - intersection / union
    if p > 1:
        jaccard = torch.cat([jaccard[:1], jaccard[1:] - jaccard[:-1]])
    return jaccard


def _mean(values, empty=0):
    values = list(values)
    if not values:
        return empty
    acc = values[0]
    for v in values[1:]:
        acc = acc + v
    return acc if len(values) == 1 else acc / len(values)


def _as_cuda(t: torch.Tensor) -> torch.Tensor:
    N.require_cuda("Lovasz loss")
    return t if t.is_cuda else t.cuda()


def _lovasz_hinge_flat(logits, labels):
    """logits [P] float, labels [P] {0,1} (lovasz.py:55-72)."""
    if len(labels) == 0:
        return logits.sum() * 0.0
    signs = 2.0 * labels.float() - 1.0
    errors = 1.0 - logits * signs
    return _LovaszSortedDot.apply(errors, labels, True)


def _flatten_binary_scores(scores, labels, ignore_index=None):
    scores = scores.reshape(-1)
    labels = labels.reshape(-1)
    if ignore_index is None:
        return scores, labels
    valid = labels != ignore_index
    return scores[valid], labels[valid]


def _lovasz_hinge(logits, labels, per_image=True, ignore_index=None):
    """Binary Lovasz hinge loss: logits [B,H,W], labels [B,H,W] in {0,1} (lovasz.py:37-52)."""
    origin = logits.device
    logits, labels = _as_cuda(logits), _as_cuda(labels)
    if per_image:
        loss = _mean(
            _lovasz_hinge_flat(*_flatten_binary_scores(lg.unsqueeze(0), lb.unsqueeze(0), ignore_index))
            for lg, lb in zip(logits, labels)
        )
    else:
        loss = _lovasz_hinge_flat(*_flatten_binary_scores(logits, labels, ignore_index))
    return loss if origin.type == "cuda" or not torch.is_tensor(loss) else loss.to(origin)


def _lovasz_softmax_flat(probas, labels, classes="present"):
    """probas [P,C], labels [P] (lovasz.py:112-140)."""
    if probas.numel() == 0:
        return probas * 0.0
    num_classes = probas.size(1)
    losses = []
    class_to_sum = list(range(num_classes)) if classes in ["all", "present"] else classes
    for c in class_to_sum:
        fg = (labels == c).type_as(probas)
        if classes == "present" and fg.sum() == 0:
            continue
        if num_classes == 1:
            if len(classes) > 1:
                raise ValueError("Sigmoid output possible only with 1 class")
            class_pred = probas[:, 0]
        else:
            class_pred = probas[:, c]
        errors = (fg - class_pred).abs()
        losses.append(_LovaszSortedDot.apply(errors, fg, False))
    return _mean(losses)


def _flatten_probas(probas, labels, ignore=None):
    if probas.dim() == 3:
        probas = probas.unsqueeze(1)
    num_classes = probas.size(1)
    probas = torch.movedim(probas, 1, -1).contiguous().view(-1, num_classes)
    labels = labels.reshape(-1)
    if ignore is None:
        return probas, labels
    valid = labels != ignore
    return probas[valid], labels[valid]


def _lovasz_softmax(probas, labels, classes="present", per_image=False, ignore_index=None):
    """Multi-class Lovasz-Softmax: probas [B,C,H,W] class PROBABILITIES (no softmax is applied, as upstream),
    labels [B,H,W] (lovasz.py:92-109)."""
    origin = probas.device
    probas, labels = _as_cuda(probas), _as_cuda(labels)
    if per_image:
        loss = _mean(
            _lovasz_softmax_flat(*_flatten_probas(prob.unsqueeze(0), lab.unsqueeze(0), ignore_index), classes=classes)
            for prob, lab in zip(probas, labels)
        )
    else:
        loss = _lovasz_softmax_flat(*_flatten_probas(probas, labels, ignore_index), classes=classes)
    return loss if origin.type == "cuda" or not torch.is_tensor(loss) else loss.to(origin)


class BinaryLovaszLoss(_Loss):
    def __init__(self, per_image: bool = False, ignore_index: Optional[Union[int, float]] = None):
        super().__init__()
        self.ignore_index = ignore_index
        self.per_image = per_image

    def forward(self, logits, target):
        return _lovasz_hinge(logits, target, per_image=self.per_image, ignore_index=self.ignore_index)


class LovaszLoss(_Loss):
    def __init__(self, per_image=False, ignore=None):
        super().__init__()
        self.ignore = ignore
        self.per_image = per_image

    def forward(self, logits, target):
        return _lovasz_softmax(logits, target, per_image=self.per_image, ignore_index=self.ignore)
