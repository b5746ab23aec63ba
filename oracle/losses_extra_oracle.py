"""CPU restatement (plain torch ops, autograd for the gradients) of the reference's remaining element-wise losses
(SURVEY 8f rank 4). TEST INFRASTRUCTURE ONLY: imported by tests/ and oracle/make_golden.py, never by the product.
Pinned by oracle/make_golden.py::golden_losses_extra against the imported reference (tests/golden/losses_extra.npz).
file:line citations are relative to /root/reference/pytorch_toolbelt/losses."""
from typing import Optional

import torch
import torch.nn.functional as TF


def soft_bce_with_logits(x, target, weight=None, ignore_index: Optional[int] = -100, reduction="mean", smooth_factor=None,
                         pos_weight=None):
    """soft_bce.py:29-50."""
    if smooth_factor is not None:
        soft = ((1 - target) * smooth_factor + target * (1 - smooth_factor)).type_as(x)
    else:
        soft = target.type_as(x)
    loss = TF.binary_cross_entropy_with_logits(x, soft, weight, pos_weight=pos_weight, reduction="none")
    if ignore_index is not None:
        loss = loss * (target != ignore_index).type_as(loss)
    if reduction == "mean":
        return loss.mean()
    if reduction == "sum":
        return loss.sum()
    return loss


def balanced_bce_with_logits(logits, targets, gamma: float = 1.0, ignore_index=None, reduction="mean"):
    """balanced_bce.py:10-49 (note: the weights are raised to gamma twice, :29 and :32-33)."""
    pos = targets.eq(1).sum()
    neg = targets.eq(0).sum()
    num = pos + neg
    pos_weight = torch.pow(neg / (num + 1e-7), gamma)
    neg_weight = 1.0 - pos_weight
    pos_term = pos_weight.pow(gamma) * targets * TF.logsigmoid(logits)
    neg_term = neg_weight.pow(gamma) * (1 - targets) * TF.logsigmoid(-logits)
    loss = -(pos_term + neg_term)
    if ignore_index is not None:
        loss = torch.masked_fill(loss, targets.eq(ignore_index), 0)
    if reduction == "mean":
        return loss.mean()
    if reduction == "sum":
        return loss.sum()
    return loss


def quality_focal_loss(predictions, targets, beta: float = 2, reduction="mean"):
    """quality_focal_loss.py:24-53."""
    predictions, targets = predictions.float(), targets.float()
    bce = TF.binary_cross_entropy_with_logits(predictions, targets, reduction="none")
    focal = TF.l1_loss(predictions.sigmoid(), targets, reduction="none").pow(beta)
    loss = focal * bce
    if reduction == "mean":
        return loss.mean()
    if reduction == "sum":
        return loss.sum()
    if reduction == "normalized":
        return loss.sum() / focal.sum()
    return loss


def soft_micro_f1(preds, targets, eps=1e-6):
    """soft_f1.py:8-29."""
    tp = torch.sum(preds * targets, dim=0)
    fp = torch.sum(preds * (1 - targets), dim=0)
    fn = torch.sum((1 - preds) * targets, dim=0)
    return (1 - 2 * tp / (2 * tp + fn + fp + eps)).mean()


def binary_soft_f1(preds, targets, ignore_index=None, eps=1e-6):
    """soft_f1.py:57-77."""
    targets, preds = targets.view(-1), preds.view(-1)
    if ignore_index is not None:
        keep = targets != ignore_index
        preds, targets = preds[keep], targets[keep]
        if targets.numel() == 0:
            return torch.tensor(0, dtype=preds.dtype, device=preds.device)
    preds = preds.sigmoid().clamp(eps, 1 - eps)
    return soft_micro_f1(preds.view(-1, 1), targets.view(-1, 1), eps=1e-6)


def soft_f1(preds, targets, ignore_index=None, eps=1e-6):
    """soft_f1.py:80-107 (with ignore_index the ONE-HOT values are compared to it and the kept elements flattened)."""
    preds = preds.softmax(dim=1).clamp(eps, 1 - eps)
    targets = TF.one_hot(targets, preds.size(1))
    if ignore_index is not None:
        keep = targets != ignore_index
        preds, targets = preds[keep], targets[keep]
        if targets.numel() == 0:
            return torch.tensor(0, dtype=preds.dtype, device=preds.device)
    return soft_micro_f1(preds, targets, eps=1e-6)
