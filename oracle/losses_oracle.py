"""Oracle for the segmentation losses (reference losses/functional.py, focal.py, dice.py, jaccard.py, lovasz.py,
wing_loss.py). torch-CPU formulas with autograd supplying the reference gradients. TEST INFRASTRUCTURE ONLY."""
import math
from typing import Optional, Sequence

import torch
import torch.nn.functional as TF


# ----------------------------------------------------------------------------------------------- focal
def one_hot_targets(labels: torch.Tensor, num_classes: int, ignore_index=None) -> torch.Tensor:
    """BinaryFocalLoss one-hot expansion, losses/focal.py:94-105: [B,*] long -> [B,C,*]; ignored pixels carry
    ignore_index in every channel."""
    if ignore_index is None:
        return torch.moveaxis(TF.one_hot(labels, num_classes), -1, 1)
    ignored = labels == ignore_index
    oh = torch.moveaxis(TF.one_hot(labels.masked_fill(ignored, 0), num_classes), -1, 1).clone()
    oh.masked_fill_(ignored.unsqueeze(1), ignore_index)
    return oh


def focal_sigmoid(x, t, gamma=2.0, alpha: Optional[float] = 0.25, reduction="mean", normalized=False,
                  reduced_threshold=None, eps=1e-6, ignore_index=None, activation="sigmoid", softmax_dim=None,
                  class_weights=None):
    """focal_loss_with_logits, losses/functional.py:19-107."""
    x = x.float()
    t = t.float()
    p = torch.sigmoid(x) if activation == "sigmoid" else torch.softmax(x, dim=softmax_dim)
    ce = TF.binary_cross_entropy_with_logits(x, t, reduction="none")
    pt = p * t + (1 - p) * (1 - t)
    if reduced_threshold is None:
        focal = (1.0 - pt).pow(gamma)
    else:
        focal = ((1.0 - pt) / (1 - reduced_threshold)).pow(gamma)
        focal = focal.masked_fill(pt < reduced_threshold, 1)
    loss = focal * ce
    if alpha is not None:
        loss = loss * (alpha * t + (1 - alpha) * (1 - t))
    if class_weights is not None:
        loss = loss * class_weights.view(1, -1, *([1] * (loss.dim() - 2)))
    if ignore_index is not None:
        ign = t.eq(ignore_index)
        loss = loss.masked_fill(ign, 0)
        if normalized:
            focal = focal.masked_fill(ign, 0)
    if normalized:
        loss = loss / focal.sum(dtype=torch.float32).clamp_min(eps)
    if reduction == "mean":
        return loss.mean()
    if reduction == "sum":
        return loss.sum()
    if reduction == "batchwise_mean":
        return loss.sum(dim=0)
    return loss


def focal_softmax(x, labels, class_weights=None, gamma=2.0, reduction="mean", normalized=False,
                  reduced_threshold=None, eps=1e-6, ignore_index=-100):
    """softmax_focal_loss_with_logits, losses/functional.py:110-173."""
    ign = labels.eq(ignore_index)
    oh = TF.one_hot(labels.masked_fill(ign, 0), num_classes=x.size(1)).moveaxis(-1, 1).float()
    probs = TF.softmax(x, dim=1)
    pt = (1 - oh) * probs + oh * (1 - probs)
    bce = TF.binary_cross_entropy_with_logits(x, oh, reduction="none")
    if reduced_threshold is None:
        focal = pt.pow(gamma)
    else:
        focal = (pt / reduced_threshold).pow(gamma)
        focal = torch.where(pt < reduced_threshold, torch.ones_like(focal), focal)
    loss = focal * bce
    if class_weights is not None:
        loss = loss * class_weights.reshape(1, -1, *([1] * (loss.dim() - 2)))
    loss = loss.sum(dim=1) * (~ign)
    if normalized:
        loss = loss / focal.sum().clamp_min(eps)
    if reduction == "mean":
        return loss.mean()
    if reduction == "sum":
        return loss.sum()
    if reduction == "batchwise_mean":
        return loss.sum(0)
    return loss


# ----------------------------------------------------------------------------------------------- dice / jaccard
def _region_prep(y_pred, y_true, mode, from_logits, ignore_index):
    """Shared front half of DiceLoss.forward (losses/dice.py:59-111) / JaccardLoss.forward (jaccard.py:48-80)."""
    if from_logits:
        y_pred = y_pred.log_softmax(dim=1).exp() if mode == "multiclass" else TF.logsigmoid(y_pred).exp()
    bs, C = y_true.size(0), y_pred.size(1)
    if mode == "binary":
        y_true = y_true.view(bs, 1, -1)
        y_pred = y_pred.view(bs, 1, -1)
        if ignore_index is not None:
            m = y_true != ignore_index
            y_pred, y_true = y_pred * m, y_true * m
    elif mode == "multiclass":
        y_true = y_true.view(bs, -1)
        y_pred = y_pred.view(bs, C, -1)
        if ignore_index is not None:
            m = y_true != ignore_index
            y_pred = y_pred * m.unsqueeze(1)
            y_true = TF.one_hot((y_true * m).to(torch.long), C).permute(0, 2, 1) * m.unsqueeze(1)
        else:
            y_true = TF.one_hot(y_true, C).permute(0, 2, 1)
    else:  # multilabel
        y_true = y_true.view(bs, C, -1)
        y_pred = y_pred.view(bs, C, -1)
        if ignore_index is not None:
            m = y_true != ignore_index
            y_pred, y_true = y_pred * m, y_true * m
    return y_pred, y_true


def dice_loss(y_pred, y_true, mode, log_loss=False, from_logits=True, smooth=0.0, ignore_index=None, eps=1e-7,
              classes: Optional[Sequence[int]] = None):
    """DiceLoss.forward, losses/dice.py:59-131 with soft_dice_score, losses/functional.py:221-247."""
    p, t = _region_prep(y_pred, y_true, mode, from_logits, ignore_index)
    t = t.type_as(p)
    inter = torch.sum(p * t, dim=(0, 2))
    card = torch.sum(p + t, dim=(0, 2))
    score = (2.0 * inter + smooth) / (card + smooth).clamp_min(eps)
    loss = -torch.log(score.clamp_min(eps)) if log_loss else 1.0 - score
    loss = loss * (t.sum((0, 2)) > 0).to(loss.dtype)
    if classes is not None:
        loss = loss[torch.as_tensor(list(classes), dtype=torch.long)]  # documented intent (reference is broken, A.6)
    return loss.mean()


def jaccard_loss(y_pred, y_true, mode, log_loss=False, from_logits=True, smooth=0.0, eps=1e-7,
                 classes: Optional[Sequence[int]] = None):
    """JaccardLoss.forward, losses/jaccard.py:48-103 with soft_jaccard_score, losses/functional.py:188-218."""
    p, t = _region_prep(y_pred, y_true, mode, from_logits, None)
    t = t.type(p.dtype)
    inter = torch.sum(p * t, dim=(0, 2))
    card = torch.sum(p + t, dim=(0, 2))
    score = (inter + smooth) / (card - inter + smooth).clamp_min(eps)
    loss = -torch.log(score.clamp_min(eps)) if log_loss else 1.0 - score
    loss = loss * (t.sum((0, 2)) > 0).float()
    if classes is not None:
        loss = loss[torch.as_tensor(list(classes), dtype=torch.long)]
    return loss.mean()


def soft_dice_score(output, target, smooth=0.0, eps=1e-7, dims=None):
    """losses/functional.py:221-247."""
    inter = torch.sum(output * target, dim=dims) if dims is not None else torch.sum(output * target)
    card = torch.sum(output + target, dim=dims) if dims is not None else torch.sum(output + target)
    return (2.0 * inter + smooth) / (card + smooth).clamp_min(eps)


def soft_jaccard_score(output, target, smooth=0.0, eps=1e-7, dims=None):
    """losses/functional.py:188-218."""
    inter = torch.sum(output * target, dim=dims) if dims is not None else torch.sum(output * target)
    card = torch.sum(output + target, dim=dims) if dims is not None else torch.sum(output + target)
    return (inter + smooth) / (card - inter + smooth).clamp_min(eps)


# ----------------------------------------------------------------------------------------------- wing
def wing_loss(output, target, width=5, curvature=0.5, reduction="mean"):
    """losses/functional.py:250-277."""
    d = (target - output).abs()
    c_const = width - width * math.log(1 + width / curvature)
    loss = torch.where(d < width, width * torch.log(1 + d / curvature), d - c_const)
    if reduction == "sum":
        return loss.sum()
    if reduction == "mean":
        return loss.mean()
    return loss


# ----------------------------------------------------------------------------------------------- lovasz
def lovasz_grad(gt_sorted: torch.Tensor) -> torch.Tensor:
    """losses/lovasz.py:23-34."""
    gts = gt_sorted.sum()
    inter = gts - gt_sorted.float().cumsum(0)
    union = gts + (1 - gt_sorted).float().cumsum(0)
    jac = 1.0 - inter / union
    if len(gt_sorted) > 1:
        jac = torch.cat([jac[:1], jac[1:] - jac[:-1]])
    return jac


def _hinge_flat(logits, labels):
    """losses/lovasz.py:55-72."""
    if len(labels) == 0:
        return logits.sum() * 0.0
    signs = 2.0 * labels.float() - 1.0
    errors = 1.0 - logits * signs
    errors_sorted, perm = torch.sort(errors, dim=0, descending=True)
    return torch.dot(TF.relu(errors_sorted), lovasz_grad(labels[perm]))


def lovasz_hinge(logits, labels, per_image=False, ignore_index=None):
    """BinaryLovaszLoss, losses/lovasz.py:37-52,75-86,187-194."""

    def flat(lg, lb):
        lg, lb = lg.reshape(-1), lb.reshape(-1)
        if ignore_index is None:
            return lg, lb
        keep = lb != ignore_index
        return lg[keep], lb[keep]

    if per_image:
        vals = [_hinge_flat(*flat(lg.unsqueeze(0), lb.unsqueeze(0))) for lg, lb in zip(logits, labels)]
        return sum(vals) / len(vals) if len(vals) > 1 else vals[0]
    return _hinge_flat(*flat(logits, labels))


def _softmax_flat(probas, labels):
    """losses/lovasz.py:112-140 with classes='present'."""
    if probas.numel() == 0:
        return probas * 0.0
    C = probas.size(1)
    losses = []
    for c in range(C):
        fg = (labels == c).type_as(probas)
        if fg.sum() == 0:
            continue
        pred = probas[:, 0] if C == 1 else probas[:, c]
        errors = (fg - pred).abs()
        errors_sorted, perm = torch.sort(errors, 0, descending=True)
        losses.append(torch.dot(errors_sorted, lovasz_grad(fg[perm])))
    if not losses:
        return 0
    return sum(losses) / len(losses) if len(losses) > 1 else losses[0]


def lovasz_softmax(probas, labels, per_image=False, ignore=None):
    """LovaszLoss, losses/lovasz.py:92-109,143-160,197-204 (input is used as probabilities: no softmax)."""

    def flat(pr, lb):
        if pr.dim() == 3:
            pr = pr.unsqueeze(1)
        C = pr.size(1)
        pr = torch.movedim(pr, 1, -1).contiguous().view(-1, C)
        lb = lb.reshape(-1)
        if ignore is None:
            return pr, lb
        keep = lb != ignore
        return pr[keep], lb[keep]

    if per_image:
        vals = [_softmax_flat(*flat(pr.unsqueeze(0), lb.unsqueeze(0))) for pr, lb in zip(probas, labels)]
        return sum(vals) / len(vals) if len(vals) > 1 else vals[0]
    return _softmax_flat(*flat(probas, labels))
