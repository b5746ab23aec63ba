"""GPU test-time augmentation for segmentation and classification.

Drop-in for ``pytorch_toolbelt.inference.tta`` (reference ``pytorch_toolbelt/inference/tta.py``). Image
augment / de-augment functions are single fused sm_100a kernels (one launch writes all A views; one launch
inverse-permutes, stacks and reduces them) instead of the reference's 6 flip copies + cat/stack + mean.
They remain autograd-transparent ("respect gradients flow", tta.py:3-4): the backward of an augment is the
transposed kernel (de-augment with ``sum``) and vice versa.

Label-only helpers ([B,C]-sized tensors) are orchestration and stay plain PyTorch, like SURVEY section 8(a11) says.
"""
from __future__ import annotations

from functools import partial
from typing import Callable, Dict, List, Mapping, Optional, Tuple, Union

import torch
from torch import Tensor, nn

from . import functional as F
from .. import _native as N
from .._native import TB
from ..utils.support import pytorch_toolbelt_deprecated

__all__ = [
    "GeneralizedTTA",
    "MultiscaleTTA",
    "d2_image_augment",
    "d2_labels_augment",
    "d2_image_deaugment",
    "d2_labels_deaugment",
    "d4_image2label",
    "d4_image2mask",
    "d4_image_augment",
    "d4_labels_augment",
    "d4_image_deaugment",
    "d4_labels_deaugment",
    "fivecrop_image2label",
    "fivecrop_image_augment",
    "fivecrop_label_deaugment",
    "fliplr_image2label",
    "fliplr_image2mask",
    "fliplr_image_augment",
    "fliplr_labels_augment",
    "fliplr_image_deaugment",
    "fliplr_labels_deaugment",
    "flips_image_augment",
    "flips_labels_augment",
    "flips_image_deaugment",
    "flips_labels_deaugment",
    "flipud_image_augment",
    "flipud_image_deaugment",
    "flipud_labels_deaugment",
    "ms_image_augment",
    "ms_labels_augment",
    "ms_image_deaugment",
    "tencrop_image2label",
]

MaybeStrOrCallable = Optional[Union[str, Callable]]

# augment view codes per TTA group, in the reference's concatenation order (tta.py:269,284,333-341,410-420,484)
_VIEW_CODES = {
    TB.TTA_NONE: [F.CODE_ID],
    TB.TTA_FLIPLR: [F.CODE_ID, F.CODE_FLIPLR],
    TB.TTA_FLIPUD: [F.CODE_ID, F.CODE_FLIPUD],
    TB.TTA_FLIPS: [F.CODE_ID, F.CODE_FLIPLR, F.CODE_FLIPUD],
    TB.TTA_D2: [F.CODE_ID, F.CODE_FLIPLR, F.CODE_FLIPUD, F.CODE_ROT180],
    TB.TTA_D4: [
        F.CODE_ID,
        F.CODE_ROT90_CW,
        F.CODE_ROT180,
        F.CODE_ROT90_CCW,
        F.CODE_TRANSPOSE,
        F.CODE_FLIPLR,
        F.CODE_ANTITRANSPOSE,
        F.CODE_FLIPUD,
    ],
}


def split_into_chunks(input: Tensor, batch_size: int) -> Tuple[Tensor, ...]:
    if input.size(0) % batch_size != 0:
        raise RuntimeError(f"Input batch size ({input.size(0)}) must be divisible by {batch_size}.")
    return torch.chunk(input, batch_size)


def _deaugment_averaging(x: Tensor, reduction: MaybeStrOrCallable) -> Tensor:
    """Average predictions of a TTA-ed model along dim 0 ([T,B,...] -> [B,...]); reference tta.py:63-96."""
    if reduction == "mean":
        x = x.mean(dim=0)
    elif reduction == "sum":
        x = x.sum(dim=0)
    elif reduction in {"gmean", "geometric_mean"}:
        x = F.geometric_mean(x, dim=0)
    elif reduction in {"hmean", "harmonic_mean"}:
        x = F.harmonic_mean(x, dim=0)
    elif reduction in {"harmonic1p"}:
        x = F.harmonic1p_mean(x, dim=0)
    elif reduction == "logodd":
        x = F.logodd_mean(x, dim=0)
    elif reduction == "log1p":
        x = F.log1p_mean(x, dim=0)
    elif callable(reduction):
        x = reduction(x, dim=0)
    elif reduction in {None, "None", "none"}:
        pass
    else:
        raise KeyError(f"Unsupported reduction mode {reduction}")
    return x


# ---------------------------------------------------------------------------------------------------------
# fused kernels + autograd
# ---------------------------------------------------------------------------------------------------------
def _to_cuda_4d(x: Tensor, what: str) -> Tuple[Tensor, bool]:
    N.require_cuda(what)
    if x.dim() != 4:
        raise ValueError(f"{what} expects a [B,C,H,W] tensor, got shape {tuple(x.shape)}")
    on_cpu = not x.is_cuda
    xd = (x.cuda() if on_cpu else x).contiguous()
    return xd, on_cpu


def _augment_launch(x: Tensor, tta: int) -> Tensor:
    xd, on_cpu = _to_cuda_4d(x, "TTA augment")
    b, c, h, w = xd.shape
    if xd.element_size() not in (1, 2, 4, 8):
        raise TypeError(f"unsupported dtype {x.dtype}")
    views = TB.TTA_VIEWS[tta]
    out = torch.empty((views * b, c, h, w), dtype=xd.dtype, device=xd.device)
    if xd.numel() > 0:
        with torch.cuda.device(xd.device):
            rc = N.lib().tb_tta_augment(xd.data_ptr(), xd.element_size(), b, c, h, w, tta, out.data_ptr(),
                                        N.stream_ptr(xd.device))
        N.check(rc, "tb_tta_augment")
    return out.to(x.device) if on_cpu else out


def _deaugment_launch(x: Tensor, tta: int, red: int) -> Tensor:
    xd, on_cpu = _to_cuda_4d(x, "TTA de-augment")
    views = TB.TTA_VIEWS[tta]
    if xd.size(0) % views != 0:
        raise RuntimeError(f"Input batch size ({xd.size(0)}) must be divisible by {views}.")
    orig_dtype = xd.dtype
    if orig_dtype in (torch.float16, torch.bfloat16):
        xd = xd.float()  # reduce in fp32, hand back in the input dtype
    elif orig_dtype not in (torch.float32, torch.float64):
        raise TypeError(f"TTA de-augment reduces floating-point predictions; got {orig_dtype}")
    b, c, h, w = xd.size(0) // views, xd.size(1), xd.size(2), xd.size(3)
    shape = (views, b, c, h, w) if red == TB.RED_NONE else (b, c, h, w)
    out = torch.empty(shape, dtype=xd.dtype, device=xd.device)
    if xd.numel() > 0:
        with torch.cuda.device(xd.device):
            rc = N.lib().tb_tta_deaugment(xd.data_ptr(), TB.TORCH_DTYPE[xd.dtype], b, c, h, w, tta, red,
                                          out.data_ptr(), N.stream_ptr(xd.device))
        N.check(rc, "tb_tta_deaugment")
    if out.dtype != orig_dtype:
        out = out.to(orig_dtype)
    return out.to(x.device) if on_cpu else out


class _Augment(torch.autograd.Function):
    @staticmethod
    def forward(ctx, image: Tensor, tta: int):
        ctx.tta = tta
        return _augment_launch(image, tta)

    @staticmethod
    def backward(ctx, grad):
        # d/dx of concat_a P_a x  =  sum_a P_a^T g_a  = de-augment with "sum"
        return _deaugment_launch(grad, ctx.tta, TB.RED_SUM), None


class _Deaugment(torch.autograd.Function):
    @staticmethod
    def forward(ctx, image: Tensor, tta: int, red: int):
        ctx.tta, ctx.red = tta, red
        return _deaugment_launch(image, tta, red)

    @staticmethod
    def backward(ctx, grad):
        views = TB.TTA_VIEWS[ctx.tta]
        if ctx.red == TB.RED_SUM:
            return _augment_launch(grad, ctx.tta), None, None
        if ctx.red == TB.RED_MEAN:
            return _augment_launch(grad / views, ctx.tta), None, None
        # RED_NONE: out[a] = P_a^T x_a  ->  grad_x[a] = P_a grad[a]
        codes = _VIEW_CODES[ctx.tta]
        parts = [F._launch_permute(grad[a], codes[a]) for a in range(views)]
        return torch.cat(parts, dim=0), None, None


def _augment(image: Tensor, tta: int) -> Tensor:
    return _Augment.apply(image, tta)


def _deaugment(image: Tensor, tta: int, reduction: MaybeStrOrCallable) -> Tensor:
    if callable(reduction):
        return reduction(_Deaugment.apply(image, tta, TB.RED_NONE), dim=0)
    if reduction not in TB.REDUCTIONS:
        raise KeyError(f"Unsupported reduction mode {reduction}")
    red = TB.REDUCTIONS[reduction]
    linear = red in (TB.RED_NONE, TB.RED_MEAN, TB.RED_SUM)
    if not linear and image.requires_grad and torch.is_grad_enabled():
        # non-linear means under autograd: fused inverse permutation, then differentiable PyTorch reduction
        return _deaugment_averaging(_Deaugment.apply(image, tta, TB.RED_NONE), reduction)
    return _Deaugment.apply(image, tta, red)


# ---------------------------------------------------------------------------------------------------------
# image augment / de-augment (reference tta.py:257-524)
# ---------------------------------------------------------------------------------------------------------
def fliplr_image_augment(image: Tensor) -> Tensor:
    """[B,C,H,W] -> [2B,C,H,W]: original, horizontally flipped"""
    return _augment(image, TB.TTA_FLIPLR)


def flipud_image_augment(image: Tensor) -> Tensor:
    """[B,C,H,W] -> [2B,C,H,W]: original, vertically flipped"""
    return _augment(image, TB.TTA_FLIPUD)


def fliplr_image_deaugment(image: Tensor, reduction: MaybeStrOrCallable = "mean") -> Tensor:
    """[2B,C,H,W] -> [B,C,H,W] (or [2,B,C,H,W] when reduction is None)"""
    return _deaugment(image, TB.TTA_FLIPLR, reduction)


def flipud_image_deaugment(image: Tensor, reduction: MaybeStrOrCallable = "mean") -> Tensor:
    return _deaugment(image, TB.TTA_FLIPUD, reduction)


def d2_image_augment(image: Tensor) -> Tensor:
    """[B,C,H,W] -> [4B,C,H,W]: original, fliplr, flipud, rot180"""
    return _augment(image, TB.TTA_D2)


def d2_image_deaugment(image: Tensor, reduction: MaybeStrOrCallable = "mean") -> Tensor:
    return _deaugment(image, TB.TTA_D2, reduction)


def d4_image_augment(image: Tensor) -> Tensor:
    """[B,C,H,W] -> [8B,C,H,W]: the 8 elements of D4 in the reference's order (tta.py:410-420)"""
    if image.size(2) != image.size(3):
        raise ValueError(
            f"Input tensor must have number of rows equal to number of cols. Got input tensor of shape {image.size()}"
        )
    return _augment(image, TB.TTA_D4)


def d4_image_deaugment(image: Tensor, reduction: MaybeStrOrCallable = "mean") -> Tensor:
    return _deaugment(image, TB.TTA_D4, reduction)


def flips_image_augment(image: Tensor) -> Tensor:
    """[B,C,H,W] -> [3B,C,H,W]: original, fliplr, flipud"""
    return _augment(image, TB.TTA_FLIPS)


def flips_image_deaugment(image: Tensor, reduction: MaybeStrOrCallable = "mean") -> Tensor:
    return _deaugment(image, TB.TTA_FLIPS, reduction)


# ---------------------------------------------------------------------------------------------------------
# label-only helpers: [B,C]-sized, no spatial permutation -> host orchestration (reference tta.py:368-382,
# 425-439, 487-500, 527-581)
# ---------------------------------------------------------------------------------------------------------
def fliplr_labels_augment(labels: Tensor) -> Tensor:
    return torch.cat([labels] * 2, dim=0)


def flips_labels_augment(labels: Tensor) -> Tensor:
    return torch.cat([labels] * 3, dim=0)


def d2_labels_augment(labels: Tensor) -> Tensor:
    return torch.cat([labels] * 4, dim=0)


def d4_labels_augment(labels: Tensor) -> Tensor:
    return torch.cat([labels] * 8, dim=0)


def _labels_deaugment(logits: Tensor, views: int, reduction: MaybeStrOrCallable) -> Tensor:
    return _deaugment_averaging(torch.stack(split_into_chunks(logits, views)), reduction=reduction)


def d2_labels_deaugment(logits: Tensor, reduction: MaybeStrOrCallable = "mean") -> Tensor:
    return _labels_deaugment(logits, 4, reduction)


def d4_labels_deaugment(image: Tensor, reduction: MaybeStrOrCallable = "mean") -> Tensor:
    """[8B,C] -> [B,C]. Upstream quirk kept for drop-in parity: chunk 7 is stacked twice and chunk 6 dropped
    (reference tta.py:437); see d4_labels_deaugment_fixed for the intended behaviour."""
    b = split_into_chunks(image, 8)
    stacked = torch.stack([b[0], b[1], b[2], b[3], b[4], b[6], b[6], b[7]])
    return _deaugment_averaging(stacked, reduction=reduction)


def d4_labels_deaugment_fixed(image: Tensor, reduction: MaybeStrOrCallable = "mean") -> Tensor:
    return _labels_deaugment(image, 8, reduction)


def fliplr_labels_deaugment(logits: Tensor, reduction: MaybeStrOrCallable = "mean") -> Tensor:
    return _labels_deaugment(logits, 2, reduction)


def flipud_labels_deaugment(logits: Tensor, reduction: MaybeStrOrCallable = "mean") -> Tensor:
    return _labels_deaugment(logits, 2, reduction)


def flips_labels_deaugment(logits: Tensor, reduction: MaybeStrOrCallable = "mean") -> Tensor:
    if logits.size(0) % 3 != 0:
        raise RuntimeError("Batch size must be divisible by 3")
    return _labels_deaugment(logits, 3, reduction)


# ---------------------------------------------------------------------------------------------------------
# crops (classification TTA; slicing views + cat, reference tta.py:99-206)
# ---------------------------------------------------------------------------------------------------------
def _five_crops(image: Tensor, crop_size: Tuple[int, int]):
    rows, cols = int(image.size(2)), int(image.size(3))
    ch, cw = crop_size
    y1, x1 = rows - ch, cols - cw
    cy, cx = (rows - ch) // 2, (cols - cw) // 2
    return [
        image[..., :ch, :cw],
        image[..., :ch, x1:],
        image[..., y1:, :cw],
        image[..., y1:, x1:],
        image[..., cy : cy + ch, cx : cx + cw],
    ]


def fivecrop_image_augment(image: Tensor, crop_size: Tuple[int, int]) -> Tensor:
    rows, cols = int(image.size(2)), int(image.size(3))
    ch, cw = crop_size
    if ch > rows:
        raise ValueError(f"Tensor height ({rows}) is less than requested crop size ({ch})")
    if cw > cols:
        raise ValueError(f"Tensor width ({cols}) is less than requested crop size ({cw})")
    return torch.cat(_five_crops(image, crop_size), dim=0)


def fivecrop_label_deaugment(logits: Tensor, reduction: MaybeStrOrCallable = "mean") -> Tensor:
    return _labels_deaugment(logits, 5, reduction)


def fivecrop_image2label(model: nn.Module, image: Tensor, crop_size: Tuple) -> Tensor:
    return fivecrop_label_deaugment(model(fivecrop_image_augment(image, crop_size)))


def tencrop_image2label(model: nn.Module, image: Tensor, crop_size: Tuple) -> Tensor:
    rows, cols = int(image.size(2)), int(image.size(3))
    ch, cw = crop_size
    assert ch <= rows
    assert cw <= cols
    output = None
    for crop in _five_crops(image, crop_size):
        for view in (crop, F.torch_fliplr(crop)):
            y = model(view)
            output = y if output is None else output + y
    return output * float(1.0 / 10.0)


# ---------------------------------------------------------------------------------------------------------
# model wrappers (reference tta.py:209-254)
# ---------------------------------------------------------------------------------------------------------
def fliplr_image2label(model: nn.Module, image: Tensor) -> Tensor:
    return fliplr_labels_deaugment(model(fliplr_image_augment(image)))


def fliplr_image2mask(model: nn.Module, image: Tensor) -> Tensor:
    return fliplr_image_deaugment(model(fliplr_image_augment(image)))


def d4_image2label(model: nn.Module, image: Tensor) -> Tensor:
    return d4_labels_deaugment(model(d4_image_augment(image)))


def d4_image2mask(model: nn.Module, image: Tensor) -> Tensor:
    return d4_image_deaugment(model(d4_image_augment(image)))


@pytorch_toolbelt_deprecated("This class is deprecated. Please use GeneralizedTTA instead")
class TTAWrapper(nn.Module):
    def __init__(self, model: nn.Module, tta_function, **kwargs):
        super().__init__()
        self.model = model
        self.tta = partial(tta_function, **kwargs)

    def forward(self, *input):
        return self.tta(self.model, *input)


# ---------------------------------------------------------------------------------------------------------
# multi-scale (bilinear resize stays torch: out of kernel scope)
# ---------------------------------------------------------------------------------------------------------
def _offset_pair(offset) -> Tuple[int, int]:
    if isinstance(offset, (tuple, list)):
        return offset[0], offset[1]
    return offset, offset


def ms_labels_augment(labels: Tensor, size_offsets: List[Union[int, Tuple[int, int]]]) -> List[Tensor]:
    return [labels] * len(size_offsets)


def ms_image_augment(
    image: Tensor, size_offsets: List[Union[int, Tuple[int, int]]], mode="bilinear", align_corners=False
) -> List[Tensor]:
    rows, cols = image.size(2), image.size(3)
    outputs = []
    for offset in size_offsets:
        dr, dc = _offset_pair(offset)
        if dr == 0 and dc == 0:
            outputs.append(image)
        else:
            outputs.append(
                torch.nn.functional.interpolate(image, size=(rows + dr, cols + dc), mode=mode, align_corners=align_corners)
            )
    return outputs


def ms_labels_deaugment(
    logits: List[Tensor], size_offsets: List[Union[int, Tuple[int, int]]], reduction: MaybeStrOrCallable = "mean"
):
    if len(logits) != len(size_offsets):
        raise ValueError("Number of images must be equal to number of size offsets")
    return _deaugment_averaging(torch.stack(logits), reduction=reduction)


def ms_image_deaugment(
    images: List[Tensor],
    size_offsets: List[Union[int, Tuple[int, int]]],
    reduction: MaybeStrOrCallable = "mean",
    mode: str = "bilinear",
    align_corners: bool = True,
    stride: int = 1,
) -> Tensor:
    if len(images) != len(size_offsets):
        raise ValueError("Number of images must be equal to number of size offsets")
    restored = []
    for feature_map, offset in zip(images, size_offsets):
        dr, dc = _offset_pair(offset)
        if dr == 0 and dc == 0:
            restored.append(feature_map)
        else:
            rows, cols = feature_map.size(2), feature_map.size(3)
            size = rows - dr // stride, cols - dc // stride
            restored.append(torch.nn.functional.interpolate(feature_map, size=size, mode=mode, align_corners=align_corners))
    return _deaugment_averaging(torch.stack(restored), reduction=reduction)


class GeneralizedTTA(nn.Module):
    """Wrap a model with augment / de-augment callables (single tensor, list, or dict dispatch; tta.py:692-756).

    Example:
        tta_model = GeneralizedTTA(model, augment_fn=tta.d2_image_augment,
                                   deaugment_fn={"mask": tta.d2_image_deaugment, "edge": tta.d2_image_deaugment})
    """

    def __init__(
        self,
        model: Union[nn.Module, nn.DataParallel],
        augment_fn: Union[Callable, Dict[str, Callable], List[Callable]],
        deaugment_fn: Union[Callable, Dict[str, Callable], List[Callable]],
    ):
        super().__init__()
        self.model = model
        self.augment_fn = augment_fn
        self.deaugment_fn = deaugment_fn

    def forward(self, *input, **kwargs):
        if isinstance(self.augment_fn, dict):
            if len(input) != 0:
                raise ValueError(
                    "Input for GeneralizedTTA must not have positional arguments when augment_fn is dictionary"
                )
            outputs = self.model(**{key: fn(kwargs[key]) for key, fn in self.augment_fn.items()})
        elif isinstance(self.augment_fn, (list, tuple)):
            if len(kwargs) != 0:
                raise ValueError("Input for GeneralizedTTA must be exactly one tensor")
            outputs = self.model(*[fn(x) for x, fn in zip(input, self.augment_fn)])
        else:
            if len(input) != 1 or len(kwargs) != 0:
                raise ValueError("Input for GeneralizedTTA must be exactly one tensor")
            outputs = self.model(self.augment_fn(input[0]))

        if isinstance(self.deaugment_fn, dict):
            if not isinstance(outputs, dict):
                raise ValueError("Output of the model must be a dict")
            return {key: fn(outputs[key]) for key, fn in self.deaugment_fn.items()}
        if isinstance(self.deaugment_fn, (list, tuple)):
            if not isinstance(outputs, (dict, tuple)):
                raise ValueError("Output of the model must be a dict")
            return [fn(value) for value, fn in zip(outputs, self.deaugment_fn)]
        return self.deaugment_fn(outputs)


class MultiscaleTTA(nn.Module):
    def __init__(
        self,
        model: nn.Module,
        size_offsets: List[int],
        mode: str = "bilinear",
        align_corners: bool = False,
        augment_fn: Callable = ms_image_augment,
        deaugment_fn: Union[Callable, Dict[str, Callable]] = ms_image_deaugment,
    ):
        super().__init__()
        self.keys = set(deaugment_fn.keys()) if isinstance(deaugment_fn, Mapping) else None
        self.model = model
        self.size_offsets = size_offsets
        self.mode = mode
        self.align_corners = align_corners
        self.augment_fn = augment_fn
        self.deaugment_fn = deaugment_fn

    def forward(self, x):
        ms_inputs = self.augment_fn(x, size_offsets=self.size_offsets, mode=self.mode, align_corners=self.align_corners)
        ms_outputs = [self.model(scaled) for scaled in ms_inputs]
        if self.keys is None:
            return self.deaugment_fn(ms_outputs, self.size_offsets)
        return {
            key: self.deaugment_fn[key](
                [out[key] for out in ms_outputs], size_offsets=self.size_offsets, mode=self.mode,
                align_corners=self.align_corners,
            )
            for key in self.keys
        }
