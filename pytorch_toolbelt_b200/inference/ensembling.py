"""Model wrappers of ``pytorch_toolbelt/inference/ensembling.py``: ``ApplySoftmaxTo``, ``ApplySigmoidTo``, ``Ensembler``,
``PickModelOutput``, ``SelectByIndex``.

Same constructors and call conventions as the reference (ensembling.py:12-162); what happens to the tensors is different:

* the activations are one in-place kernel (``tb_activation_inplace``) instead of ``mul`` + ``sigmoid_`` / ``softmax``;
* ``Ensembler`` reduces the M model outputs with ``tb_stack_reduce``: M reads + 1 write of the output size, where the
  reference stacks them into a new ``[M, ...]`` tensor and then reduces it (``torch.stack`` + ``_deaugment_averaging``,
  ensembling.py:108-118); the reductions (mean, sum, gmean, hmean, harmonic1p, logodd, log1p) and their sequential
  summation order are those of ``tta._deaugment_averaging``;
* when the wrapped model feeds a ``TiledInferenceEngine``, prefer ``TiledInferenceEngine(activation="sigmoid", temperature=T)``:
  the sigmoid then runs inside the merge kernel, before the TTA mean, with no pass of its own.

Inference-time helpers: the fused ops do not record autograd history (the reference's do; train-time ensembling is not part
of the path). ``average_checkpoints`` (file I/O) is out of scope.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Iterable, List, Optional, Tuple, Union

import torch
from torch import Tensor, nn

from .. import _native as N
from .._native import TB

__all__ = ["ApplySoftmaxTo", "ApplySigmoidTo", "Ensembler", "PickModelOutput", "SelectByIndex", "stack_reduce"]


def _activate_(x: Tensor, mode: int, temperature: float, dim: int = 1) -> Tensor:
    """sigmoid(x * T) / softmax(x * T, dim) computed by one kernel; returns a float32 tensor on x's device."""
    N.require_cuda("activation")
    origin = x.device
    xd = x.detach()
    xd = (xd if xd.is_cuda else xd.cuda()).to(torch.float32)
    if mode == 1 and dim % xd.dim() != 1:
        xd = xd.movedim(dim, 1)
    xd = xd.contiguous()
    if xd is x or xd.data_ptr() == x.data_ptr():
        xd = xd.clone()  # never mutate the model's own output tensor behind its back
    if xd.numel():
        b = xd.shape[0] if xd.dim() >= 1 else 1
        c = xd.shape[1] if xd.dim() >= 2 else 1
        inner = xd.numel() // (b * c)
        with torch.cuda.device(xd.device):
            rc = N.lib().tb_activation_inplace(xd.data_ptr(), b, c, inner, mode, float(temperature), N.stream_ptr(xd.device))
        N.check(rc, "tb_activation_inplace")
    if mode == 1 and dim % xd.dim() != 1:
        xd = xd.movedim(1, dim)
    return xd if origin.type == "cuda" else xd.to(origin)


def stack_reduce(tensors: List[Tensor], reduction: Optional[str] = "mean") -> Tensor:
    """``_deaugment_averaging(torch.stack(tensors), reduction)`` without materialising the stack (reference ensembling.py:108-118)."""
    if reduction in (None, "None", "none"):
        return torch.stack(tensors)
    if callable(reduction):
        return reduction(torch.stack(tensors), dim=0)
    if reduction not in TB.REDUCTIONS:
        raise KeyError(f"Unsupported reduction mode {reduction}")
    N.require_cuda("Ensembler")
    origin = tensors[0].device
    shape = tensors[0].shape
    if any(t.shape != shape for t in tensors):
        raise RuntimeError("stack expects each tensor to be equal size")
    dev_tensors = [(t.detach() if t.is_cuda else t.detach().cuda()).to(torch.float32).contiguous() for t in tensors]
    out = torch.empty(shape, dtype=torch.float32, device=dev_tensors[0].device)
    ptrs = (C.c_void_p * len(dev_tensors))(*[t.data_ptr() for t in dev_tensors])
    with torch.cuda.device(out.device):
        rc = N.lib().tb_stack_reduce(ptrs, len(dev_tensors), out.numel(), TB.REDUCTIONS[reduction], out.data_ptr(), N.stream_ptr(out.device))
    N.check(rc, "tb_stack_reduce")
    return out if origin.type == "cuda" else out.to(origin)


def _keys(output_key) -> Tuple:
    return (output_key,) if isinstance(output_key, (str, int)) else tuple(set(output_key))


class ApplySoftmaxTo(nn.Module):
    output_keys: Tuple
    temperature: float
    dim: int

    def __init__(self, model: nn.Module, output_key: Union[str, int, Iterable[str]] = "logits", dim: int = 1, temperature: float = 1):
        """Softmax (with temperature) on the given output(s) of ``model`` (reference ensembling.py:12-43)."""
        super().__init__()
        self.output_keys = _keys(output_key)
        self.model, self.dim, self.temperature = model, dim, temperature

    def forward(self, *input, **kwargs):
        output = self.model(*input, **kwargs)
        for key in self.output_keys:
            output[key] = _activate_(output[key], 1, self.temperature, self.dim)
        return output


class ApplySigmoidTo(nn.Module):
    output_keys: Tuple
    temperature: float

    def __init__(self, model: nn.Module, output_key: Union[str, int, Iterable[str]] = "logits", temperature=1):
        """Sigmoid (with temperature) on the given output(s) of ``model`` (reference ensembling.py:46-68)."""
        super().__init__()
        self.output_keys = _keys(output_key)
        self.model, self.temperature = model, temperature

    def forward(self, *input, **kwargs):
        output = self.model(*input, **kwargs)
        for key in self.output_keys:
            output[key] = _activate_(output[key], 0, self.temperature)
        return output


class Ensembler(nn.Module):
    """Sum / average (or any ``_deaugment_averaging`` reduction) of the outputs of several models (reference ensembling.py:71-123)."""

    def __init__(self, models: List[nn.Module], reduction: str = "mean", outputs: Optional[Iterable[str]] = None):
        super().__init__()
        self.return_some_outputs = outputs is not None
        self.outputs = tuple(outputs) if outputs else tuple()
        self.models = nn.ModuleList(models)
        self.reduction = reduction

    def forward(self, *input, **kwargs):
        outputs = [model(*input, **kwargs) for model in self.models]
        output_is_dict = isinstance(outputs[0], dict)
        output_is_list = isinstance(outputs[0], (list, tuple))
        if self.return_some_outputs:
            keys = self.outputs
        elif output_is_dict:
            keys = outputs[0].keys()
        elif output_is_list:
            keys = list(range(len(outputs[0])))
        elif torch.is_tensor(outputs[0]):
            keys = None
        else:
            raise RuntimeError()
        if keys is None:
            return stack_reduce(outputs, self.reduction)
        averaged = {} if output_is_dict else []
        for key in keys:
            value = stack_reduce([output[key] for output in outputs], self.reduction)
            if output_is_dict:
                averaged[key] = value
            else:
                averaged.append(value)
        return averaged


class PickModelOutput(nn.Module):
    """Wraps a model that returns a dict or list and returns only one element (reference ensembling.py:125-144)."""

    def __init__(self, model: nn.Module, key: Union[str, int]):
        super().__init__()
        self.model, self.target_key = model, key

    def forward(self, *input, **kwargs) -> Tensor:
        return self.model(*input, **kwargs)[self.target_key]


class SelectByIndex(nn.Module):
    """Selects a single tensor from a dict or list of output tensors (reference ensembling.py:147-162)."""

    def __init__(self, key: Union[str, int]):
        super().__init__()
        self.target_key = key

    def forward(self, outputs: Dict[str, Tensor]) -> Tensor:
        return outputs[self.target_key]
