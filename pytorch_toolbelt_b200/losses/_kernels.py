"""Thin launch helpers shared by the loss modules: workspace handling, dtype/device marshalling, the
"scale stored gradient by the upstream scalar" step of backward. No arithmetic happens here."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch

from .. import _native as N
from .._native import TB

_REDUCTION = {"none": TB.LOSS_NONE, None: TB.LOSS_NONE, "mean": TB.LOSS_MEAN, "sum": TB.LOSS_SUM,
              "batchwise_mean": TB.LOSS_BATCHWISE}


def reduction_code(reduction) -> int:
    # the reference silently returns the unreduced tensor for unknown strings (functional.py:100-107)
    return _REDUCTION.get(reduction, TB.LOSS_NONE)


def to_cuda_f32(x: torch.Tensor, what: str) -> Tuple[torch.Tensor, torch.device]:
    """Logits for the kernels: contiguous float32 on the GPU (the reference also forces .float(), functional.py:58)."""
    N.require_cuda(what)
    origin = x.device
    xd = x.detach()
    if not xd.is_cuda:
        xd = xd.cuda()
    return xd.to(torch.float32).contiguous(), origin


def bci(x: torch.Tensor) -> Tuple[int, int, int]:
    """View an arbitrary-rank tensor as [B, C, inner]."""
    if x.dim() >= 2:
        b, c = x.shape[0], x.shape[1]
        inner = 1
        for s in x.shape[2:]:
            inner *= s
        return b, c, inner
    return 1, 1, x.numel()


def workspace(b: int, c: int, inner: int, device) -> torch.Tensor:
    n = N.lib().tb_loss_workspace_bytes(b, c, inner)
    return torch.empty(n, dtype=torch.uint8, device=device)


def scale_stored_grad(grad: torch.Tensor, upstream: torch.Tensor) -> torch.Tensor:
    """grad *= upstream (0-dim device tensor) in place, skipped on the device when upstream == 1."""
    up = upstream.detach().to(device=grad.device, dtype=torch.float32).reshape(1).contiguous()
    with torch.cuda.device(grad.device):
        rc = N.lib().tb_scale_by_device_scalar(grad.data_ptr(), grad.numel(), up.data_ptr(), N.stream_ptr(grad.device))
    N.check(rc, "tb_scale_by_device_scalar")
    return grad


def class_weights_ptr(cw: Optional[torch.Tensor], c: int, device):
    if cw is None:
        return None, None
    w = cw.detach().to(device=device, dtype=torch.float32).contiguous()
    if w.numel() != c:
        raise ValueError(f"class_weights has {w.numel()} entries, expected {c}")
    return w, w.data_ptr()
