"""Geometry primitives on dims (2,3) of [B,C,H,W] tensors and the generalised means used by TTA.

Drop-in for the geometry / mean part of ``pytorch_toolbelt/inference/functional.py`` (reference lines
47-145 and 250-333). Every primitive is one element of the dihedral group D4 and runs as a single
index-permuting copy kernel (``tb_view_permute``); all are autograd-transparent (the backward of a
permutation is the inverse permutation). pad/unpad helpers are out of scope (trivial ``F.pad``).
"""
from __future__ import annotations

import torch
from torch import Tensor

from .. import _native as N
from .._native import TB
from ..utils.support import pytorch_toolbelt_deprecated

__all__ = [
    "geometric_mean",
    "harmonic_mean",
    "harmonic1p_mean",
    "logodd_mean",
    "log1p_mean",
    "torch_fliplr",
    "torch_flipud",
    "torch_none",
    "torch_rot180",
    "torch_rot270",
    "torch_rot90",
    "torch_rot90_ccw",
    "torch_rot90_ccw_transpose",
    "torch_rot90_cw",
    "torch_rot90_cw_transpose",
    "torch_rot180_transpose",
    "torch_transpose",
    "torch_transpose2",
    "torch_transpose_",
    "torch_transpose_rot90_ccw",
    "torch_transpose_rot90_cw",
    "torch_transpose_rot180",
]

_T, _FI, _FJ = TB.V_T, TB.V_FI, TB.V_FJ

# view codes: out[i,j] = x[map(i,j)], (a,b) = T ? (j,i) : (i,j); si = FI ? H-1-a : a; sj = FJ ? W-1-b : b
CODE_ID = 0
CODE_FLIPLR = _FJ  # x[i, W-1-j]
CODE_FLIPUD = _FI  # x[H-1-i, j]
CODE_ROT180 = _FI | _FJ
CODE_TRANSPOSE = _T  # x[j, i]
CODE_ROT90_CW = _T | _FI  # rot90(k=-1): x[M-j, i]
CODE_ROT90_CCW = _T | _FJ  # rot90(k=1):  x[j, M-i]
CODE_ANTITRANSPOSE = _T | _FI | _FJ  # x[M-j, M-i]


def _inverse_code(code: int) -> int:
    if code & _T:
        return _T | (_FI if code & _FJ else 0) | (_FJ if code & _FI else 0)
    return code


def _compose(first: int, then: int) -> int:
    """Code of ``then(first(x))`` (apply ``first``, then ``then``) via a 2x2 probe -- tiny, host only."""
    import numpy as np

    def apply(code, a):
        if code & _T:
            # out[i,j] = a[(M-j or j), (M-i or i)]
            b = a.T
            if code & _FI:  # source row = M - j  -> flips along output columns
                b = b[:, ::-1]
            if code & _FJ:  # source col = M - i  -> flips along output rows
                b = b[::-1, :]
            return b
        b = a
        if code & _FI:
            b = b[::-1, :]
        if code & _FJ:
            b = b[:, ::-1]
        return b

    probe = np.arange(9).reshape(3, 3)
    target = apply(then, apply(first, probe))
    for code in range(8):
        if np.array_equal(apply(code, probe), target):
            return code
    raise AssertionError("D4 is closed")


def _launch_permute(x: Tensor, code: int) -> Tensor:
    N.require_cuda("torch_* geometry primitives")
    if x.dim() < 2:
        raise ValueError("expected a tensor with at least 2 dims")
    on_cpu = not x.is_cuda
    xd = x.cuda() if on_cpu else x
    xd = xd.contiguous()
    if xd.element_size() not in (1, 2, 4, 8):
        raise TypeError(f"unsupported dtype {x.dtype}")
    h, w = xd.shape[-2], xd.shape[-1]
    planes = xd.numel() // max(h * w, 1)
    out_shape = list(xd.shape)
    if code & _T:
        out_shape[-2], out_shape[-1] = w, h
    out = torch.empty(out_shape, dtype=xd.dtype, device=xd.device)
    if xd.numel() > 0:
        with torch.cuda.device(xd.device):
            rc = N.lib().tb_view_permute(xd.data_ptr(), xd.element_size(), planes, h, w, code, out.data_ptr(),
                                         N.stream_ptr(xd.device))
        N.check(rc, "tb_view_permute")
    return out.to(x.device) if on_cpu else out


class _ViewPermute(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x: Tensor, code: int):
        ctx.code = code
        return _launch_permute(x, code)

    @staticmethod
    def backward(ctx, grad):
        return _launch_permute(grad, _inverse_code(ctx.code)), None


def view_permute(x: Tensor, code: int) -> Tensor:
    """Apply one D4 element (3-bit code) to the last two dims of ``x``."""
    if code == CODE_ID:
        return x
    return _ViewPermute.apply(x, code)


def torch_none(x: Tensor) -> Tensor:
    """Return input argument without any modifications"""
    return x


def torch_rot90_ccw(x: Tensor) -> Tensor:
    return view_permute(x, CODE_ROT90_CCW)


def torch_rot90_cw(x: Tensor) -> Tensor:
    return view_permute(x, CODE_ROT90_CW)


def torch_rot90_ccw_transpose(x: Tensor) -> Tensor:
    return view_permute(x, _compose(CODE_ROT90_CCW, CODE_TRANSPOSE))


def torch_rot90_cw_transpose(x: Tensor) -> Tensor:
    return view_permute(x, _compose(CODE_ROT90_CW, CODE_TRANSPOSE))


def torch_transpose_rot90_ccw(x: Tensor) -> Tensor:
    return view_permute(x, _compose(CODE_TRANSPOSE, CODE_ROT90_CCW))


def torch_transpose_rot90_cw(x: Tensor) -> Tensor:
    return view_permute(x, _compose(CODE_TRANSPOSE, CODE_ROT90_CW))


@pytorch_toolbelt_deprecated("Function torch_rot90 has been marked as deprecated. Please use torch_rot90_ccw instead")
def torch_rot90(x: Tensor):
    return torch_rot90_ccw(x)


def torch_rot180(x: Tensor) -> Tensor:
    """Rotate 4D image tensor by 180 degrees"""
    return view_permute(x, CODE_ROT180)


def torch_rot180_transpose(x: Tensor) -> Tensor:
    return view_permute(x, _compose(CODE_ROT180, CODE_TRANSPOSE))


def torch_transpose_rot180(x: Tensor) -> Tensor:
    return view_permute(x, _compose(CODE_TRANSPOSE, CODE_ROT180))


@pytorch_toolbelt_deprecated("Function torch_rot270 has been marked as deprecated. Please use torch_rot90_cw instead")
def torch_rot270(x: Tensor):
    return torch_rot90_cw(x)


def torch_flipud(x: Tensor) -> Tensor:
    """Flip 4D image tensor vertically"""
    return view_permute(x, CODE_FLIPUD)


def torch_fliplr(x: Tensor) -> Tensor:
    """Flip 4D image tensor horizontally"""
    return view_permute(x, CODE_FLIPLR)


def torch_transpose(x: Tensor) -> Tensor:
    """Transpose 4D image tensor by main image diagonal (materialised copy, not a view)"""
    return view_permute(x, CODE_TRANSPOSE)


def torch_transpose_(x: Tensor) -> Tensor:
    # in-place metadata transpose has no kernel work; kept for API parity (functional.py:135-136)
    return x.transpose_(2, 3)


def torch_transpose2(x: Tensor) -> Tensor:
    """Same permutation as torch_transpose (functional.py:139-145 swaps the same two dims)"""
    return view_permute(x, CODE_TRANSPOSE)


# ---------------------------------------------------------------------------------------------------------
# Generalised means over an arbitrary dim (functional.py:250-333). These operate on already de-augmented
# stacks; the fused kernels implement the same formulas for the TTA dimension (csrc/common.cuh red_pre/red_post).
# ---------------------------------------------------------------------------------------------------------
def geometric_mean(x: Tensor, dim: int) -> Tensor:
    return x.log().mean(dim=dim).exp()


def harmonic_mean(x: Tensor, dim: int, eps: float = 1e-6) -> Tensor:
    inv = torch.reciprocal(x.clamp_min(eps))
    return torch.reciprocal(torch.mean(inv, dim=dim).clamp_min(eps))


def harmonic1p_mean(x: Tensor, dim: int) -> Tensor:
    return torch.reciprocal(torch.mean(torch.reciprocal(x + 1), dim=dim)) - 1


def logodd_mean(x: Tensor, dim: int, eps: float = 1e-6) -> Tensor:
    x = x.clamp(min=eps, max=1.0 - eps)
    m = torch.mean(torch.log(x / (1 - x)), dim=dim)
    e = torch.exp(m)
    return e / (1 + e)


def log1p_mean(x: Tensor, dim: int) -> Tensor:
    return torch.exp(torch.mean(torch.log1p(x), dim=dim)) - 1
