"""ctypes binding of the tb200 C-ABI library (``include/tb200.h``).

This is the "reference-side stub" of INTEGRATION.md: plain pointers and sizes, no torch types cross the
boundary. torch is used only as the allocator / stream provider. There is deliberately NO fallback: if the
shared library is missing or CUDA is unavailable every entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np
import torch

__all__ = ["lib", "check", "TB", "stream_ptr", "require_cuda", "Weight", "Grid", "MergeExtra", "FocalParams", "DiceParams"]

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libtb200.so")


class TB:
    """Enum values mirrored from include/tb200.h."""

    # dtypes
    U8, I8, U16, I16, I32, I64, F16, BF16, F32, F64 = range(10)
    HWC, CHW = 0, 1
    TTA_NONE, TTA_FLIPLR, TTA_FLIPUD, TTA_FLIPS, TTA_D2, TTA_D4 = range(6)
    BORDER_CONSTANT, BORDER_REPLICATE, BORDER_REFLECT, BORDER_WRAP, BORDER_REFLECT_101 = range(5)
    RED_NONE, RED_MEAN, RED_SUM, RED_GMEAN, RED_HMEAN, RED_HARMONIC1P, RED_LOGODD, RED_LOG1P = range(8)
    W_TABLE, W_ONES, W_PYRAMID, W_GAUSSIAN = range(4)
    LOSS_NONE, LOSS_MEAN, LOSS_SUM, LOSS_BATCHWISE, LOSS_NORMALIZED = range(5)
    EL_SOFT_BCE, EL_BALANCED_BCE, EL_QFL = 0, 1, 2
    PRE_NONE, PRE_SIGMOID = 0, 1
    OUT_LOGITS, OUT_ARGMAX, OUT_MASK = 0, 1, 2
    T_DENSE_F32, T_LABELS_I64 = 0, 1
    # view code bits (tb_view_permute)
    V_T, V_FI, V_FJ = 1, 2, 4

    TTA_VIEWS = {0: 1, 1: 2, 2: 2, 3: 3, 4: 4, 5: 8}

    TORCH_DTYPE = {
        torch.uint8: U8,
        torch.int8: I8,
        torch.int16: I16,
        torch.int32: I32,
        torch.int64: I64,
        torch.float16: F16,
        torch.bfloat16: BF16,
        torch.float32: F32,
        torch.float64: F64,
    }
    if hasattr(torch, "uint16"):
        TORCH_DTYPE[torch.uint16] = U16

    REDUCTIONS = {
        "mean": RED_MEAN,
        "sum": RED_SUM,
        "gmean": RED_GMEAN,
        "geometric_mean": RED_GMEAN,
        "hmean": RED_HMEAN,
        "harmonic_mean": RED_HMEAN,
        "harmonic1p": RED_HARMONIC1P,
        "logodd": RED_LOGODD,
        "log1p": RED_LOG1P,
        None: RED_NONE,
        "None": RED_NONE,
        "none": RED_NONE,
    }


class Weight(C.Structure):
    _fields_ = [("kind", C.c_int), ("table", C.c_void_p), ("param", C.c_double)]


class Grid(C.Structure):
    _fields_ = [
        ("nx", C.c_int),
        ("ny", C.c_int),
        ("Sx", C.c_int),
        ("Sy", C.c_int),
        ("Th", C.c_int),
        ("Tw", C.c_int),
        ("n_slots", C.c_int),
        ("tile_slot", C.c_void_p),
        ("tile_ptr", C.c_void_p),
        ("tile_vstride", C.c_void_p),
    ]


class PeerPlan(C.Structure):
    """tb_peer_plan (include/tb200.h): multi-GPU ragged strips -- which block classes a launch processes and where the
    seed (read) / spill (peer write) partial sums live."""

    _fields_ = [("phase_mask", C.c_int), ("t0", C.c_int), ("t1", C.c_int), ("seed_num", C.c_void_p), ("seed_den", C.c_void_p),
                ("seed_y0", C.c_int), ("seed_rows", C.c_int), ("spill_num", C.c_void_p), ("spill_den", C.c_void_p),
                ("spill_y0", C.c_int), ("spill_rows", C.c_int),
                ("fused", C.c_int), ("chained", C.c_int), ("has_next", C.c_int), ("has_prev", C.c_int), ("epoch", C.c_uint64),
                ("counters", C.c_void_p), ("ready_flag", C.c_void_p), ("ack_flag", C.c_void_p), ("next_ready", C.c_void_p),
                ("prev_ack", C.c_void_p), ("status", C.c_void_p), ("timeout_ms", C.c_int)]


class MergeExtra(C.Structure):
    """tb_merge_extra (include/tb200.h): pre-activation before the view reduction, label / mask output modes, peer plan."""

    _fields_ = [("pre", C.c_int), ("temperature", C.c_float), ("out_mode", C.c_int), ("threshold", C.c_float),
                ("out_u8", C.c_void_p), ("peer", C.POINTER(PeerPlan))]


class FocalParams(C.Structure):
    _fields_ = [
        ("gamma", C.c_float),
        ("alpha", C.c_float),
        ("has_alpha", C.c_int),
        ("reduced_threshold", C.c_float),
        ("has_reduced_threshold", C.c_int),
        ("ignore_index", C.c_float),
        ("has_ignore", C.c_int),
        ("normalized", C.c_int),
        ("eps", C.c_float),
        ("reduction", C.c_int),
        ("softmax", C.c_int),
    ]


class ElemLossParams(C.Structure):
    _fields_ = [("kind", C.c_int), ("reduction", C.c_int), ("smooth", C.c_float), ("has_smooth", C.c_int),
                ("ignore_index", C.c_float), ("has_ignore", C.c_int), ("weight_n", C.c_int), ("pos_weight_n", C.c_int),
                ("gamma", C.c_float), ("beta", C.c_float)]


class DiceParams(C.Structure):
    _fields_ = [
        ("jaccard", C.c_int),
        ("mode", C.c_int),
        ("from_logits", C.c_int),
        ("log_loss", C.c_int),
        ("smooth", C.c_float),
        ("eps", C.c_float),
        ("has_ignore", C.c_int),
        ("ignore_index", C.c_float),
    ]


_vp, _i, _i64, _f, _d, _sz = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_double, C.c_size_t

# name -> (restype, argtypes); must list every symbol include/tb200.h declares (tests check this)
SIGNATURES = {
    "tb_version": (_i, []),
    "tb_last_error": (C.c_char_p, []),
    "tb_tta_views": (_i, [_i]),
    "tb_crop_tiles": (_i, [_vp, _i, _i, _i, _i, _vp, _i, _i, _i, _i, _vp, _i, _vp, _vp]),
    "tb_slice_augment": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "tb_tta_augment": (_i, [_vp, _i, _i, _i, _i, _i, _i, _vp, _vp]),
    "tb_view_permute": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _vp]),
    "tb_tta_deaugment": (_i, [_vp, _i, _i, _i, _i, _i, _i, _i, _vp, _vp]),
    "tb_deaugment_merge": (_i, [_vp, _i, _vp, _i, _i, _i, _i, _i, _i, _i, C.POINTER(Weight), _vp, _vp, _i, _i, _vp]),
    "tb_normalize_crop": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _d, _vp, _vp]),
    "tb_merge_all": (
        _i,
        [_vp, C.POINTER(Grid), _i, _i, _i, C.POINTER(Weight), _i, _i, _i, _i, _i, _i, _vp, _vp, _i, _i, _i, _i, _i,
         _vp, _vp, _vp],
    ),
    "tb_merge_all_batched": (
        _i,
        [_vp, _i, _i, _i, _i, C.POINTER(Grid), _i, _i, _i, C.POINTER(Weight), _i, _i, _i, _i, _i, _i, _vp, _vp, _i, _i, _i,
         _i, _i, _vp, _vp, _vp],
    ),
    "tb_merge_all_ex": (
        _i,
        [_vp, C.POINTER(Grid), _i, _i, _i, C.POINTER(Weight), _i, _i, _i, _i, _i, _i, _vp, _vp, _i, _i, _i, _i, _i,
         _vp, _vp, C.POINTER(MergeExtra), _vp],
    ),
    "tb_merge_all_batched_ex": (
        _i,
        [_vp, _i, _i, _i, _i, C.POINTER(Grid), _i, _i, _i, C.POINTER(Weight), _i, _i, _i, _i, _i, _i, _vp, _vp, _i, _i, _i,
         _i, _i, _vp, _vp, C.POINTER(MergeExtra), _vp],
    ),
    "tb_stack_reduce": (_i, [_vp, _i, _i64, _i, _vp, _vp]),
    "tb_activation_inplace": (_i, [_vp, _i64, _i, _i64, _i, _f, _vp]),
    "tb_volume_crop_tiles": (_i, [_vp, _i, _i, _i, _i, _vp, _i, _i, _i, _i, C.c_uint64, _vp, _vp]),
    "tb_volume_integrate": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "tb_volume_normalize": (_i, [_vp, _vp, _i, _i64, _vp, _vp]),
    "tb_volume_merge_all": (_i, [_vp, _vp] + [_i] * 16 + [_vp, _vp]),
    "tb_peer_alloc": (_i, [_sz, C.POINTER(C.c_void_p)]),
    "tb_peer_free": (_i, [_vp]),
    "tb_peer_export": (_i, [_vp, _vp]),
    "tb_peer_open": (_i, [_vp, C.POINTER(C.c_void_p)]),
    "tb_peer_close": (_i, [_vp]),
    "tb_peer_signal": (_i, [_vp, C.c_uint64, _vp]),
    "tb_peer_wait": (_i, [_vp, C.c_uint64, _i, _vp, _vp]),
    "tb_peer_read": (_i, [_vp, C.POINTER(C.c_uint64)]),
    "tb_loss_workspace_bytes": (_sz, [_i64, _i, _i64]),
    "tb_focal_fwd_bwd": (_i, [_vp, _vp, _i, _i64, _i, _i64, C.POINTER(FocalParams), _vp, _vp, _vp, _vp, _sz, _vp]),
    "tb_softmax_focal_fwd_bwd": (
        _i,
        [_vp, _vp, _i64, _i, _i64, _f, _f, _i, _i64, _i, _f, _i, _vp, _vp, _vp, _vp, _sz, _vp],
    ),
    "tb_dice_stats": (_i, [_vp, _vp, _i, _i64, _i, _i64, C.POINTER(DiceParams), _vp, _vp, _sz, _vp]),
    "tb_dice_loss_grad": (_i, [_vp, _vp, _i, _i64, _i, _i64, C.POINTER(DiceParams), _vp, _vp, _i, _vp, _vp, _vp]),
    "tb_logits_to_labels": (_i, [_vp, _i, _i64, _i64, _i, _f, _vp, _i64, _vp]),
    "tb_focal_dice_fwd_bwd": (
        _i,
        [_vp, _vp, _i64, _i, _i64, C.POINTER(FocalParams), _vp, C.POINTER(DiceParams), _vp, _i, _f, _f, _vp, _vp, _vp, _sz, _vp],
    ),
    "tb_wing_fwd_bwd": (_i, [_vp, _vp, _i64, _f, _f, _i, _vp, _vp, _vp, _sz, _vp]),
    "tb_scale_by_device_scalar": (_i, [_vp, _i64, _vp, _vp]),
    "tb_elem_loss_fwd_bwd": (_i, [_vp, _vp, _i64, _i, _i64, C.POINTER(ElemLossParams), _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "tb_soft_f1_workspace_bytes": (_sz, [_i64, _i]),
    "tb_soft_f1_fwd_bwd": (_i, [_vp, _vp, _i64, _i, _i, _i, _f, _f, _vp, _vp, _vp, _sz, _vp]),
    "tb_lovasz_fused_workspace_bytes": (_sz, [_i64, _i, _i64, _i]),
    "tb_lovasz_fwd_bwd": (_i, [_vp, _vp, _i, _i64, _i, _i64, _i, _i, _i, _f, _vp, _vp, _vp, _sz, _vp]),
    "tb_standin_model": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
}

_lib: Optional[C.CDLL] = None


def load_library(path: str = _LIB_PATH) -> C.CDLL:
    """dlopen libtb200.so and attach the prototypes. Raises (never falls back) if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(path):
        raise RuntimeError(
            f"tb200 native library not found at {path}. Build it with "
            f"`python -c 'import __graft_entry__ as g; g.build()'` or `make -C pytorch_toolbelt_b200/csrc`. "
            "There is no CPU / eager fallback for this engine."
        )
    handle = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(handle, name, None)
        if fn is None:
            continue  # reported by exported_symbols(); calling it raises AttributeError
        fn.restype = res
        fn.argtypes = args
    _lib = handle
    return handle


def lib() -> C.CDLL:
    return load_library()


def exported_symbols() -> Sequence[str]:
    handle = load_library()
    return [name for name in SIGNATURES if hasattr(handle, name)]


class TB200Error(RuntimeError):
    pass


def check(rc: int, what: str = "") -> None:
    if rc == 0:
        return
    msg = lib().tb_last_error()
    msg = msg.decode("utf-8", "replace") if msg else ""
    kind = "argument error" if rc < 0 else "CUDA error"
    raise TB200Error(f"tb200 {what}: {kind} {rc}: {msg}")


def require_cuda(what: str = "this operation") -> None:
    if not torch.cuda.is_available():
        raise RuntimeError(
            f"pytorch_toolbelt_b200: {what} runs only as hand-written CUDA kernels on a B200-class GPU; "
            "CUDA is not available and there is no CPU fallback."
        )


def stream_ptr(device: Optional[torch.device] = None) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def float_array(values: Optional[Sequence[float]], n: int):
    if values is None:
        return None
    arr = (C.c_float * n)()
    for k in range(n):
        arr[k] = float(values[k]) if k < len(values) else 0.0
    return arr


def as_dev_xy(coords, device: torch.device) -> torch.Tensor:
    """[n,>=2] (x, y, ...) rows as a contiguous int32 device tensor [n,2]."""
    if torch.is_tensor(coords):
        xy = coords.detach()[:, :2].to(device="cpu", dtype=torch.int32)
    else:
        xy = torch.from_numpy(np.ascontiguousarray(np.asarray(coords)[:, :2], dtype=np.int32))
    return xy.contiguous().to(device, non_blocking=False)
