"""Oracle for the TTA path (reference inference/tta.py and inference/functional.py). TEST INFRASTRUCTURE ONLY.

A view is a D4 element (transpose?, flip rows?, flip cols?) with gather semantics
    out[i, j] = x[si, sj],  (a, b) = (j, i) if transpose else (i, j),  si = M-a if flip_rows else a,
                                                                     sj = M-b if flip_cols else b
The tables below restate, view by view, what the reference builds from flip / rot90 / transpose calls:
    fliplr tta.py:269, flipud :284, d2 :333-341, flips :484, d4 :410-420  (forward maps: SURVEY Appendix A.3)
"""
from typing import Callable, Optional, Union

import torch

ID, LR, UD, R180 = (0, 0, 0), (0, 0, 1), (0, 1, 0), (0, 1, 1)
TR, CW, CCW, ANTI = (1, 0, 0), (1, 1, 0), (1, 0, 1), (1, 1, 1)

GROUPS = {
    "none": [ID],
    "fliplr": [ID, LR],
    "flipud": [ID, UD],
    "flips": [ID, LR, UD],
    "d2": [ID, LR, UD, R180],
    # id, rot90_cw, rot180, rot90_ccw, transpose, rot90_cw(T)=fliplr, rot180(T)=anti-transpose, rot90_ccw(T)=flipud
    "d4": [ID, CW, R180, CCW, TR, LR, ANTI, UD],
}


def apply_view(x: torch.Tensor, view) -> torch.Tensor:
    """Gather semantics above, realised with the same ATen ops the reference uses (transpose view + flip copies)."""
    t, fi, fj = view
    if t:
        y = x.transpose(2, 3)  # y[i,j] = x[j,i]; then out[i,j] = y[(M-i if fj), (M-j if fi)]
        dims = ([2] if fj else []) + ([3] if fi else [])
    else:
        y = x
        dims = ([2] if fi else []) + ([3] if fj else [])
    return y.flip(dims) if dims else y


def inverse_view(view):
    t, fi, fj = view
    return (t, fj, fi) if t else view


def augment(x: torch.Tensor, group: str) -> torch.Tensor:
    """*_image_augment: cat of the views along dim 0 (aug-major)."""
    if group == "d4" and x.size(2) != x.size(3):
        raise ValueError("d4 needs square input")
    return torch.cat([apply_view(x, v) for v in GROUPS[group]], dim=0)


def averaging(x: torch.Tensor, reduction: Union[str, Callable, None]) -> torch.Tensor:
    """_deaugment_averaging, tta.py:63-96, with the means of inference/functional.py:250-333."""
    if reduction == "mean":
        return x.mean(dim=0)
    if reduction == "sum":
        return x.sum(dim=0)
    if reduction in ("gmean", "geometric_mean"):
        return x.log().mean(dim=0).exp()
    if reduction in ("hmean", "harmonic_mean"):
        return torch.reciprocal(torch.reciprocal(x.clamp_min(1e-6)).mean(dim=0).clamp_min(1e-6))
    if reduction == "harmonic1p":
        return torch.reciprocal(torch.reciprocal(x + 1).mean(dim=0)) - 1
    if reduction == "logodd":
        c = x.clamp(min=1e-6, max=1.0 - 1e-6)
        m = torch.log(c / (1 - c)).mean(dim=0)
        return torch.exp(m) / (1 + torch.exp(m))
    if reduction == "log1p":
        return torch.exp(torch.log1p(x).mean(dim=0)) - 1
    if callable(reduction):
        return reduction(x, dim=0)
    if reduction in (None, "None", "none"):
        return x
    raise KeyError(f"Unsupported reduction mode {reduction}")


def deaugment(x: torch.Tensor, group: str, reduction="mean") -> torch.Tensor:
    """*_image_deaugment: chunk, inverse view per chunk, stack, reduce (tta.py:287-316,344-365,442-467,503-524)."""
    views = GROUPS[group]
    if x.size(0) % len(views) != 0:
        raise RuntimeError(f"Input batch size ({x.size(0)}) must be divisible by {len(views)}.")
    chunks = torch.chunk(x, len(views))
    stacked = torch.stack([apply_view(c, inverse_view(v)) for c, v in zip(chunks, views)])
    return averaging(stacked, reduction)


def gather_reference(x: torch.Tensor, view) -> torch.Tensor:
    """Slow, loop-free index-map evaluation of the gather semantics: independent check of apply_view."""
    t, fi, fj = view
    H, W = x.shape[2], x.shape[3]
    Ho, Wo = (W, H) if t else (H, W)
    i = torch.arange(Ho)[:, None].expand(Ho, Wo)
    j = torch.arange(Wo)[None, :].expand(Ho, Wo)
    a, b = (j, i) if t else (i, j)
    si = (H - 1 - a) if fi else a
    sj = (W - 1 - b) if fj else b
    return x[:, :, si, sj]
