"""The reference's canonical tiled-inference loop (README.md:196-227, tests/test_tiles.py:59-85) restated on
torch-CPU / NumPy with the oracle pieces. Used as the end-to-end checker and as the timed CPU baseline.
TEST INFRASTRUCTURE ONLY."""
import time
from typing import Callable, Dict, Optional

import numpy as np
import torch

from . import tiles_oracle as TO
from . import tta_oracle as AO


def standin_model_weights(K: int, Cin: int) -> torch.Tensor:
    """Wm = randn(K, Cin; seed 0)."""
    return torch.randn(K, Cin, generator=torch.Generator().manual_seed(0))


def standin_model(x: torch.Tensor, Wm: torch.Tensor) -> torch.Tensor:
    """y[b,k,i,j] = sum_c Wm[k,c] x[b,c,i,j] + 0.01 (i - 2j + 7k)/T : deterministic and NOT D4-equivariant."""
    B, C, H, W = x.shape
    K = Wm.shape[0]
    dev = x.device
    i = torch.arange(H, dtype=torch.float32, device=dev).view(1, 1, H, 1)
    j = torch.arange(W, dtype=torch.float32, device=dev).view(1, 1, 1, W)
    k = torch.arange(K, dtype=torch.float32, device=dev).view(1, K, 1, 1)
    return torch.einsum("kc,bchw->bkhw", Wm.to(dev), x) + 0.01 * (i - 2 * j + 7 * k) / float(H)


def tiled_inference_cpu(image: np.ndarray, tile, step, K: int, weight="pyramid", group="none", batch_size=8,
                        model: Optional[Callable] = None, timings: Optional[Dict[str, float]] = None,
                        max_tiles: Optional[int] = None, device: str = "cpu") -> torch.Tensor:
    """split -> HWC->CHW float -> [augment] -> model -> [de-augment] -> integrate_batch -> merge -> crop.
    Returns [K,H,W] float32. ``max_tiles`` bounds the work for the timed CPU baseline (the merged result is then
    partial and only the timings are meaningful). ``device="cuda"`` runs the same torch ops as CUDA eager (the
    reference's own GPU path: tiles cross to the device per batch like README.md:219, the merger lives on the GPU)."""
    on_gpu = str(device).startswith("cuda")

    def clock():
        if on_gpu:
            torch.cuda.synchronize()
        return time.perf_counter()

    H, W = image.shape[:2]
    tile = (tile, tile) if np.isscalar(tile) else tuple(tile)
    step = (step, step) if np.isscalar(step) else tuple(step)
    grid = TO.tile_grid(H, W, tile, step)
    if isinstance(weight, str):
        w = {"pyramid": TO.pyramid_weight, "gaussian": TO.gaussian_weight,
             "mean": lambda h, ww: np.ones((h, ww), np.float32)}[weight](*tile)
    else:
        w = weight
    Cin = 1 if image.ndim == 2 else image.shape[2]
    Wm = standin_model_weights(K, Cin)
    model = model or (lambda x: standin_model(x, Wm))
    tm = timings if timings is not None else {}
    for key in ("split", "augment", "model", "deaugment", "integrate", "merge"):
        tm.setdefault(key, 0.0)

    t0 = time.perf_counter()
    tiles = TO.split(image, grid)
    tm["split"] += time.perf_counter() - t0
    merger = TO.TorchMerger(grid["target_shape"], K, w, device=device)
    n = len(tiles) if max_tiles is None else min(len(tiles), max_tiles)
    for s in range(0, n, batch_size):
        e = min(s + batch_size, n)
        t0 = time.perf_counter()
        batch = torch.stack([torch.from_numpy(np.ascontiguousarray(np.moveaxis(t.reshape(tile[0], tile[1], Cin), -1, 0)))
                             for t in tiles[s:e]]).float().to(device)
        tm["split"] += clock() - t0
        t0 = clock()
        aug = AO.augment(batch, group) if group != "none" else batch
        tm["augment"] += clock() - t0
        t0 = clock()
        pred = model(aug)
        tm["model"] += clock() - t0
        t0 = clock()
        pred = AO.deaugment(pred, group, "mean") if group != "none" else pred
        tm["deaugment"] += clock() - t0
        t0 = clock()
        merger.integrate_batch(pred, grid["crops"][s:e])
        tm["integrate"] += clock() - t0
    t0 = clock()
    merged = merger.merge()
    left, right, top, bottom = grid["margins"]
    out = merged[:, top : top + H, left : left + W].contiguous()
    tm["merge"] += clock() - t0
    return out
