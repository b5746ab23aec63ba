"""The reference arm: the UNMODIFIED pytorch_toolbelt installed under baseline/_ref (see install_ref.sh), driven
through its own public API exactly like README.md:196-227 / tests/test_tiles.py:59-85 of the reference:

    ImageSlicer.split -> image_to_tensor -> .float() -> tta.<group>_image_augment -> model ->
    tta.<group>_image_deaugment -> TileMerger.integrate_batch -> TileMerger.merge -> crop

Only the model between augment and de-augment is ours (the deterministic, D4-non-equivariant stand-in of SURVEY 8d,
plain torch ops) -- the reference ships no model for this path. Nothing from pytorch_toolbelt_b200 is imported here.
Used by `bench.py --impl reference`, by bench.py's `cpu_baseline` leg and by tests that compare against the real
reference when baseline/_ref is present. TEST / BASELINE INFRASTRUCTURE ONLY.
"""
import importlib
import os
import sys
import time
import warnings
from typing import Callable, Dict, Optional

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(_HERE, "_ref")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_DIR, "pytorch_toolbelt"))


def load():
    """Import the installed reference (never the repo's own package). Returns the `pytorch_toolbelt` module."""
    if not available():
        raise ImportError(f"the reference is not installed under {REF_DIR}; run baseline/install_ref.sh")
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        mod = importlib.import_module("pytorch_toolbelt")
        importlib.import_module("pytorch_toolbelt.inference.tiles")
        importlib.import_module("pytorch_toolbelt.inference.tta")
        importlib.import_module("pytorch_toolbelt.losses")
    if not os.path.abspath(mod.__file__).startswith(REF_DIR):
        raise ImportError(f"`pytorch_toolbelt` resolved to {mod.__file__}, not to the install under {REF_DIR}")
    return mod


def standin_model_weights(K: int, Cin: int) -> torch.Tensor:
    return torch.randn(K, Cin, generator=torch.Generator().manual_seed(0))


def standin_model(x: torch.Tensor, Wm: torch.Tensor) -> torch.Tensor:
    """y[b,k,i,j] = sum_c Wm[k,c] x[b,c,i,j] + 0.01 (i - 2j + 7k)/T."""
    B, C, H, W = x.shape
    K = Wm.shape[0]
    i = torch.arange(H, dtype=torch.float32, device=x.device).view(1, 1, H, 1)
    j = torch.arange(W, dtype=torch.float32, device=x.device).view(1, 1, 1, W)
    k = torch.arange(K, dtype=torch.float32, device=x.device).view(1, K, 1, 1)
    return torch.einsum("kc,bchw->bkhw", Wm.to(x.device), x) + 0.01 * (i - 2 * j + 7 * k) / float(H)


def gaussian_weight(th: int, tw: int) -> np.ndarray:
    """cfg 5's builder-defined weight: exp(-((i-(T-1)/2)^2 + (j-(T-1)/2)^2) / (2 sigma^2)), sigma = T/4."""
    i = np.arange(th, dtype=np.float64)[:, None] - (th - 1) / 2.0
    j = np.arange(tw, dtype=np.float64)[None, :] - (tw - 1) / 2.0
    return np.exp(-(i * i / (2.0 * (th / 4.0) ** 2) + j * j / (2.0 * (tw / 4.0) ** 2)))


def tiled_inference(image: np.ndarray, tile: int, step: int, K: int, weight="pyramid", group: Optional[str] = "d4",
                    batch_size: int = 8, model: Optional[Callable] = None, timings: Optional[Dict[str, float]] = None,
                    max_tiles: Optional[int] = None, device: str = "cpu") -> torch.Tensor:
    """The reference's loop over ALL tiles (or the first `max_tiles`, for bounded timing samples).
    Returns the merged, cropped [K,H,W] float32 tensor on the CPU; fills `timings` per stage (seconds)."""
    load()
    from pytorch_toolbelt.inference import tta as RT
    from pytorch_toolbelt.inference.tiles import ImageSlicer, TileMerger
    from pytorch_toolbelt.utils.torch_utils import image_to_tensor

    on_gpu = str(device).startswith("cuda")

    def clock():
        if on_gpu:
            torch.cuda.synchronize()
        return time.perf_counter()

    aug_fn = deaug_fn = None
    if group not in (None, "none"):
        aug_fn = getattr(RT, f"{group}_image_augment")
        deaug_fn = getattr(RT, f"{group}_image_deaugment")
    H, W = image.shape[:2]
    Cin = 1 if image.ndim == 2 else image.shape[2]
    if isinstance(weight, str) and weight == "gaussian":
        weight = gaussian_weight(tile, tile)
    Wm = standin_model_weights(K, Cin)
    model = model or (lambda x: standin_model(x, Wm))
    tm = timings if timings is not None else {}
    for key in ("split", "augment", "model", "deaugment", "integrate", "merge"):
        tm.setdefault(key, 0.0)

    t0 = time.perf_counter()
    tiler = ImageSlicer(image.shape, tile_size=(tile, tile), tile_step=(step, step), weight=weight)
    tiles = [image_to_tensor(t) for t in tiler.split(image)]
    tm["split"] += time.perf_counter() - t0
    merger = TileMerger(tiler.target_shape, K, tiler.weight, device=device)
    n = len(tiles) if max_tiles is None else min(len(tiles), max_tiles)
    for s in range(0, n, batch_size):
        e = min(s + batch_size, n)
        t0 = time.perf_counter()
        batch = torch.stack(tiles[s:e]).float().to(device)
        tm["split"] += clock() - t0
        t0 = clock()
        aug = aug_fn(batch) if aug_fn else batch
        tm["augment"] += clock() - t0
        t0 = clock()
        pred = model(aug)
        tm["model"] += clock() - t0
        t0 = clock()
        pred = deaug_fn(pred) if deaug_fn else pred
        tm["deaugment"] += clock() - t0
        t0 = clock()
        merger.integrate_batch(pred, tiler.crops[s:e])
        tm["integrate"] += clock() - t0
    t0 = clock()
    merged = merger.merge()
    out = merged[:, tiler.margin_top : tiler.margin_top + H, tiler.margin_left : tiler.margin_left + W].contiguous()
    tm["merge"] += clock() - t0
    return out.cpu()
