"""Oracle for the tiling path. Each function cites the reference lines it restates
(paths relative to /root/reference/pytorch_toolbelt). TEST INFRASTRUCTURE ONLY."""
import math
from typing import List, Sequence, Tuple

import numpy as np
import torch

BORDER_CONSTANT, BORDER_REPLICATE, BORDER_REFLECT, BORDER_WRAP, BORDER_REFLECT_101 = 0, 1, 2, 3, 4


def pyramid_weight(tile_h: int, tile_w: int) -> np.ndarray:
    """inference/tiles.py:16-50 called as (tile_h, tile_w) by :285-287 -> float64 [tile_h, tile_w]."""
    i = np.arange(tile_h, dtype=np.float64)[:, None]
    j = np.arange(tile_w, dtype=np.float64)[None, :]
    dc = np.sqrt((i - tile_h * 0.5 + 0.5) ** 2 + (j - tile_w * 0.5 + 0.5) ** 2)
    de_i = np.sqrt(np.minimum((i + 0.5) ** 2 + 0.25, (i - tile_h + 0.5) ** 2 + 0.25))
    de_j = np.sqrt(np.minimum(0.25 + (j + 0.5) ** 2, 0.25 + (j - tile_w + 0.5) ** 2))
    de = np.minimum(de_i, de_j)
    ratio = de / (dc + de)
    return (tile_h * tile_w) / np.sum(ratio) * ratio


def gaussian_weight(tile_h: int, tile_w: int, sigma=None) -> np.ndarray:
    """Builder-defined (not in the reference): exp(-r^2 / (2 sigma^2)), sigma = min(h,w)/4."""
    sigma = min(tile_h, tile_w) / 4.0 if sigma is None else sigma
    i = np.arange(tile_h, dtype=np.float64)[:, None] - (tile_h - 1) / 2.0
    j = np.arange(tile_w, dtype=np.float64)[None, :] - (tile_w - 1) / 2.0
    return np.exp(-(i * i + j * j) * (1.0 / (2.0 * sigma * sigma)))


def tile_grid(height: int, width: int, tile: Tuple[int, int], step: Tuple[int, int], margin=0):
    """inference/tiles.py:97-142. Returns dict(crops, bbox_crops, margins=(left,right,top,bottom), target_shape)."""
    th, tw = tile
    sh, sw = step
    if margin == 0:
        nw = max(1, math.ceil((width - (tw - sw)) / sw))
        nh = max(1, math.ceil((height - (th - sh)) / sh))
        extra_w = sw * nw - (width - (tw - sw))
        extra_h = sh * nh - (height - (th - sh))
        left, top = extra_w // 2, extra_h // 2
        right, bottom = extra_w - left, extra_h - top
    elif isinstance(margin, (tuple, list)):
        left, right, top, bottom = margin
    else:
        left = right = top = bottom = margin
    crops, bbox = [], []
    y = 0
    while y <= height + top + bottom - th:
        x = 0
        while x <= width + left + right - tw:
            crops.append((x, y, tw, th))
            bbox.append((x - left, y - top, tw, th))
            x += sw
        y += sh
    return dict(
        crops=np.array(crops),
        bbox_crops=np.array(bbox),
        margins=(left, right, top, bottom),
        target_shape=(height + top + bottom, width + left + right),
    )


def border_index(p: np.ndarray, length: int, mode: int) -> np.ndarray:
    """OpenCV borderInterpolate semantics (cv2.copyMakeBorder is what inference/tiles.py:161-169,182-190 call).
    Returns source indices in [0,length) or -1 where the constant fill applies."""
    p = np.asarray(p, dtype=np.int64).copy()
    inside = (p >= 0) & (p < length)
    if mode == BORDER_CONSTANT:
        return np.where(inside, p, -1)
    if mode == BORDER_REPLICATE:
        return np.clip(p, 0, length - 1)
    if mode == BORDER_WRAP:
        return np.mod(p, length)
    if length == 1:
        return np.zeros_like(p)
    delta = 1 if mode == BORDER_REFLECT_101 else 0
    out = p
    while True:
        bad = (out < 0) | (out >= length)
        if not bad.any():
            return out
        neg = out < 0
        out = np.where(neg, -out - 1 + delta, out)
        big = out >= length
        out = np.where(big & ~neg, length - 1 - (out - length) - delta, out)


def _fill_value(value, channels: int, dtype) -> np.ndarray:
    """cv2 Scalar semantics: bare number -> channel 0 only; tuple -> up to 4 channels."""
    vals = np.zeros(max(channels, 4), dtype=np.float64)
    if isinstance(value, (tuple, list, np.ndarray)):
        for k, v in enumerate(list(value)[:4]):
            vals[k] = v
    else:
        vals[0] = value
    if np.issubdtype(np.dtype(dtype), np.integer):
        info = np.iinfo(dtype)
        vals = np.clip(np.rint(vals), info.min, info.max)
    return vals[:channels].astype(dtype)


def crop_with_border(image: np.ndarray, x: int, y: int, tw: int, th: int, mode: int, value, clip_to_tile: bool):
    """One tile. clip_to_tile=False: pad relative to the whole image (split, tiles.py:177-204);
    True: relative to the clipped tile (iter_split / cut_patch, tiles.py:144-175,206-234)."""
    H, W = image.shape[:2]
    img = image if image.ndim == 3 else image[:, :, None]
    C = img.shape[2]
    if clip_to_tile:
        y_lo, y_hi = max(y, 0), min(H, y + th)
        x_lo, x_hi = max(x, 0), min(W, x + tw)
    else:
        y_lo, y_hi, x_lo, x_hi = 0, H, 0, W
    ys = border_index(np.arange(y, y + th) - y_lo, y_hi - y_lo, mode)
    xs = border_index(np.arange(x, x + tw) - x_lo, x_hi - x_lo, mode)
    fill = _fill_value(value, C, image.dtype)
    out = np.empty((th, tw, C), dtype=image.dtype)
    out[:] = fill
    vy, vx = ys >= 0, xs >= 0
    out[np.ix_(vy, vx)] = img[np.ix_(ys[vy] + y_lo, xs[vx] + x_lo)]
    return out if image.ndim == 3 else out[:, :, 0]


def pad_whole_image(image: np.ndarray, grid: dict, mode=BORDER_CONSTANT, value=0) -> np.ndarray:
    """cv2.copyMakeBorder(image, top, bottom, left, right) of ImageSlicer.split, inference/tiles.py:182-190:
    one gather of the whole image through the border index maps."""
    H, W = image.shape[:2]
    left, right, top, bottom = grid["margins"]
    img = image if image.ndim == 3 else image[:, :, None]
    ys = border_index(np.arange(-top, H + bottom), H, mode)
    xs = border_index(np.arange(-left, W + right), W, mode)
    if mode == BORDER_CONSTANT:
        padded = np.empty((len(ys), len(xs), img.shape[2]), dtype=image.dtype)
        padded[:] = _fill_value(value, img.shape[2], image.dtype)
        padded[top : top + H, left : left + W] = img
    else:
        padded = img[np.ix_(ys, xs)]
    return padded if image.ndim == 3 else padded[:, :, 0]


def split(image: np.ndarray, grid: dict, mode=BORDER_CONSTANT, value=0, clip_to_tile=False) -> List[np.ndarray]:
    """clip_to_tile=False: ImageSlicer.split (pad the whole image once, tiles are views, tiles.py:177-204);
    True: iter_split / cut_patch semantics (pad each clipped tile, tiles.py:144-175)."""
    if clip_to_tile:
        return [crop_with_border(image, x, y, w, h, mode, value, True) for (x, y, w, h) in grid["bbox_crops"]]
    padded = pad_whole_image(image, grid, mode, value)
    return [padded[y : y + h, x : x + w] for (x, y, w, h) in grid["crops"]]


def merge_numpy_f64(tiles: Sequence[np.ndarray], grid: dict, weight: np.ndarray, height: int, width: int, dtype):
    """ImageSlicer.merge, inference/tiles.py:244-269 (+ crop :271-280): float64 overlap-add, eps clip."""
    hp, wp = grid["target_shape"]
    first = np.asarray(tiles[0])
    C = 1 if first.ndim == 2 else first.shape[2]
    acc = np.zeros((hp, wp, C), dtype=np.float64)
    norm = np.zeros((hp, wp, C), dtype=np.float64)
    w3 = np.repeat(np.asarray(weight)[:, :, None], C, axis=2)
    for tile, (x, y, tw, th) in zip(tiles, grid["crops"]):
        acc[y : y + th, x : x + tw] += np.asarray(tile).reshape(th, tw, C) * w3
        norm[y : y + th, x : x + tw] += w3
    norm = np.maximum(norm, np.finfo(np.float64).eps)
    left, right, top, bottom = grid["margins"]
    return (acc / norm).astype(dtype)[top : top + height, left : left + width]


class TorchMerger:
    """TileMerger, inference/tiles.py:290-350, on torch CPU (float32 by default): per tile `image += tile*w`."""

    def __init__(self, canvas_hw, channels: int, weight: np.ndarray, dtype=torch.float32, device="cpu"):
        self.weight = torch.from_numpy(np.asarray(weight)[None]).to(device=device, dtype=dtype)
        self.image = torch.zeros((channels, canvas_hw[0], canvas_hw[1]), dtype=dtype, device=device)
        self.norm_mask = torch.zeros((1, canvas_hw[0], canvas_hw[1]), dtype=dtype, device=device)

    def integrate_batch(self, batch: torch.Tensor, coords):
        if len(batch) != len(coords):
            raise ValueError("Number of images in batch does not correspond to number of coordinates")
        batch = batch.to(device=self.image.device, dtype=self.image.dtype)
        for tile, c in zip(batch, coords):
            x, y, tw, th = (int(v) for v in c)
            self.image[:, y : y + th, x : x + tw] += tile * self.weight
            self.norm_mask[:, y : y + th, x : x + tw] += self.weight

    def merge(self) -> torch.Tensor:
        return self.image / self.norm_mask
