"""Tile-based inference for huge images: slice into overlapping tiles, merge predictions back.

Drop-in for ``pytorch_toolbelt.inference.tiles`` (reference file ``pytorch_toolbelt/inference/tiles.py``):
same class names, constructor arguments, attributes and exception types. The grid arithmetic stays on the
host (integer, bit-exact with the reference); every pixel-touching operation -- crop with border synthesis,
weighted accumulation, normalisation -- runs as a hand-written sm_100a kernel behind the C-ABI of
``include/tb200.h``. There is no CPU compute path: without CUDA these methods raise.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Iterable, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from .. import _native as N
from .._native import TB

__all__ = [
    "ImageSlicer",
    "TileMerger",
    "CudaTileMerger",
    "compute_pyramid_patch_weight_loss",
    "gaussian_patch_weight",
]

# cv2.BORDER_* numeric values (the reference forwards these to cv2.copyMakeBorder; tiles.py:145,177,206)
BORDER_CONSTANT, BORDER_REPLICATE, BORDER_REFLECT, BORDER_WRAP, BORDER_REFLECT_101 = 0, 1, 2, 3, 4


def compute_pyramid_patch_weight_loss(width: int, height: int):
    """Pyramid tile weight: large in the tile centre, small on the border.

    Restates reference tiles.py:16-50 (float64 NumPy, host side, one-off). Axis 0 of the result has extent
    ``width`` -- the reference calls it as ``(tile_h, tile_w)`` (tiles.py:285-287), so the result is
    ``[tile_h, tile_w]``. Returns ``(W, Dc, De)`` like the reference.
    """
    n0, n1 = int(width), int(height)
    a0 = np.arange(n0)
    a1 = np.arange(n1)
    # distance to the tile centre
    dc0 = np.square(a0 - n0 * 0.5 + 0.5)
    dc1 = np.square(a1 - n1 * 0.5 + 0.5)
    Dc = np.sqrt(dc0[:, None] + dc1[None, :])
    # distance to the nearest edge along each axis
    quarter = np.square(0.5)
    e0 = np.sqrt(np.minimum(np.square(a0 + 0.5) + quarter, np.square(a0 - n0 + 0.5) + quarter))
    e1 = np.sqrt(np.minimum(quarter + np.square(a1 + 0.5), quarter + np.square(a1 - n1 + 0.5)))
    De = np.minimum(e0[:, None], e1[None, :])
    ratio = np.divide(De, np.add(Dc, De))
    alpha = (n0 * n1) / np.sum(ratio)
    return alpha * ratio, Dc, De


def _pyramid_alpha(tile_h: int, tile_w: int) -> float:
    _, Dc, De = compute_pyramid_patch_weight_loss(tile_h, tile_w)
    return float((tile_h * tile_w) / np.sum(np.divide(De, np.add(Dc, De))))


def gaussian_patch_weight(tile_h: int, tile_w: int, sigma: Optional[float] = None) -> np.ndarray:
    """Gaussian tile weight ``exp(-((i-(h-1)/2)^2 + (j-(w-1)/2)^2) / (2 sigma^2))``, float64 ``[h, w]``.

    Not in the reference (it only has "mean"/"pyramid", tiles.py:88, and accepts any ndarray, :90); defined here
    for BASELINE config 5. ``sigma`` defaults to ``min(h, w) / 4``.
    """
    sigma = float(min(tile_h, tile_w)) / 4.0 if sigma is None else float(sigma)
    di = np.arange(tile_h, dtype=np.float64) - (tile_h - 1) * 0.5
    dj = np.arange(tile_w, dtype=np.float64) - (tile_w - 1) * 0.5
    return np.exp(-(np.square(di)[:, None] + np.square(dj)[None, :]) * (1.0 / (2.0 * sigma * sigma)))


def _pair(value, what: str) -> Tuple[int, int]:
    if isinstance(value, (np.ndarray, Sequence)):
        if len(value) != 2:
            raise ValueError(f"Tile size must have exactly 2 elements. Got: {what}={value}")
        return int(value[0]), int(value[1])
    return int(value), int(value)


def _border_bits(value, dtype: np.dtype, channels: int) -> "C.Array":
    """cv2 Scalar semantics: a bare number fills channel 0 only, a tuple fills <= 4 channels."""
    vals = [0.0, 0.0, 0.0, 0.0]
    if isinstance(value, (tuple, list, np.ndarray)):
        for k, v in enumerate(list(value)[:4]):
            vals[k] = float(v)
    else:
        vals[0] = float(value)
    arr = np.asarray(vals)
    if np.issubdtype(dtype, np.integer):
        info = np.iinfo(dtype)
        arr = np.clip(np.rint(arr), info.min, info.max)  # cv2 saturate_cast
    arr = arr.astype(dtype)
    bits = arr.view({1: np.uint8, 2: np.uint16, 4: np.uint32, 8: np.uint64}[dtype.itemsize]).astype(np.uint64)
    out = (C.c_uint64 * 4)()
    for k in range(4):
        out[k] = int(bits[k])
    return out


class ImageSlicer:
    """Slice an image into (overlapping) tiles and merge tiles back (reference tiles.py:53-287)."""

    tile_size: Tuple[int, int]
    tile_step: Tuple[int, int]

    def __init__(self, image_shape: Tuple[int, int], tile_size, tile_step=0, image_margin=0, weight="mean"):
        """
        :param image_shape: shape of the source image (H, W[, C]); only [0] and [1] are read
        :param tile_size: scalar or (H, W)
        :param tile_step: scalar or (H, W); must satisfy 1 <= step <= size (the default 0 raises, as upstream)
        :param image_margin: 0 = centre the padded grid automatically; int or (left, right, top, bottom)
        :param weight: "mean", "pyramid", "gaussian" (new) or an ndarray [tile_h, tile_w]
        """
        self.image_height = image_shape[0]
        self.image_width = image_shape[1]
        self.tile_size = _pair(tile_size, "tile_size")
        self.tile_step = _pair(tile_step, "tile_step")

        # analytic description of the weight, used by the fused kernels to regenerate it in registers
        self.weight_kind = TB.W_TABLE
        self.weight_param = 0.0
        if isinstance(weight, np.ndarray):
            self.weight = weight
        else:
            makers = {"mean": self._mean, "pyramid": self._pyramid, "gaussian": self._gaussian}
            self.weight = makers[weight](self.tile_size)  # KeyError for unknown names, as upstream

        th, tw = self.tile_size
        sh, sw = self.tile_step
        if sh < 1 or sh > th:
            raise ValueError()
        if sw < 1 or sw > tw:
            raise ValueError()

        if image_margin == 0:
            # centre the grid: pad so that an integer number of steps covers the image (tiles.py:104-116)
            lap_h, lap_w = th - sh, tw - sw
            count_w = max(1, math.ceil((self.image_width - lap_w) / sw))
            count_h = max(1, math.ceil((self.image_height - lap_h) / sh))
            slack_w = sw * count_w - (self.image_width - lap_w)
            slack_h = sh * count_h - (self.image_height - lap_h)
            self.margin_left = slack_w // 2
            self.margin_right = slack_w - self.margin_left
            self.margin_top = slack_h // 2
            self.margin_bottom = slack_h - self.margin_top
        elif isinstance(image_margin, Sequence):
            self.margin_left, self.margin_right, self.margin_top, self.margin_bottom = image_margin
        else:
            self.margin_left = self.margin_right = self.margin_top = self.margin_bottom = image_margin

        padded_h, padded_w = self.target_shape
        ys = range(0, padded_h - th + 1, sh)
        xs = range(0, padded_w - tw + 1, sw)
        self.grid_shape = (len(ys), len(xs))  # (tile rows, tile cols); new attribute
        crops = [(x, y, tw, th) for y in ys for x in xs]
        self.crops = np.array(crops)
        self.bbox_crops = np.array([(x - self.margin_left, y - self.margin_top, w, h) for (x, y, w, h) in crops])

    # ------------------------------------------------------------------ geometry helpers
    @property
    def target_shape(self) -> Tuple[int, int]:
        return (
            self.image_height + self.margin_bottom + self.margin_top,
            self.image_width + self.margin_right + self.margin_left,
        )

    def crop_to_orignal_size(self, image):
        assert image.shape[0] == self.target_shape[0]
        assert image.shape[1] == self.target_shape[1]
        top, left = self.margin_top, self.margin_left
        crop = image[top : top + self.image_height, left : left + self.image_width]
        assert crop.shape[0] == self.image_height
        assert crop.shape[1] == self.image_width
        return crop

    # ------------------------------------------------------------------ crop (CUDA)
    def _crop(self, image: np.ndarray, tile_indices: np.ndarray, border_type: int, value, clip_to_tile: bool):
        """Crop the listed tiles on the GPU in the image's own dtype; returns ndarray [n, Th, Tw(, C)]."""
        N.require_cuda("ImageSlicer crop")
        if image.ndim not in (2, 3):
            raise ValueError(f"Image must have shape [H,W] or [H,W,C]. Got {image.shape}")
        dtype = image.dtype
        if dtype.itemsize not in (1, 2, 4, 8) or dtype.kind not in "uifb":
            raise TypeError(f"Unsupported image dtype {dtype}")
        channels = 1 if image.ndim == 2 else image.shape[2]
        th, tw = self.tile_size
        host = np.ascontiguousarray(image)
        device = torch.device("cuda", torch.cuda.current_device())
        # raw bytes cross the boundary: the kernel is a pure gather and never interprets the element type
        dev_img = torch.from_numpy(host.reshape(-1).view(np.uint8)).to(device)
        n = len(tile_indices)
        xy = N.as_dev_xy(self.bbox_crops[tile_indices], device)
        out = torch.empty(n * th * tw * channels * dtype.itemsize, dtype=torch.uint8, device=device)
        rc = N.lib().tb_crop_tiles(
            dev_img.data_ptr(), dtype.itemsize, self.image_height, self.image_width, channels, xy.data_ptr(), n,
            th, tw, int(border_type), _border_bits(value, dtype, channels), 1 if clip_to_tile else 0,
            out.data_ptr(), N.stream_ptr(device),
        )
        N.check(rc, "tb_crop_tiles")
        tiles = out.cpu().numpy().view(dtype)
        shape = (n, th, tw) if image.ndim == 2 else (n, th, tw, channels)
        return tiles.reshape(shape)

    def iter_split(self, image: np.ndarray, border_type=BORDER_CONSTANT, value=0) -> Iterable[Tuple[np.ndarray, np.ndarray]]:
        """Yield ``(tile, crops_row)`` pairs; edge tiles are padded relative to the clipped tile (tiles.py:144-175)."""
        if (image.shape[0] != self.image_height) or (image.shape[1] != self.image_width):
            raise ValueError()
        tiles = self._crop(image, np.arange(len(self.crops)), border_type, value, clip_to_tile=True)
        for tile, coords in zip(tiles, self.crops):
            yield tile, coords

    def split(self, image, border_type=BORDER_CONSTANT, value=0) -> List[np.ndarray]:
        """All tiles; the border is synthesised against the whole image (tiles.py:177-204).

        ``image`` may also be a CUDA tensor ``[H,W,C]``/``[H,W]``: then the tiles stay on the device and one
        tensor ``[N,Th,Tw(,C)]`` is returned.
        """
        assert image.shape[0] == self.image_height
        assert image.shape[1] == self.image_width
        if torch.is_tensor(image):
            return self.split_tensor(image, border_type=border_type, value=value)
        tiles = self._crop(image, np.arange(len(self.crops)), border_type, value, clip_to_tile=False)
        return list(tiles)

    def cut_patch(self, image: np.ndarray, slice_index, border_type=BORDER_CONSTANT, value=0):
        assert image.shape[0] == self.image_height
        assert image.shape[1] == self.image_width
        index = np.arange(len(self.crops))[slice_index]
        tiles = self._crop(image, np.atleast_1d(index), border_type, value, clip_to_tile=True)
        return tiles[0]

    def split_tensor(self, image: torch.Tensor, border_type=BORDER_CONSTANT, value=0) -> torch.Tensor:
        """Device-resident split: CUDA tensor [H,W(,C)] -> CUDA tensor [N,Th,Tw(,C)], same dtype."""
        N.require_cuda("ImageSlicer.split_tensor")
        if not image.is_cuda:
            raise ValueError("split_tensor expects a CUDA tensor; pass a numpy array to split() for host images")
        img = image.contiguous()
        channels = 1 if img.dim() == 2 else img.shape[2]
        th, tw = self.tile_size
        n = len(self.crops)
        xy = N.as_dev_xy(self.bbox_crops, img.device)
        out = torch.empty((n, th, tw) if img.dim() == 2 else (n, th, tw, channels), dtype=img.dtype, device=img.device)
        np_dtype = np.dtype(str(img.dtype).replace("torch.", "")) if img.dtype != torch.bfloat16 else np.dtype(np.uint16)
        with torch.cuda.device(img.device):
            rc = N.lib().tb_crop_tiles(
                img.data_ptr(), img.element_size(), self.image_height, self.image_width, channels, xy.data_ptr(), n,
                th, tw, int(border_type), _border_bits(value, np_dtype, channels), 0, out.data_ptr(),
                N.stream_ptr(img.device),
            )
        N.check(rc, "tb_crop_tiles")
        return out

    # ------------------------------------------------------------------ merge (CUDA, float64 accumulators)
    def merge(self, tiles: List[np.ndarray], dtype=np.float32):
        """Weighted overlap-add of tiles in float64, then clip-normalise, cast and crop (tiles.py:244-269)."""
        if len(tiles) != len(self.crops):
            raise ValueError
        N.require_cuda("ImageSlicer.merge")
        device = torch.device("cuda", torch.cuda.current_device())
        first = np.asarray(tiles[0])
        channels = 1 if first.ndim == 2 else first.shape[2]
        th, tw = self.tile_size
        stack = np.stack([np.asarray(t).reshape(th, tw, channels) for t in tiles])  # [N,Th,Tw,C]
        pred = torch.from_numpy(stack).to(device).permute(0, 3, 1, 2).to(torch.float64).contiguous()
        hp, wp = self.target_shape
        merger = TileMerger((hp, wp), channels, np.asarray(self.weight, dtype=np.float64), device=device, dtype=torch.float64)
        merger.integrate_batch(pred, self.crops)
        out = merger._normalized(eps=float(np.finfo(np.float64).eps), top=self.margin_top, left=self.margin_left,
                                 height=self.image_height, width=self.image_width)
        return out.permute(1, 2, 0).cpu().numpy().astype(dtype)

    # ------------------------------------------------------------------ weights
    def _mean(self, tile_size):
        self.weight_kind = TB.W_ONES
        return np.ones((tile_size[0], tile_size[1]), dtype=np.float32)

    def _pyramid(self, tile_size):
        w, Dc, De = compute_pyramid_patch_weight_loss(tile_size[0], tile_size[1])
        self.weight_kind = TB.W_PYRAMID
        self.weight_param = float((tile_size[0] * tile_size[1]) / np.sum(np.divide(De, np.add(Dc, De))))
        return w

    def _gaussian(self, tile_size):
        self.weight_kind = TB.W_GAUSSIAN
        self.weight_param = float(min(tile_size)) / 4.0
        return gaussian_patch_weight(tile_size[0], tile_size[1], self.weight_param)


def _coords_to_host_xy(crop_coords, tile_h: int, tile_w: int) -> np.ndarray:
    """Rows of (x, y, w, h) given as ndarray / list / int64 tensor (DataLoader collate) -> int32 [n,2]."""
    if torch.is_tensor(crop_coords):
        arr = crop_coords.detach().cpu().numpy()
    else:
        arr = np.asarray([np.asarray(c.cpu() if torch.is_tensor(c) else c).reshape(-1) for c in crop_coords])
    arr = np.asarray(arr).reshape(len(arr), -1)
    if arr.shape[1] >= 4:
        if np.any(arr[:, 2] != tile_w) or np.any(arr[:, 3] != tile_h):
            raise ValueError(f"Tile coordinates carry a size different from the weight's {tile_w}x{tile_h}")
    return np.ascontiguousarray(arr[:, :2], dtype=np.int32)


class TileMerger:
    """Accumulate weighted tile predictions into a full-resolution canvas on the GPU (reference tiles.py:290-350).

    ``device="cpu"`` (the reference default) is honoured as "hand results back on the CPU": accumulation still
    runs in the CUDA engine on the current CUDA device (there is no CPU implementation), ``image`` / ``norm_mask``
    / ``merge()`` return CPU copies. Those copies are cached per accumulator state: reading ``merger.image`` twice
    downloads the canvas once (it is 419 MB at BASELINE config 2), and again only after the next ``integrate_batch``
    / ``merge_`` / assignment. Treat the returned CPU tensors as read-only snapshots (in-place edits do not reach the GPU
    state -- assign to ``merger.image`` for that).
    """

    def __init__(self, image_shape, channels, weight, device="cpu", dtype=torch.float32):
        """
        :param image_shape: shape of the PADDED canvas (``ImageSlicer.target_shape``)
        :param channels: number of prediction channels K
        :param weight: weighting matrix [tile_h, tile_w]
        """
        N.require_cuda("TileMerger")
        if dtype not in (torch.float32, torch.float64):
            raise TypeError(f"TileMerger accumulates in float32 or float64 on the GPU; got {dtype}")
        self.image_height = image_shape[0]
        self.image_width = image_shape[1]
        self.channels = channels
        self._out_device = torch.device(device)
        if self._out_device.type == "cuda":
            index = self._out_device.index if self._out_device.index is not None else torch.cuda.current_device()
            self._dev = torch.device("cuda", index)
        else:
            self._dev = torch.device("cuda", torch.cuda.current_device())
        self._dtype = dtype
        weight = np.asarray(weight)
        self._weight = torch.from_numpy(np.expand_dims(weight, axis=0)).to(device=self._dev, dtype=dtype).contiguous()
        self._image = torch.zeros((channels, self.image_height, self.image_width), device=self._dev, dtype=dtype)
        self._norm_mask = torch.zeros((1, self.image_height, self.image_width), device=self._dev, dtype=dtype)
        self._version = 0      # bumped whenever the accumulator changes
        self._host_cache = {}  # name -> (version, CPU copy)

    # plain tensors, like the reference's attributes (tiles.py:306-308); CPU-requesting mergers get (cached) copies
    def _present(self, t: torch.Tensor, name: Optional[str] = None) -> torch.Tensor:
        if self._out_device.type == "cuda":
            return t
        if name is None:
            return t.to(self._out_device)
        hit = self._host_cache.get(name)
        if hit is None or hit[0] != self._version:
            hit = (self._version, t.to(self._out_device))
            self._host_cache[name] = hit
        return hit[1]

    def _touch(self):
        self._version += 1
        self._host_cache.clear()

    @property
    def image(self) -> torch.Tensor:
        return self._present(self._image, "image")

    @image.setter
    def image(self, value: torch.Tensor):
        self._image = value.to(device=self._dev, dtype=self._dtype).contiguous()
        self._touch()

    @property
    def norm_mask(self) -> torch.Tensor:
        return self._present(self._norm_mask, "norm_mask")

    @norm_mask.setter
    def norm_mask(self, value: torch.Tensor):
        self._norm_mask = value.to(device=self._dev, dtype=self._dtype).contiguous()
        self._touch()

    @property
    def weight(self) -> torch.Tensor:
        return self._present(self._weight, "weight")

    @property
    def device(self) -> torch.device:
        return self._out_device if self._out_device.type != "cuda" else self._dev

    def _weight_spec(self) -> N.Weight:
        return N.Weight(TB.W_TABLE, self._weight.data_ptr(), 0.0)

    def accumulate_single(self, tile: torch.Tensor, coords):
        """
        Accumulate a single prediction
        :param tile: predicted image of shape [C,H,W]
        :param coords: (x, y, w, h) of the tile w.r.t. the padded canvas
        """
        self.integrate_batch(tile.unsqueeze(0), [coords])

    def integrate_batch(self, batch: torch.Tensor, crop_coords, tta: int = TB.TTA_NONE, reduction="mean"):
        """
        Accumulate a batch of tile predictions
        :param batch: predicted tiles [B,C,H,W] (or [A*B,C,H,W] TTA-augmented predictions when ``tta`` is given)
        :param crop_coords: B rows of (x, y, w, h) w.r.t. the padded canvas
        """
        views = TB.TTA_VIEWS[tta]
        if len(batch) != len(crop_coords) * views:
            raise ValueError("Number of images in batch does not correspond to number of coordinates")
        if batch.device != self._dev:
            batch = batch.to(device=self._dev)
        if batch.dtype != self._dtype:
            batch = batch.to(self._dtype)
        batch = batch.contiguous()
        th, tw = self._weight.shape[1], self._weight.shape[2]
        if batch.dim() != 4 or batch.shape[1] != self.channels or batch.shape[2] != th or batch.shape[3] != tw:
            raise ValueError(f"Expected batch of shape [B,{self.channels},{th},{tw}], got {tuple(batch.shape)}")
        n = len(crop_coords)
        if n == 0:
            return
        xy = _coords_to_host_xy(crop_coords, th, tw)
        if xy[:, 0].min() < 0 or xy[:, 1].min() < 0 or (xy[:, 0] + tw).max() > self.image_width or (
            xy[:, 1] + th
        ).max() > self.image_height:
            raise ValueError("Tile coordinates fall outside the merger canvas")
        dtype_code = TB.F32 if self._dtype == torch.float32 else TB.F64
        red = TB.REDUCTIONS[reduction]
        wspec = self._weight_spec()
        with torch.cuda.device(self._dev):
            if n <= 64:
                xy_arg, is_host, keep = xy.ctypes.data_as(C.c_void_p), 1, xy
            else:
                keep = torch.from_numpy(xy).to(self._dev)
                xy_arg, is_host = keep.data_ptr(), 0
            rc = N.lib().tb_deaugment_merge(
                batch.data_ptr(), dtype_code, xy_arg, is_host, n, self.channels, th, tw, tta, red, C.byref(wspec),
                self._image.data_ptr(), self._norm_mask.data_ptr(), self.image_height, self.image_width,
                N.stream_ptr(self._dev),
            )
        N.check(rc, "tb_deaugment_merge")
        self._touch()
        del keep

    def _normalized(self, eps: float = -1.0, top: int = 0, left: int = 0, height: Optional[int] = None,
                    width: Optional[int] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        height = self.image_height if height is None else height
        width = self.image_width if width is None else width
        if out is None:
            out = torch.empty((self.channels, height, width), device=self._dev, dtype=self._dtype)
        dtype_code = TB.F32 if self._dtype == torch.float32 else TB.F64
        with torch.cuda.device(self._dev):
            rc = N.lib().tb_normalize_crop(
                self._image.data_ptr(), self._norm_mask.data_ptr(), dtype_code, self.channels, self.image_height,
                self.image_width, top, left, height, width, float(eps), out.data_ptr(), N.stream_ptr(self._dev),
            )
        N.check(rc, "tb_normalize_crop")
        return out

    def merge(self) -> torch.Tensor:
        return self._present(self._normalized())

    def merge_(self) -> torch.Tensor:
        self._normalized(out=self._image)
        self._touch()
        return self._present(self._image, "image")


class CudaTileMerger(TileMerger):
    """``TileMerger`` with the accumulator on the GPU by default (README.md:201,215 of the reference)."""

    def __init__(self, image_shape, channels, weight, device="cuda", dtype=torch.float32):
        super().__init__(image_shape, channels, weight, device=device, dtype=dtype)
