"""3-D tiles: drop-in for ``pytorch_toolbelt/inference/tiles_3d.py`` (``VolumeSlicer``, ``VolumeMerger``), SURVEY 8f rank 3.

Grid arithmetic stays on the host (NumPy, identical to the reference, tiles_3d.py:53-104); voxels are touched only by the
CUDA kernels of ``csrc/volume.cu``: ``split`` / ``iter_split`` gather the constant-padded tiles on the GPU, ``VolumeMerger``
accumulates with vector reductions, and ``VolumeSlicer.merge_all`` / ``VolumeMerger.merge_tiles`` run the whole
integrate -> normalise -> crop loop as ONE output-stationary launch (deterministic, no canvas, no atomics).

Deviations from the reference, deliberate: ``VolumeSlicer.merge`` (dead 2-D code copied from ImageSlicer, tiles_3d.py:140-166,
it references attributes that do not exist) is implemented as the documented intent -- float64 weighted overlap-add of host
tiles; ``VolumeMerger.accumulate_single`` indexes ``[:, roi]`` with a tuple of slices, which is a 4-D index error in the
reference (:192-193): here it does what ``integrate_batch`` does for one tile.
"""
from __future__ import annotations

import ctypes as C
from typing import Any, Iterable, List, Sequence, Tuple, Union

import numpy as np
import torch

from .. import _native as N
from .tiles import compute_pyramid_patch_weight_loss

__all__ = ["VolumeSlicer", "VolumeMerger"]


def _triple(v, what):
    if isinstance(v, (tuple, list)):
        if len(v) != 3:
            raise ValueError(f"{what} must be a scalar or have 3 entries")
        return np.array(v, dtype=int)
    return np.array([int(v)] * 3)


def _roi_origin(roi) -> Tuple[int, int, int]:
    return tuple(int(s.start) for s in roi)


class VolumeSlicer:
    """Slices a [D,H,W] volume into overlapping [d,h,w] tiles and merges tile predictions back (reference tiles_3d.py:10-177)."""

    def __init__(self, volume_shape: Tuple[int, int, int], voxel_size: Union[int, Tuple[int, int, int]],
                 voxel_step: Union[int, Tuple[int, int, int]], weight="mean"):
        self.volume_shape = np.array(volume_shape)[:3]
        self.tile_size = _triple(voxel_size, "voxel_size")
        self.tile_step = _triple(voxel_step, "voxel_step")
        self.weight = weight
        for axis in range(3):
            if self.tile_step[axis] < 1 or self.tile_step[axis] > self.tile_size[axis]:
                raise ValueError()
        overlap = self.tile_size - self.tile_step
        self.num_tiles = np.maximum(1, np.ceil((self.volume_shape - overlap) / self.tile_step)).astype(int)
        self.extra_pad = self.tile_step * self.num_tiles - (self.volume_shape - overlap)
        self.pad_before = self.extra_pad // 2
        self.pad_after = self.extra_pad - self.pad_before
        self.orignal_image_roi = tuple(slice(self.pad_before[a], self.pad_before[a] + self.volume_shape[a]) for a in range(3))
        self.orignal_mask_roi = (slice(None),) + self.orignal_image_roi
        rois, bbox_crops = [], []
        ranges = [range(0, self.volume_shape[a] + self.extra_pad[a] - self.tile_size[a] + 1, self.tile_step[a]) for a in range(3)]
        for i in ranges[0]:
            for j in ranges[1]:
                for k in ranges[2]:
                    o = (i, j, k)
                    rois.append(tuple(slice(o[a], o[a] + self.tile_size[a]) for a in range(3)))
                    bbox_crops.append(tuple(slice(o[a] - self.pad_before[a], o[a] + self.tile_size[a]) for a in range(3)))
        self.crops = rois
        self.bbox_crops = bbox_crops

    # ---------------------------------------------------------------- tiles out
    def _split_device(self, volume: np.ndarray, value=0) -> torch.Tensor:
        if (np.array(volume.shape) != self.volume_shape).any():
            raise ValueError(f"Volume shape {volume.shape} is not equal to the expected {self.volume_shape}")
        N.require_cuda("VolumeSlicer.split")
        vol = torch.from_numpy(np.ascontiguousarray(volume)).cuda()
        d, h, w = (int(v) for v in self.tile_size)
        n = len(self.crops)
        out = torch.empty((n, d, h, w), dtype=vol.dtype, device=vol.device)
        origins = np.array([[roi[a].start - self.pad_before[a] for a in range(3)] for roi in self.crops], dtype=np.int32)
        zyx = torch.from_numpy(origins).to(vol.device)
        fill = np.zeros(1, dtype=np.uint64)
        fill.view(volume.dtype)[0] = np.asarray(value).astype(volume.dtype)
        D, H, W = (int(v) for v in self.volume_shape)
        rc = N.lib().tb_volume_crop_tiles(vol.data_ptr(), vol.element_size(), D, H, W, zyx.data_ptr(), n, d, h, w, int(fill[0]),
                                          out.data_ptr(), N.stream_ptr(vol.device))
        N.check(rc, "tb_volume_crop_tiles")
        return out

    def split(self, volume: np.ndarray, value=0) -> List[np.ndarray]:
        """Constant-padded tiles of ``volume`` as host arrays, in ``crops`` order (reference tiles_3d.py:106-118)."""
        tiles = self._split_device(volume, value).cpu().numpy()
        return [tiles[t] for t in range(tiles.shape[0])]

    def iter_split(self, volume, value=0) -> Iterable[Tuple[np.ndarray, Any]]:
        tiles = self._split_device(volume, value).cpu().numpy()
        for t, roi in enumerate(self.crops):
            yield tiles[t], roi

    @property
    def target_shape(self):
        return self.volume_shape + self.extra_pad

    def crop_to_orignal_size(self, volume):
        return volume[self.orignal_image_roi]

    # ---------------------------------------------------------------- weights (helpers the reference exposes)
    def _mean(self, volume_size):
        return np.ones(volume_size, dtype=np.float32)

    def _pyramid(self, tile_size):
        w, _, _ = compute_pyramid_patch_weight_loss(tile_size[0], tile_size[1])
        return w

    # ---------------------------------------------------------------- tiles in
    def merge(self, tiles: Sequence[np.ndarray], weight: np.ndarray = None, dtype=np.float32) -> np.ndarray:
        """Weighted overlap-add of per-tile [d,h,w] host arrays back to the [D,H,W] volume (float64 accumulation, then
        ``dtype``): the documented intent of the reference's ``merge`` (its body is dead 2-D code, tiles_3d.py:140-166)."""
        if len(tiles) != len(self.crops):
            raise ValueError
        w = np.ones(tuple(self.tile_size), dtype=np.float64) if weight is None else np.asarray(weight, dtype=np.float64)
        pred = torch.from_numpy(np.stack([np.asarray(t, dtype=np.float32) for t in tiles])[:, None]).cuda()
        merged = self.merge_all(pred, w.astype(np.float32))
        return merged[0].cpu().numpy().astype(dtype)

    def merge_all(self, pred: torch.Tensor, weight: np.ndarray) -> torch.Tensor:
        """All tile predictions [n,K,d,h,w] (``crops`` order, on the GPU) -> merged, normalised, cropped [K,D,H,W] in one
        output-stationary launch (= ``VolumeMerger.integrate_batch`` over all tiles + ``merge`` + crop)."""
        N.require_cuda("VolumeSlicer.merge_all")
        n = len(self.crops)
        d, h, w = (int(v) for v in self.tile_size)
        if not pred.is_cuda or pred.dim() != 5 or tuple(pred.shape[2:]) != (d, h, w) or pred.shape[0] != n:
            raise ValueError(f"pred must be a CUDA tensor [{n},K,{d},{h},{w}], got {tuple(pred.shape)}")
        pred = pred.to(torch.float32).contiguous()
        k = pred.shape[1]
        wt = torch.from_numpy(np.ascontiguousarray(np.asarray(weight, dtype=np.float32))).to(pred.device)
        if tuple(wt.shape) != (d, h, w):
            raise ValueError(f"weight must be [{d},{h},{w}]")
        D, H, W = (int(v) for v in self.volume_shape)
        out = torch.empty((k, D, H, W), dtype=torch.float32, device=pred.device)
        sd, sh, sw = (int(v) for v in self.tile_step)
        nd, nh, nw = (int(v) for v in self.num_tiles)
        pd, ph, pw = (int(v) for v in self.pad_before)
        with torch.cuda.device(pred.device):
            rc = N.lib().tb_volume_merge_all(pred.data_ptr(), wt.data_ptr(), k, d, h, w, sd, sh, sw, nd, nh, nw, pd, ph, pw, D, H, W,
                                             out.data_ptr(), N.stream_ptr(pred.device))
        N.check(rc, "tb_volume_merge_all")
        return out


class VolumeMerger:
    """Streaming merger of volume predictions (reference tiles_3d.py:180-211): ``volume [K,Dp,Hp,Wp]``, ``norm_mask
    [1,Dp,Hp,Wp]``, ``weight [1,d,h,w]`` are plain tensors on the GPU; the arithmetic runs in the CUDA kernels."""

    def __init__(self, volume_shape, channels: int, weight, device="cuda", dtype=torch.float32):
        N.require_cuda("VolumeMerger")
        if dtype != torch.float32:
            raise NotImplementedError("VolumeMerger accumulates in float32")
        dev = torch.device(device)
        if dev.type != "cuda":
            dev = torch.device("cuda", torch.cuda.current_device())  # results live on the GPU; .cpu() them if needed
        self.channels = channels
        shape = tuple(int(v) for v in volume_shape)
        self.weight = torch.from_numpy(np.expand_dims(np.asarray(weight), axis=0)).to(device=dev, dtype=dtype).contiguous()
        self.volume = torch.zeros((channels, *shape), device=dev, dtype=dtype)
        self.norm_mask = torch.zeros((1, *shape), device=dev, dtype=dtype)

    def integrate_batch(self, batch: torch.Tensor, rois):
        """``batch`` [B,C,d,h,w]; ``rois``: B tuples of 3 slices in the padded frame (``VolumeSlicer.crops`` entries)."""
        if len(batch) != len(rois):
            raise ValueError("Number of images in batch does not correspond to number of coordinates")
        if len(batch) == 0:
            return
        dev = self.volume.device
        batch = batch.to(device=dev, dtype=torch.float32).contiguous()
        d, h, w = self.weight.shape[1:]
        if tuple(batch.shape[1:]) != (self.channels, d, h, w):
            raise ValueError(f"batch must be [B,{self.channels},{d},{h},{w}], got {tuple(batch.shape)}")
        Dp, Hp, Wp = self.volume.shape[1:]
        origins = np.array([_roi_origin(r) for r in rois], dtype=np.int32)
        if (origins < 0).any() or (origins + np.array([d, h, w]) > np.array([Dp, Hp, Wp])).any():
            raise IndexError("a tile leaves the padded volume")
        zyx = torch.from_numpy(origins).to(dev)
        with torch.cuda.device(dev):
            rc = N.lib().tb_volume_integrate(batch.data_ptr(), zyx.data_ptr(), len(rois), self.channels, d, h, w,
                                             self.weight.data_ptr(), self.volume.data_ptr(), self.norm_mask.data_ptr(), Dp, Hp, Wp,
                                             N.stream_ptr(dev))
        N.check(rc, "tb_volume_integrate")
        self._keep = (batch, zyx)

    def accumulate_single(self, tile: torch.Tensor, roi):
        self.integrate_batch(tile.unsqueeze(0), [roi])

    def merge(self) -> torch.Tensor:
        out = torch.empty_like(self.volume)
        dev = self.volume.device
        with torch.cuda.device(dev):
            rc = N.lib().tb_volume_normalize(self.volume.data_ptr(), self.norm_mask.data_ptr(), self.channels,
                                             self.norm_mask.numel(), out.data_ptr(), N.stream_ptr(dev))
        N.check(rc, "tb_volume_normalize")
        return out
