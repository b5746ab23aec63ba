"""Fused B200 tiled-inference pipeline (new API; the reference has only the Python loop of README.md:196-227).

    engine = TiledInferenceEngine(image.shape, tile_size=512, tile_step=256, channels=4, weight="pyramid", tta="d4")
    logits = engine.predict(image, model, batch_size=8)          # [K,H,W] float32 on the GPU

Two kernels bracket the (black-box) model:

* ``slice``  : image resident in HBM -> crop + border + HWC->CHW + cast + per-channel affine + all TTA views,
               one launch, written once as the model's input batch (tb_slice_augment);
* ``merge``  : all model outputs -> inverse TTA permutation + view reduction + tile weight + overlap-add +
               normalisation + margin crop, one launch, output-stationary: every prediction element is read
               once, every output pixel written once, no atomics and no canvas / norm_mask in memory
               (tb_merge_all). Accumulation follows the reference's tile order, so the result matches
               ``TileMerger`` fed in row-major order bit for bit (given the same per-tile values).
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from .. import _native as N
from .._native import TB
from .tiles import BORDER_CONSTANT, ImageSlicer

__all__ = ["TiledInferenceEngine", "PreparedLaunch", "tta_code"]

_TTA_NAMES = {
    None: TB.TTA_NONE,
    "none": TB.TTA_NONE,
    "fliplr": TB.TTA_FLIPLR,
    "flipud": TB.TTA_FLIPUD,
    "flips": TB.TTA_FLIPS,
    "d2": TB.TTA_D2,
    "d4": TB.TTA_D4,
}


def tta_code(tta) -> int:
    if isinstance(tta, int) and not isinstance(tta, bool):
        if tta not in TB.TTA_VIEWS:
            raise ValueError(f"Unknown TTA code {tta}")
        return tta
    if tta not in _TTA_NAMES:
        raise ValueError(f"Unknown TTA group {tta!r}; expected one of {sorted(k for k in _TTA_NAMES if k)}")
    return _TTA_NAMES[tta]


class PreparedLaunch:
    """One kernel launch with all arguments already marshalled (ctypes structs, pointer arrays, tensor maps inputs).

    ``run()`` costs a single ctypes call: use it for loops that launch the same slice / merge on the same buffers
    many times (benchmarks, the multi-GPU strip merger, steady-state serving). The launch goes to the CUDA stream that
    was current when the plan was built and reads / writes the tensors it was built with.
    """

    __slots__ = ("fn", "args", "keep", "result", "device", "what", "fallback")

    def __init__(self, fn, args, keep, result, device, what, fallback=None):
        self.fn, self.args, self.keep, self.result, self.device, self.what = fn, args, keep, result, device, what
        self.fallback = fallback  # () -> (fn, args, what): taken once if the first choice answers TB_ERR_UNSUPPORTED

    def _call(self):
        if torch.cuda.current_device() == self.device.index:
            return self.fn(*self.args)
        with torch.cuda.device(self.device):
            return self.fn(*self.args)

    def run(self):
        rc = self._call()
        if rc == -2 and self.fallback is not None:  # TB_ERR_UNSUPPORTED (nothing was launched): switch kernels for good
            self.fn, self.args, self.what = self.fallback()
            self.fallback = None
            rc = self._call()
        if rc != 0:
            N.check(rc, self.what)
        return self.result


class TiledInferenceEngine:
    def __init__(
        self,
        image_shape: Sequence[int],
        tile_size,
        tile_step,
        channels: int,
        weight: Union[str, np.ndarray] = "pyramid",
        tta=None,
        reduction: str = "mean",
        image_margin=0,
        border_type: int = BORDER_CONSTANT,
        border_value=0,
        scale: Optional[Sequence[float]] = None,
        bias: Optional[Sequence[float]] = None,
        analytic_weight: bool = False,
        device: Optional[Union[str, torch.device]] = None,
        activation: Optional[str] = None,
        temperature: float = 1.0,
    ):
        """
        :param image_shape: (H, W[, C]) of the source image
        :param channels: number of model output channels K
        :param weight: "mean" | "pyramid" | "gaussian" | ndarray [Th,Tw]
        :param tta: None | "fliplr" | "flipud" | "flips" | "d2" | "d4"
        :param reduction: how TTA views are averaged ("mean", "sum", "gmean", "hmean", "harmonic1p", "logodd", "log1p")
        :param scale, bias: optional per-input-channel affine applied while slicing: x = float(pixel)*scale[c] + bias[c]
        :param analytic_weight: regenerate "pyramid"/"gaussian"/"mean" weights in registers (fp32, <= 3e-7 relative
               to the float64->float32 table) instead of reading the table
        :param activation, temperature: None | "sigmoid": applied to every prediction BEFORE the TTA view reduction, inside
               the merge kernel -- what wrapping the model in the reference's ``ApplySigmoidTo(model, temperature=...)``
               does with two more passes over the prediction stream (inference/ensembling.py:46-68). ("softmax" couples
               the channels of a pixel: use ``ensembling.ApplySoftmaxTo`` on the model, one extra in-place pass.)
        """
        N.require_cuda("TiledInferenceEngine")
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.slicer = ImageSlicer(image_shape, tile_size, tile_step, image_margin=image_margin, weight=weight)
        self.channels = int(channels)
        self.tta = tta_code(tta)
        self.views = TB.TTA_VIEWS[self.tta]
        if reduction not in TB.REDUCTIONS or TB.REDUCTIONS[reduction] == TB.RED_NONE:
            raise KeyError(f"Unsupported reduction mode {reduction}")
        self.reduction = TB.REDUCTIONS[reduction]
        th, tw = self.slicer.tile_size
        if self.tta == TB.TTA_D4 and th != tw:
            raise ValueError(f"D4 TTA needs square tiles, got {th}x{tw}")
        self.border_type = int(border_type)
        vals = list(border_value)[:4] if isinstance(border_value, (tuple, list, np.ndarray)) else [border_value]
        self._border_value = N.float_array(vals, 4)
        self._scale, self._bias = scale, bias
        self.num_tiles = len(self.slicer.crops)
        self._bbox_xy = N.as_dev_xy(self.slicer.bbox_crops, self.device)
        self._weight_table = torch.from_numpy(np.asarray(self.slicer.weight, dtype=np.float32)).to(self.device).contiguous()
        if analytic_weight and self.slicer.weight_kind != TB.W_TABLE:
            self._weight = N.Weight(self.slicer.weight_kind, None, float(self.slicer.weight_param))
        else:
            self._weight = N.Weight(TB.W_TABLE, self._weight_table.data_ptr(), 0.0)
        ny, nx = self.slicer.grid_shape
        sh, sw = self.slicer.tile_step
        self._grid_dims = (nx, ny, sw, sh, th, tw)
        self._keepalive: List[torch.Tensor] = []
        if activation not in (None, "none", "sigmoid"):
            raise ValueError(f"activation must be None or 'sigmoid', got {activation!r}")
        self._pre = TB.PRE_SIGMOID if activation == "sigmoid" else TB.PRE_NONE
        self._temperature = float(temperature)

    # ------------------------------------------------------------------ image upload
    def to_device(self, image: Union[np.ndarray, torch.Tensor]) -> torch.Tensor:
        """Host ndarray / CPU tensor -> contiguous CUDA tensor (HWC or HW), dtype preserved."""
        if isinstance(image, np.ndarray):
            image = torch.from_numpy(np.ascontiguousarray(image))
        if not image.is_cuda:
            image = image.to(self.device, non_blocking=True)
        return image.contiguous()

    # ------------------------------------------------------------------ slice (+augment)
    def slice_plan(self, image: torch.Tensor, start: int = 0, stop: Optional[int] = None, layout: str = "HWC",
                   out: Optional[torch.Tensor] = None) -> PreparedLaunch:
        """Prepared launch of ``slice`` (see PreparedLaunch)."""
        stop = self.num_tiles if stop is None else stop
        n = stop - start
        if not (0 <= start <= stop <= self.num_tiles):
            raise IndexError(f"tile range [{start},{stop}) outside [0,{self.num_tiles})")
        if not image.is_cuda:
            raise ValueError("slice() expects the image resident on the GPU; call to_device() first")
        image = image.contiguous()
        if layout == "HWC":
            h, w = image.shape[0], image.shape[1]
            c = 1 if image.dim() == 2 else image.shape[2]
            layout_code = TB.HWC
        elif layout == "CHW":
            c, h, w = image.shape
            layout_code = TB.CHW
        else:
            raise ValueError("layout must be 'HWC' or 'CHW'")
        if h != self.slicer.image_height or w != self.slicer.image_width:
            raise ValueError(f"image is {h}x{w}, engine was built for {self.slicer.image_height}x{self.slicer.image_width}")
        if image.dtype not in TB.TORCH_DTYPE:
            raise TypeError(f"unsupported image dtype {image.dtype}")
        th, tw = self.slicer.tile_size
        shape = (self.views * n, c, th, tw)
        if out is None:
            out = torch.empty(shape, dtype=torch.float32, device=image.device)
        elif tuple(out.shape) != shape or out.dtype != torch.float32 or not out.is_contiguous():
            raise ValueError(f"out must be a contiguous float32 tensor of shape {shape}")
        if n == 0:
            return PreparedLaunch(lambda: 0, (), [], out, image.device, "tb_slice_augment")
        xy = self._bbox_xy[start:stop]
        scale = N.float_array(self._scale, c) if self._scale is not None else None
        bias = N.float_array(self._bias, c) if self._bias is not None else None
        args = (image.data_ptr(), TB.TORCH_DTYPE[image.dtype], layout_code, h, w, c, xy.data_ptr(), n, th, tw, self.tta,
                self.border_type, self._border_value, scale, bias, out.data_ptr(), N.stream_ptr(image.device))
        return PreparedLaunch(N.lib().tb_slice_augment, args, [image, xy, out, scale, bias], out, image.device,
                              "tb_slice_augment")

    def slice(self, image: torch.Tensor, start: int = 0, stop: Optional[int] = None, layout: str = "HWC",
              out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Tiles [start, stop) of the grid as the model input batch: float32 [A*n, C, Th, Tw], aug-major."""
        return self.slice_plan(image, start, stop, layout, out).run()

    # ------------------------------------------------------------------ merge
    def _grid(self, n_slots: int) -> N.Grid:
        nx, ny, sw, sh, th, tw = self._grid_dims
        return N.Grid(nx, ny, sw, sh, th, tw, n_slots, None, None, None)

    def merge(
        self,
        preds: Union[torch.Tensor, Sequence[torch.Tensor]],
        out: Optional[torch.Tensor] = None,
        rows: Optional[Tuple[int, int]] = None,
        tile_rows: Optional[Tuple[int, int]] = None,
        raw: bool = False,
        init: Optional[Tuple[torch.Tensor, torch.Tensor]] = None,
        output: str = "logits",
        threshold: float = 0.5,
    ):
        """All predictions -> merged, normalised, cropped logits float32 [K,H,W] (see ``merge_plan`` for the arguments)."""
        plan = self.merge_plan(preds, out, rows, tile_rows, raw, init, output, threshold)
        result = plan.run()
        # tables / casts must outlive the asynchronous launch: park them until the next call on this engine
        self._keepalive = plan.keep
        return result

    def merge_plan(
        self,
        preds: Union[torch.Tensor, Sequence[torch.Tensor]],
        out: Optional[torch.Tensor] = None,
        rows: Optional[Tuple[int, int]] = None,
        tile_rows: Optional[Tuple[int, int]] = None,
        raw: bool = False,
        init: Optional[Tuple[torch.Tensor, torch.Tensor]] = None,
        output: str = "logits",
        threshold: float = 0.5,
        first_tile: Optional[int] = None,
        peer: Optional[dict] = None,
    ) -> PreparedLaunch:
        """Prepared launch: all predictions -> merged, normalised, cropped logits float32 [K,H,W].

        ``preds`` is either one tensor [A*N,K,Th,Tw] (aug-major over the whole grid) or the list of the model's
        per-batch outputs ([A*b,K,Th,Tw] each, batches in grid order) -- the latter is merged in place (one TMA
        tensor map per batch, or a device pointer table off the 32-pixel lattice), without concatenating.

        Band mode (multi-GPU strips): ``rows=(y0,y1)`` canvas rows to produce, ``tile_rows=(r0,r1)`` contributing
        tile rows, ``raw=True`` returns un-normalised ``(num [K,y1-y0,Wp], den [y1-y0,Wp])``, ``init`` seeds the
        accumulation with such a pair (the halo received from the previous strip).

        ``output="labels"`` writes ``merged.argmax(0)`` as uint8 [H,W], ``output="mask"`` writes
        ``merged.sigmoid() > threshold`` as uint8 [K,H,W] straight from the merge epilogue: the float canvas is never
        written (README.md:225 does this with two more full-canvas passes) and one byte per pixel leaves the GPU.
        """
        if output not in ("logits", "labels", "mask"):
            raise ValueError("output must be 'logits', 'labels' or 'mask'")
        if output != "logits" and raw:
            raise ValueError("label / mask output needs the normalised (raw=False) mode")
        nx, ny, sw, sh, th, tw = self._grid_dims
        hp, wp = self.slicer.target_shape
        k = self.channels
        y0, y1 = rows if rows is not None else (0, hp)
        r0, r1 = tile_rows if tile_rows is not None else (0, ny)
        grid = self._grid(self.num_tiles)
        keep = []
        plane_set = k * th * tw
        if torch.is_tensor(preds):
            preds = [preds]
        batches = []
        for batch in preds:
            if not batch.is_cuda:
                raise ValueError("predictions must live on the GPU")
            batch = batch.to(torch.float32).contiguous()  # no-ops for the expected layout
            if tuple(batch.shape[1:]) != (k, th, tw) or batch.shape[0] % self.views != 0:
                raise ValueError(f"predictions must be [A*b,{k},{th},{tw}], got {tuple(batch.shape)}")
            if batch.data_ptr() % 16 != 0:
                raise ValueError("prediction tensors must be 16-byte aligned")
            batches.append(batch)
        keep += batches
        sizes = [b.shape[0] // self.views for b in batches]
        total = sum(sizes)
        if first_tile is not None:  # an explicit row-major tile range is resident (multi-GPU ragged strips)
            if first_tile < 0 or first_tile + total > self.num_tiles:
                raise ValueError(f"tiles [{first_tile},{first_tile + total}) outside the grid of {self.num_tiles}")
        elif total == self.num_tiles:
            first_tile = 0  # the whole grid is resident; tile_rows (if any) only restricts who contributes
        elif tile_rows is not None and total == (r1 - r0) * nx:
            first_tile = r0 * nx  # only this strip's tiles are resident (multi-GPU)
        else:
            raise ValueError(f"predictions cover {total} tiles; expected the whole grid ({self.num_tiles})"
                             + (f" or the strip's {(r1 - r0) * nx}" if tile_rows is not None else ""))
        device = batches[0].device
        uniform = all(sz == sizes[0] for sz in sizes[:-1]) and sizes[-1] <= sizes[0]
        whole_grid_single = len(batches) == 1 and sizes[0] == self.num_tiles and first_tile == 0
        mode = "single" if whole_grid_single else ("batched" if uniform else "table")

        def pointer_table():
            ptrs = np.zeros(self.num_tiles, dtype=np.int64)
            vstr = np.zeros(self.num_tiles, dtype=np.int64)
            t = first_tile
            for batch, b in zip(batches, sizes):
                ptrs[t : t + b] = batch.data_ptr() + np.arange(b, dtype=np.int64) * (plane_set * 4)
                vstr[t : t + b] = b * plane_set
                t += b
            ptr_dev = torch.from_numpy(ptrs).to(device)
            vstr_dev = torch.from_numpy(vstr).to(device)
            keep.extend([ptr_dev, vstr_dev])
            grid.tile_ptr = ptr_dev.data_ptr()
            grid.tile_vstride = vstr_dev.data_ptr()

        pred_ptr = batches[0].data_ptr() if mode == "single" else None
        if mode == "table":
            pointer_table()
        top, left = self.slicer.margin_top, self.slicer.margin_left
        h, w = self.slicer.image_height, self.slicer.image_width
        init_num = init_den = None
        if init is not None:
            init_num, init_den = init
            if tuple(init_num.shape) != (k, y1 - y0, wp) or tuple(init_den.shape) != (y1 - y0, wp):
                raise ValueError("init must be (num [K,rows,Wp], den [rows,Wp])")
            init_num, init_den = init_num.contiguous(), init_den.contiguous()
            keep.extend([init_num, init_den])
        if raw:
            num = torch.empty((k, y1 - y0, wp), dtype=torch.float32, device=device)
            den = torch.empty((y1 - y0, wp), dtype=torch.float32, device=device)
            out_num, out_den, result = num.data_ptr(), den.data_ptr(), (num, den)
        elif output == "logits":
            if out is None:
                out = torch.empty((k, h, w), dtype=torch.float32, device=device)
            elif tuple(out.shape) != (k, h, w) or out.dtype != torch.float32 or not out.is_contiguous():
                raise ValueError(f"out must be a contiguous float32 tensor [{k},{h},{w}]")
            out_num, out_den, result = out.data_ptr(), None, out
        else:
            u8_shape = (h, w) if output == "labels" else (k, h, w)
            if out is None:
                out = torch.empty(u8_shape, dtype=torch.uint8, device=device)
            elif tuple(out.shape) != u8_shape or out.dtype != torch.uint8 or not out.is_contiguous():
                raise ValueError(f"out must be a contiguous uint8 tensor {u8_shape} for output={output!r}")
            out_num, out_den, result = None, None, out
        peer_ref = None
        if peer is not None:
            if peer["t0"] != first_tile or peer["t1"] != first_tile + total:
                raise ValueError("peer plan tile range must equal the resident tile range")
            peer_struct = N.PeerPlan(int(peer["phase_mask"]), int(peer["t0"]), int(peer["t1"]), peer.get("seed_num"),
                                     peer.get("seed_den"), int(peer.get("seed_y0", 0)), int(peer.get("seed_rows", 0)),
                                     peer.get("spill_num"), peer.get("spill_den"), int(peer.get("spill_y0", 0)),
                                     int(peer.get("spill_rows", 0)),
                                     int(peer.get("fused", 0)), int(peer.get("chained", 0)), int(peer.get("has_next", 0)),
                                     int(peer.get("has_prev", 0)), int(peer.get("epoch", 0)), peer.get("counters"),
                                     peer.get("ready_flag"), peer.get("ack_flag"), peer.get("next_ready"), peer.get("prev_ack"),
                                     peer.get("status"), int(peer.get("timeout_ms", 5000)))
            keep.append(peer_struct)
            peer["struct"] = peer_struct  # the caller advances `.epoch` before every run of a fused plan
            peer_ref = C.pointer(peer_struct)
        extra = N.MergeExtra(self._pre, self._temperature,
                             {"logits": TB.OUT_LOGITS, "labels": TB.OUT_ARGMAX, "mask": TB.OUT_MASK}[output], float(threshold),
                             out.data_ptr() if (output != "logits" and not raw) else None, peer_ref)
        grid_ref = C.byref(grid)
        common = (k, self.tta, self.reduction, C.byref(self._weight), hp, wp, y0, y1, r0, r1, N.ptr(init_num), N.ptr(init_den),
                  1 if raw else 0, top, left, h, w, out_num, out_den, C.byref(extra), N.stream_ptr(device))
        keep += [grid, self._weight, self._weight_table, result, extra]
        if mode == "batched":
            ptr_arr = (C.c_void_p * len(batches))(*[b.data_ptr() for b in batches])
            keep.append(ptr_arr)
            args = (ptr_arr, len(batches), sizes[0], sizes[-1], first_tile, grid_ref) + common

            def to_table():  # geometry off the TMA lattice -> pointer-table kernel
                pointer_table()
                return N.lib().tb_merge_all_ex, (None, grid_ref) + common, "tb_merge_all_ex"

            return PreparedLaunch(N.lib().tb_merge_all_batched_ex, args, keep, result, device, "tb_merge_all_batched_ex", to_table)
        args = (pred_ptr, grid_ref) + common
        return PreparedLaunch(N.lib().tb_merge_all_ex, args, keep, result, device, "tb_merge_all_ex")

    # ------------------------------------------------------------------ output side
    def labels(self, logits: torch.Tensor, mode: str = "argmax", threshold: float = 0.5,
               out: Optional[torch.Tensor] = None, rows: Optional[Tuple[int, int]] = None) -> torch.Tensor:
        """uint8 class map [H,W] (``mode="argmax"``, = ``logits.argmax(0).byte()``) or uint8 masks [K,H,W]
        (``mode="threshold"``, = ``(logits.sigmoid() > threshold).byte()``) from merged float32 logits [K,H,W];
        ``rows=(i0,i1)`` restricts the work to those image rows (the pipelined host path converts band by band)."""
        if mode not in ("argmax", "threshold"):
            raise ValueError("mode must be 'argmax' or 'threshold'")
        if not logits.is_cuda or logits.dtype != torch.float32 or logits.dim() != 3 or not logits.is_contiguous():
            raise ValueError("labels() expects contiguous float32 CUDA logits [K,H,W]")
        k, h, w = logits.shape
        shape = (h, w) if mode == "argmax" else (k, h, w)
        if out is None:
            out = torch.empty(shape, dtype=torch.uint8, device=logits.device)
        elif tuple(out.shape) != shape or out.dtype != torch.uint8 or not out.is_contiguous() or out.device != logits.device:
            raise ValueError(f"out must be a contiguous uint8 tensor of shape {shape} on {logits.device}")
        i0, i1 = rows if rows is not None else (0, h)
        if not (0 <= i0 <= i1 <= h):
            raise IndexError(f"rows {rows} outside [0,{h}]")
        with torch.cuda.device(logits.device):
            rc = N.lib().tb_logits_to_labels(
                logits.data_ptr() + i0 * w * 4, k, h * w, (i1 - i0) * w, 0 if mode == "argmax" else 1, float(threshold),
                out.data_ptr() + i0 * w, h * w, N.stream_ptr(logits.device))
        N.check(rc, "tb_logits_to_labels")
        return out

    # ------------------------------------------------------------------ whole pipeline
    @torch.no_grad()
    def predict(self, image: Union[np.ndarray, torch.Tensor], model: Callable[[torch.Tensor], torch.Tensor],
                batch_size: Optional[int] = None, layout: str = "HWC", output: str = "logits",
                threshold: float = 0.5) -> torch.Tensor:
        """image (host or device) -> [K,H,W] float32 logits on the GPU (or the uint8 class map / masks, see ``merge_plan``).
        ``batch_size`` counts tiles (each contributes A views to the model batch); None = all tiles in one model call."""
        image = self.to_device(image)
        batch_size = self.num_tiles if batch_size is None else int(batch_size)
        outputs = []
        for start in range(0, self.num_tiles, batch_size):
            stop = min(start + batch_size, self.num_tiles)
            pred = model(self.slice(image, start, stop, layout=layout))
            if pred.dtype != torch.float32:
                pred = pred.float()
            outputs.append(pred.contiguous())
        return self.merge(outputs[0] if len(outputs) == 1 else outputs, output=output, threshold=threshold)

    # ------------------------------------------------------------------ host -> host, pipelined over tile rows
    @torch.no_grad()
    def predict_to_host(self, image_host: torch.Tensor, model: Callable[[torch.Tensor], torch.Tensor],
                        out_host: Optional[torch.Tensor] = None, rows_per_chunk: int = 2, layout: str = "HWC",
                        output: str = "logits", threshold: float = 0.5) -> torch.Tensor:
        """Pinned host image -> pinned host logits [K,H,W], software-pipelined over chunks of tile rows.

        ``output="labels"`` / ``"mask"``: every finished band leaves the merge epilogue as the uint8 class map [H,W] /
        masks [K,H,W] (no float canvas at all), so 1 byte per pixel crosses PCIe instead of 4*K.

        For every chunk: (copy stream A) upload the image rows the chunk needs; (compute stream) slice+augment, model,
        merge the canvas rows that are now final (band mode of tb_merge_all_batched over the model outputs so far);
        (copy stream B) download those output rows. PCIe transfers in both directions overlap the kernels; the
        arithmetic and its order are identical to ``predict`` (bit-identical result).
        """
        if layout != "HWC":
            raise ValueError("predict_to_host streams HWC images")
        dev = self.device
        k = self.channels
        h, w = self.slicer.image_height, self.slicer.image_width
        nx, ny, sw, sh, th, tw = self._grid_dims
        hp, wp = self.slicer.target_shape
        top = self.slicer.margin_top
        if output not in ("logits", "labels", "mask"):
            raise ValueError("output must be 'logits', 'labels' or 'mask'")
        if not image_host.is_pinned():
            image_host = image_host.pin_memory()
        u8_shape = (h, w) if output == "labels" else (k, h, w)
        if out_host is None:
            out_host = (torch.empty((k, h, w), dtype=torch.float32) if output == "logits"
                        else torch.empty(u8_shape, dtype=torch.uint8)).pin_memory()
        elif output != "logits" and (tuple(out_host.shape) != u8_shape or out_host.dtype != torch.uint8):
            raise ValueError(f"out_host must be uint8 {u8_shape} for output={output!r}")
        if not hasattr(self, "_pipe"):
            self._pipe = dict(h2d=torch.cuda.Stream(dev), d2h=torch.cuda.Stream(dev),
                              image=torch.empty(image_host.shape, dtype=image_host.dtype, device=dev), out=None)
        pipe = self._pipe
        if output == "logits" and pipe["out"] is None:  # the float canvas exists only when float logits are asked for
            pipe["out"] = torch.empty((k, h, w), dtype=torch.float32, device=dev)
        image_dev, out_dev = pipe["image"], pipe["out"]
        u8_dev = None
        if output != "logits":
            u8_dev = pipe.get("u8")
            if u8_dev is None or tuple(u8_dev.shape) != u8_shape:
                u8_dev = pipe["u8"] = torch.empty(u8_shape, dtype=torch.uint8, device=dev)
        if image_dev.shape != image_host.shape or image_dev.dtype != image_host.dtype:
            image_dev = pipe["image"] = torch.empty(image_host.shape, dtype=image_host.dtype, device=dev)
        main = torch.cuda.current_stream(dev)
        pipe["h2d"].wait_stream(main)  # the previous call's kernels may still read image_dev
        pipe["d2h"].wait_stream(main)
        # border modes whose source rows can lie far from the tile (WRAP) need the whole image up front
        stream_rows = self.border_type != 3
        uploaded, merged_until = 0, 0
        outputs: List[torch.Tensor] = []
        for r in range(0, ny, rows_per_chunk):
            rb = min(r + rows_per_chunk, ny)
            need = h if (rb == ny or not stream_rows) else min(h, max(0, (rb - 1) * sh + th - top))
            if need > uploaded:
                with torch.cuda.stream(pipe["h2d"]):
                    image_dev[uploaded:need].copy_(image_host[uploaded:need], non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(pipe["h2d"])
                main.wait_event(ev)
                uploaded = need
            x = self.slice(image_dev, r * nx, rb * nx)
            y = model(x)
            outputs.append(y if y.dtype == torch.float32 and y.is_contiguous() else y.float().contiguous())
            y_cur = hp if rb == ny else rb * sh
            # finished band -> float logits, or (fused in the merge epilogue) the uint8 class map / masks: 1 byte per pixel
            self.merge(outputs, out=out_dev if u8_dev is None else u8_dev, rows=(merged_until, y_cur), tile_rows=(0, rb),
                       output=output, threshold=threshold)
            o0, o1 = max(merged_until - top, 0), min(y_cur - top, h)
            if o1 > o0:
                ev2 = torch.cuda.Event()
                ev2.record(main)
                pipe["d2h"].wait_event(ev2)
                with torch.cuda.stream(pipe["d2h"]):
                    if output == "logits":
                        for kk in range(k):  # per-channel row blocks are contiguous on both sides: truly async copies
                            out_host[kk, o0:o1].copy_(out_dev[kk, o0:o1], non_blocking=True)
                    elif output == "labels":
                        out_host[o0:o1].copy_(u8_dev[o0:o1], non_blocking=True)
                    else:
                        for kk in range(k):
                            out_host[kk, o0:o1].copy_(u8_dev[kk, o0:o1], non_blocking=True)
            merged_until = y_cur
        main.wait_stream(pipe["d2h"])
        self._keepalive_outputs = outputs  # kernels of this call may still be reading them
        # the result is a HOST tensor: do not hand it back before the asynchronous downloads have landed (stream order
        # alone does not hold the host back). Only the d2h stream is waited for, not the whole device.
        pipe["d2h"].synchronize()
        return out_host
