"""CPU restatement (NumPy / torch-CPU) of the reference's 3-D tile path: inference/tiles_3d.py (VolumeSlicer grid :53-104,
split :106-118, VolumeMerger.integrate_batch / merge :195-211). TEST INFRASTRUCTURE ONLY.
Pinned by oracle/make_golden.py::golden_tiles3d against the imported reference (tests/golden/tiles3d.npz)."""
import numpy as np
import torch


def volume_grid(volume_shape, tile, step):
    """tiles_3d.py:53-104 -> dict(num_tiles, extra_pad, pad_before, pad_after, origins [n,3] (padded frame), target_shape)."""
    vs = np.array(volume_shape)[:3]
    tile = np.array([tile] * 3 if np.isscalar(tile) else tile, dtype=int)
    step = np.array([step] * 3 if np.isscalar(step) else step, dtype=int)
    overlap = tile - step
    num = np.maximum(1, np.ceil((vs - overlap) / step)).astype(int)
    extra = step * num - (vs - overlap)
    before = extra // 2
    after = extra - before
    origins = [(i, j, k)
               for i in range(0, vs[0] + extra[0] - tile[0] + 1, step[0])
               for j in range(0, vs[1] + extra[1] - tile[1] + 1, step[1])
               for k in range(0, vs[2] + extra[2] - tile[2] + 1, step[2])]
    return dict(num_tiles=num, extra_pad=extra, pad_before=before, pad_after=after, origins=np.array(origins, dtype=np.int64),
                target_shape=vs + extra, tile=tile, step=step)


def split(volume, grid, value=0):
    """tiles_3d.py:106-118."""
    pad = np.stack([grid["pad_before"], grid["pad_after"]], axis=-1)
    padded = np.pad(volume, pad, mode="constant", constant_values=value)
    d, h, w = grid["tile"]
    return [padded[i : i + d, j : j + h, k : k + w].copy() for i, j, k in grid["origins"]]


def merge(tiles: torch.Tensor, grid, weight: np.ndarray, channels: int) -> torch.Tensor:
    """VolumeMerger(...).integrate_batch(all tiles, crops order) + merge() + crop (tiles_3d.py:185-211,168-169)."""
    w = torch.from_numpy(np.expand_dims(np.asarray(weight), 0)).float()
    tgt = tuple(int(v) for v in grid["target_shape"])
    volume = torch.zeros((channels, *tgt))
    norm = torch.zeros((1, *tgt))
    d, h, wd = (int(v) for v in grid["tile"])
    for tile, (i, j, k) in zip(tiles, grid["origins"]):
        volume[:, i : i + d, j : j + h, k : k + wd] += tile * w
        norm[:, i : i + d, j : j + h, k : k + wd] += w
    merged = volume / norm
    b = grid["pad_before"]
    D, H, W = (int(v) for v in (np.array(tgt) - grid["extra_pad"]))
    return merged[:, b[0] : b[0] + D, b[1] : b[1] + H, b[2] : b[2] + W].contiguous()
