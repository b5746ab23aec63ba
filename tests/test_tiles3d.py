"""3-D tiles. CPU part: oracle/tiles3d_oracle.py and the host grid of the drop-in VolumeSlicer against the
reference's outputs (tests/golden/tiles3d.npz, written by oracle/make_golden.py::golden_tiles3d from the imported reference;
the seeded inputs are regenerated here, only reference OUTPUTS are stored). GPU part: the CUDA kernels against both."""
import numpy as np
import pytest
import torch

from oracle import tiles3d_oracle as T3
from oracle.make_golden_inputs import VOLUME_CASES, volume_case_inputs
from pytorch_toolbelt_b200.inference.tiles_3d import VolumeMerger, VolumeSlicer


def test_oracle_and_host_grid_match_the_reference(golden):
    g = golden("tiles3d")
    for ci, (shape, tile, step) in enumerate(VOLUME_CASES):
        grid = T3.volume_grid(shape, tile, step)
        sl = VolumeSlicer(shape, tile, step)
        origins = np.array([[r.start for r in roi] for roi in sl.crops])
        assert np.array_equal(grid["origins"], g[f"c{ci}_origins"]) and np.array_equal(origins, g[f"c{ci}_origins"])
        assert np.array_equal(sl.target_shape, g[f"c{ci}_target"]) and np.array_equal(sl.pad_before, g[f"c{ci}_pad_before"])
        bbox = np.array([[r.start for r in roi] for roi in sl.bbox_crops])
        assert np.array_equal(bbox, origins - sl.pad_before)
        d, h, w = (int(v) for v in grid["tile"])
        vol, weight, preds = volume_case_inputs(ci, shape, len(origins), (d, h, w))
        tiles = T3.split(vol, grid, value=7)
        for slot, t in enumerate(g[f"c{ci}_tile_idx"]):
            assert np.array_equal(tiles[int(t)], g[f"c{ci}_tiles"][slot])
        np.testing.assert_array_equal(np.array([float(np.asarray(t, dtype=np.float64).sum()) for t in tiles]), g[f"c{ci}_tile_sums"])
        assert np.array_equal(T3.merge(torch.from_numpy(preds), grid, weight, 2).numpy(), g[f"c{ci}_merged"])
    with pytest.raises(ValueError):
        VolumeSlicer((10, 10, 10), 8, 9)
    with pytest.raises(ValueError):
        VolumeSlicer((10, 10, 10), (8, 8), 4)


@pytest.mark.gpu
def test_volume_kernels_match_the_reference(golden):
    g = golden("tiles3d")
    for ci, (shape, tile, step) in enumerate(VOLUME_CASES):
        sl = VolumeSlicer(shape, tile, step)
        d, h, w = (int(v) for v in sl.tile_size)
        vol, weight, preds = volume_case_inputs(ci, shape, len(sl.crops), (d, h, w))
        tiles = sl.split(vol, value=7)
        assert tiles[0].dtype == vol.dtype and len(tiles) == len(sl.crops)
        for slot, t in enumerate(g[f"c{ci}_tile_idx"]):
            assert np.array_equal(tiles[int(t)], g[f"c{ci}_tiles"][slot])
        np.testing.assert_array_equal(np.array([float(np.asarray(t, dtype=np.float64).sum()) for t in tiles]), g[f"c{ci}_tile_sums"])
        for (tile_arr, roi), want in zip(sl.iter_split(vol, value=7), tiles):
            assert np.array_equal(tile_arr, want) and roi in sl.crops
        # one-launch output-stationary merge: bit-identical to integrate_batch in crops order + merge + crop
        pred = torch.from_numpy(preds).cuda()
        merged = sl.merge_all(pred, weight)
        assert torch.equal(merged.cpu(), torch.from_numpy(g[f"c{ci}_merged"])), ci
        # streaming VolumeMerger (vector reductions: tile order is not fixed -> compare to rounding)
        merger = VolumeMerger(sl.target_shape, 2, weight)
        for s0 in range(0, len(sl.crops), 3):
            merger.integrate_batch(pred[s0 : s0 + 3], sl.crops[s0 : s0 + 3])
        streamed = merger.merge()[sl.orignal_mask_roi]
        torch.testing.assert_close(streamed.cpu(), torch.from_numpy(g[f"c{ci}_merged"]), rtol=1e-5, atol=1e-5)
        single = VolumeMerger(sl.target_shape, 2, weight)
        for t, roi in enumerate(sl.crops):
            single.accumulate_single(pred[t], roi)
        torch.testing.assert_close(single.merge()[sl.orignal_mask_roi], streamed, rtol=1e-5, atol=1e-5)
        # host-array merge (the documented intent of the reference's dead merge()): identity round trip with any weight
        back = sl.merge(tiles, weight=weight, dtype=np.float32)
        np.testing.assert_allclose(back, vol.astype(np.float32), rtol=1e-5, atol=1e-3)
    with pytest.raises(ValueError):
        sl.merge_all(pred[:-1], weight)
    with pytest.raises(ValueError):
        merger.integrate_batch(pred[:2], sl.crops[:3])


@pytest.mark.gpu
def test_volume_merge_seeded_larger_against_the_oracle():
    shape, tile, step = (96, 128, 160), (64, 64, 64), (32, 32, 32)
    sl = VolumeSlicer(shape, tile, step)
    grid = T3.volume_grid(shape, tile, step)
    rng = np.random.default_rng(3)
    weight = rng.random(tile).astype(np.float32) + 0.05
    preds = rng.standard_normal((len(sl.crops), 3, *tile)).astype(np.float32)
    want = T3.merge(torch.from_numpy(preds), grid, weight, 3)
    got = sl.merge_all(torch.from_numpy(preds).cuda(), weight)
    assert torch.equal(got.cpu(), want)
    vol = rng.integers(0, 65535, shape).astype(np.uint16)
    tiles = sl.split(vol)
    for a, b in zip(tiles, T3.split(vol, grid)):
        assert np.array_equal(a, b)
