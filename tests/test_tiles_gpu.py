"""GPU: CUDA tile crop / merge against the oracle and the golden vectors of the real reference.
Bit-exact for crops; merged float buffers within 1e-5 relative to buffer scale (BASELINE north_star)."""
import numpy as np
import pytest
import torch

from oracle import tiles_oracle as TO
from pytorch_toolbelt_b200.inference.tiles import CudaTileMerger, ImageSlicer, TileMerger

pytestmark = pytest.mark.gpu

BORDER_IDS = {"constant": 0, "replicate": 1, "reflect": 2, "wrap": 3, "reflect101": 4}


def assert_buffer_close(ours, ref, rel=1e-5):
    ours, ref = np.asarray(ours, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    scale = np.abs(ref).max()
    assert np.abs(ours - ref).max() <= rel * scale, (np.abs(ours - ref).max(), scale)
    np.testing.assert_allclose(ours, ref, rtol=1e-5, atol=1e-6 * scale)


def test_split_bit_exact_vs_golden(golden):
    g = golden("tiles_split")
    checked = 0
    for key in g.files:
        for kind in ("_split_", "_itersplit_"):
            if kind in key and not key.startswith("big_"):
                tag, rest = key.split(kind)
                bname, vname = rest.split("_")
                value = 7 if vname == "s7" else (3, 5, 250, 9)
                image = g[f"{tag}_image"]
                s = ImageSlicer(image.shape, (24, 16), (12, 8), weight="pyramid")
                if kind == "_split_":
                    tiles = np.stack(s.split(image, border_type=BORDER_IDS[bname], value=value))
                else:
                    pairs = list(s.iter_split(image, border_type=BORDER_IDS[bname], value=value))
                    tiles = np.stack([t for t, _ in pairs])
                    assert all(np.array_equal(c, ref) for (_, c), ref in zip(pairs, s.crops))
                assert tiles.dtype == g[key].dtype and np.array_equal(tiles, g[key]), key
                checked += 1
    assert checked >= 15
    for tag in ("u8c3", "u8c1", "u8hw", "f32c3", "u16c4"):
        image = g[f"{tag}_image"]
        s = ImageSlicer(image.shape, (24, 16), (12, 8))
        patch = s.cut_patch(image, len(s.crops) - 1, border_type=4)
        assert np.array_equal(patch, g[f"{tag}_cutpatch_last_reflect101"])
        assert patch.ndim == image.ndim


def test_split_tile_larger_than_image(golden):
    g = golden("tiles_split")
    image = g["big_image"]
    s = ImageSlicer(image.shape, 64, 64, weight="mean")
    assert len(s.crops) == 1 and (s.margin_left, s.margin_right, s.margin_top, s.margin_bottom) == (7, 7, 12, 12)
    for bname, mode in BORDER_IDS.items():
        assert np.array_equal(np.stack(s.split(image, border_type=mode, value=11)), g[f"big_split_{bname}"]), bname
        it = np.stack([t for t, _ in s.iter_split(image, border_type=mode, value=11)])
        assert np.array_equal(it, g[f"big_itersplit_{bname}"]), bname


def test_split_vs_oracle_seeded_large():
    rng = np.random.default_rng(5)
    image = rng.integers(0, 256, (1000, 777, 3)).astype(np.uint8)
    s = ImageSlicer(image.shape, (256, 192), (128, 100), weight="mean")
    grid = TO.tile_grid(1000, 777, (256, 192), (128, 100))
    for mode in (0, 2, 4):
        ours = np.stack(s.split(image, border_type=mode, value=(9, 8, 7)))
        ref = np.stack(TO.split(image, grid, mode, (9, 8, 7)))
        assert np.array_equal(ours, ref)
    dev = torch.from_numpy(image).cuda()
    t = s.split(dev)
    assert t.is_cuda and t.shape == (len(s.crops), 256, 192, 3)
    assert np.array_equal(t.cpu().numpy(), np.stack(TO.split(image, grid, 0, 0)))


def test_reference_roundtrip_tests():
    """reference tests/test_tiles.py:13-26,47-55 with non-trivial data and mean weights (exact in float64)."""
    rng = np.random.default_rng(0)
    for shape, tile, step in [((500, 500, 3), 51, 26), ((563, 512, 3), (128, 128), (128, 128))]:
        image = rng.integers(0, 256, shape).astype(np.uint8)
        s = ImageSlicer(image.shape, tile_size=tile, tile_step=step, weight="mean")
        merged = s.merge(s.split(image), dtype=np.uint8)
        np.testing.assert_equal(merged, image)


def test_merger_vs_golden_and_oracle(golden):
    g = golden("tiles_merge")
    for wname in ("mean", "pyramid"):
        s = ImageSlicer((61, 53, 3), (24, 16), (12, 8), weight=wname)
        preds = torch.from_numpy(g[f"{wname}_preds"])
        for device, cls in (("cuda", TileMerger), ("cuda", CudaTileMerger), ("cpu", TileMerger)):
            m = cls(s.target_shape, 2, s.weight, device=device)
            for b in range(0, len(preds), 5):
                m.integrate_batch(preds[b : b + 5].cuda(), s.crops[b : b + 5])
            assert m.image.device.type == device and m.norm_mask.shape == (1,) + tuple(s.target_shape)
            assert_buffer_close(m.image.cpu().numpy(), g[f"{wname}_merger_image"])
            assert_buffer_close(m.norm_mask.cpu().numpy(), g[f"{wname}_merger_norm"])
            merged = m.merge()
            assert merged.device.type == device
            assert_buffer_close(merged.cpu().numpy(), g[f"{wname}_merged"])
        # in-place variant and accumulate_single
        m = CudaTileMerger(s.target_shape, 2, s.weight)
        for p, c in zip(preds, s.crops):
            m.accumulate_single(p.cuda(), c)
        out = m.merge_()
        assert out.data_ptr() == m.image.data_ptr()
        assert_buffer_close(out.cpu().numpy(), g[f"{wname}_merged"])
        # float64 NumPy-path twin
        tiles = [np.moveaxis(p.numpy(), 0, -1) for p in preds]
        m64 = s.merge(tiles, dtype=np.float32)
        np.testing.assert_allclose(m64, g[f"{wname}_merge_f64"], rtol=1e-6, atol=1e-6)


def test_merger_coordinate_forms_and_errors():
    s = ImageSlicer((200, 300, 3), (64, 32), (32, 16), weight="pyramid")
    m = CudaTileMerger(s.target_shape, 3, s.weight)
    batch = torch.rand(4, 3, 64, 32)
    with pytest.raises(ValueError, match="Number of images in batch"):
        m.integrate_batch(batch.cuda(), s.crops[:3])
    # ndarray rows, list of tuples, int64 tensor (DataLoader collate) all accepted
    m.integrate_batch(batch.cuda(), s.crops[:4])
    m.integrate_batch(batch.double(), [tuple(int(v) for v in c) for c in s.crops[4:8]])  # dtype + device cast
    m.integrate_batch(batch.half().cuda(), torch.from_numpy(s.crops[8:12]))
    ref = TO.TorchMerger(s.target_shape, 3, s.weight)
    ref.integrate_batch(batch, s.crops[:4])
    ref.integrate_batch(batch, s.crops[4:8])
    ref.integrate_batch(batch.half().float(), s.crops[8:12])
    assert_buffer_close(m.image.cpu().numpy(), ref.image.numpy())
    with pytest.raises(ValueError):
        m.integrate_batch(torch.rand(1, 3, 64, 32), [(10_000, 0, 32, 64)])
    with pytest.raises(TypeError):
        TileMerger(s.target_shape, 3, s.weight, device="cuda", dtype=torch.float16)
    # uncovered canvas -> NaN like the reference's 0/0
    m2 = CudaTileMerger((128, 128), 1, np.ones((32, 32), np.float32))
    m2.integrate_batch(torch.ones(1, 1, 32, 32), [(0, 0, 32, 32)])
    out = m2.merge()
    assert torch.isnan(out[0, 100, 100]) and out[0, 5, 5] == 1.0


def test_merger_many_tiles_device_coords_and_unaligned():
    """> 64 tiles per call (device coordinate path), odd tile size / odd positions (scalar reduction path)."""
    rng = np.random.default_rng(2)
    s = ImageSlicer((333, 417, 1), (51, 37), (26, 19), weight="pyramid")
    n = len(s.crops)
    assert n > 64
    preds = torch.from_numpy(rng.standard_normal((n, 2, 51, 37)).astype(np.float32)) * 3
    m = CudaTileMerger(s.target_shape, 2, s.weight)
    m.integrate_batch(preds.cuda(), s.crops)
    ref = TO.TorchMerger(s.target_shape, 2, s.weight)
    ref.integrate_batch(preds, s.crops)
    assert_buffer_close(m.image.cpu().numpy(), ref.image.numpy())
    assert_buffer_close(m.norm_mask.cpu().numpy(), ref.norm_mask.numpy())
    assert_buffer_close(m.merge().cpu().numpy(), ref.merge().numpy())


def test_reference_cuda_loop_max_channel():
    """reference tests/test_tiles.py:58-85 (README loop) on non-trivial data: DataLoader-collated int64 coords,
    merger constructed with the CPU default while predictions live on the GPU."""
    from torch.utils.data import DataLoader

    from pytorch_toolbelt_b200.utils.torch_utils import image_to_tensor, to_numpy

    image = np.random.default_rng(1).integers(0, 256, (700, 650, 3)).astype(np.uint8)
    tiler = ImageSlicer(image.shape, tile_size=(256, 256), tile_step=(128, 128), weight="pyramid")
    tiles = [image_to_tensor(tile) for tile in tiler.split(image)]
    merger = TileMerger(tiler.target_shape, 1, tiler.weight)
    for tiles_batch, coords_batch in DataLoader(list(zip(tiles, tiler.crops)), batch_size=8, pin_memory=True):
        pred_batch, _ = torch.max(tiles_batch.float().cuda(), dim=1, keepdim=True)
        merger.integrate_batch(pred_batch, coords_batch)
    merged = np.moveaxis(to_numpy(merger.merge()), 0, -1)
    merged = tiler.crop_to_orignal_size(merged)
    expected = image.max(axis=2, keepdims=True).astype(np.float32)
    np.testing.assert_allclose(merged, expected, rtol=1e-5, atol=1e-3)
    assert np.array_equal(np.rint(merged).astype(np.uint8), image.max(axis=2, keepdims=True))


def test_cpu_merger_downloads_the_canvas_once_per_state():
    """TileMerger(device="cpu") used to copy the whole canvas to the host on every attribute read."""
    from pytorch_toolbelt_b200.inference.tiles import ImageSlicer, TileMerger

    slicer = ImageSlicer((200, 180, 3), 64, 32, weight="pyramid")
    merger = TileMerger(slicer.target_shape, 2, slicer.weight, device="cpu")
    batch = torch.randn(4, 2, 64, 64)
    merger.integrate_batch(batch, slicer.crops[:4])
    a, b = merger.image, merger.image
    assert a.device.type == "cpu" and a.data_ptr() == b.data_ptr()            # cached: one download
    assert merger.norm_mask.data_ptr() == merger.norm_mask.data_ptr()
    merger.integrate_batch(batch, slicer.crops[4:8])
    c = merger.image
    assert c.data_ptr() != a.data_ptr() and not torch.equal(a, c)             # new state -> new snapshot
    merger.image = torch.zeros_like(c)
    assert float(merger.image.abs().max()) == 0.0
