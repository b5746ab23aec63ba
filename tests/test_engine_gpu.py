"""GPU: the fused pipeline (slice+augment kernel, output-stationary merge kernel) against the oracle's restatement
of the reference loop, plus size-independent properties at the BASELINE sizes."""
import numpy as np
import pytest
import torch

from oracle import pipeline as OP
from oracle import tiles_oracle as TO
from oracle import tta_oracle as AO
from pytorch_toolbelt_b200 import _native as N
from pytorch_toolbelt_b200.inference.engine import TiledInferenceEngine
from pytorch_toolbelt_b200.inference.tiles import CudaTileMerger, ImageSlicer

pytestmark = pytest.mark.gpu


def standin_cuda(x, Wm):
    B, C, T, T2 = x.shape
    y = torch.empty((B, Wm.shape[0], T, T2), dtype=torch.float32, device=x.device)
    import ctypes as C_

    arr = (C_.c_float * Wm.numel())(*[float(v) for v in Wm.flatten().tolist()])
    N.check(N.lib().tb_standin_model(x.data_ptr(), B, C, Wm.shape[0], T, T2, arr, y.data_ptr(), N.stream_ptr(x.device)),
            "standin")
    return y


def buffer_close(ours, ref, rel=1e-5):
    ours, ref = ours.double().cpu(), ref.double().cpu()
    scale = ref.abs().max().item()
    err = (ours - ref).abs().max().item()
    assert err <= rel * scale, (err, scale)


@pytest.mark.parametrize(
    "shape,tile,step,group,K,weight",
    [
        ((300, 260, 3), 64, 32, "d4", 4, "pyramid"),
        ((300, 260, 3), 64, 32, "none", 3, "pyramid"),
        ((257, 301, 1), (64, 32), (32, 32), "flips", 2, "mean"),
        ((200, 200, 3), 51, 26, "d2", 2, "pyramid"),       # odd tile size: scalar (unaligned) kernels
        ((200, 200, 3), 50, 20, "d4", 2, "gaussian"),      # stride not a multiple of 4, 3-fold overlap
        ((130, 120, 4), 128, 128, "fliplr", 1, "pyramid"),  # single tile bigger than the image
    ],
)
def test_pipeline_vs_oracle(shape, tile, step, group, K, weight):
    rng = np.random.default_rng(0)
    image = rng.integers(0, 256, shape).astype(np.uint8)
    Cin = shape[2]
    Wm = OP.standin_model_weights(K, Cin)
    # slice + augment: bit-exact against split -> HWC->CHW -> float -> *_image_augment
    eng = TiledInferenceEngine(image.shape, tile, step, channels=K, weight=weight, tta=None if group == "none" else group)
    dev_img = eng.to_device(image)
    aug = eng.slice(dev_img)
    t2 = (tile, tile) if np.isscalar(tile) else tile
    s2 = (step, step) if np.isscalar(step) else step
    grid = TO.tile_grid(shape[0], shape[1], t2, s2)
    tiles = TO.split(image, grid)
    batch = torch.stack([torch.from_numpy(np.ascontiguousarray(np.moveaxis(t.reshape(t2[0], t2[1], Cin), -1, 0))) for t in tiles]).float()
    ref_aug = AO.augment(batch, group) if group != "none" else batch
    assert torch.equal(aug.cpu(), ref_aug)
    # identical predictions on both sides (computed once on the CPU)
    pred = OP.standin_model(ref_aug, Wm)
    ours = eng.merge(pred.cuda())
    ref = OP.tiled_inference_cpu(image, t2, s2, K, weight=weight, group=group, batch_size=5,
                                 model=lambda x: OP.standin_model(x, Wm))
    assert ours.shape == ref.shape == (K, shape[0], shape[1])
    buffer_close(ours, ref, rel=2e-6)
    # batch-list form (pointer table) gives the same bits as the single-tensor form
    A = eng.views
    n = eng.num_tiles
    batches = []
    for s in range(0, n, 5):
        e = min(s + 5, n)
        batches.append(pred.view(A, n, K, t2[0], t2[1])[:, s:e].reshape(-1, K, t2[0], t2[1]).contiguous().cuda())
    assert torch.equal(eng.merge(batches), ours)
    # streaming TileMerger path with fused de-augment agrees too
    slicer = ImageSlicer(image.shape, tile, step, weight=weight)
    m = CudaTileMerger(slicer.target_shape, K, slicer.weight)
    for b, s in zip(batches, range(0, n, 5)):
        m.integrate_batch(b, slicer.crops[s : s + 5], tta=eng.tta)
    top, left = slicer.margin_top, slicer.margin_left
    streamed = m.merge()[:, top : top + shape[0], left : left + shape[1]]
    buffer_close(streamed, ref, rel=1e-5)


def test_predict_end_to_end_with_cuda_model():
    rng = np.random.default_rng(4)
    image = rng.integers(0, 256, (500, 420, 3)).astype(np.uint8)
    K = 4
    Wm = OP.standin_model_weights(K, 3)
    eng = TiledInferenceEngine(image.shape, 128, 64, channels=K, weight="pyramid", tta="d4")
    out = eng.predict(image, lambda x: standin_cuda(x, Wm), batch_size=7)
    ref = OP.tiled_inference_cpu(image, 128, 64, K, weight="pyramid", group="d4")
    buffer_close(out, ref, rel=1e-5)


def test_analytic_weights_match_tables():
    rng = np.random.default_rng(8)
    image = rng.integers(0, 256, (300, 300, 3)).astype(np.uint8)
    for weight in ("pyramid", "gaussian", "mean"):
        a = TiledInferenceEngine(image.shape, 64, 32, channels=2, weight=weight, tta="fliplr", analytic_weight=True)
        b = TiledInferenceEngine(image.shape, 64, 32, channels=2, weight=weight, tta="fliplr", analytic_weight=False)
        pred = torch.randn(2 * a.num_tiles, 2, 64, 64, generator=torch.Generator().manual_seed(0)).cuda() * 3
        buffer_close(a.merge(pred), b.merge(pred), rel=1e-5)


def test_slice_options_borders_affine_layout():
    rng = np.random.default_rng(9)
    image = rng.integers(0, 256, (150, 170, 3)).astype(np.uint8)
    grid = TO.tile_grid(150, 170, (64, 64), (48, 48))
    for mode in (0, 1, 2, 3, 4):
        eng = TiledInferenceEngine(image.shape, 64, 48, channels=1, tta=None, border_type=mode, border_value=(5, 6, 7))
        out = eng.slice(eng.to_device(image))
        ref = np.stack(TO.split(image, grid, mode, (5, 6, 7)))
        assert np.array_equal(out.cpu().numpy(), np.moveaxis(ref, -1, 1).astype(np.float32)), mode
    scale, bias = [1 / 58.0, 1 / 57.0, 1 / 57.5], [-2.1, -2.0, -1.8]
    eng = TiledInferenceEngine(image.shape, 64, 48, channels=1, tta="d2", scale=scale, bias=bias)
    out = eng.slice(eng.to_device(image))
    base = np.moveaxis(np.stack(TO.split(image, grid)), -1, 1).astype(np.float32)
    norm = base * np.array(scale, np.float32)[None, :, None, None] + np.array(bias, np.float32)[None, :, None, None]
    assert torch.equal(out.cpu(), AO.augment(torch.from_numpy(norm), "d2"))
    chw = torch.from_numpy(np.ascontiguousarray(np.moveaxis(image, -1, 0))).float().cuda()
    eng = TiledInferenceEngine(image.shape, 64, 48, channels=1, tta=None)
    assert np.array_equal(eng.slice(chw, layout="CHW").cpu().numpy(), base)
    assert np.array_equal(eng.slice(eng.to_device(image), 3, 7).cpu().numpy(), base[3:7])


def test_band_mode_reproduces_full_merge():
    """Strip partition used by the multi-GPU path: bands + halo seed == single full merge, bit for bit."""
    rng = np.random.default_rng(12)
    image_shape = (700, 500, 3)
    K = 3
    eng = TiledInferenceEngine(image_shape, 128, 64, channels=K, weight="pyramid", tta="flips")
    ny, nx = eng.slicer.grid_shape
    pred = torch.randn(eng.views * eng.num_tiles, K, 128, 128, generator=torch.Generator().manual_seed(1)).cuda()
    full = eng.merge(pred)
    hp, wp = eng.slicer.target_shape
    r_split = ny // 2
    y_split = r_split * 64
    # strip 0 owns tile rows [0,r_split) and canvas rows [0,y_split); its tiles spill into [y_split, y_split+64)
    top_part = eng.merge(pred, rows=(0, y_split), tile_rows=(0, r_split))
    halo = eng.merge(pred, rows=(y_split, y_split + 64), tile_rows=(0, r_split), raw=True)
    # strip 1 owns the rest; the first 64 rows start from the halo
    seeded = eng.merge(pred, rows=(y_split, y_split + 64), tile_rows=(r_split, ny), init=halo)
    rest = eng.merge(pred, rows=(y_split + 64, hp), tile_rows=(r_split, ny))
    top = eng.slicer.margin_top
    H = image_shape[0]

    def rows_of(t, y0, y1):
        a, b = max(y0 - top, 0), min(y1 - top, H)
        return t[:, a:b]

    assert torch.equal(rows_of(top_part, 0, y_split), rows_of(full, 0, y_split))
    assert torch.equal(rows_of(seeded, y_split, y_split + 64), rows_of(full, y_split, y_split + 64))
    assert torch.equal(rows_of(rest, y_split + 64, hp), rows_of(full, y_split + 64, hp))


@pytest.mark.gpu
def test_prepared_plans_rerun_on_refilled_buffers():
    """slice_plan / merge_plan (the steady-state form used by the multi-GPU loop): marshalled once, run() re-launches on
    whatever the bound buffers hold now. Three strips are emulated on one GPU; each strip's predictions are its own
    tensor (only its tiles resident), the halo buffers are refilled every step."""
    from pytorch_toolbelt_b200.distributed.strips import StripPlan

    rng = np.random.default_rng(5)
    image_shape = (900, 500, 3)
    K, T, S = 2, 128, 64
    eng = TiledInferenceEngine(image_shape, T, S, channels=K, weight="pyramid", tta="fliplr")
    ny, nx = eng.slicer.grid_shape
    hp, wp = eng.slicer.target_shape
    world = 3
    plans = [StripPlan.build(ny, T, S, hp, g, world) for g in range(world)]
    image = eng.to_device(rng.integers(0, 256, image_shape).astype(np.uint8))
    A = eng.views
    strips = []
    out = torch.zeros((K, image_shape[0], image_shape[1]), device="cuda")
    for p in plans:
        n_own = (p.r1 - p.r0) * nx
        pred = torch.empty((A * n_own, K, T, T), device="cuda")
        aug = torch.empty((A * n_own, 3, T, T), device="cuda")
        band = lambda rows, raw=False, init=None, pred=pred, p=p: eng.merge_plan(  # noqa: E731
            pred, out=None if raw else out, rows=rows, tile_rows=(p.r0, p.r1), raw=raw, init=init)
        top_rows = p.halo_in
        recv = (torch.empty((K, top_rows, wp), device="cuda"), torch.empty((top_rows, wp), device="cuda")) if top_rows else None
        strips.append(dict(
            plan=p, pred=pred, aug=aug, recv=recv,
            slicer=eng.slice_plan(image, p.r0 * nx, p.r1 * nx, out=aug),
            spill=band((p.y1, p.y1 + p.halo), raw=True) if p.halo else None,
            top=band((p.y0, p.y0 + top_rows), init=recv) if top_rows else None,
            interior=band((p.y0 + top_rows, p.y1)),
        ))
    for step in range(2):
        full_pred = torch.randn(A * eng.num_tiles, K, T, T, generator=torch.Generator().manual_seed(step)).cuda()
        want = eng.merge(full_pred)
        want_aug = eng.slice(image)
        out.fill_(float("nan"))
        by_view = full_pred.view(A, eng.num_tiles, K, T, T)
        for st in strips:
            p = st["plan"]
            st["pred"].view(A, -1, K, T, T).copy_(by_view[:, p.r0 * nx : p.r1 * nx])
            got_aug = st["slicer"].run()
            assert got_aug is st["aug"]
            assert torch.equal(got_aug.view(A, -1, 3, T, T), want_aug.view(A, eng.num_tiles, 3, T, T)[:, p.r0 * nx : p.r1 * nx])
        for g, st in enumerate(strips):
            if st["spill"] is not None:
                num, den = st["spill"].run()
                strips[g + 1]["recv"][0].copy_(num)
                strips[g + 1]["recv"][1].copy_(den)
            if st["top"] is not None:
                st["top"].run()
            st["interior"].run()
        assert torch.equal(out, want), f"step {step}"


def test_full_size_properties_cfg2():
    """BASELINE cfg 2 geometry (5000x5000x3, 512/256, d4, K=4) through size-independent properties:
    constant predictions merge to the constant; D4 de-augment of an augmented constant-per-view field is exact;
    slice->identity-model->merge reproduces the image (weights cancel)."""
    H = W = 5000
    rng = np.random.default_rng(0)
    image = rng.integers(0, 256, (H, W, 3)).astype(np.uint8)
    eng = TiledInferenceEngine(image.shape, 512, 256, channels=3, weight="pyramid", tta="d4")
    assert eng.num_tiles == 361 and eng.slicer.target_shape == (5120, 5120)
    dev = eng.to_device(image)
    aug = eng.slice(dev)  # [8*361,3,512,512] = 9.1 GB
    assert aug.shape == (8 * 361, 3, 512, 512)
    # identity model: de-augmented views all equal the tile, the weighted mean of equal values is the value
    out = eng.merge(aug)
    ref = torch.from_numpy(np.ascontiguousarray(np.moveaxis(image, -1, 0))).float()
    torch.testing.assert_close(out.cpu(), ref, rtol=1e-5, atol=1e-3)
    del aug, out
    torch.cuda.empty_cache()
    # checksum of the tile stream: every view holds a permutation of the same pixels
    eng1 = TiledInferenceEngine(image.shape, 512, 256, channels=1, weight="mean", tta="d4")
    a = eng1.slice(dev, 100, 104)
    sums = a.view(8, 4, -1).double().sum(-1)
    assert torch.equal(sums, sums[0:1].expand(8, 4))


def test_tma_and_generic_merge_kernels_agree_bitwise():
    """tb_merge_all has two implementations (TMA-pipelined and plain-load); same arithmetic, same bits."""
    import os

    for tta_name, K in (("d4", 4), ("flips", 2), (None, 3), ("d2", 1)):
        eng = TiledInferenceEngine((700, 900, 3), 128, 64, channels=K, weight="pyramid", tta=tta_name)
        pred = torch.randn(eng.views * eng.num_tiles, K, 128, 128, generator=torch.Generator().manual_seed(5)).cuda()
        fast = eng.merge(pred)
        os.environ["TB200_FORCE_GENERIC_MERGE"] = "1"
        try:
            slow = eng.merge(pred)
        finally:
            del os.environ["TB200_FORCE_GENERIC_MERGE"]
        assert torch.equal(fast, slow), tta_name


def test_batched_prediction_lists_take_the_tma_path_and_agree():
    """predict(batch_size=b): per-batch model outputs (separate allocations) merged through per-batch tensor maps."""
    K = 3
    eng = TiledInferenceEngine((900, 700, 3), 128, 64, channels=K, weight="pyramid", tta="d4")
    A, n = eng.views, eng.num_tiles
    pred = torch.randn(A, n, K, 128, 128, generator=torch.Generator().manual_seed(9)).cuda()
    whole = eng.merge(pred.reshape(-1, K, 128, 128))
    for bsz in (1, 7, 8, n):
        batches = [pred[:, s : s + bsz].reshape(-1, K, 128, 128).contiguous() for s in range(0, n, bsz)]
        assert torch.equal(eng.merge(batches), whole), bsz
    # non-uniform batch sizes -> pointer-table kernel, same bits
    cuts = [0, 5, 6, 20, n]
    batches = [pred[:, a:b].reshape(-1, K, 128, 128).contiguous() for a, b in zip(cuts, cuts[1:])]
    assert torch.equal(eng.merge(batches), whole)
    with pytest.raises(ValueError):
        eng.merge(batches[:-1])


def test_predict_to_host_pipelined_is_bit_identical_to_predict():
    rng = np.random.default_rng(6)
    image = rng.integers(0, 256, (1000, 840, 3)).astype(np.uint8)
    K = 4
    Wm = OP.standin_model_weights(K, 3)
    for border in (0, 4):
        eng = TiledInferenceEngine(image.shape, 128, 64, channels=K, weight="pyramid", tta="d4", border_type=border)
        ref = eng.predict(image, lambda x: standin_cuda(x, Wm), batch_size=None)
        host_img = torch.from_numpy(image).pin_memory()
        for rows in (1, 2, 5, 100):
            out = eng.predict_to_host(host_img, lambda x: standin_cuda(x, Wm), rows_per_chunk=rows)
            torch.cuda.synchronize()
            assert out.is_pinned() and torch.equal(out, ref.cpu()), (border, rows)
