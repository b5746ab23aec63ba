"""GPU: round-2 additions to the merge path.
* single-view kernels (csrc/single_view.cu) against the TMA / generic kernels, bit for bit, and against the oracle;
* label / mask output fused into the merge epilogue against argmax / sigmoid>thr of the float result;
* sigmoid pre-activation before the TTA mean against the reference's ApplySigmoidTo semantics."""
import os

import numpy as np
import pytest
import torch

from oracle import pipeline as OP
from oracle import tta_oracle as AO
from pytorch_toolbelt_b200.inference.engine import TiledInferenceEngine

pytestmark = pytest.mark.gpu


class env:
    def __init__(self, **kv):
        self.kv = kv

    def __enter__(self):
        self.old = {k: os.environ.get(k) for k in self.kv}
        os.environ.update(self.kv)

    def __exit__(self, *exc):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


@pytest.mark.parametrize("shape,T,S,K,weight,analytic", [
    ((700, 900, 3), 128, 64, 3, "pyramid", False),     # 32-wide units (stride 64)
    ((1500, 1300, 1), 256, 128, 1, "gaussian", False),  # 128-wide units, table weight
    ((1500, 1300, 1), 256, 128, 2, "gaussian", True),   # separable in-register gaussian
    ((600, 520, 2), 256, 256, 4, "mean", False),        # no overlap
    ((900, 800, 3), 512, 128, 2, "pyramid", False),     # 4x4 covering tiles per pixel (several load chunks)
])
def test_single_view_kernels_match_the_general_kernels(shape, T, S, K, weight, analytic):
    rng = np.random.default_rng(3)
    image = rng.integers(0, 256, shape).astype(np.uint8)
    eng = TiledInferenceEngine(shape, T, S, channels=K, weight=weight, tta=None, analytic_weight=analytic)
    dev = eng.to_device(image)
    aug = eng.slice(dev)
    pred = torch.randn(eng.num_tiles, K, T, T, generator=torch.Generator().manual_seed(5)).cuda() * 3
    fast = eng.merge(pred)
    lab = eng.merge(pred, output="labels")
    msk = eng.merge(pred, output="mask", threshold=0.4)
    with env(TB200_SINGLE_VIEW="0"):
        aug_general = eng.slice(dev)
        general = eng.merge(pred)                       # TMA ring (lattice geometry) or generic
        with env(TB200_FORCE_GENERIC_MERGE="1"):
            generic = eng.merge(pred)
    assert torch.equal(aug, aug_general)
    if analytic:  # separable exp(a)*exp(b) vs exp(a+b): a few ulp
        torch.testing.assert_close(fast, general, rtol=2e-6, atol=2e-6)
    else:
        assert torch.equal(fast, general) and torch.equal(fast, generic)
        assert torch.equal(lab, fast.argmax(0).to(torch.uint8))
        assert torch.equal(msk, (torch.sigmoid(fast) > 0.4).to(torch.uint8))
    # list-of-batches form (per-batch pointers by value in the kernel parameters) and band mode
    batches = [pred[s : s + 5].contiguous() for s in range(0, eng.num_tiles, 5)]
    assert torch.equal(eng.merge(batches), fast)
    hp, wp = eng.slicer.target_shape
    ny, nx = eng.slicer.grid_shape
    if T > S and ny >= 2:  # band mode (multi-GPU strips): halo seed + bands == one merge, bit for bit
        r = ny // 2
        num, den = eng.merge(pred, rows=(r * S, r * S + T - S), tile_rows=(0, r), raw=True)
        out = torch.full_like(fast, float("nan"))
        eng.merge(pred, out=out, rows=(0, r * S), tile_rows=(0, r))
        eng.merge(pred, out=out, rows=(r * S, r * S + T - S), tile_rows=(r, ny), init=(num, den))
        eng.merge(pred, out=out, rows=(r * S + T - S, hp), tile_rows=(r, ny))
        assert torch.equal(out, fast)


def test_single_view_slice_all_dtypes_and_borders():
    from oracle import tiles_oracle as TO

    rng = np.random.default_rng(4)
    for dtype, C in ((np.uint8, 1), (np.uint8, 3), (np.float32, 2), (np.int16, 1)):
        image = rng.integers(0, 200, (300, 280, C)).astype(dtype)
        grid = TO.tile_grid(300, 280, (128, 128), (96, 96))
        for mode in (0, 1, 2, 4):
            eng = TiledInferenceEngine(image.shape, 128, 96, channels=1, tta=None, border_type=mode, border_value=(3, 4, 5))
            out = eng.slice(eng.to_device(image))
            ref = np.stack(TO.split(image, grid, mode, (3, 4, 5)))
            assert np.array_equal(out.cpu().numpy(), np.moveaxis(ref.reshape(-1, 128, 128, C), -1, 1).astype(np.float32)), (dtype, C, mode)


@pytest.mark.parametrize("group,K", [("d4", 4), ("flips", 8), ("d2", 2), ("fliplr", 1)])
def test_fused_label_and_mask_outputs_multi_view(group, K):
    eng = TiledInferenceEngine((700, 820, 3), 128, 64, channels=K, weight="pyramid", tta=group)
    pred = torch.randn(eng.views * eng.num_tiles, K, 128, 128, generator=torch.Generator().manual_seed(2)).cuda() * 3
    logits = eng.merge(pred)
    if K > 1:
        assert torch.equal(eng.merge(pred, output="labels"), logits.argmax(0).to(torch.uint8))
    assert torch.equal(eng.merge(pred, output="mask", threshold=0.5), (torch.sigmoid(logits) > 0.5).to(torch.uint8))
    with env(TB200_FORCE_GENERIC_MERGE="1"):
        if K > 1:
            assert torch.equal(eng.merge(pred, output="labels"), logits.argmax(0).to(torch.uint8))
        assert torch.equal(eng.merge(pred, output="mask", threshold=0.5), (torch.sigmoid(logits) > 0.5).to(torch.uint8))
    with pytest.raises(ValueError):
        eng.merge(pred, output="labels", raw=True, rows=(0, 64))
    # NaN handling follows torch.argmax (NaN is maximal, first one wins)
    if K > 2:
        bad = pred.clone()
        bad.view(eng.views, eng.num_tiles, K, 128, 128)[:, 0, 1] = float("nan")
        want = eng.merge(bad).argmax(0).to(torch.uint8)
        assert torch.equal(eng.merge(bad, output="labels"), want)


@pytest.mark.parametrize("group", ["d4", "flips", None])
def test_sigmoid_pre_activation_before_the_view_mean(group):
    """ApplySigmoidTo(model, temperature) + TTA mean + TileMerger of the reference (ensembling.py:46-68, tta.py:63-96)
    == engine(activation="sigmoid", temperature=...) on the raw logits."""
    K, T, S = 3, 128, 64
    shape = (500, 420, 3)
    temperature = 1.7
    rng = np.random.default_rng(6)
    image = rng.integers(0, 256, shape).astype(np.uint8)
    Wm = OP.standin_model_weights(K, 3)
    ref = OP.tiled_inference_cpu(image, T, S, K, weight="pyramid", group=group or "none",
                                 model=lambda x: OP.standin_model(x, Wm).mul(temperature).sigmoid_())
    eng = TiledInferenceEngine(shape, T, S, channels=K, weight="pyramid", tta=group, activation="sigmoid", temperature=temperature)
    plain = TiledInferenceEngine(shape, T, S, channels=K, weight="pyramid", tta=group)
    aug = plain.slice(plain.to_device(image))
    logits = OP.standin_model(aug.cpu(), Wm).cuda()
    out = eng.merge(logits)
    scale = ref.abs().max().item()
    assert (out.cpu() - ref).abs().max().item() <= 1e-5 * scale
    # same thing with the activation applied by the caller and a plain engine
    want = plain.merge(torch.sigmoid(logits * temperature))
    torch.testing.assert_close(out, want, rtol=2e-6, atol=2e-7)
    with env(TB200_FORCE_GENERIC_MERGE="1"):
        torch.testing.assert_close(eng.merge(logits), want, rtol=2e-6, atol=2e-7)
    with pytest.raises(ValueError):
        TiledInferenceEngine(shape, T, S, channels=K, activation="softmax")
