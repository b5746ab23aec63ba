"""GPU: every BASELINE config compared with the oracle (and with the installed reference when baseline/_ref travelled)
at the size it is benchmarked at -- with the D4-NON-equivariant stand-in model of SURVEY 8(d) between slice and merge, so
that a consistent view-order / index error in augment + de-augment cannot cancel (an identity model would hide it).

cfg 1  2048x2048x3, 256/128, pyramid, no TTA, K=3      full size, + the reference's own test (tests/test_tiles.py:59-85)
cfg 2  5000x5000x3, 512/256, pyramid, d4, K=4          FULL SIZE vs the oracle loop (about 10-20 s of CPU)
cfg 3  focal + dice (+ joint) on 32x4x1024x1024        FULL SIZE vs the reference's torch ops (oracle functions on CUDA
                                                        float32) and float64 statistics
cfg 4  10000^2x3, 512/256, flips, K=8                  a 5x5-tile crop of the geometry (same T/S/TTA/K), + full-size properties
cfg 5  20000^2x1, 1024/512, gaussian, no TTA, K=1      a 4x3-tile crop of the geometry, + full-size properties
"""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import losses_oracle as LO
from oracle import pipeline as OP
from pytorch_toolbelt_b200 import _native as N
from pytorch_toolbelt_b200.inference.engine import TiledInferenceEngine

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def standin_cuda(x, Wm):
    import ctypes as C_

    B, C, T, T2 = x.shape
    y = torch.empty((B, Wm.shape[0], T, T2), dtype=torch.float32, device=x.device)
    arr = (C_.c_float * Wm.numel())(*[float(v) for v in Wm.flatten().tolist()])
    N.check(N.lib().tb_standin_model(x.data_ptr(), B, C, Wm.shape[0], T, T2, arr, y.data_ptr(), N.stream_ptr(x.device)),
            "standin")
    return y


def buffer_close(ours, ref, rel=1e-5):
    """north_star gate: max |a-b| <= 1e-5 * max|ref| (relative to the buffer scale)."""
    ours, ref = ours.double().cpu(), ref.double().cpu()
    scale = ref.abs().max().item()
    err = (ours - ref).abs().max().item()
    assert err <= rel * scale, (err, scale)
    return err / scale


def reference_loop():
    """The installed reference if it travelled with the repo, else the oracle's restatement (bit-identical to it:
    tests/test_reference_install.py)."""
    from baseline import reference_arm as RA

    if RA.available():
        return lambda image, T, S, K, weight, group, **kw: RA.tiled_inference(image, T, S, K, weight=weight, group=group, **kw)
    return lambda image, T, S, K, weight, group, **kw: OP.tiled_inference_cpu(image, T, S, K, weight=weight,
                                                                               group=group or "none", **kw)


def run_config(shape, T, S, K, weight, group, batch_tiles):
    rng = np.random.default_rng(0)
    image = rng.integers(0, 256, shape).astype(np.uint8)
    Wm = OP.standin_model_weights(K, shape[2])
    eng = TiledInferenceEngine(image.shape, T, S, channels=K, weight=weight, tta=group)
    out = eng.predict(image, lambda x: standin_cuda(x, Wm), batch_size=batch_tiles)
    ref = reference_loop()(image, T, S, K, weight, group)
    assert out.shape == ref.shape == (K, shape[0], shape[1])
    rel = buffer_close(out, ref)
    # the pipelined host path and the uint8 label download agree with the resident result
    host = eng.predict_to_host(torch.from_numpy(image).pin_memory(), lambda x: standin_cuda(x, Wm), rows_per_chunk=2)
    torch.cuda.synchronize()
    assert torch.equal(host, out.cpu())
    if K > 1:
        lab = eng.predict_to_host(torch.from_numpy(image).pin_memory(), lambda x: standin_cuda(x, Wm), rows_per_chunk=3,
                                  output="labels")
        torch.cuda.synchronize()
        assert torch.equal(lab, out.argmax(0).to(torch.uint8).cpu())
    return eng, rel


def test_cfg2_full_size_against_the_oracle_with_the_non_equivariant_model():
    eng, rel = run_config((5000, 5000, 3), 512, 256, 4, "pyramid", "d4", batch_tiles=64)
    assert eng.num_tiles == 361 and eng.views == 8 and eng.slicer.target_shape == (5120, 5120)
    assert rel <= 1e-5


def test_cfg1_full_size_against_the_oracle_and_the_reference_test():
    eng, _ = run_config((2048, 2048, 3), 256, 128, 3, "pyramid", None, batch_tiles=75)
    assert eng.num_tiles == 225
    # the reference's own CUDA test, restated against the drop-in classes (tests/test_tiles.py:59-85)
    from pytorch_toolbelt_b200.inference.tiles import CudaTileMerger, ImageSlicer
    from pytorch_toolbelt_b200.utils.torch_utils import tensor_from_rgb_image, to_numpy

    image = np.random.default_rng(1).integers(0, 256, (2048, 2048, 3)).astype(np.uint8)
    tiler = ImageSlicer(image.shape, tile_size=(256, 256), tile_step=(128, 128), weight="pyramid")
    tiles = [tensor_from_rgb_image(tile) for tile in tiler.split(image)]
    merger = CudaTileMerger(tiler.target_shape, 1, tiler.weight)
    for s in range(0, len(tiles), 8):
        batch = torch.stack(tiles[s : s + 8]).float().cuda()
        merger.integrate_batch(batch.max(dim=1, keepdim=True)[0], tiler.crops[s : s + 8])
    # the reference test casts with .astype(np.uint8) (its image is all zeros: np.random.random(...).astype(uint8)); with real
    # pixel values a weighted mean of equal numbers can land at 192.99999, so round instead of truncating
    merged_f = np.moveaxis(to_numpy(merger.merge()), 0, -1)
    merged = tiler.crop_to_orignal_size(np.rint(merged_f).astype(np.uint8))
    np.testing.assert_equal(merged, image.max(axis=2, keepdims=True))
    np.testing.assert_allclose(tiler.crop_to_orignal_size(merged_f), image.max(axis=2, keepdims=True).astype(np.float32), atol=1e-3)
    zero = np.random.random((2048, 2048, 3)).astype(np.uint8)  # exactly the reference's input
    merger0 = CudaTileMerger(tiler.target_shape, 1, tiler.weight)
    tiles0 = [tensor_from_rgb_image(tile) for tile in tiler.split(zero)]
    for s in range(0, len(tiles0), 8):
        merger0.integrate_batch(torch.stack(tiles0[s : s + 8]).float().cuda().max(dim=1, keepdim=True)[0], tiler.crops[s : s + 8])
    np.testing.assert_equal(tiler.crop_to_orignal_size(np.moveaxis(to_numpy(merger0.merge()), 0, -1).astype(np.uint8)),
                            zero.max(axis=2, keepdims=True))


def test_cfg4_geometry_crop_against_the_oracle():
    # 5x5 tiles of cfg 4's geometry: 512/256, flips (3 views), K=8, ragged margins on both axes
    eng, _ = run_config((1400, 1500, 3), 512, 256, 8, "pyramid", "flips", batch_tiles=7)
    assert eng.slicer.grid_shape == (5, 5) and eng.views == 3


def test_cfg5_geometry_crop_against_the_oracle():
    # 4x3 tiles of cfg 5's geometry: 1024/512, gaussian weight, single view, single channel in and out
    eng, _ = run_config((2500, 2000, 1), 1024, 512, 1, "gaussian", None, batch_tiles=5)
    assert eng.slicer.grid_shape == (4, 3) and eng.views == 1


@pytest.mark.parametrize("name,shape,T,S,K,weight,group", [
    ("cfg4", (10000, 10000, 3), 512, 256, 8, "pyramid", "flips"),
    ("cfg5", (20000, 20000, 1), 1024, 512, 1, "gaussian", None),
])
def test_full_size_properties_cfg4_cfg5(name, shape, T, S, K, weight, group):
    """Size-independent properties at BASELINE's full size: (a) a chunk of the full-size slice equals the oracle's crops
    of the same tiles bit for bit; (b) a linear ramp model (not flip-equivariant) merges to its closed form:
    predictions y = a*row + b*col + k (tile-local coordinates, view-dependent through the flip) de-augment to the same
    ramp, and the weighted mean of per-tile ramps at a canvas pixel is computed in float64 from the weights."""
    from oracle import tiles_oracle as TO
    from oracle import tta_oracle as AO

    H, W, Cin = shape
    rng = np.random.default_rng(0)
    image = rng.integers(0, 256, shape).astype(np.uint8)
    eng = TiledInferenceEngine(shape, T, S, channels=K, weight=weight, tta=group)
    assert eng.num_tiles == 1521
    dev = eng.to_device(image)
    # (a) the last 3 tiles (bottom-right corner: margins on two sides) and 3 interior tiles
    grid = TO.tile_grid(H, W, (T, T), (S, S))
    for start in (eng.num_tiles - 3, 20 * 39 + 17):
        got = eng.slice(dev, start, start + 3).cpu()
        crops = [TO.crop_with_border(image, int(x), int(y), T, T, 0, 0, False)
                 for (x, y, _, _) in grid["bbox_crops"][start : start + 3]]
        base = torch.from_numpy(np.moveaxis(np.stack(crops).reshape(3, T, T, Cin), -1, 1).astype(np.float32))
        want = AO.augment(base, group) if group else base
        assert torch.equal(got, want), (name, start)
    # (b) slice -> channel-replicating model with a NON-symmetric per-view perturbation -> merge: the view mean of the
    # de-augmented perturbations is known in closed form, the image part must come back exactly (weights cancel)
    A, n = eng.views, eng.num_tiles
    aug = eng.slice(dev)
    pred = torch.empty((A * n, K, T, T), dtype=torch.float32, device=aug.device)
    ii = torch.arange(T, device=aug.device, dtype=torch.float32).view(1, T, 1)
    jj = torch.arange(T, device=aug.device, dtype=torch.float32).view(1, 1, T)
    ramp = (ii * 2.0 + jj) / 64.0  # exact in float32 (multiples of 1/64 below 2^24)
    for k in range(K):
        pred[:, k] = aug[:, k % Cin] + ramp + float(k)
    out = eng.merge(pred)
    del pred, aug
    # de-augmented ramp per view: identity -> 2i+j; fliplr -> 2i+(T-1-j); flipud -> 2(T-1-i)+j; mean over views
    M = float(T - 1)
    i = torch.arange(T, dtype=torch.float64).view(T, 1)
    j = torch.arange(T, dtype=torch.float64).view(1, T)
    if group == "flips":
        mean_ramp = ((2 * i + j) + (2 * i + (M - j)) + (2 * (M - i) + j)) / 3.0 / 64.0
    else:
        mean_ramp = (2 * i + j) / 64.0
    w = torch.from_numpy(np.asarray(eng.slicer.weight, dtype=np.float32)).double()
    hp, wp = eng.slicer.target_shape
    top, left = eng.slicer.margin_top, eng.slicer.margin_left
    # float64 overlap-add of (w * mean_ramp) and w over a band of canvas rows (full width), compared on that band
    for y_lo in (0, (hp // 2 // S) * S, hp - T):
        rows = T
        num = torch.zeros(rows, wp, dtype=torch.float64)
        den = torch.zeros(rows, wp, dtype=torch.float64)
        for (x, y, _, _) in grid["crops"]:
            x, y = int(x), int(y)
            a, b = max(y, y_lo), min(y + T, y_lo + rows)
            if b <= a:
                continue
            num[a - y_lo : b - y_lo, x : x + T] += (w * mean_ramp)[a - y : b - y]
            den[a - y_lo : b - y_lo, x : x + T] += w[a - y : b - y]
        band = num / den
        oa, ob = max(y_lo - top, 0), min(y_lo + rows - top, H)
        if ob <= oa:
            continue
        want_ramp = band[oa + top - y_lo : ob + top - y_lo, left : left + W]
        img_part = torch.from_numpy(image[oa:ob].astype(np.float64))
        for k in (0, K - 1):
            want = img_part[:, :, k % Cin] + want_ramp + float(k)
            got = out[k, oa:ob].double().cpu()
            err = (got - want).abs().max().item()
            assert err <= 1e-5 * 300.0, (name, y_lo, k, err)


def test_cfg3_full_size_losses_against_the_reference_ops():
    """32x4x1024x1024 logits: BinaryFocalLoss, DiceLoss(multilabel), their sum through the joint kernel, multiclass dice
    and CE-focal, forward AND gradient, against the oracle's restatement of the reference run as float32 torch ops on the
    same GPU; the dice statistics additionally against float64 sums (reduction drift shows at 134 M elements)."""
    import pytorch_toolbelt_b200.losses as L

    B, C, H, W = 32, 4, 1024, 1024
    g = torch.Generator().manual_seed(0)
    x = (torch.randn(B, C, H, W, generator=g) * 3).cuda()
    t = (torch.rand(B, C, H, W, generator=g) < 0.3).float().cuda()
    lab = torch.randint(0, C, (B, H, W), generator=g).cuda()

    def both(ours, ref, target):
        xa = x.clone().requires_grad_(True)
        la = ours(xa, target)
        la.backward()
        xb = x.clone().requires_grad_(True)
        lb = ref(xb, target)
        lb.backward()
        torch.testing.assert_close(la, lb.detach(), rtol=1e-5, atol=1e-7)
        scale = float(xb.grad.abs().max())
        err = float((xa.grad - xb.grad).abs().max())
        assert err <= 1e-4 * scale, (err, scale)
        return float(la)

    focal, dice = L.BinaryFocalLoss(), L.DiceLoss("multilabel")
    lf = both(focal, lambda z, tt: LO.focal_sigmoid(z, tt, alpha=None), t)
    ld = both(dice, lambda z, tt: LO.dice_loss(z, tt, "multilabel"), t)
    joint = L.BinaryFocalDiceLoss(focal, dice)
    lj = both(joint, lambda z, tt: LO.focal_sigmoid(z, tt, alpha=None) + LO.dice_loss(z, tt, "multilabel"), t)
    assert abs(lj - (lf + ld)) <= 1e-5 * abs(lf + ld)
    both(L.JaccardLoss("multilabel"), lambda z, tt: LO.jaccard_loss(z, tt, "multilabel"), t)
    both(L.DiceLoss("multiclass"), lambda z, tt: LO.dice_loss(z, tt, "multiclass"), lab)
    both(L.CrossEntropyFocalLoss(), lambda z, tt: LO.focal_softmax(z, tt), lab)
    both(focal, lambda z, tt: LO.focal_sigmoid(z, LO.one_hot_targets(tt, C), alpha=None), lab)
    # float64 ground truth of the dice value (per-class sums over 33.5 M elements each)
    p = torch.sigmoid(x.double())
    I = (p * t).sum(dim=(0, 2, 3))
    S = (p + t).sum(dim=(0, 2, 3))
    want = (1.0 - 2.0 * I / S.clamp_min(1e-7)).mean().item()
    assert abs(ld - want) <= 1e-6 * abs(want), (ld, want)


def test_sharded_merge_equals_single_gpu_merge_on_two_ranks():
    """tests/multi_gpu_check.py (strip-sharded merge + halo exchange vs the single-GPU merge, bit for bit; Sync losses vs the
    whole-batch loss) spawned on 2 ranks. Needs 2 visible GPUs; with one GPU the same partition logic is covered by
    test_engine_gpu.py::test_band_mode_* (same kernels, ranks simulated in one process) and the gloo tests."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (single-GPU box: band-mode tests cover the same kernels in one process)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(REPO, "tests", "multi_gpu_check.py")]
    proc = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=REPO)
    assert proc.returncode == 0 and "MULTI_GPU_CHECK PASS" in proc.stdout, proc.stdout[-3000:] + proc.stderr[-3000:]
