"""CPU: host-side logic of the drop-in (bit-exact grid arithmetic, argument validation, error types) and the
C-ABI library: loads and exports every symbol include/tb200.h declares. No compute calls (no GPU here)."""
import os
import re

import numpy as np
import pytest
import torch

from pytorch_toolbelt_b200 import _native as N
from pytorch_toolbelt_b200.inference import functional as F
from pytorch_toolbelt_b200.inference import tta
from pytorch_toolbelt_b200.inference.tiles import (
    ImageSlicer,
    compute_pyramid_patch_weight_loss,
    gaussian_patch_weight,
)

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_are_exported():
    header = open(os.path.join(REPO, "include", "tb200.h")).read()
    declared = re.findall(r"TB_API\s+[\w\s\*]+?\b(tb_[a-z0-9_]+)\s*\(", header)
    assert len(declared) >= 20
    handle = N.load_library()
    missing = [name for name in declared if not hasattr(handle, name)]
    assert not missing, f"declared in include/tb200.h but not exported by libtb200.so: {missing}"
    unbound = [name for name in declared if name not in N.SIGNATURES]
    assert not unbound, f"no ctypes prototype for: {unbound}"
    assert handle.tb_version() == 100
    assert handle.tb_tta_views(N.TB.TTA_D4) == 8 and handle.tb_tta_views(N.TB.TTA_FLIPS) == 3
    assert handle.tb_tta_views(99) < 0
    assert b"unknown tta" in handle.tb_last_error()


def test_grid_is_bit_exact_with_reference(golden):
    g = golden("tiles_geometry")
    for idx in range(int(g["geom_count"])):
        H, W, th, tw, sh, sw = (int(v) for v in g[f"geom{idx}_args"])
        m = g[f"geom{idx}_margin_arg"]
        margin = int(m[0]) if len(m) == 1 else tuple(int(v) for v in m)
        s = ImageSlicer((H, W, 3), (th, tw), (sh, sw), image_margin=margin, weight="mean")
        assert s.crops.dtype == np.int64 and np.array_equal(s.crops, g[f"geom{idx}_crops"])
        assert np.array_equal(s.bbox_crops, g[f"geom{idx}_bbox"])
        assert (s.margin_left, s.margin_right, s.margin_top, s.margin_bottom) == tuple(g[f"geom{idx}_margins"])
        assert tuple(s.target_shape) == tuple(g[f"geom{idx}_target"])
        assert s.grid_shape[0] * s.grid_shape[1] == len(s.crops)


def test_scalar_and_pair_arguments_agree():
    a = ImageSlicer((500, 400), 51, 26)
    b = ImageSlicer((500, 400, 3), (51, 51), np.array([26, 26]))
    assert np.array_equal(a.crops, b.crops) and a.tile_size == (51, 51) and b.tile_step == (26, 26)


def test_pyramid_weight_bit_exact(golden):
    g = golden("tiles_geometry")
    for th, tw in [(32, 24), (24, 16), (16, 16), (51, 51)]:
        w, dc, de = compute_pyramid_patch_weight_loss(th, tw)
        assert w.dtype == np.float64 and w.shape == (th, tw)
        assert np.array_equal(w, g[f"pyramid_{th}x{tw}"])
    for t in (256, 512, 1024):
        s = ImageSlicer((t, t), t, t, weight="pyramid")
        assert s.weight_param == float(g[f"pyramid_alpha_{t}"])
        assert s.weight_kind == N.TB.W_PYRAMID
    s = ImageSlicer((5000, 5000, 3), (512, 512), (256, 256), weight="pyramid")
    np.testing.assert_allclose(s.weight, s.weight.T)  # reference tests/test_tiles.py:51
    np.testing.assert_allclose(
        [s.weight.min(), s.weight.max(), s.weight.mean(), s.weight[3, 7], s.weight[256, 170]], g["pyramid_stats_512"],
        rtol=0, atol=0)


def test_weight_options():
    s = ImageSlicer((100, 100), 32, 16, weight="mean")
    assert s.weight.dtype == np.float32 and np.all(s.weight == 1) and s.weight_kind == N.TB.W_ONES
    user = np.random.default_rng(0).random((32, 32))
    s = ImageSlicer((100, 100), 32, 16, weight=user)
    assert s.weight is user and s.weight_kind == N.TB.W_TABLE
    s = ImageSlicer((100, 100), (32, 16), (16, 8), weight="gaussian")
    assert s.weight.shape == (32, 16) and s.weight_kind == N.TB.W_GAUSSIAN and s.weight_param == 4.0
    assert np.array_equal(s.weight, gaussian_patch_weight(32, 16))
    assert s.weight.max() <= 1.0 and s.weight[15, 7] == s.weight[16, 8]
    with pytest.raises(KeyError):
        ImageSlicer((100, 100), 32, 16, weight="nope")


def test_argument_errors_match_reference():
    with pytest.raises(ValueError):
        ImageSlicer((100, 100), 32)  # default tile_step=0 always raises (tiles.py:62,92)
    with pytest.raises(ValueError):
        ImageSlicer((100, 100), 32, 33)
    with pytest.raises(ValueError):
        ImageSlicer((100, 100), (32, 32, 32), 16)
    with pytest.raises(ValueError):
        ImageSlicer((100, 100), 32, (16,))
    s = ImageSlicer((100, 90), 32, 16)
    with pytest.raises(ValueError):
        s.merge([np.zeros((32, 32, 1))])  # wrong number of tiles
    with pytest.raises(AssertionError):
        s.crop_to_orignal_size(np.zeros((5, 5)))
    padded = np.zeros(s.target_shape + (2,))
    assert s.crop_to_orignal_size(padded).shape == (100, 90, 2)


def test_no_cpu_fallback_without_cuda():
    if torch.cuda.is_available():
        pytest.skip("CUDA present: the no-CUDA error path cannot be exercised")
    x = torch.rand(2, 3, 8, 8)
    for fn in (tta.d4_image_augment, tta.fliplr_image_augment, F.torch_fliplr, F.torch_rot90_cw):
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            fn(x)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        tta.d4_image_deaugment(torch.rand(8, 3, 8, 8))
    s = ImageSlicer((64, 64, 3), 32, 16)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        s.split(np.zeros((64, 64, 3), np.uint8))
    from pytorch_toolbelt_b200.inference.tiles import TileMerger

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        TileMerger((64, 64), 1, s.weight)


def test_tta_host_side_semantics(golden):
    g = golden("tta")
    with pytest.raises(ValueError):
        tta.d4_image_augment(torch.rand(1, 1, 4, 6))  # non-square (tta.py:403-407)
    lg = torch.from_numpy(g["labels_logits"])
    assert np.array_equal(tta.d4_labels_deaugment(lg).numpy(), g["labels_d4_deaug"])  # keeps the b7-twice quirk
    assert np.array_equal(tta.d2_labels_deaugment(lg).numpy(), g["labels_d2_deaug"])
    assert np.array_equal(tta.fliplr_labels_deaugment(lg).numpy(), g["labels_fliplr_deaug"])
    with pytest.raises(RuntimeError):
        tta.d4_labels_deaugment(torch.rand(7, 3))
    with pytest.raises(RuntimeError):
        tta.flips_labels_deaugment(torch.rand(7, 3))
    with pytest.raises(KeyError):
        tta._deaugment_averaging(torch.rand(2, 3), "bogus")
    assert tta.d4_labels_augment(lg).shape == (128, 5)
    x = torch.arange(16.0).view(1, 1, 4, 4)
    assert tta.fivecrop_image_augment(x, (2, 2)).shape == (5, 1, 2, 2)
    with pytest.raises(ValueError):
        tta.fivecrop_image_augment(x, (5, 2))


def test_d4_view_codes_form_the_group():
    codes = tta._VIEW_CODES[N.TB.TTA_D4]
    assert sorted(codes) == list(range(8))
    for c in range(8):
        assert F._compose(c, F._inverse_code(c)) == F.CODE_ID
    # the compositions named by the reference primitives (inference/functional.py:55-95)
    assert F._compose(F.CODE_ROT90_CCW, F.CODE_TRANSPOSE) == F.CODE_FLIPLR
    assert F._compose(F.CODE_ROT90_CW, F.CODE_TRANSPOSE) == F.CODE_FLIPUD
    assert F._compose(F.CODE_ROT180, F.CODE_TRANSPOSE) == F.CODE_ANTITRANSPOSE


def test_prepared_launch_falls_back_once_on_unsupported():
    from pytorch_toolbelt_b200.inference.engine import PreparedLaunch

    class HostLaunch(PreparedLaunch):  # no CUDA device to select on this host
        __slots__ = ()

        def _call(self):
            return self.fn(*self.args)

    calls = []

    def first(*args):
        calls.append(("first", args))
        return -2  # TB_ERR_UNSUPPORTED: nothing launched

    def second(*args):
        calls.append(("second", args))
        return 0

    plan = HostLaunch(first, (1, 2), [], "result", None, "first", fallback=lambda: (second, (3,), "second"))
    assert plan.run() == "result" and plan.run() == "result"
    assert [c[0] for c in calls] == ["first", "second", "second"] and calls[1][1] == (3,)
    with pytest.raises(N.TB200Error):
        HostLaunch(first, (), [], None, None, "tb_merge_all_batched").run()


def test_clock_sampler_summary_and_fallbacks():
    from pytorch_toolbelt_b200.utils.clocks import ClockSampler

    s = ClockSampler()
    assert s.summary()["reasons"] == ["unavailable"] and s.summary()["samples"] == 0
    s.sm, s.max_mhz, s.source = [1965.0, 1950.0, 1965.0], 1965.0, "nvml"
    s.reasons.add("sw_power_cap")
    out = s.summary()
    assert out == {"sm_mhz": 1965.0, "sm_max_mhz": 1965.0, "reasons": ["sw_power_cap"], "samples": 3, "source": "nvml"}
    with ClockSampler() as live:  # no NVML / nvidia-smi on the build host: must degrade, not raise
        pass
    assert live.summary()["samples"] >= 0


def test_joint_loss_fusability_rules():
    import pytorch_toolbelt_b200.losses as L

    x, t = torch.zeros(2, 3, 8, 8), torch.zeros(2, 3, 8, 8)
    labels = torch.zeros(2, 8, 8, dtype=torch.long)
    assert L.BinaryFocalDiceLoss(L.BinaryFocalLoss(), L.DiceLoss("multilabel")).fusable(x, t)
    assert L.BinaryFocalDiceLoss(L.BinaryFocalLoss(reduction="sum"), L.JaccardLoss("binary")).fusable(x, t)
    assert not L.BinaryFocalDiceLoss(L.BinaryFocalLoss(), L.DiceLoss("multiclass")).fusable(x, labels)
    assert not L.BinaryFocalDiceLoss(L.BinaryFocalLoss(normalized=True), L.DiceLoss("multilabel")).fusable(x, t)
    assert not L.BinaryFocalDiceLoss(L.BinaryFocalLoss(reduction="none"), L.DiceLoss("multilabel")).fusable(x, t)
    assert not L.BinaryFocalDiceLoss(L.BinaryFocalLoss(), L.DiceLoss("multilabel", from_logits=False)).fusable(x, t)
    assert not L.BinaryFocalDiceLoss(L.BinaryFocalLoss(class_weights=[1.0, 2.0, 3.0]), L.DiceLoss("binary")).fusable(x, t)
    with pytest.raises(TypeError):
        L.BinaryFocalDiceLoss(L.DiceLoss("binary"), L.DiceLoss("binary"))
