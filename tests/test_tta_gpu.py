"""GPU: fused TTA kernels against the golden vectors of the real reference and the oracle; reference's own
tests/test_tta.py restated; autograd transparency."""
import numpy as np
import pytest
import torch

from oracle import tta_oracle as AO
from pytorch_toolbelt_b200.inference import functional as F
from pytorch_toolbelt_b200.inference import tta

pytestmark = pytest.mark.gpu

AUG = {
    "fliplr": (tta.fliplr_image_augment, tta.fliplr_image_deaugment),
    "flipud": (tta.flipud_image_augment, tta.flipud_image_deaugment),
    "flips": (tta.flips_image_augment, tta.flips_image_deaugment),
    "d2": (tta.d2_image_augment, tta.d2_image_deaugment),
    "d4": (tta.d4_image_augment, tta.d4_image_deaugment),
}
REDUCTIONS = ["mean", "sum", "gmean", "hmean", "harmonic1p", "logodd", "log1p", None]
PRIMITIVES = [
    "torch_rot90_ccw", "torch_rot90_cw", "torch_rot90_ccw_transpose", "torch_rot90_cw_transpose",
    "torch_transpose_rot90_ccw", "torch_transpose_rot90_cw", "torch_rot180", "torch_rot180_transpose",
    "torch_transpose_rot180", "torch_flipud", "torch_fliplr", "torch_transpose", "torch_transpose2",
]


def test_augment_deaugment_vs_golden(golden):
    g = golden("tta")
    for xname in ("sq", "rect", "odd"):
        x = torch.from_numpy(g[f"x_{xname}"]).cuda()
        for group, (aug, deaug) in AUG.items():
            if f"{xname}_{group}_aug" not in g.files:
                continue
            a = aug(x)
            assert a.is_cuda and np.array_equal(a.cpu().numpy(), g[f"{xname}_{group}_aug"]), (xname, group)
            y = torch.from_numpy(g[f"{xname}_{group}_y"]).cuda()
            for red in REDUCTIONS:
                r = deaug(y, reduction=red).cpu().numpy()
                ref = g[f"{xname}_{group}_deaug_{red}"]
                assert r.shape == ref.shape
                if red is None:
                    assert np.array_equal(r, ref), (xname, group, red)
                elif red in ("mean", "sum") and xname != "odd":
                    # sequential view-order sum == torch CPU's order for all but tiny planes (exactness at
                    # realistic sizes is asserted in test_d4_vs_oracle_seeded)
                    assert np.array_equal(r, ref), (xname, group, red)
                else:
                    np.testing.assert_allclose(r, ref, rtol=2e-6, atol=1e-7, err_msg=f"{xname} {group} {red}")


def test_primitives_vs_golden(golden):
    g = golden("tta")
    x = torch.from_numpy(g["x_sq"]).cuda()
    for name in PRIMITIVES:
        assert np.array_equal(getattr(F, name)(x).cpu().numpy(), g[f"prim_{name}"]), name
    xr = torch.from_numpy(g["x_rect"]).cuda()
    for name in ("torch_flipud", "torch_fliplr", "torch_rot180", "torch_transpose", "torch_rot90_cw", "torch_rot90_ccw"):
        out = getattr(F, name)(xr)
        assert np.array_equal(out.cpu().numpy(), g[f"primrect_{name}"]), name
    assert F.torch_none(x) is x


@pytest.mark.parametrize("shape", [(4, 3, 224, 224), (2, 5, 96, 96), (1, 1, 33, 33), (3, 2, 130, 130)])
def test_d4_vs_oracle_seeded(shape):
    g = torch.Generator().manual_seed(sum(shape))
    x = torch.rand(shape, generator=g)
    a = tta.d4_image_augment(x.cuda())
    assert torch.equal(a.cpu(), AO.augment(x, "d4"))
    y = torch.randn((8 * shape[0], 4) + shape[2:], generator=g)
    for red in ("mean", "sum"):
        ours, ref = tta.d4_image_deaugment(y.cuda(), reduction=red).cpu(), AO.deaugment(y, "d4", red)
        if shape[2] % 32 == 0:
            assert torch.equal(ours, ref)  # same sequential view-order sum as torch CPU
        else:
            torch.testing.assert_close(ours, ref, rtol=1e-6, atol=1e-6)  # torch CPU re-associates tiny planes


@pytest.mark.parametrize("group", ["fliplr", "flipud", "flips", "d2"])
def test_flip_groups_nonsquare_vs_oracle(group):
    g = torch.Generator().manual_seed(7)
    x = torch.rand((3, 2, 75, 132), generator=g)
    aug, deaug = AUG[group]
    assert torch.equal(aug(x.cuda()).cpu(), AO.augment(x, group))
    y = torch.randn((len(AO.GROUPS[group]) * 3, 5, 75, 132), generator=g)
    assert torch.equal(deaug(y.cuda()).cpu(), AO.deaugment(y, group, "mean"))
    assert torch.equal(deaug(y.cuda(), reduction=None).cpu(), AO.deaugment(y, group, None))


def test_other_dtypes_are_pure_permutations():
    g = torch.Generator().manual_seed(3)
    for dtype in (torch.uint8, torch.int16, torch.int32, torch.int64, torch.float16, torch.bfloat16, torch.float64):
        x = (torch.rand((2, 3, 20, 20), generator=g) * 200).to(dtype)
        assert torch.equal(tta.d4_image_augment(x.cuda()).cpu(), AO.augment(x, "d4")), dtype
        assert torch.equal(F.torch_rot90_cw(x.cuda()).cpu(), x.rot90(k=-1, dims=(2, 3))), dtype
    y = torch.randn((8, 2, 16, 16), generator=g, dtype=torch.float64)
    assert torch.equal(tta.d4_image_deaugment(y.cuda()).cpu(), AO.deaugment(y, "d4", "mean"))


def test_reference_test_suite_restated():
    """reference tests/test_tta.py:31-108 against the drop-in (CPU tensors in, CPU tensors out)."""
    x = torch.rand((4, 3, 224, 224))
    noop = torch.nn.Identity()
    for out in (tta.d4_image2mask(noop, x), tta.d4_image_deaugment(tta.d4_image_augment(x)),
                tta.flips_image_deaugment(tta.flips_image_augment(x)), tta.d2_image_deaugment(tta.d2_image_augment(x)),
                tta.fliplr_image2mask(noop, x)):
        assert out.device == x.device
        np.testing.assert_allclose(out.numpy(), x.numpy(), atol=1e-6, rtol=1e-6)

    class SumAll(torch.nn.Module):
        def forward(self, input):
            return input.sum(dim=[1, 2, 3])

    x44 = torch.tensor([[1, 2, 3, 4], [5, 6, 7, 8], [9, 0, 1, 2], [3, 4, 5, 6]]).unsqueeze(0).unsqueeze(0).float()
    assert int(tta.d4_image2label(SumAll(), x44)) == int(x44.sum())
    assert int(tta.fliplr_image2label(SumAll(), x44)) == int(x44.sum())
    assert int(tta.fivecrop_image2label(SumAll(), x44, (2, 2))) == (14 + 22 + 16 + 14 + 14) / 5
    assert int(tta.tencrop_image2label(SumAll(), x44, (2, 2))) == (2 * (14 + 22 + 16 + 14 + 14)) / 10


def test_errors():
    with pytest.raises(RuntimeError):
        tta.d4_image_deaugment(torch.rand(7, 3, 8, 8).cuda())
    with pytest.raises(ValueError):
        tta.d4_image_augment(torch.rand(1, 1, 8, 9).cuda())
    with pytest.raises(KeyError):
        tta.d4_image_deaugment(torch.rand(8, 3, 8, 8).cuda(), reduction="bogus")


def test_autograd_matches_reference_ops():
    """Gradients flow through the fused kernels exactly as through the reference's flip/rot90/cat/stack graph."""
    g = torch.Generator().manual_seed(11)
    for group, (aug, deaug) in AUG.items():
        x = torch.rand((2, 3, 16, 16), generator=g)
        w = torch.randn((len(AO.GROUPS[group]) * 2, 3, 16, 16), generator=g)
        xo = x.clone().requires_grad_(True)
        (AO.augment(xo, group) * w).sum().backward()
        xc = x.clone().cuda().requires_grad_(True)
        (aug(xc) * w.cuda()).sum().backward()
        torch.testing.assert_close(xc.grad.cpu(), xo.grad, rtol=1e-6, atol=1e-6)
        for red in ("mean", "sum", None, "gmean", "hmean"):
            y = torch.rand((len(AO.GROUPS[group]) * 2, 3, 16, 16), generator=g) * 0.9 + 0.05
            yo = y.clone().requires_grad_(True)
            ro = AO.deaugment(yo, group, red)
            wr = torch.randn(ro.shape, generator=g)
            (ro * wr).sum().backward()
            yc = y.clone().cuda().requires_grad_(True)
            (deaug(yc, reduction=red) * wr.cuda()).sum().backward()
            torch.testing.assert_close(yc.grad.cpu(), yo.grad, rtol=1e-5, atol=1e-6)
    xo = torch.rand((1, 2, 8, 8), generator=g)
    xc = xo.clone().cuda().requires_grad_(True)
    wr = torch.randn((1, 2, 8, 8), generator=g)
    (F.torch_rot90_cw(xc) * wr.cuda()).sum().backward()
    xr = xo.clone().requires_grad_(True)
    (xr.rot90(k=-1, dims=(2, 3)) * wr).sum().backward()
    torch.testing.assert_close(xc.grad.cpu(), xr.grad)


def test_generalized_tta_wrapper():
    model = torch.nn.Conv2d(3, 2, 3, padding=1).cuda()
    wrapped = tta.GeneralizedTTA(model, tta.d4_image_augment, tta.d4_image_deaugment)
    x = torch.rand(2, 3, 32, 32).cuda()
    with torch.no_grad():
        out = wrapped(x)
        ref = AO.deaugment(model(AO.augment(x, "d4")).cpu(), "d4", "mean")
    torch.testing.assert_close(out.cpu(), ref, rtol=1e-5, atol=1e-6)
    ms = tta.MultiscaleTTA(model, [0, 16, -16])
    with torch.no_grad():
        assert ms(x).shape == (2, 2, 32, 32)


def test_tma_and_plain_deaugment_kernels_agree_bitwise():
    """tb_tta_deaugment / tb_deaugment_merge have a TMA-ring and a plain-load implementation; same arithmetic."""
    import os

    from pytorch_toolbelt_b200.inference.tiles import CudaTileMerger, ImageSlicer
    from pytorch_toolbelt_b200._native import TB

    g = torch.Generator().manual_seed(21)
    for group, (aug, deaug) in AUG.items():
        A = len(AO.GROUPS[group])
        y = (torch.rand((A * 3, 5, 96, 96), generator=g) * 0.9 + 0.05).cuda()
        for red in ("mean", "sum", "gmean", "logodd"):
            fast = deaug(y, reduction=red)
            os.environ["TB200_FORCE_GENERIC_MERGE"] = "1"
            try:
                slow = deaug(y, reduction=red)
            finally:
                del os.environ["TB200_FORCE_GENERIC_MERGE"]
            assert torch.equal(fast, slow), (group, red)
    s = ImageSlicer((700, 500, 3), 128, 64, weight="pyramid")
    pred = torch.randn(8 * len(s.crops), 3, 128, 128, generator=g).cuda()
    results = []
    for force in (False, True):
        if force:
            os.environ["TB200_FORCE_GENERIC_MERGE"] = "1"
        try:
            m = CudaTileMerger(s.target_shape, 3, s.weight)
            for b0 in range(0, len(s.crops), 9):
                idx = torch.arange(b0, min(b0 + 9, len(s.crops)))
                chunk = pred.view(8, len(s.crops), 3, 128, 128)[:, idx].reshape(-1, 3, 128, 128).contiguous()
                m.integrate_batch(chunk, s.crops[idx.numpy()], tta=TB.TTA_D4)
            results.append((m.image.clone(), m.norm_mask.clone()))
        finally:
            os.environ.pop("TB200_FORCE_GENERIC_MERGE", None)
    # float atomics: order differs between runs, compare to buffer scale
    scale = results[1][0].abs().max()
    assert (results[0][0] - results[1][0]).abs().max() <= 1e-5 * scale
    assert (results[0][1] - results[1][1]).abs().max() <= 1e-5 * results[1][1].abs().max()
