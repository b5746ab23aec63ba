"""GPU: the fused kernels of SURVEY 8f rank 4 (SoftBCE, BalancedBCE, QualityFocal, BinarySoftF1, SoftF1) against the
reference's outputs (tests/golden/losses_extra.npz) and, on larger seeded inputs, against the oracle restatement."""
import numpy as np
import pytest
import torch

import pytorch_toolbelt_b200.losses as L
from oracle import losses_extra_oracle as LE

pytestmark = pytest.mark.gpu

W3 = lambda: torch.tensor([1.5, 0.7, 1.0]).view(3, 1, 1)  # noqa: E731
PW3 = lambda: torch.tensor([0.5, 2.0, 1.5]).view(3, 1, 1)  # noqa: E731
OURS = {
    "softbce_default": (lambda: L.SoftBCEWithLogitsLoss(), "x", "t_ign"),
    "softbce_smooth_sum": (lambda: L.SoftBCEWithLogitsLoss(smooth_factor=0.1, reduction="sum"), "x", "t_ign"),
    "softbce_weights": (lambda: L.SoftBCEWithLogitsLoss(weight=W3(), pos_weight=PW3(), ignore_index=None), "x", "t"),
    "softbce_none": (lambda: L.SoftBCEWithLogitsLoss(reduction="none", pos_weight=torch.tensor([3.0])), "x", "t_ign"),
    "balbce_default": (lambda: L.BalancedBCEWithLogitsLoss(), "x", "t"),
    "balbce_gamma_ignore": (lambda: L.BalancedBCEWithLogitsLoss(gamma=0.7, ignore_index=-100), "x", "t_ign"),
    "balbce_sum": (lambda: L.BalancedBCEWithLogitsLoss(gamma=2.0, reduction="sum"), "x", "t"),
    "qfl_mean": (lambda: L.QualityFocalLoss(), "x", "t_soft"),
    "qfl_sum_beta1p5": (lambda: L.QualityFocalLoss(beta=1.5, reduction="sum"), "x", "t_soft"),
    "qfl_normalized": (lambda: L.QualityFocalLoss(reduction="normalized"), "x", "t_soft"),
    "qfl_none": (lambda: L.QualityFocalLoss(reduction="none"), "x", "t_soft"),
    "bsf1": (lambda: L.BinarySoftF1Loss(), "x", "t"),
    "bsf1_ignore": (lambda: L.BinarySoftF1Loss(ignore_index=-100), "x", "t_ign"),
    "sf1": (lambda: L.SoftF1Loss(), "xc", "tc"),
    "sf1_ignore0": (lambda: L.SoftF1Loss(ignore_index=0), "xc", "tc"),
    "sf1_ignore7": (lambda: L.SoftF1Loss(ignore_index=7), "xc", "tc"),
}


@pytest.mark.parametrize("name", sorted(OURS))
def test_against_the_reference_outputs(name, golden):
    g = golden("losses_extra")
    factory, xk, tk = OURS[name]
    module = factory().cuda()
    x = torch.from_numpy(g[xk]).cuda().requires_grad_(True)
    loss = module(x, torch.from_numpy(g[tk]).cuda())
    (loss.sum() if loss.dim() else loss).backward()
    want, want_grad = g[f"{name}_loss"], g[f"{name}_grad"]
    np.testing.assert_allclose(loss.detach().cpu().numpy(), want, rtol=1e-5, atol=1e-6 * max(1.0, float(np.abs(want).max())), err_msg=name)
    scale = float(np.abs(want_grad).max())
    np.testing.assert_allclose(x.grad.cpu().numpy(), want_grad, rtol=1e-4, atol=2e-6 * scale + 1e-12, err_msg=name)


def test_seeded_large_and_odd_shapes_against_the_oracle():
    g = torch.Generator().manual_seed(9)
    for shape in ((4, 4, 256, 256), (3, 5, 7, 9)):  # vectorised / scalar kernels
        x = torch.randn(*shape, generator=g) * 3
        t = (torch.rand(*shape, generator=g) < 0.3).float()
        ti = t.clone()
        ti[torch.rand(*shape, generator=g) < 0.02] = 255
        ts = t * torch.rand(*shape, generator=g)
        C = shape[1]
        pw = torch.rand(C, generator=g).view(C, 1, 1) + 0.5
        cases = [
            (L.SoftBCEWithLogitsLoss(ignore_index=255, smooth_factor=0.05, pos_weight=pw), lambda z: LE.soft_bce_with_logits(z, ti, ignore_index=255, smooth_factor=0.05, pos_weight=pw), ti),
            (L.BalancedBCEWithLogitsLoss(gamma=1.3, ignore_index=255), lambda z: LE.balanced_bce_with_logits(z, ti, 1.3, 255), ti),
            (L.QualityFocalLoss(reduction="normalized"), lambda z: LE.quality_focal_loss(z, ts, 2, "normalized"), ts),
            (L.QualityFocalLoss(beta=1.0), lambda z: LE.quality_focal_loss(z, ts, 1.0), ts),
            (L.BinarySoftF1Loss(ignore_index=255), lambda z: LE.binary_soft_f1(z, ti, 255), ti),
        ]
        for module, oracle, target in cases:
            xo = x.clone().requires_grad_(True)
            lo = oracle(xo)
            lo.backward()
            xc = x.clone().cuda().requires_grad_(True)
            lc = module.cuda()(xc, target.cuda())
            lc.backward()
            np.testing.assert_allclose(float(lc), float(lo), rtol=1e-5, err_msg=type(module).__name__)
            scale = xo.grad.abs().max().item()
            np.testing.assert_allclose(xc.grad.cpu().numpy(), xo.grad.numpy(), rtol=1e-4, atol=2e-6 * scale + 1e-12,
                                       err_msg=type(module).__name__)
    xc_, tc_ = torch.randn(1000, 17, generator=g) * 2, torch.randint(0, 17, (1000,), generator=g)
    xo = xc_.clone().requires_grad_(True)
    lo = LE.soft_f1(xo, tc_)
    lo.backward()
    xg = xc_.clone().cuda().requires_grad_(True)
    lg = L.SoftF1Loss()(xg, tc_.cuda())
    lg.backward()
    np.testing.assert_allclose(float(lg), float(lo), rtol=1e-5)
    np.testing.assert_allclose(xg.grad.cpu().numpy(), xo.grad.numpy(), rtol=1e-4, atol=2e-6 * xo.grad.abs().max().item())


def test_argument_errors_and_edge_cases():
    x = torch.randn(2, 3, 8, 8).cuda()
    t = (torch.rand(2, 3, 8, 8) < 0.5).float().cuda()
    with pytest.raises(ValueError):
        L.SoftBCEWithLogitsLoss()(x, t[:, :2])
    with pytest.raises(NotImplementedError):
        L.SoftBCEWithLogitsLoss(pos_weight=torch.ones(8))(x, t)  # per-column weights: not a channel vector
    with pytest.raises(ValueError):
        L.SoftF1Loss()(x, t)
    # everything ignored: BinarySoftF1Loss returns 0 with a zero gradient (soft_f1.py:71-72)
    xr = x.clone().requires_grad_(True)
    loss = L.BinarySoftF1Loss(ignore_index=1)(xr, torch.ones_like(t))
    loss.backward()
    assert float(loss) == 0.0 and float(xr.grad.abs().max()) == 0.0
    # CPU in -> CPU out
    xc = x.cpu().clone().requires_grad_(True)
    lc = L.QualityFocalLoss()(xc, t.cpu() * 0.5)
    lc.backward()
    assert lc.device.type == "cpu" and xc.grad.device.type == "cpu"
    np.testing.assert_allclose(float(lc), float(LE.quality_focal_loss(x.cpu(), t.cpu() * 0.5)), rtol=1e-5)
