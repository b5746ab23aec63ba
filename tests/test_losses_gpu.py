"""GPU: fused loss kernels against the golden vectors of the real reference (values and gradients), against the
oracle on larger seeded inputs, and the reference's own known-answer tests (tests/test_losses.py) restated.
Tolerances: loss values rtol 1e-5; gradients rtol 1e-4 / atol 1e-7*scale (BASELINE north_star)."""
import numpy as np
import pytest
import torch

from oracle import losses_oracle as LO
from pytorch_toolbelt_b200 import losses as L
from pytorch_toolbelt_b200.losses import functional as LF

pytestmark = pytest.mark.gpu


def run(fn, x):
    x = torch.from_numpy(np.asarray(x)).cuda().requires_grad_(True)
    loss = fn(x)
    (loss.sum() if loss.dim() else loss).backward()
    return loss.detach().cpu().numpy(), x.grad.cpu().numpy()


def check(name, g, fn, xkey="x", loss_rtol=1e-5, grad_rtol=1e-4):
    loss, grad = run(fn, g[xkey])
    ref_l, ref_g = g[f"{name}_loss"], g[f"{name}_grad"]
    np.testing.assert_allclose(loss, ref_l, rtol=loss_rtol, atol=1e-6 * max(1.0, np.abs(ref_l).max()), err_msg=name)
    scale = max(np.abs(ref_g).max(), 1e-12)
    np.testing.assert_allclose(grad, ref_g, rtol=grad_rtol, atol=2e-7 * scale + 1e-9, err_msg=name)


def test_sigmoid_focal_vs_golden(golden):
    g = golden("losses")
    T = lambda k: torch.from_numpy(g[k]).cuda()  # noqa: E731
    t_ml, t_ml_ign, t_mc, t_mc_ign, cw = T("t_ml"), T("t_ml_ign"), T("t_mc"), T("t_mc_ign"), T("cw")
    cfgs = {
        "default": dict(),
        "alpha": dict(alpha=0.25),
        "gamma0": dict(gamma=0.0),
        "gamma1p5_sum": dict(gamma=1.5, reduction="sum"),
        "batchwise": dict(reduction="batchwise_mean"),
        "none": dict(reduction="none"),
        "normalized": dict(normalized=True),
        "reduced": dict(reduced_threshold=0.5, reduction="sum"),
        "reduced03_norm": dict(reduced_threshold=0.3, normalized=True),
        "classw": dict(alpha=0.6, class_weights=cw),
        "softmax_act": dict(activation="softmax", softmax_dim=1),
    }
    for name, kw in cfgs.items():
        k = dict(alpha=None)
        k.update(kw)
        check(f"focal_{name}", g, lambda z, k=k: LF.focal_loss_with_logits(z, t_ml, **k))
    check("focal_ignore", g, lambda z: LF.focal_loss_with_logits(z, t_ml_ign, alpha=0.25, ignore_index=255))
    check("focal_ignore_norm", g, lambda z: LF.focal_loss_with_logits(z, t_ml_ign, alpha=None, ignore_index=255, normalized=True))
    check("binaryfocal_labels", g, lambda z: L.BinaryFocalLoss()(z, t_mc))
    check("binaryfocal_labels_ignore", g, lambda z: L.BinaryFocalLoss(ignore_index=-100, alpha=0.3)(z, t_mc_ign))
    assert "BinaryFocalLoss(alpha=None" in repr(L.BinaryFocalLoss())  # the reference's __repr__ crashes here (A.6)


def test_softmax_focal_vs_golden(golden):
    g = golden("losses")
    t_mc, t_mc_ign = torch.from_numpy(g["t_mc"]).cuda(), torch.from_numpy(g["t_mc_ign"]).cuda()
    cw = torch.from_numpy(g["cw"]).cuda()
    cfgs = {"default": dict(), "gamma1_sum": dict(gamma=1.0, reduction="sum"), "none": dict(reduction="none"),
            "batchwise": dict(reduction="batchwise_mean"), "normalized": dict(normalized=True),
            "reduced": dict(reduced_threshold=0.5), "classw": dict(class_weights=cw)}
    for name, kw in cfgs.items():
        check(f"smfocal_{name}", g, lambda z, k=kw: LF.softmax_focal_loss_with_logits(z, t_mc_ign, **k))
    check("cefocal_module", g, lambda z: L.CrossEntropyFocalLoss()(z, t_mc))


def test_dice_jaccard_vs_golden(golden):
    g = golden("losses")
    T = lambda k: torch.from_numpy(g[k]).cuda()  # noqa: E731
    t_ml, t_ml_ign, t_mc, t_mc_ign, tb = T("t_ml"), T("t_ml_ign"), T("t_mc"), T("t_mc_ign"), T("tb")
    dice = {
        "ml": ("multilabel", t_ml, dict()),
        "ml_log_smooth": ("multilabel", t_ml, dict(log_loss=True, smooth=1.0)),
        "ml_ignore": ("multilabel", t_ml_ign, dict(ignore_index=255)),
        "mc": ("multiclass", t_mc, dict()),
        "mc_smooth": ("multiclass", t_mc, dict(smooth=0.5)),
        "mc_ignore": ("multiclass", t_mc_ign, dict(ignore_index=-100)),
        "mc_log": ("multiclass", t_mc, dict(log_loss=True)),
    }
    for name, (mode, tt, kw) in dice.items():
        check(f"dice_{name}", g, lambda z, m=mode, t_=tt, k=kw: L.DiceLoss(m, **k)(z, t_))
    check("dice_bin", g, lambda z: L.DiceLoss("binary")(z, tb), xkey="xb")
    check("dice_bin_ignore", g, lambda z: L.DiceLoss("binary", ignore_index=255)(z, t_ml_ign[:, :1]), xkey="xb")
    check("dice_ml_probs", g, lambda z: L.DiceLoss("multilabel", from_logits=False)(z, t_ml), xkey="probs")
    check("dice_ml_emptyclass", g, lambda z: L.DiceLoss("multilabel")(z, T("t_empty")))
    jac = {"ml": ("multilabel", t_ml, dict()), "ml_log_smooth": ("multilabel", t_ml, dict(log_loss=True, smooth=1.0)),
           "mc": ("multiclass", t_mc, dict()), "mc_smooth": ("multiclass", t_mc, dict(smooth=0.5))}
    for name, (mode, tt, kw) in jac.items():
        check(f"jaccard_{name}", g, lambda z, m=mode, t_=tt, k=kw: L.JaccardLoss(m, **k)(z, t_))
    check("jaccard_bin", g, lambda z: L.JaccardLoss("binary")(z, tb), xkey="xb")
    probs = T("probs").reshape(2, 3, -1)
    np.testing.assert_allclose(LF.soft_dice_score(probs, t_ml.reshape(2, 3, -1), dims=(0, 2)).cpu().numpy(), g["score_dice"], rtol=1e-6)
    np.testing.assert_allclose(LF.soft_jaccard_score(probs, t_ml.reshape(2, 3, -1), dims=(0, 2)).cpu().numpy(), g["score_jaccard"], rtol=1e-6)


def test_wing_and_lovasz_vs_golden(golden):
    g = golden("losses")
    tw = torch.from_numpy(g["t_wing"]).cuda()
    for name, kw in {"mean": dict(), "sum": dict(reduction="sum"), "none": dict(reduction="none"),
                     "w10c2": dict(width=10, curvature=2.0)}.items():
        check(f"wing_{name}", g, lambda z, k=kw: LF.wing_loss(z, tw, **k))
    check("wing_mean", g, lambda z: L.WingLoss()(z, tw))
    lb, lb_ign = torch.from_numpy(g["lb"]).cuda(), torch.from_numpy(g["lb_ign"]).cuda()
    t_mc, t_mc255 = torch.from_numpy(g["t_mc"]).cuda(), torch.from_numpy(g["t_mc_ign255"]).cuda()
    check("lovasz_hinge", g, lambda z: L.BinaryLovaszLoss()(z, lb), xkey="xl")
    check("lovasz_hinge_perimage", g, lambda z: L.BinaryLovaszLoss(per_image=True)(z, lb), xkey="xl")
    check("lovasz_hinge_ignore", g, lambda z: L.BinaryLovaszLoss(ignore_index=255)(z, lb_ign), xkey="xl")
    check("lovasz_softmax", g, lambda z: L.LovaszLoss()(z, t_mc), xkey="pr")
    check("lovasz_softmax_perimage", g, lambda z: L.LovaszLoss(per_image=True)(z, t_mc), xkey="pr")
    check("lovasz_softmax_ignore", g, lambda z: L.LovaszLoss(ignore=255)(z, t_mc255), xkey="pr")


def _cmp_oracle(ours_fn, oracle_fn, x, loss_rtol=1e-5, grad_rtol=1e-4, grad_atol=None, tag=""):
    import inspect

    tag = tag or f"case at line {inspect.currentframe().f_back.f_lineno}"
    xo = x.clone().requires_grad_(True)
    lo = oracle_fn(xo)
    lo.backward()
    xc = x.clone().cuda().requires_grad_(True)
    lc = ours_fn(xc)
    lc.backward()
    np.testing.assert_allclose(lc.item(), lo.item(), rtol=loss_rtol, err_msg=tag)
    scale = xo.grad.abs().max().item()
    atol = 2e-7 * scale + 1e-12 if grad_atol is None else grad_atol
    np.testing.assert_allclose(xc.grad.cpu().numpy(), xo.grad.numpy(), rtol=grad_rtol, atol=atol, err_msg=tag)


def test_losses_vs_oracle_seeded_large():
    """SURVEY 8(c) inputs at a size the oracle finishes in seconds: x = randn*3, Bernoulli(0.3) / randint targets,
    1 % ignore pixels."""
    g = torch.Generator().manual_seed(0)
    B, C, H, W = 4, 4, 256, 256
    x = torch.randn(B, C, H, W, generator=g) * 3
    t_ml = (torch.rand(B, C, H, W, generator=g) < 0.3).float()
    t_mc = torch.randint(0, C, (B, H, W), generator=g)
    t_mc_ign = t_mc.clone()
    t_mc_ign[torch.rand(B, H, W, generator=g) < 0.01] = -100
    _cmp_oracle(lambda z: L.BinaryFocalLoss()(z, t_ml.cuda()), lambda z: LO.focal_sigmoid(z, t_ml, alpha=None), x)
    _cmp_oracle(lambda z: L.BinaryFocalLoss(alpha=0.25, gamma=1.5)(z, t_mc.cuda()),
                lambda z: LO.focal_sigmoid(z, LO.one_hot_targets(t_mc, C), alpha=0.25, gamma=1.5), x)
    _cmp_oracle(lambda z: L.CrossEntropyFocalLoss()(z, t_mc_ign.cuda()), lambda z: LO.focal_softmax(z, t_mc_ign), x)
    _cmp_oracle(lambda z: L.DiceLoss("multilabel")(z, t_ml.cuda()), lambda z: LO.dice_loss(z, t_ml, "multilabel"), x)
    _cmp_oracle(lambda z: L.DiceLoss("multiclass", ignore_index=-100)(z, t_mc_ign.cuda()),
                lambda z: LO.dice_loss(z, t_mc_ign, "multiclass", ignore_index=-100), x)
    _cmp_oracle(lambda z: L.JaccardLoss("multilabel", log_loss=True)(z, t_ml.cuda()),
                lambda z: LO.jaccard_loss(z, t_ml, "multilabel", log_loss=True), x)
    _cmp_oracle(lambda z: L.JaccardLoss("multiclass")(z, t_mc.cuda()), lambda z: LO.jaccard_loss(z, t_mc, "multiclass"), x)
    tw = torch.randn(B, C, H, W, generator=g) * 4
    _cmp_oracle(lambda z: L.WingLoss()(z, tw.cuda()), lambda z: LO.wing_loss(z, tw), x)
    # focal + dice summed, as in BASELINE config 3
    _cmp_oracle(lambda z: L.BinaryFocalLoss()(z, t_ml.cuda()) + L.DiceLoss("multilabel")(z, t_ml.cuda()),
                lambda z: LO.focal_sigmoid(z, t_ml, alpha=None) + LO.dice_loss(z, t_ml, "multilabel"), x)
    # upstream gradient other than 1 goes through the on-device scaling
    _cmp_oracle(lambda z: 0.37 * L.DiceLoss("multilabel")(z, t_ml.cuda()), lambda z: 0.37 * LO.dice_loss(z, t_ml, "multilabel"), x)
    # lovasz: every gradient entry is a DIFFERENCE of two O(1) float32 Jaccard values (lovasz.py:31-33), so the
    # reference itself carries +-1 ulp(1.0) = 1.2e-7 of absolute noise per entry, and so do we: two differences of
    # rounded values can be 4 ulp apart, and torch.sort's tie order (not deterministic on the GPU) decides which
    # entries meet -> atol 6e-7
    lb = (torch.rand(B, H, W, generator=g) < 0.4).long()
    _cmp_oracle(lambda z: L.BinaryLovaszLoss()(z, lb.cuda()), lambda z: LO.lovasz_hinge(z, lb), x[:, 0].contiguous(),
                grad_rtol=1e-3, grad_atol=6e-7)
    pr = torch.softmax(x, dim=1)
    _cmp_oracle(lambda z: L.LovaszLoss()(z, t_mc.cuda()), lambda z: LO.lovasz_softmax(z, t_mc), pr, grad_rtol=1e-3,
                grad_atol=6e-7)


def test_odd_shapes_and_many_classes():
    g = torch.Generator().manual_seed(3)
    x = torch.randn(3, 5, 7, 9, generator=g) * 2  # inner = 63: scalar kernels
    t = (torch.rand(3, 5, 7, 9, generator=g) < 0.4).float()
    lab = torch.randint(0, 5, (3, 7, 9), generator=g)
    _cmp_oracle(lambda z: LF.focal_loss_with_logits(z, t.cuda(), alpha=0.25), lambda z: LO.focal_sigmoid(z, t), x)
    _cmp_oracle(lambda z: L.DiceLoss("multilabel")(z, t.cuda()), lambda z: LO.dice_loss(z, t, "multilabel"), x)
    _cmp_oracle(lambda z: L.DiceLoss("multiclass")(z, lab.cuda()), lambda z: LO.dice_loss(z, lab, "multiclass"), x)
    x19 = torch.randn(2, 19, 32, 32, generator=g)
    l19 = torch.randint(0, 19, (2, 32, 32), generator=g)
    _cmp_oracle(lambda z: L.CrossEntropyFocalLoss()(z, l19.cuda()), lambda z: LO.focal_softmax(z, l19), x19)
    _cmp_oracle(lambda z: L.JaccardLoss("multiclass")(z, l19.cuda()), lambda z: LO.jaccard_loss(z, l19, "multiclass"), x19)
    _cmp_oracle(lambda z: L.DiceLoss("multiclass", classes=[1, 3, 18])(z, l19.cuda()),
                lambda z: LO.dice_loss(z, l19, "multiclass", classes=[1, 3, 18]), x19)


def test_reference_known_answer_tests():
    """reference tests/test_losses.py:11-34,83-254 restated against the drop-in (CPU tensors in, like upstream)."""
    eps = 1e-5
    inp_good = torch.tensor([10, -10, 10]).float()
    inp_bad = torch.tensor([-1, 2, 0]).float()
    tgt = torch.tensor([1, 0, 1])
    assert LF.focal_loss_with_logits(inp_good, tgt) < LF.focal_loss_with_logits(inp_bad, tgt)
    inp_good = torch.tensor([[0, 10, 0], [10, 0, 0], [0, 0, 10]]).float()
    inp_bad = torch.tensor([[0, -10, 0], [0, 10, 0], [0, 0, 10]]).float()
    tgt = torch.tensor([1, 0, 2]).long()
    assert LF.softmax_focal_loss_with_logits(inp_good, tgt) < LF.softmax_focal_loss_with_logits(inp_bad, tgt)

    crit = L.DiceLoss(mode="binary", from_logits=False)
    cases = [([1.0, 1.0, 1.0], [1, 1, 1], 0.0), ([1.0, 0.0, 1.0], [1, 0, 1], 0.0), ([0, 0, 0], [0, 0, 0], 0.0),
             ([1, 1, 1], [0, 0, 0], 0.0), ([1, 0, 1], [0, 1, 0], 1.0), ([0, 0, 0], [1, 1, 1], 1.0)]
    for yp, yt, want in cases:
        yp = torch.tensor(yp).float().view(1, 1, 1, -1)
        yt = torch.tensor(yt).view(1, 1, 1, -1)
        assert float(crit(yp, yt)) == pytest.approx(want, abs=eps)

    crit = L.JaccardLoss(mode="binary", from_logits=False)
    cases = [([1.0, 1.0, 1.0], [1, 1, 1], 0.0), ([1.0, 0.0, 1.0], [1, 0, 1], 0.0), ([1, 1, 1], [0, 0, 0], 0.0),
             ([1, 0, 1], [0, 1, 0], 1.0), ([0, 0, 0], [1, 1, 1], 1.0)]
    for yp, yt, want in cases:
        yp = torch.tensor(yp).float().view(1, 1, 1, -1)
        yt = torch.tensor(yt).view(1, 1, 1, -1)
        assert float(crit(yp, yt)) == pytest.approx(want, abs=eps)

    # reference tests/test_losses.py:160-209 (same vectors)
    crit = L.JaccardLoss(mode="multiclass", from_logits=False)
    y_pred = torch.tensor([[[1.0, 1.0, 0.0, 0.0], [0.0, 0.0, 1.0, 1.0]]])
    assert float(crit(y_pred, torch.tensor([[0, 0, 1, 1]]))) == pytest.approx(0.0, abs=eps)
    assert float(crit(y_pred, torch.tensor([[1, 1, 0, 0]]))) == pytest.approx(1.0, abs=eps)
    y_pred = torch.tensor([[[1.0, 0.0, 1.0, 0.0], [0.0, 1.0, 0.0, 1.0]]])
    assert float(crit(y_pred, torch.tensor([[1, 1, 0, 0]]))) == pytest.approx(1.0 - 1.0 / 3.0, abs=eps)

    crit = L.JaccardLoss(mode="multilabel", from_logits=False)
    y_pred = torch.tensor([[[1.0, 1.0, 0.0, 0.0], [0.0, 0.0, 1.0, 1.0]]])
    assert float(crit(y_pred, y_pred.clone())) == pytest.approx(0.0, abs=eps)
    assert float(crit(y_pred, 1 - y_pred)) == pytest.approx(1.0, abs=eps)
    y_pred = torch.tensor([[[0.0, 1.0, 1.0, 0.0], [0.0, 1.0, 1.0, 0.0]]])
    y_true = torch.tensor([[[1.0, 1.0, 0.0, 0.0], [1.0, 1.0, 0.0, 0.0]]])
    assert float(crit(y_pred, y_true)) == pytest.approx(1.0 - 1.0 / 3.0, abs=eps)

    x = torch.randn((4, 3, 4, 4))
    y = torch.randint(0, 3, (4, 4, 4))
    for crit in (L.BinaryFocalLoss(), L.CrossEntropyFocalLoss()):
        loss = crit(x, y)
        assert loss.dim() == 0 and torch.isfinite(loss) and loss.device.type == "cpu"
    xg = x.clone().requires_grad_(True)
    L.BinaryFocalLoss()(xg, y).backward()
    assert xg.grad is not None and xg.grad.device.type == "cpu"


def test_more_than_65535_batch_class_rows():
    """(batch, class) rows used to ride on gridDim.y (<= 65535); the stats kernels now fold them into gridDim.x."""
    g = torch.Generator().manual_seed(11)
    x = torch.randn(70000, 1, 4, 4, generator=g) * 2
    t = (torch.rand(70000, 1, 4, 4, generator=g) < 0.4).float()
    _cmp_oracle(lambda z: L.DiceLoss("binary")(z, t.cuda()), lambda z: LO.dice_loss(z, t, "binary"), x)
    _cmp_oracle(lambda z: L.BinaryFocalDiceLoss(L.BinaryFocalLoss(), L.DiceLoss("binary"))(z, t.cuda()),
                lambda z: LO.focal_sigmoid(z, t, alpha=None) + LO.dice_loss(z, t, "binary"), x)


def test_cross_entropy_focal_with_150_classes():
    """the pixel-wise focal kernels held <= 64 channels in registers; larger label sets now take the any-C kernel."""
    g = torch.Generator().manual_seed(12)
    C = 150
    x = torch.randn(2, C, 24, 20, generator=g) * 2
    lab = torch.randint(0, C, (2, 24, 20), generator=g)
    lab[torch.rand(2, 24, 20, generator=g) < 0.05] = -100
    cw = torch.rand(C, generator=g) + 0.5
    _cmp_oracle(lambda z: L.CrossEntropyFocalLoss()(z, lab.cuda()), lambda z: LO.focal_softmax(z, lab), x)
    _cmp_oracle(lambda z: L.CrossEntropyFocalLoss(gamma=1.5, normalized=True, class_weights=cw)(z, lab.cuda()),
                lambda z: LO.focal_softmax(z, lab, gamma=1.5, normalized=True, class_weights=cw), x)
    _cmp_oracle(lambda z: L.CrossEntropyFocalLoss(reduced_threshold=0.4, reduction="sum")(z, lab.cuda()),
                lambda z: LO.focal_softmax(z, lab, reduced_threshold=0.4, reduction="sum"), x)
