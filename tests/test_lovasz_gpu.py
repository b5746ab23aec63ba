"""GPU: the fused Lovasz path (error/key build -> segmented radix sort -> scan -> Jaccard increments -> dot -> gradient
scatter, csrc/lovasz.cu) against the oracle's restatement of the reference (losses/lovasz.py) and, at cfg-3 size, against
the same reference ops run with torch.sort on the GPU."""
import numpy as np
import pytest
import torch

import pytorch_toolbelt_b200.losses as L
from oracle import losses_oracle as LO

pytestmark = pytest.mark.gpu


def _both(ours_fn, oracle_fn, x, loss_rtol=1e-5, grad_atol=6e-7, grad_rtol=1e-3, oracle_device="cpu"):
    xo = x.clone().to(oracle_device).requires_grad_(True)
    lo = oracle_fn(xo)
    xc = x.clone().cuda().requires_grad_(True)
    lc = ours_fn(xc)
    np.testing.assert_allclose(float(lc), float(lo), rtol=loss_rtol, atol=1e-7)
    if torch.is_tensor(lo) and lo.requires_grad:
        lo.backward()
        lc.backward()
        scale = max(xo.grad.abs().max().item(), 1e-30)
        np.testing.assert_allclose(xc.grad.cpu().numpy(), xo.grad.cpu().numpy(), rtol=grad_rtol, atol=grad_atol * max(1.0, scale))
    return float(lc)


@pytest.mark.parametrize("per_image", [False, True])
@pytest.mark.parametrize("ignore", [None, 255])
def test_hinge_variants(per_image, ignore):
    g = torch.Generator().manual_seed(1)
    B, H, W = 3, 70, 61  # 4270 pixels per image: radix tiles with a ragged tail, several apply tiles
    x = torch.randn(B, H, W, generator=g) * 2
    lab = (torch.rand(B, H, W, generator=g) < 0.35).long()
    if ignore is not None:
        lab[torch.rand(B, H, W, generator=g) < 0.1] = ignore
        lab[1] = ignore  # a whole image of void pixels
    for labels in (lab, lab.float()):
        _both(lambda z: L.BinaryLovaszLoss(per_image=per_image, ignore_index=ignore)(z, labels.cuda()),
              lambda z: LO.lovasz_hinge(z, labels, per_image=per_image, ignore_index=ignore), x)
    # [B,1,H,W] logits against [B,H,W] labels
    _both(lambda z: L.BinaryLovaszLoss(per_image=per_image, ignore_index=ignore)(z, lab.cuda()),
          lambda z: LO.lovasz_hinge(z[:, 0], lab, per_image=per_image, ignore_index=ignore), x.unsqueeze(1))


@pytest.mark.parametrize("per_image", [False, True])
@pytest.mark.parametrize("ignore", [None, 7])
def test_softmax_variants(per_image, ignore):
    g = torch.Generator().manual_seed(2)
    B, C, H, W = 3, 5, 33, 47
    pr = torch.softmax(torch.randn(B, C, H, W, generator=g) * 2, dim=1)
    lab = torch.randint(0, C - 1, (B, H, W), generator=g)  # class C-1 is absent everywhere: skipped ("present")
    lab[0][lab[0] == 2] = 1                                 # class 2 absent from image 0 only
    if ignore is not None:
        lab[torch.rand(B, H, W, generator=g) < 0.15] = ignore
    _both(lambda z: L.LovaszLoss(per_image=per_image, ignore=ignore)(z, lab.cuda()),
          lambda z: LO.lovasz_softmax(z, lab, per_image=per_image, ignore=ignore), pr)
    # probabilities outside [0,1] are legal inputs (no softmax is applied): sign(fg - p) matters for the gradient
    raw = torch.randn(B, C, H, W, generator=g)
    _both(lambda z: L.LovaszLoss(per_image=per_image, ignore=ignore)(z, lab.cuda()),
          lambda z: LO.lovasz_softmax(z, lab, per_image=per_image, ignore=ignore), raw)


def test_sigmoid_output_single_plane_and_ties():
    g = torch.Generator().manual_seed(3)
    B, H, W = 2, 40, 40
    p = torch.sigmoid(torch.randn(B, H, W, generator=g) * 3)
    lab = (torch.rand(B, H, W, generator=g) < 0.5).long()
    # C == 1 ("sigmoid output", lovasz.py:129-132,145-148): foreground = (label == 0)
    _both(lambda z: L.LovaszLoss()(z, lab.cuda()), lambda z: LO.lovasz_softmax(z, lab), p)
    # saturated probabilities: thousands of exactly equal errors; the loss does not depend on the order inside a tie
    # group, the gradient does (sub-gradient) -> value only
    sat = (p > 0.5).float()
    ours = L.LovaszLoss()(sat.cuda(), lab.cuda())
    want = LO.lovasz_softmax(sat, lab)
    np.testing.assert_allclose(float(ours), float(want), rtol=1e-5)
    quant = torch.round(p * 8) / 8  # 9 distinct values: long tie groups with both labels inside
    ours = L.BinaryLovaszLoss()((quant * 4 - 2).cuda(), lab.cuda())
    want = LO.lovasz_hinge(quant * 4 - 2, lab)
    np.testing.assert_allclose(float(ours), float(want), rtol=1e-5)


def test_argument_handling_matches_the_reference():
    g = torch.Generator().manual_seed(4)
    x = torch.randn(2, 16, 16, generator=g)
    lab = (torch.rand(2, 16, 16, generator=g) < 0.5).long()
    with pytest.raises(TypeError):
        L.BinaryLovaszLoss()(x.cuda().half(), lab.cuda())
    with pytest.raises(ValueError):
        L.BinaryLovaszLoss()(x.cuda(), lab[:1].cuda())
    # CPU tensors in -> CPU loss and CPU gradient out (the arithmetic still runs in the CUDA kernels)
    xc = x.clone().requires_grad_(True)
    loss = L.BinaryLovaszLoss()(xc, lab)
    loss.backward()
    assert loss.device.type == "cpu" and xc.grad.device.type == "cpu"
    np.testing.assert_allclose(float(loss), float(LO.lovasz_hinge(x, lab)), rtol=1e-5)
    # only void pixels: zero loss, zero gradient (lovasz.py:62-64)
    xv = x.clone().cuda().requires_grad_(True)
    lv = L.BinaryLovaszLoss(ignore_index=1)(xv, torch.ones_like(lab).cuda())
    lv.backward()
    assert float(lv) == 0.0 and float(xv.grad.abs().max()) == 0.0
    # upstream gradient other than 1
    xa = x.clone().cuda().requires_grad_(True)
    (0.25 * L.BinaryLovaszLoss()(xa, lab.cuda())).backward()
    xb = x.clone().requires_grad_(True)
    (0.25 * LO.lovasz_hinge(xb, lab)).backward()
    np.testing.assert_allclose(xa.grad.cpu().numpy(), xb.grad.numpy(), rtol=1e-3, atol=2e-7)


def _hinge_float64(x, lab):
    """Lovasz hinge of ONE flat segment with float64 cumsums / Jaccard values (the reference's float32 cumsum stops
    counting at 2**24 = 16.7 M, lovasz.py:29-31): loss and d loss / d x."""
    x, lab = x.reshape(-1).double(), lab.reshape(-1).double()
    signs = 2.0 * lab - 1.0
    errors = 1.0 - x * signs
    es, perm = torch.sort(errors, descending=True)
    gt = lab[perm]
    gts = gt.sum()
    inter = gts - gt.cumsum(0)
    union = gts + (1.0 - gt).cumsum(0)
    jac = 1.0 - inter / union
    jac[1:] = jac[1:] - jac[:-1].clone()
    loss = torch.dot(torch.relu(es), jac)
    grad = torch.zeros_like(x)
    grad[perm] = -signs[perm] * jac * (es > 0)
    return loss, grad


def test_cfg3_sized_segments_against_the_reference_ops_on_the_gpu():
    """33.5 M keys in one segment (hinge over 32x1024x1024), 32 x 1 M (per image) and 4 x 8.4 M (softmax over
    8x4x1024x1024). Checker: the reference's ops (oracle functions, torch.sort) on the same GPU while segments stay below
    2**24 elements; above that the reference's float32 cumsum no longer counts (its value is off by 0.6 % here), so the
    33.5 M-key segment is checked against the same algorithm in float64."""
    g = torch.Generator().manual_seed(0)
    x = (torch.randn(32, 1024, 1024, generator=g) * 3).cuda()
    lab = (torch.rand(32, 1024, 1024, generator=g) < 0.3).long().cuda()
    xa = x.clone().requires_grad_(True)
    la = L.BinaryLovaszLoss(per_image=False)(xa, lab)
    la.backward()
    want, want_grad = _hinge_float64(x, lab)
    np.testing.assert_allclose(float(la), float(want), rtol=1e-5)
    assert float((xa.grad.double().reshape(-1) - want_grad).abs().max()) <= 6e-7
    del xa, want_grad
    xa = x.clone().requires_grad_(True)
    la = L.BinaryLovaszLoss(per_image=True)(xa, lab)
    la.backward()
    xb = x.clone().requires_grad_(True)
    lb = LO.lovasz_hinge(xb, lab, per_image=True)
    lb.backward()
    np.testing.assert_allclose(float(la), float(lb), rtol=1e-5)
    assert float((xa.grad - xb.grad).abs().max()) <= 6e-7
    del x, lab, xa, xb
    pr = torch.softmax(torch.randn(8, 4, 1024, 1024, generator=g) * 3, dim=1).cuda()
    t = torch.randint(0, 4, (8, 1024, 1024), generator=g).cuda()
    pa = pr.clone().requires_grad_(True)
    la = L.LovaszLoss()(pa, t)
    la.backward()
    pb = pr.clone().requires_grad_(True)
    lb = LO.lovasz_softmax(pb, t)
    lb.backward()
    np.testing.assert_allclose(float(la), float(lb), rtol=1e-5)
    assert float((pa.grad - pb.grad).abs().max()) <= 6e-7
