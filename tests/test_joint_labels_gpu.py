"""Parity tests for the joint focal+dice kernel (BASELINE config 3 as stated), the logits->labels output kernel and the
uint8 download path of predict_to_host. First run on a B200 in round 2 (10 passed; compute-sanitizer memcheck: 0 errors,
profiles/r2_validation.md), since then part of the default -m gpu suite."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode,shape", [("multilabel", (4, 3, 64, 48)), ("binary", (5, 1, 33, 31)), ("multilabel", (2, 4, 17, 19)),
                                        ("binary", (3, 2, 16, 16))])
def test_joint_focal_region_matches_the_two_separate_losses(mode, shape):
    import pytorch_toolbelt_b200.losses as L
    from oracle import losses_oracle as LO

    g = torch.Generator().manual_seed(3)
    x = (torch.randn(*shape, generator=g) * 3).cuda()
    t = (torch.rand(*shape, generator=g) < 0.3).float().cuda()
    variants = [
        (L.BinaryFocalLoss(), L.DiceLoss(mode), 1.0, 1.0),
        (L.BinaryFocalLoss(alpha=0.25, gamma=1.5, reduction="sum"), L.DiceLoss(mode, log_loss=True, smooth=1.0), 0.7, 1.3),
        (L.BinaryFocalLoss(gamma=2.0, reduced_threshold=0.5), L.JaccardLoss(mode, smooth=0.5), 2.0, 0.25),
    ]
    for focal, region, wf, wr in variants:
        joint = L.BinaryFocalDiceLoss(focal, region, wf, wr)
        assert joint.fusable(x, t)
        x1 = x.clone().requires_grad_(True)
        l1 = joint(x1, t)
        l1.backward()
        x2 = x.clone().requires_grad_(True)
        a, b = focal(x2, t), region(x2, t)
        l2 = wf * a + wr * b
        l2.backward()
        torch.testing.assert_close(l1, l2, rtol=1e-5, atol=1e-7)
        torch.testing.assert_close(joint.last_terms, torch.stack([a.detach(), b.detach()]), rtol=1e-5, atol=1e-7)
        scale = float(x2.grad.abs().max())
        torch.testing.assert_close(x1.grad, x2.grad, rtol=1e-4, atol=1e-6 * scale)
    # against the oracle (the reference's torch ops) for the BASELINE config-3 pair
    xo = x.detach().cpu().clone().requires_grad_(True)
    lo = LO.focal_sigmoid(xo, t.cpu(), alpha=None) + LO.dice_loss(xo, t.cpu(), mode)
    lo.backward()
    xj = x.clone().requires_grad_(True)
    lj = L.BinaryFocalDiceLoss(L.BinaryFocalLoss(), L.DiceLoss(mode))(xj, t)
    lj.backward()
    torch.testing.assert_close(lj.cpu(), lo.detach(), rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(xj.grad.cpu(), xo.grad, rtol=1e-4, atol=1e-6 * float(xo.grad.abs().max()))


def test_joint_falls_back_to_two_kernels_when_not_fusable():
    import pytorch_toolbelt_b200.losses as L

    g = torch.Generator().manual_seed(4)
    x = (torch.randn(2, 4, 32, 32, generator=g) * 3).cuda()
    labels = torch.randint(0, 4, (2, 32, 32), generator=g).cuda()
    joint = L.BinaryFocalDiceLoss(L.BinaryFocalLoss(), L.DiceLoss("multiclass"))
    assert not joint.fusable(x, labels)
    x1 = x.clone().requires_grad_(True)
    joint(x1, labels).backward()
    x2 = x.clone().requires_grad_(True)
    (L.BinaryFocalLoss()(x2, labels) + L.DiceLoss("multiclass")(x2, labels)).backward()
    torch.testing.assert_close(x1.grad, x2.grad)


@pytest.mark.parametrize("shape", [(4, 96, 128), (3, 37, 53), (1, 16, 20), (7, 64, 64)])
def test_logits_to_labels_matches_torch(shape):
    from pytorch_toolbelt_b200.inference.engine import TiledInferenceEngine

    k, h, w = shape
    eng = TiledInferenceEngine((h, w, 3), 16, 16, channels=k, weight="mean")
    g = torch.Generator().manual_seed(9)
    logits = (torch.randn(k, h, w, generator=g) * 3)
    logits[:, 1, 2] = 1.5  # a tie: the first maximum wins
    logits = logits.cuda()
    assert torch.equal(eng.labels(logits), logits.argmax(dim=0).to(torch.uint8))
    want = (logits.sigmoid() > 0.3).to(torch.uint8)
    got = eng.labels(logits, mode="threshold", threshold=0.3)
    boundary = (logits.sigmoid() - 0.3).abs() < 1e-6  # 1-ulp differences of the sigmoid may flip exactly-at-threshold pixels
    assert torch.equal(got[~boundary], want[~boundary])
    # band by band into one buffer == whole image
    out = torch.zeros((h, w), dtype=torch.uint8, device="cuda")
    cuts = [0, h // 3, h // 2, h]
    for a, b in zip(cuts, cuts[1:]):
        eng.labels(logits, out=out, rows=(a, b))
    assert torch.equal(out, logits.argmax(dim=0).to(torch.uint8))


def test_predict_to_host_labels_and_masks():
    import numpy as np

    from pytorch_toolbelt_b200.inference.engine import TiledInferenceEngine

    rng = np.random.default_rng(2)
    image = rng.integers(0, 256, (600, 520, 3)).astype(np.uint8)
    k = 3
    eng = TiledInferenceEngine(image.shape, 128, 64, channels=k, weight="pyramid", tta="flips")
    wm = torch.randn(k, 3, generator=torch.Generator().manual_seed(0)).cuda()

    def model(x):
        return torch.einsum("kc,bchw->bkhw", wm, x) * 0.02 - 1.0

    logits = eng.predict(eng.to_device(image), model, batch_size=8)
    host = torch.from_numpy(image).pin_memory()
    labels = eng.predict_to_host(host, model, output="labels")
    assert labels.dtype == torch.uint8 and tuple(labels.shape) == (600, 520)
    torch.cuda.synchronize()
    assert torch.equal(labels, logits.argmax(dim=0).to(torch.uint8).cpu())
    masks = eng.predict_to_host(host, model, output="mask", threshold=0.4)
    torch.cuda.synchronize()
    want = (logits.sigmoid() > 0.4).to(torch.uint8).cpu()
    boundary = ((logits.sigmoid() - 0.4).abs() < 1e-6).cpu()
    assert torch.equal(masks[~boundary], want[~boundary])
    again = eng.predict_to_host(host, model)  # the default path is untouched by the option
    torch.cuda.synchronize()
    assert torch.equal(again, logits.cpu())
