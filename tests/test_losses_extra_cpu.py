"""CPU: oracle/losses_extra_oracle.py (the checker of the GPU tests) against the reference's outputs stored in
tests/golden/losses_extra.npz (written by oracle/make_golden.py::golden_losses_extra from the imported reference)."""
import numpy as np
import torch

from oracle import losses_extra_oracle as LE

ORACLE = {
    "softbce_default": (lambda z, t: LE.soft_bce_with_logits(z, t), "x", "t_ign"),
    "softbce_smooth_sum": (lambda z, t: LE.soft_bce_with_logits(z, t, smooth_factor=0.1, reduction="sum"), "x", "t_ign"),
    "softbce_weights": (lambda z, t: LE.soft_bce_with_logits(z, t, weight=torch.tensor([1.5, 0.7, 1.0]).view(3, 1, 1),
                                                             pos_weight=torch.tensor([0.5, 2.0, 1.5]).view(3, 1, 1), ignore_index=None), "x", "t"),
    "softbce_none": (lambda z, t: LE.soft_bce_with_logits(z, t, reduction="none", pos_weight=torch.tensor([3.0])), "x", "t_ign"),
    "balbce_default": (lambda z, t: LE.balanced_bce_with_logits(z, t), "x", "t"),
    "balbce_gamma_ignore": (lambda z, t: LE.balanced_bce_with_logits(z, t, 0.7, -100), "x", "t_ign"),
    "balbce_sum": (lambda z, t: LE.balanced_bce_with_logits(z, t, 2.0, None, "sum"), "x", "t"),
    "qfl_mean": (lambda z, t: LE.quality_focal_loss(z, t), "x", "t_soft"),
    "qfl_sum_beta1p5": (lambda z, t: LE.quality_focal_loss(z, t, 1.5, "sum"), "x", "t_soft"),
    "qfl_normalized": (lambda z, t: LE.quality_focal_loss(z, t, 2, "normalized"), "x", "t_soft"),
    "qfl_none": (lambda z, t: LE.quality_focal_loss(z, t, 2, "none"), "x", "t_soft"),
    "bsf1": (lambda z, t: LE.binary_soft_f1(z, t), "x", "t"),
    "bsf1_ignore": (lambda z, t: LE.binary_soft_f1(z, t, -100), "x", "t_ign"),
    "sf1": (lambda z, t: LE.soft_f1(z, t), "xc", "tc"),
    "sf1_ignore0": (lambda z, t: LE.soft_f1(z, t, 0), "xc", "tc"),
    "sf1_ignore7": (lambda z, t: LE.soft_f1(z, t, 7), "xc", "tc"),
}


def test_extra_loss_oracle_reproduces_the_reference(golden):
    g = golden("losses_extra")
    for name, (fn, xk, tk) in ORACLE.items():
        x = torch.from_numpy(g[xk]).clone().requires_grad_(True)
        loss = fn(x, torch.from_numpy(g[tk]))
        (loss.sum() if loss.dim() else loss).backward()
        np.testing.assert_allclose(loss.detach().numpy(), g[f"{name}_loss"], rtol=1e-6, atol=1e-8, err_msg=name)
        np.testing.assert_allclose(x.grad.numpy(), g[f"{name}_grad"], rtol=1e-5, atol=1e-9, err_msg=name)
