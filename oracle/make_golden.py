"""Generate tests/golden/*.npz from the REAL reference (run in the build container only).

    python oracle/make_golden.py            # needs /root/reference (read-only) on PYTHONPATH-able disk

Every vector stored here is an output of the unmodified reference (BloodAxe/pytorch-toolbelt @ 05e07ac) on
seeded inputs; while generating, the oracle restatement is checked against the same outputs (assertions
below), which is what pins the oracle. The fixtures travel to the GPU box; the reference does not.
TEST INFRASTRUCTURE ONLY.
"""
import os
import sys
import warnings

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, "/root/reference")
warnings.simplefilter("ignore")

import cv2  # noqa: E402
from pytorch_toolbelt.inference import functional as RF  # noqa: E402
from pytorch_toolbelt.inference import tiles as RT  # noqa: E402
from pytorch_toolbelt.inference import tta as RA  # noqa: E402
from pytorch_toolbelt import losses as RL  # noqa: E402
from pytorch_toolbelt.losses import functional as RLF  # noqa: E402

from oracle import losses_extra_oracle as LE  # noqa: E402
from oracle import losses_oracle as LO  # noqa: E402
from oracle import tiles_oracle as TO  # noqa: E402
from oracle import tta_oracle as AO  # noqa: E402

OUT = os.path.join(REPO, "tests", "golden")
os.makedirs(OUT, exist_ok=True)


def save(name, **arrays):
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **arrays)
    print(f"{name}.npz: {len(arrays)} arrays, {os.path.getsize(path) / 1024:.1f} KiB")


# ------------------------------------------------------------------------------------------------ tiles
GEOMETRY_CASES = [
    # (H, W, tile, step, margin)
    (2048, 2048, 256, 128, 0),
    (5000, 5000, 512, 256, 0),
    (10000, 10000, 512, 256, 0),
    (20000, 20000, 1024, 512, 0),
    (500, 500, 51, 26, 0),
    (563, 512, (128, 128), (128, 128), 0),
    (5632, 5120, (1280, 1280), (1280, 1280), 0),
    (40, 50, 64, 64, 0),
    (61, 53, (24, 16), (12, 8), 0),
    (97, 83, (32, 24), (16, 12), 0),
    (97, 83, (32, 24), (16, 12), 5),
    (97, 83, (32, 24), (20, 24), (3, 9, 0, 7)),
    (100, 100, 33, 17, 0),
]

BORDERS = {
    "constant": cv2.BORDER_CONSTANT,
    "replicate": cv2.BORDER_REPLICATE,
    "reflect": cv2.BORDER_REFLECT,
    "wrap": cv2.BORDER_WRAP,
    "reflect101": cv2.BORDER_REFLECT_101,
}
assert [BORDERS[k] for k in ("constant", "replicate", "reflect", "wrap", "reflect101")] == [0, 1, 2, 3, 4]


def golden_tiles():
    out = {}
    for idx, (H, W, tile, step, margin) in enumerate(GEOMETRY_CASES):
        ref = RT.ImageSlicer((H, W, 3), tile, step, image_margin=margin, weight="mean")
        t2 = (tile, tile) if np.isscalar(tile) else tile
        s2 = (step, step) if np.isscalar(step) else step
        g = TO.tile_grid(H, W, t2, s2, margin)
        assert np.array_equal(g["crops"], ref.crops) and np.array_equal(g["bbox_crops"], ref.bbox_crops)
        assert g["target_shape"] == ref.target_shape
        assert g["margins"] == (ref.margin_left, ref.margin_right, ref.margin_top, ref.margin_bottom)
        out[f"geom{idx}_args"] = np.array([H, W, t2[0], t2[1], s2[0], s2[1]])
        out[f"geom{idx}_margin_arg"] = np.array(margin if isinstance(margin, tuple) else [margin])
        out[f"geom{idx}_crops"] = ref.crops
        out[f"geom{idx}_bbox"] = ref.bbox_crops
        out[f"geom{idx}_margins"] = np.array(g["margins"])
        out[f"geom{idx}_target"] = np.array(ref.target_shape)
    out["geom_count"] = np.array(len(GEOMETRY_CASES))

    # pyramid weights
    for (th, tw) in [(32, 24), (24, 16), (16, 16), (51, 51)]:
        w, _, _ = RT.compute_pyramid_patch_weight_loss(th, tw)
        assert np.array_equal(w, TO.pyramid_weight(th, tw)) or np.allclose(w, TO.pyramid_weight(th, tw), rtol=1e-15, atol=0)
        out[f"pyramid_{th}x{tw}"] = w
    for t in (256, 512, 1024):
        w, dc, de = RT.compute_pyramid_patch_weight_loss(t, t)
        out[f"pyramid_alpha_{t}"] = np.array((t * t) / np.sum(de / (dc + de)))
        out[f"pyramid_stats_{t}"] = np.array([w.min(), w.max(), w.mean(), w[3, 7], w[t // 2, t // 3]])
    save("tiles_geometry", **out)

    # pixel-level split / merge vectors
    rng = np.random.default_rng(0)
    out = {}
    H, W, tile, step = 61, 53, (24, 16), (12, 8)
    for C, dtype, tag in [(3, np.uint8, "u8c3"), (1, np.uint8, "u8c1"), (None, np.uint8, "u8hw"), (3, np.float32, "f32c3"),
                          (4, np.uint16, "u16c4")]:
        shape = (H, W) if C is None else (H, W, C)
        if np.issubdtype(dtype, np.integer):
            image = rng.integers(0, np.iinfo(dtype).max + 1, shape).astype(dtype)
        else:
            image = rng.standard_normal(shape).astype(dtype)
        out[f"{tag}_image"] = image
        ref = RT.ImageSlicer(image.shape, tile, step, weight="pyramid")
        grid = TO.tile_grid(H, W, tile, step)
        for bname, btype in BORDERS.items():
            for vname, value in (("s7", 7), ("t", (3, 5, 250, 9))):
                if bname != "constant" and vname == "t":
                    continue
                if tag != "u8c3" and bname not in ("constant", "reflect101"):
                    continue
                tiles_ref = np.stack(ref.split(image, border_type=btype, value=value))
                tiles_orc = np.stack(TO.split(image, grid, btype, value, clip_to_tile=False))
                assert np.array_equal(tiles_ref, tiles_orc), (tag, bname, vname, "split")
                out[f"{tag}_split_{bname}_{vname}"] = tiles_ref
                it_ref = np.stack([t for t, _ in ref.iter_split(image, border_type=btype, value=value)])
                it_orc = np.stack(TO.split(image, grid, btype, value, clip_to_tile=True))
                assert np.array_equal(it_ref, it_orc), (tag, bname, vname, "iter_split")
                if not np.array_equal(it_ref, tiles_ref):
                    out[f"{tag}_itersplit_{bname}_{vname}"] = it_ref
        # cut_patch on a corner tile
        cp = ref.cut_patch(image, len(ref.crops) - 1, border_type=cv2.BORDER_REFLECT_101)
        assert np.array_equal(cp, TO.crop_with_border(image, *ref.bbox_crops[-1], cv2.BORDER_REFLECT_101, 0, True))
        out[f"{tag}_cutpatch_last_reflect101"] = cp
    # tile larger than the image, wrap border differs between split and iter_split
    image = rng.integers(0, 256, (40, 50, 3)).astype(np.uint8)
    ref = RT.ImageSlicer(image.shape, 64, 64, weight="mean")
    grid = TO.tile_grid(40, 50, (64, 64), (64, 64))
    out["big_image"] = image
    for bname, btype in BORDERS.items():
        a = np.stack(ref.split(image, border_type=btype, value=11))
        b = np.stack([t for t, _ in ref.iter_split(image, border_type=btype, value=11)])
        assert np.array_equal(a, np.stack(TO.split(image, grid, btype, 11, False))), bname
        assert np.array_equal(b, np.stack(TO.split(image, grid, btype, 11, True))), bname
        out[f"big_split_{bname}"] = a
        out[f"big_itersplit_{bname}"] = b
    save("tiles_split", **out)

    # merges
    out = {}
    H, W, tile, step, K = 61, 53, (24, 16), (12, 8), 2
    for wname in ("mean", "pyramid"):
        ref = RT.ImageSlicer((H, W, 3), tile, step, weight=wname)
        grid = TO.tile_grid(H, W, tile, step)
        n = len(ref.crops)
        preds = (torch.randn(n, K, tile[0], tile[1], generator=torch.Generator().manual_seed(1)) * 3)
        merger = RT.TileMerger(ref.target_shape, K, ref.weight, device="cpu")
        for s in range(0, n, 5):
            merger.integrate_batch(preds[s : s + 5], ref.crops[s : s + 5])
        merged = merger.merge()
        om = TO.TorchMerger(grid["target_shape"], K, ref.weight)
        om.integrate_batch(preds, grid["crops"])
        assert torch.equal(om.merge(), merged), wname
        out[f"{wname}_preds"] = preds.numpy()
        out[f"{wname}_merger_image"] = merger.image.numpy()
        out[f"{wname}_merger_norm"] = merger.norm_mask.numpy()
        out[f"{wname}_merged"] = merged.numpy()
        # float64 NumPy path
        tiles_hw = [np.moveaxis(p.numpy(), 0, -1) for p in preds]
        m64 = ref.merge(tiles_hw, dtype=np.float32)
        o64 = TO.merge_numpy_f64(tiles_hw, grid, ref.weight, H, W, np.float32)
        assert np.array_equal(m64, o64), wname
        out[f"{wname}_merge_f64"] = m64
    # split -> merge round trip on a non-trivial uint8 image with mean weights (exact)
    image = np.random.default_rng(3).integers(0, 256, (H, W, 3)).astype(np.uint8)
    ref = RT.ImageSlicer(image.shape, tile, step, weight="mean")
    out["roundtrip_image"] = image
    out["roundtrip_merged_u8"] = ref.merge(ref.split(image), dtype=np.uint8)
    save("tiles_merge", **out)


# ------------------------------------------------------------------------------------------------ tta
REF_AUG = {
    "fliplr": (RA.fliplr_image_augment, RA.fliplr_image_deaugment),
    "flipud": (RA.flipud_image_augment, RA.flipud_image_deaugment),
    "flips": (RA.flips_image_augment, RA.flips_image_deaugment),
    "d2": (RA.d2_image_augment, RA.d2_image_deaugment),
    "d4": (RA.d4_image_augment, RA.d4_image_deaugment),
}
REDUCTIONS = ["mean", "sum", "gmean", "hmean", "harmonic1p", "logodd", "log1p", None]
PRIMITIVES = [
    "torch_rot90_ccw", "torch_rot90_cw", "torch_rot90_ccw_transpose", "torch_rot90_cw_transpose",
    "torch_transpose_rot90_ccw", "torch_transpose_rot90_cw", "torch_rot180", "torch_rot180_transpose",
    "torch_transpose_rot180", "torch_flipud", "torch_fliplr", "torch_transpose", "torch_transpose2",
]


def golden_tta():
    out = {}
    g = torch.Generator().manual_seed(0)
    xs = {"sq": torch.rand(2, 3, 8, 8, generator=g), "rect": torch.rand(2, 3, 6, 10, generator=g),
          "odd": torch.rand(1, 2, 7, 7, generator=g)}
    for xname, x in xs.items():
        out[f"x_{xname}"] = x.numpy()
        for group, (aug, deaug) in REF_AUG.items():
            if group == "d4" and x.size(2) != x.size(3):
                continue
            a = aug(x)
            assert torch.equal(a, AO.augment(x, group)), (xname, group)
            for v, view in enumerate(AO.GROUPS[group]):
                assert torch.equal(AO.apply_view(x, view), AO.gather_reference(x, view))
            out[f"{xname}_{group}_aug"] = a.numpy()
            A = len(AO.GROUPS[group])
            y = torch.rand(A * x.size(0), 4, x.size(2), x.size(3), generator=g) * 0.98 + 0.01
            out[f"{xname}_{group}_y"] = y.numpy()
            for red in REDUCTIONS:
                r = deaug(y, reduction=red)
                o = AO.deaugment(y, group, red)
                assert torch.equal(r, o), (xname, group, red)
                out[f"{xname}_{group}_deaug_{red}"] = r.numpy()
    x = xs["sq"]
    for name in PRIMITIVES:
        out[f"prim_{name}"] = getattr(RF, name)(x).contiguous().numpy()
    xr = xs["rect"]
    for name in ("torch_flipud", "torch_fliplr", "torch_rot180", "torch_transpose", "torch_rot90_cw", "torch_rot90_ccw"):
        out[f"primrect_{name}"] = getattr(RF, name)(xr).contiguous().numpy()
    # known-answer label tests of the reference (tests/test_tta.py:71-108)
    x44 = torch.tensor([[1, 2, 3, 4], [5, 6, 7, 8], [9, 0, 1, 2], [3, 4, 5, 6]]).unsqueeze(0).unsqueeze(0).float()

    class SumAll(torch.nn.Module):
        def forward(self, input):
            return input.sum(dim=[1, 2, 3])

    out["label_d4"] = RA.d4_image2label(SumAll(), x44).numpy()
    out["label_fliplr"] = RA.fliplr_image2label(SumAll(), x44).numpy()
    out["label_fivecrop"] = RA.fivecrop_image2label(SumAll(), x44, (2, 2)).numpy()
    out["label_tencrop"] = RA.tencrop_image2label(SumAll(), x44, (2, 2)).numpy()
    lg = torch.rand(16, 5, generator=g)
    out["labels_logits"] = lg.numpy()
    out["labels_d4_deaug"] = RA.d4_labels_deaugment(lg).numpy()
    out["labels_d2_deaug"] = RA.d2_labels_deaugment(lg).numpy()
    out["labels_fliplr_deaug"] = RA.fliplr_labels_deaugment(lg).numpy()
    save("tta", **out)


# ------------------------------------------------------------------------------------------------ losses
def _grad(fn, x):
    x = x.clone().requires_grad_(True)
    loss = fn(x)
    (loss.sum() if loss.dim() else loss).backward()
    return loss.detach(), x.grad.detach()


def golden_losses():
    out = {}
    g = torch.Generator().manual_seed(0)
    B, C, H, W = 2, 3, 16, 16
    x = torch.randn(B, C, H, W, generator=g) * 3
    t_ml = (torch.rand(B, C, H, W, generator=g) < 0.3).float()
    t_mc = torch.randint(0, C, (B, H, W), generator=g)
    t_mc_ign = t_mc.clone()
    t_mc_ign[torch.rand(B, H, W, generator=g) < 0.05] = -100
    t_ml_ign = t_ml.clone()
    t_ml_ign[torch.rand(B, C, H, W, generator=g) < 0.05] = 255
    cw = torch.tensor([0.5, 2.0, 1.25])
    out.update(x=x.numpy(), t_ml=t_ml.numpy(), t_mc=t_mc.numpy(), t_mc_ign=t_mc_ign.numpy(), t_ml_ign=t_ml_ign.numpy(),
               cw=cw.numpy())

    def record(name, ref_fn, orc_fn, xin=x, rtol=1e-6):
        lr, gr = _grad(ref_fn, xin)
        lo, go = _grad(orc_fn, xin)
        assert torch.allclose(lr, lo, rtol=rtol, atol=1e-7), (name, lr, lo)
        assert torch.allclose(gr, go, rtol=1e-5, atol=1e-8), (name, (gr - go).abs().max())
        out[f"{name}_loss"] = lr.numpy()
        out[f"{name}_grad"] = gr.numpy()

    # ---- sigmoid focal (functional + module) ----
    focal_cfgs = {
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
    for name, kw in focal_cfgs.items():
        kwr = dict(alpha=None)
        kwr.update(kw)
        record(f"focal_{name}", lambda z, k=kwr: RLF.focal_loss_with_logits(z, t_ml, **k),
               lambda z, k=kwr: LO.focal_sigmoid(z, t_ml, **k))
    kwr = dict(alpha=0.25, ignore_index=255)
    record("focal_ignore", lambda z: RLF.focal_loss_with_logits(z, t_ml_ign, **kwr),
           lambda z: LO.focal_sigmoid(z, t_ml_ign, **kwr))
    kwr = dict(alpha=None, ignore_index=255, normalized=True)
    record("focal_ignore_norm", lambda z: RLF.focal_loss_with_logits(z, t_ml_ign, **kwr),
           lambda z: LO.focal_sigmoid(z, t_ml_ign, **kwr))
    # module with long targets (one-hot expansion, focal.py:88-89) incl. ignore
    record("binaryfocal_labels", lambda z: RL.BinaryFocalLoss()(z, t_mc),
           lambda z: LO.focal_sigmoid(z, LO.one_hot_targets(t_mc, C), alpha=None))
    record("binaryfocal_labels_ignore", lambda z: RL.BinaryFocalLoss(ignore_index=-100, alpha=0.3)(z, t_mc_ign),
           lambda z: LO.focal_sigmoid(z, LO.one_hot_targets(t_mc_ign, C, -100), alpha=0.3, ignore_index=-100))

    # ---- softmax focal ----
    sm_cfgs = {
        "default": dict(),
        "gamma1_sum": dict(gamma=1.0, reduction="sum"),
        "none": dict(reduction="none"),
        "batchwise": dict(reduction="batchwise_mean"),
        "normalized": dict(normalized=True),
        "reduced": dict(reduced_threshold=0.5),
        "classw": dict(class_weights=cw),
    }
    for name, kw in sm_cfgs.items():
        record(f"smfocal_{name}", lambda z, k=kw: RLF.softmax_focal_loss_with_logits(z, t_mc_ign, **k),
               lambda z, k=kw: LO.focal_softmax(z, t_mc_ign, **k))
    record("cefocal_module", lambda z: RL.CrossEntropyFocalLoss()(z, t_mc), lambda z: LO.focal_softmax(z, t_mc))

    # ---- dice / jaccard ----
    xb = x[:, :1].contiguous()
    tb = t_ml[:, :1].contiguous()
    out["xb"] = xb.numpy()
    out["tb"] = tb.numpy()
    probs = torch.sigmoid(x)
    out["probs"] = probs.numpy()
    dice_cfgs = {
        "ml": ("multilabel", t_ml, dict()),
        "ml_log_smooth": ("multilabel", t_ml, dict(log_loss=True, smooth=1.0)),
        "ml_ignore": ("multilabel", t_ml_ign, dict(ignore_index=255)),
        "mc": ("multiclass", t_mc, dict()),
        "mc_smooth": ("multiclass", t_mc, dict(smooth=0.5)),
        "mc_ignore": ("multiclass", t_mc_ign, dict(ignore_index=-100)),
        "mc_log": ("multiclass", t_mc, dict(log_loss=True)),
    }
    for name, (mode, tt, kw) in dice_cfgs.items():
        record(f"dice_{name}", lambda z, m=mode, t_=tt, k=kw: RL.DiceLoss(m, **k)(z, t_),
               lambda z, m=mode, t_=tt, k=kw: LO.dice_loss(z, t_, m, **k))
    record("dice_bin", lambda z: RL.DiceLoss("binary")(z, tb), lambda z: LO.dice_loss(z, tb, "binary"), xin=xb)
    record("dice_bin_ignore", lambda z: RL.DiceLoss("binary", ignore_index=255)(z, t_ml_ign[:, :1]),
           lambda z: LO.dice_loss(z, t_ml_ign[:, :1], "binary", ignore_index=255), xin=xb)
    record("dice_ml_probs", lambda z: RL.DiceLoss("multilabel", from_logits=False)(z, t_ml),
           lambda z: LO.dice_loss(z, t_ml, "multilabel", from_logits=False), xin=probs)
    jac_cfgs = {
        "ml": ("multilabel", t_ml, dict()),
        "ml_log_smooth": ("multilabel", t_ml, dict(log_loss=True, smooth=1.0)),
        "mc": ("multiclass", t_mc, dict()),
        "mc_smooth": ("multiclass", t_mc, dict(smooth=0.5)),
    }
    for name, (mode, tt, kw) in jac_cfgs.items():
        record(f"jaccard_{name}", lambda z, m=mode, t_=tt, k=kw: RL.JaccardLoss(m, **k)(z, t_),
               lambda z, m=mode, t_=tt, k=kw: LO.jaccard_loss(z, t_, m, **k))
    record("jaccard_bin", lambda z: RL.JaccardLoss("binary")(z, tb), lambda z: LO.jaccard_loss(z, tb, "binary"), xin=xb)
    # empty class: zero targets in channel 2
    t_empty = t_ml.clone()
    t_empty[:, 2] = 0
    out["t_empty"] = t_empty.numpy()
    record("dice_ml_emptyclass", lambda z: RL.DiceLoss("multilabel")(z, t_empty),
           lambda z: LO.dice_loss(z, t_empty, "multilabel"))
    # functional scores (known answers of tests/test_losses.py:37-80 are re-tested directly in tests/)
    out["score_dice"] = RLF.soft_dice_score(probs, t_ml, smooth=0.0, eps=1e-7, dims=(0, 2, 3)).numpy()
    out["score_jaccard"] = RLF.soft_jaccard_score(probs, t_ml, smooth=0.0, eps=1e-7, dims=(0, 2, 3)).numpy()
    assert torch.equal(LO.soft_dice_score(probs, t_ml, dims=(0, 2, 3)), torch.from_numpy(out["score_dice"]))

    # ---- wing ----
    tw_ = torch.randn(B, C, H, W, generator=g) * 4
    out["t_wing"] = tw_.numpy()
    for name, kw in {"mean": dict(), "sum": dict(reduction="sum"), "none": dict(reduction="none"),
                     "w10c2": dict(width=10, curvature=2.0)}.items():
        record(f"wing_{name}", lambda z, k=kw: RLF.wing_loss(z, tw_, **k), lambda z, k=kw: LO.wing_loss(z, tw_, **k))

    # ---- lovasz ----
    lb = (torch.rand(B, H, W, generator=g) < 0.4).long()
    lb_ign = lb.clone()
    lb_ign[torch.rand(B, H, W, generator=g) < 0.1] = 255
    xl = x[:, 0].contiguous()
    out.update(xl=xl.numpy(), lb=lb.numpy(), lb_ign=lb_ign.numpy())
    record("lovasz_hinge", lambda z: RL.BinaryLovaszLoss()(z, lb), lambda z: LO.lovasz_hinge(z, lb), xin=xl)
    record("lovasz_hinge_perimage", lambda z: RL.BinaryLovaszLoss(per_image=True)(z, lb),
           lambda z: LO.lovasz_hinge(z, lb, per_image=True), xin=xl)
    record("lovasz_hinge_ignore", lambda z: RL.BinaryLovaszLoss(ignore_index=255)(z, lb_ign),
           lambda z: LO.lovasz_hinge(z, lb_ign, ignore_index=255), xin=xl)
    pr = torch.softmax(x, dim=1)
    out["pr"] = pr.numpy()
    t_mc_ign255 = t_mc.clone()
    t_mc_ign255[t_mc_ign == -100] = 255
    out["t_mc_ign255"] = t_mc_ign255.numpy()
    record("lovasz_softmax", lambda z: RL.LovaszLoss()(z, t_mc), lambda z: LO.lovasz_softmax(z, t_mc), xin=pr)
    record("lovasz_softmax_perimage", lambda z: RL.LovaszLoss(per_image=True)(z, t_mc),
           lambda z: LO.lovasz_softmax(z, t_mc, per_image=True), xin=pr)
    record("lovasz_softmax_ignore", lambda z: RL.LovaszLoss(ignore=255)(z, t_mc_ign255),
           lambda z: LO.lovasz_softmax(z, t_mc_ign255, ignore=255), xin=pr)
    save("losses", **out)


EXTRA_CASES = {
    # name: (reference module factory, oracle function, input key, target key)
    "softbce_default": (lambda: RL.SoftBCEWithLogitsLoss(), lambda z, t: LE.soft_bce_with_logits(z, t), "x", "t_ign"),
    "softbce_smooth_sum": (lambda: RL.SoftBCEWithLogitsLoss(smooth_factor=0.1, reduction="sum"),
                           lambda z, t: LE.soft_bce_with_logits(z, t, smooth_factor=0.1, reduction="sum"), "x", "t_ign"),
    "softbce_weights": (lambda: RL.SoftBCEWithLogitsLoss(weight=torch.tensor([1.5, 0.7, 1.0]).view(3, 1, 1),
                                                         pos_weight=torch.tensor([0.5, 2.0, 1.5]).view(3, 1, 1), ignore_index=None),
                        lambda z, t: LE.soft_bce_with_logits(z, t, weight=torch.tensor([1.5, 0.7, 1.0]).view(3, 1, 1),
                                                             pos_weight=torch.tensor([0.5, 2.0, 1.5]).view(3, 1, 1), ignore_index=None),
                        "x", "t"),
    "softbce_none": (lambda: RL.SoftBCEWithLogitsLoss(reduction="none", pos_weight=torch.tensor([3.0])),
                     lambda z, t: LE.soft_bce_with_logits(z, t, reduction="none", pos_weight=torch.tensor([3.0])), "x", "t_ign"),
    "balbce_default": (lambda: RL.BalancedBCEWithLogitsLoss(), lambda z, t: LE.balanced_bce_with_logits(z, t), "x", "t"),
    "balbce_gamma_ignore": (lambda: RL.BalancedBCEWithLogitsLoss(gamma=0.7, ignore_index=-100),
                            lambda z, t: LE.balanced_bce_with_logits(z, t, 0.7, -100), "x", "t_ign"),
    "balbce_sum": (lambda: RL.BalancedBCEWithLogitsLoss(gamma=2.0, reduction="sum"),
                   lambda z, t: LE.balanced_bce_with_logits(z, t, 2.0, None, "sum"), "x", "t"),
    "qfl_mean": (lambda: RQ.QualityFocalLoss(), lambda z, t: LE.quality_focal_loss(z, t), "x", "t_soft"),
    "qfl_sum_beta1p5": (lambda: RQ.QualityFocalLoss(beta=1.5, reduction="sum"), lambda z, t: LE.quality_focal_loss(z, t, 1.5, "sum"),
                        "x", "t_soft"),
    "qfl_normalized": (lambda: RQ.QualityFocalLoss(reduction="normalized"), lambda z, t: LE.quality_focal_loss(z, t, 2, "normalized"),
                       "x", "t_soft"),
    "qfl_none": (lambda: RQ.QualityFocalLoss(reduction="none"), lambda z, t: LE.quality_focal_loss(z, t, 2, "none"), "x", "t_soft"),
    "bsf1": (lambda: RL.BinarySoftF1Loss(), lambda z, t: LE.binary_soft_f1(z, t), "x", "t"),
    "bsf1_ignore": (lambda: RL.BinarySoftF1Loss(ignore_index=-100), lambda z, t: LE.binary_soft_f1(z, t, -100), "x", "t_ign"),
    "sf1": (lambda: RL.SoftF1Loss(), lambda z, t: LE.soft_f1(z, t), "xc", "tc"),
    "sf1_ignore0": (lambda: RL.SoftF1Loss(ignore_index=0), lambda z, t: LE.soft_f1(z, t, 0), "xc", "tc"),
    "sf1_ignore7": (lambda: RL.SoftF1Loss(ignore_index=7), lambda z, t: LE.soft_f1(z, t, 7), "xc", "tc"),
}


from oracle.make_golden_inputs import VOLUME_CASES, volume_case_inputs  # noqa: E402


def golden_tiles3d():
    """3-D tiles: VolumeSlicer grid / split and VolumeMerger integrate+merge outputs of the reference; pins tiles3d_oracle."""
    from pytorch_toolbelt.inference.tiles_3d import VolumeMerger, VolumeSlicer

    from oracle import tiles3d_oracle as T3

    out = {}
    for ci, (shape, tile, step) in enumerate(VOLUME_CASES):
        ref = VolumeSlicer(shape, tile, step)
        grid = T3.volume_grid(shape, tile, step)
        origins = np.array([[r.start for r in roi] for roi in ref.crops])
        assert np.array_equal(origins, grid["origins"]) and np.array_equal(ref.target_shape, grid["target_shape"])
        assert np.array_equal(ref.pad_before, grid["pad_before"]) and np.array_equal(ref.num_tiles, grid["num_tiles"])
        d, h, w = (int(v) for v in ref.tile_size)
        vol, weight, preds = volume_case_inputs(ci, shape, len(ref.crops), (d, h, w))
        tiles_ref = ref.split(vol, value=7)
        tiles_orc = T3.split(vol, grid, value=7)
        assert all(np.array_equal(a, b) for a, b in zip(tiles_ref, tiles_orc))
        preds_t = torch.from_numpy(preds)
        merger = VolumeMerger(ref.target_shape, 2, weight)
        for s0 in range(0, len(ref.crops), 3):
            merger.integrate_batch(preds_t[s0 : s0 + 3], ref.crops[s0 : s0 + 3])
        merged = merger.merge()[ref.orignal_mask_roi]
        assert torch.equal(merged, T3.merge(preds_t, grid, weight, 2))
        pick = sorted({0, len(tiles_ref) // 2, len(tiles_ref) - 1})
        out.update({f"c{ci}_origins": origins, f"c{ci}_target": np.array(ref.target_shape), f"c{ci}_pad_before": ref.pad_before,
                    f"c{ci}_tile_idx": np.array(pick), f"c{ci}_tiles": np.stack([tiles_ref[t] for t in pick]),
                    f"c{ci}_tile_sums": np.array([float(np.asarray(t, dtype=np.float64).sum()) for t in tiles_ref]),
                    f"c{ci}_merged": merged.numpy()})
    save("tiles3d", **out)


def golden_losses_extra():
    """SURVEY 8f rank 4 losses: reference outputs + gradients on seeded inputs; pins oracle/losses_extra_oracle.py."""
    global RQ
    from pytorch_toolbelt.losses import quality_focal_loss as RQ  # noqa: F811

    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, 3, 16, 16, generator=g) * 3
    t = (torch.rand(2, 3, 16, 16, generator=g) < 0.3).float()
    t_ign = t.clone()
    t_ign[torch.rand(2, 3, 16, 16, generator=g) < 0.08] = -100
    t_soft = t * torch.rand(2, 3, 16, 16, generator=g)  # IoU-like quality targets in [0,1]
    xc = torch.randn(41, 5, generator=g) * 2
    tc = torch.randint(0, 5, (41,), generator=g)
    data = dict(x=x, t=t, t_ign=t_ign, t_soft=t_soft, xc=xc, tc=tc)
    out = {k: v.numpy() for k, v in data.items()}
    for name, (ref_factory, orc, xk, tk) in EXTRA_CASES.items():
        ref = ref_factory()
        lr, gr = _grad(lambda z: ref(z, data[tk]), data[xk])
        lo, go = _grad(lambda z: orc(z, data[tk]), data[xk])
        assert torch.allclose(lr, lo, rtol=1e-6, atol=1e-8), (name, lr, lo)
        assert torch.allclose(gr, go, rtol=1e-5, atol=1e-9), (name, (gr - go).abs().max())
        out[f"{name}_loss"] = lr.numpy()
        out[f"{name}_grad"] = gr.numpy()
    save("losses_extra", **out)


if __name__ == "__main__":
    golden_tiles()
    golden_tta()
    golden_losses()
    golden_losses_extra()
    golden_tiles3d()
    print("golden vectors written to", OUT)
