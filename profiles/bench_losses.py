"""BASELINE config 3: BinaryFocalLoss + DiceLoss forward+backward on 32x4x1024x1024 logits, one B200.
Prints one JSON object: our fused kernels vs the reference's torch ops run as CUDA eager (oracle on CUDA tensors),
with achieved algorithmic GB/s against the measured HBM peak. Not part of bench.py's contract line (cfg 2 is)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import losses_oracle as LO
from pytorch_toolbelt_b200 import losses as L

B, C, H, W = 32, 4, 1024, 1024
dev = torch.device("cuda", 0)
g = torch.Generator(device="cpu").manual_seed(0)
x = (torch.randn(B, C, H, W, generator=g) * 3).to(dev)
t = (torch.rand(B, C, H, W, generator=g) < 0.3).float().to(dev)
lab = torch.randint(0, C, (B, H, W), generator=g).to(dev)
N = x.numel()
peak = 6570.6
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass

def timed(fn, steps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / steps

def fwd_bwd(loss_fn):
    def run():
        xr = x.detach().requires_grad_(True)
        loss_fn(xr).backward()
        return xr.grad
    return run

focal, dice, dice_mc, jac = L.BinaryFocalLoss(), L.DiceLoss("multilabel"), L.DiceLoss("multiclass"), L.JaccardLoss("multilabel")
joint = L.BinaryFocalDiceLoss(focal, dice)
cases = {
    "focal": (lambda z: focal(z, t), lambda z: LO.focal_sigmoid(z, t, alpha=None), 3 * N * 4),
    "focal_labels": (lambda z: focal(z, lab), lambda z: LO.focal_sigmoid(z, LO.one_hot_targets(lab, C), alpha=None), 2 * N * 4 + C * lab.numel() * 8),
    "dice_multilabel": (lambda z: dice(z, t), lambda z: LO.dice_loss(z, t, "multilabel"), 5 * N * 4),
    "jaccard_multilabel": (lambda z: jac(z, t), lambda z: LO.jaccard_loss(z, t, "multilabel"), 5 * N * 4),
    "dice_multiclass": (lambda z: dice_mc(z, lab), lambda z: LO.dice_loss(z, lab, "multiclass"), 3 * N * 4 + 2 * lab.numel() * 8),
    "focal_plus_dice": (lambda z: focal(z, t) + dice(z, t), lambda z: LO.focal_sigmoid(z, t, alpha=None) + LO.dice_loss(z, t, "multilabel"), 8 * N * 4 + 3 * N * 4),
    # the same pair through ONE fused two-pass kernel (tb_focal_dice_fwd_bwd; 5N*4 algorithmic bytes)
    "focal_plus_dice_fused": (lambda z: joint(z, t), lambda z: LO.focal_sigmoid(z, t, alpha=None) + LO.dice_loss(z, t, "multilabel"), 5 * N * 4),
    "cross_entropy_focal": (lambda z: L.CrossEntropyFocalLoss()(z, lab), lambda z: LO.focal_softmax(z, lab), 2 * N * 4 + lab.numel() * 8),
    "wing": (lambda z: L.WingLoss()(z, t), lambda z: LO.wing_loss(z, t), 3 * N * 4),
}
out = {"config": "cfg3: 32x4x1024x1024 f32 logits, fwd+bwd to x.grad", "peak_gbs": peak, "cases": {}}
only = set(sys.argv[1:])  # optional case filter
for name, (ours, ref, alg_bytes) in cases.items():
    if only and name not in only:
        continue
    try:
        ms_ours = timed(fwd_bwd(ours))
    except Exception as e:  # an experimental kernel must not take the other measurements down with it
        out["cases"][name] = {"error": repr(e)[:300]}
        continue
    ms_ref = timed(fwd_bwd(ref), steps=3, warm=1)
    lo = ours(x.detach()).item(); lr = ref(x.detach()).item()
    out["cases"][name] = {"ours_ms": round(ms_ours, 4), "torch_cuda_eager_ms": round(ms_ref, 4), "speedup": round(ms_ref / ms_ours, 2),
                          "algorithmic_bytes": alg_bytes, "achieved_gbs": round(alg_bytes / ms_ours / 1e6, 1),
                          "frac_of_measured_peak": round(alg_bytes / ms_ours / 1e6 / peak, 3),
                          "loss_ours": lo, "loss_ref": lr, "gelem_per_s": round(N / ms_ours / 1e6, 1)}
print(json.dumps(out))
