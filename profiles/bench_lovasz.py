"""Lovasz losses at BASELINE-config-3-sized inputs on one B200: fused kernel path (csrc/lovasz.cu) vs the reference's ops
(oracle functions: torch.sort + float cumsums, as CUDA eager). Reports keys/s and the fraction of the ideal 4-pass 32-bit
radix sort bound of SURVEY 8(d): 64 B per key (4 passes x (read + write) x (4 B key + 4 B payload))."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import losses_oracle as LO
from pytorch_toolbelt_b200 import losses as L

dev = torch.device("cuda", 0)
peak = 6570.6
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass

def timed(fn, steps=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / steps

def fwd_bwd(fn, x):
    def run():
        xr = x.detach().requires_grad_(True)
        fn(xr).backward()
    return run

g = torch.Generator().manual_seed(0)
out = {"peak_gbs": peak, "cases": {}}
x = (torch.randn(32, 1024, 1024, generator=g) * 3).to(dev)
lab = (torch.rand(32, 1024, 1024, generator=g) < 0.3).long().to(dev)
pr = torch.softmax(torch.randn(32, 4, 1024, 1024, generator=g) * 3, dim=1).to(dev)
t = torch.randint(0, 4, (32, 1024, 1024), generator=g).to(dev)
cases = {
    "hinge_batch_33.5M_x1": (lambda z: L.BinaryLovaszLoss()(z, lab), lambda z: LO.lovasz_hinge(z, lab), x, 32 * 1024 * 1024),
    "hinge_per_image_1M_x32": (lambda z: L.BinaryLovaszLoss(per_image=True)(z, lab), lambda z: LO.lovasz_hinge(z, lab, per_image=True), x, 32 * 1024 * 1024),
    "softmax_batch_33.5M_x4": (lambda z: L.LovaszLoss()(z, t), lambda z: LO.lovasz_softmax(z, t), pr, 4 * 32 * 1024 * 1024),
    "softmax_per_image_1M_x128": (lambda z: L.LovaszLoss(per_image=True)(z, t), lambda z: LO.lovasz_softmax(z, t, per_image=True), pr, 4 * 32 * 1024 * 1024),
}
for name, (ours, ref, inp, keys) in cases.items():
    ms = timed(fwd_bwd(ours, inp))
    ms_ref = timed(fwd_bwd(ref, inp), steps=2, warm=1)
    out["cases"][name] = {"keys": keys, "ours_ms": round(ms, 3), "torch_cuda_eager_ms": round(ms_ref, 3), "speedup": round(ms_ref / ms, 2),
                          "gkeys_per_s": round(keys / ms / 1e6, 2), "ideal_sort_bytes": keys * 64,
                          "frac_of_4pass_sort_bound": round(keys * 64 / ms / 1e6 / peak, 3)}
print(json.dumps(out))
