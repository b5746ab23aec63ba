"""Engine-only timings of the other BASELINE configs on one B200 (cfg2 is bench.py's contract line):
   cfg1 2048^2x3 256/128 K=3 no TTA pyramid; cfg4 10000^2x3 512/256 flips K=8 pyramid; cfg5 20000^2x1 1024/512 K=1 gaussian.
For each: slice+augment kernel, merge kernel (predictions from the stand-in model, resident), algorithmic GB/s vs the
measured HBM peak, and a full-size property check (identity model -> merged output == image)."""
import ctypes as C, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from pytorch_toolbelt_b200 import _native as N
from pytorch_toolbelt_b200.inference.engine import TiledInferenceEngine

dev = torch.device("cuda", 0)
pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
peak = float(json.load(open(pk))["hbm_gbs"]) if os.path.exists(pk) else 6570.6

def timed(fn, steps=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / steps

CFGS = {
    "cfg1": dict(H=2048, W=2048, Cin=3, T=256, S=128, K=3, tta=None, weight="pyramid", analytic=False),
    "cfg4": dict(H=10000, W=10000, Cin=3, T=512, S=256, K=8, tta="flips", weight="pyramid", analytic=False),
    "cfg5": dict(H=20000, W=20000, Cin=1, T=1024, S=512, K=1, tta=None, weight="gaussian", analytic=True),
    "cfg5_table_weight": dict(H=20000, W=20000, Cin=1, T=1024, S=512, K=1, tta=None, weight="gaussian", analytic=False),
}
out = {"peak_gbs": peak}
sel = sys.argv[1:] or list(CFGS)
for name, c in CFGS.items():
    if name not in sel:
        continue
    H, W, Cin, T, S, K = c["H"], c["W"], c["Cin"], c["T"], c["S"], c["K"]
    rng = np.random.default_rng(0)
    image = torch.from_numpy(rng.integers(0, 256, (H, W, Cin)).astype(np.uint8)).to(dev)
    eng = TiledInferenceEngine((H, W, Cin), T, S, channels=K, weight=c["weight"], tta=c["tta"], analytic_weight=c["analytic"])
    A, n = eng.views, eng.num_tiles
    aug = torch.empty((A * n, Cin, T, T), dtype=torch.float32, device=dev)
    pred = torch.empty((A * n, K, T, T), dtype=torch.float32, device=dev)
    res = torch.empty((K, H, W), dtype=torch.float32, device=dev)
    Wm = torch.randn(K, Cin, generator=torch.Generator().manual_seed(0))
    wm = (C.c_float * (K * Cin))(*[float(v) for v in Wm.flatten().tolist()])
    eng.slice(image, out=aug)
    N.check(N.lib().tb_standin_model(aug.data_ptr(), A * n, Cin, K, T, T, wm, pred.data_ptr(), N.stream_ptr(dev)), "m")
    ms_slice = timed(lambda: eng.slice(image, out=aug))
    ms_merge = timed(lambda: eng.merge(pred, out=res))
    b_slice = H * W * Cin + n * A * Cin * T * T * 4
    b_merge = n * A * K * T * T * 4 + K * H * W * 4
    eng_id = TiledInferenceEngine((H, W, Cin), T, S, channels=Cin, weight=c["weight"], tta=c["tta"], analytic_weight=c["analytic"])
    back = eng_id.merge(aug)
    ref = image.permute(2, 0, 1).float()
    ok = bool(torch.allclose(back, ref, rtol=1e-5, atol=1e-3))
    del back, ref
    out[name] = {"tiles": n, "views": A, "canvas": list(eng.slicer.target_shape),
                 "slice_ms": round(ms_slice, 4), "merge_ms": round(ms_merge, 4), "step_ms": round(ms_slice + ms_merge, 4),
                 "MP_per_s": round(H * W / 1e6 / ((ms_slice + ms_merge) * 1e-3), 1),
                 "slice_gbs": round(b_slice / ms_slice / 1e6, 1), "merge_gbs": round(b_merge / ms_merge / 1e6, 1),
                 "slice_frac": round(b_slice / ms_slice / 1e6 / peak, 3), "merge_frac": round(b_merge / ms_merge / 1e6 / peak, 3),
                 "step_frac": round((b_slice + b_merge) / (ms_slice + ms_merge) / 1e6 / peak, 3),
                 "identity_roundtrip_ok": ok}
    del image, aug, pred, res, eng, eng_id
    torch.cuda.empty_cache()
print(json.dumps(out))
