import ctypes as C, os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pytorch_toolbelt_b200 import _native as N
from pytorch_toolbelt_b200.inference.engine import TiledInferenceEngine
H = W = 5000; Cin = 3; T = 512; S = 256; K = 4
dev = torch.device("cuda", 0)
image = torch.from_numpy(np.random.default_rng(0).integers(0, 256, (H, W, Cin)).astype(np.uint8)).to(dev)
eng = TiledInferenceEngine((H, W, Cin), T, S, channels=K, weight="pyramid", tta="d4")
A, n = eng.views, eng.num_tiles
aug = torch.empty((A * n, Cin, T, T), dtype=torch.float32, device=dev)
pred = torch.empty((A * n, K, T, T), dtype=torch.float32, device=dev)
out = torch.empty((K, H, W), dtype=torch.float32, device=dev)
Wm = torch.randn(K, Cin, generator=torch.Generator().manual_seed(0))
wm = (C.c_float * (K * Cin))(*[float(v) for v in Wm.flatten().tolist()])
eng.slice(image, out=aug)
N.check(N.lib().tb_standin_model(aug.data_ptr(), A * n, Cin, K, T, T, wm, pred.data_ptr(), N.stream_ptr(dev)), "m")
def timed(fn, steps=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / steps
ref_out = eng.merge(pred).clone()
ref_aug_sum = aug.double().sum().item()
for v in ("0", "1", "2", "3"):
    os.environ["TB200_SLICE_VARIANT"] = v
    ms = timed(lambda: eng.slice(image, out=aug))
    print("slice variant", v, round(ms, 4), "ms", "ok" if abs(aug.double().sum().item() - ref_aug_sum) == 0 else "MISMATCH", flush=True)
os.environ["TB200_SLICE_VARIANT"] = "0"
for v in ("0", "2"):
    os.environ["TB200_MERGE_VARIANT"] = v
    ms = timed(lambda: eng.merge(pred, out=out))
    print("merge variant", v, round(ms, 4), "ms", "ok" if torch.equal(out, ref_out) else "MISMATCH", flush=True)
