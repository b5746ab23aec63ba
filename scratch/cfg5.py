import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pytorch_toolbelt_b200.inference.engine import TiledInferenceEngine
dev = torch.device("cuda", 0)
def timed(fn, steps=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / steps
H = W = 20000; T = 1024; S = 512
image = torch.from_numpy(np.random.default_rng(0).integers(0, 256, (H, W, 1)).astype(np.uint8)).to(dev)
eng = TiledInferenceEngine((H, W, 1), T, S, channels=1, weight="gaussian", tta=None)
aug = eng.slice(image)
print("slice_plain ms", round(timed(lambda: eng.slice(image, out=aug)), 3))
out = torch.empty((1, H, W), dtype=torch.float32, device=dev)
ref = eng.merge(aug, out=out).clone()
print("merge tma ms", round(timed(lambda: eng.merge(aug, out=out)), 3), "roundtrip", bool(torch.allclose(out, image.permute(2,0,1).float(), atol=1e-3)))
os.environ["TB200_MERGE_VARIANT"] = "4"
print("merge tma forced ms", round(timed(lambda: eng.merge(aug, out=out)), 3), "same bits", bool(torch.equal(out, ref)))
os.environ["TB200_MERGE_VARIANT"] = "0"
os.environ["TB200_FORCE_GENERIC_MERGE"] = "1"
print("merge generic ms", round(timed(lambda: eng.merge(aug, out=out)), 3), "same bits", bool(torch.equal(out, ref)))
