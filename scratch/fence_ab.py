import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pytorch_toolbelt_b200.inference.engine import TiledInferenceEngine
dev = torch.device("cuda", 0)
def timed(fn, steps=8, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / steps
for name, (H, T, S, K, tta) in {"cfg2": (5000, 512, 256, 4, "d4"), "cfg4": (10000, 512, 256, 8, "flips"), "d2": (5000, 512, 256, 4, "d2"),
                                "fliplr": (5000, 512, 256, 4, "fliplr"), "none_k4": (5000, 512, 256, 4, None)}.items():
    eng = TiledInferenceEngine((H, H, 3), T, S, channels=K, weight="pyramid", tta=tta)
    pred = torch.randn(eng.views * eng.num_tiles, K, T, T, device=dev)
    out = torch.empty((K, H, H), dtype=torch.float32, device=dev)
    for dbg, label in (("0", "grouped stages"), ("1", "one tile per stage")):
        os.environ["TB200_MERGE_VARIANT"] = dbg
        print(name, label, round(timed(lambda: eng.merge(pred, out=out)), 3), "ms", flush=True)
    del pred, out, eng
    torch.cuda.empty_cache()
