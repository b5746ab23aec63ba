"""README-loop (reference README.md:196-227) through the drop-in classes on one B200, cfg2 geometry, batch 8:
   tiles (resident on the GPU as uint8 CHW, like the DataLoader output after .cuda()) -> .float() -> tta.d4_image_augment
   -> stand-in model -> tta.d4_image_deaugment -> CudaTileMerger.integrate_batch -> merge.
Times each stage with CUDA events; prints one JSON object. Also times the fused integrate_batch(tta=D4) variant."""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pytorch_toolbelt_b200 import _native as N
from pytorch_toolbelt_b200._native import TB
from pytorch_toolbelt_b200.inference import tta
from pytorch_toolbelt_b200.inference.tiles import CudaTileMerger, ImageSlicer
from pytorch_toolbelt_b200.inference.engine import TiledInferenceEngine

H = W = 5000; Cin = 3; T = 512; S = 256; K = 4; B = 8
dev = torch.device("cuda", 0)
rng = np.random.default_rng(0)
image = rng.integers(0, 256, (H, W, Cin)).astype(np.uint8)
tiler = ImageSlicer(image.shape, T, S, weight="pyramid")
eng = TiledInferenceEngine(image.shape, T, S, channels=K, weight="pyramid", tta=None)
tiles_dev = eng.slice(eng.to_device(image)).to(torch.uint8)  # [361,3,512,512] u8 CHW: what the DataLoader delivers
n = tiles_dev.shape[0]
Wm = torch.randn(K, Cin, generator=torch.Generator().manual_seed(0))
wm = (C.c_float * (K * Cin))(*[float(v) for v in Wm.flatten().tolist()])
def model(x):
    y = torch.empty((x.shape[0], K, T, T), dtype=torch.float32, device=dev)
    N.check(N.lib().tb_standin_model(x.data_ptr(), x.shape[0], Cin, K, T, T, wm, y.data_ptr(), N.stream_ptr(dev)), "m")
    return y
def ev(): return torch.cuda.Event(enable_timing=True)
def loop(fused):
    merger = CudaTileMerger(tiler.target_shape, K, tiler.weight)
    acc = dict(float=0.0, augment=0.0, model=0.0, deaugment=0.0, integrate=0.0, merge=0.0)
    marks = []
    t_all0, t_all1 = ev(), ev()
    t_all0.record()
    for s in range(0, n, B):
        e = [ev() for _ in range(5)]
        e[0].record(); x = tiles_dev[s:s + B].float()
        e[1].record(); xa = tta.d4_image_augment(x)
        e[2].record(); y = model(xa)
        e[3].record()
        if fused:
            merger.integrate_batch(y, tiler.crops[s:s + B], tta=TB.TTA_D4)
            e[4].record(); marks.append((e, None))
        else:
            yd = tta.d4_image_deaugment(y)
            e[4].record(); merger.integrate_batch(yd, tiler.crops[s:s + B])
            e5 = ev(); e5.record(); marks.append((e, e5))
    m0, m1 = ev(), ev()
    m0.record(); out = merger.merge(); m1.record()
    t_all1.record()
    torch.cuda.synchronize()
    for e, e5 in marks:
        acc["float"] += e[0].elapsed_time(e[1]); acc["augment"] += e[1].elapsed_time(e[2]); acc["model"] += e[2].elapsed_time(e[3])
        if e5 is None:
            acc["integrate"] += e[3].elapsed_time(e[4])
        else:
            acc["deaugment"] += e[3].elapsed_time(e[4]); acc["integrate"] += e[4].elapsed_time(e5)
    acc["merge"] = m0.elapsed_time(m1)
    acc["total_ms"] = t_all0.elapsed_time(t_all1)
    return acc, out
for fused in (False, True):
    loop(fused)
res = {}
for fused in (False, True):
    acc, out = loop(fused)
    eng_ms = acc["total_ms"] - acc["model"]
    res["fused_integrate" if fused else "separate_deaugment"] = {**{k: round(v, 3) for k, v in acc.items()},
        "engine_only_ms": round(eng_ms, 3), "engine_only_MPs": round(H * W / 1e6 / (eng_ms * 1e-3), 1)}
# parity of the two variants against the fused engine
ref = TiledInferenceEngine(image.shape, T, S, channels=K, weight="pyramid", tta="d4").predict(image, model, batch_size=8)
top, left = tiler.margin_top, tiler.margin_left
err = (out[:, top:top + H, left:left + W] - ref).abs().max().item() / ref.abs().max().item()
res["rel_err_vs_engine"] = err
print(json.dumps(res))
