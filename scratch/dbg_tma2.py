import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
from pytorch_toolbelt_b200.inference.engine import TiledInferenceEngine

def run(tta_name, K, shape=(700, 900, 3), T=128, S=64, reps=3, decode=False):
    eng = TiledInferenceEngine(shape, T, S, channels=K, weight="mean", tta=tta_name)
    n = eng.num_tiles
    if decode:
        pred = (torch.arange(eng.views * n).view(-1, 1, 1, 1) * 10.0 + torch.arange(K).view(1, K, 1, 1)).expand(-1, -1, T, T).contiguous().float().cuda()
    else:
        pred = torch.randn(eng.views * n, K, T, T, generator=torch.Generator().manual_seed(5)).cuda()
    os.environ["TB200_FORCE_GENERIC_MERGE"] = "1"
    slow = eng.merge(pred); torch.cuda.synchronize()
    del os.environ["TB200_FORCE_GENERIC_MERGE"]
    res = []
    for rep in range(reps):
        fast = eng.merge(pred); torch.cuda.synchronize()
        d = (fast != slow)
        res.append(int(d.sum()))
        if decode and res[-1]:
            nz = d.nonzero()
            k0, y0, x0 = nz[0].tolist()
            print("   decode first bad", nz[0].tolist(), "fast", float(fast[k0, y0, x0]), "slow", float(slow[k0, y0, x0]),
                  "bad ks", nz[:, 0].unique().tolist(), "n", len(nz))
            # show distinct (fast-slow) values
            diffs = (fast - slow)[d]
            print("   distinct diffs", diffs.unique()[:12].tolist())
    return res

for dbg in ("0", "4", "1", "2"):
    os.environ["TB200_TMA_DEBUG"] = dbg
    print("DEBUG", dbg)
    for cfg in ((None, 3), (None, 2), (None, 4), (None, 5), ("fliplr", 3), ("d4", 3)):
        print("  ", cfg, run(*cfg))
os.environ["TB200_TMA_DEBUG"] = "0"
print("decode run")
run(None, 3, decode=True, reps=2)
