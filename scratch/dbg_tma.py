import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
from pytorch_toolbelt_b200.inference.engine import TiledInferenceEngine
from oracle import tiles_oracle as TO
for tta_name, K in ((None, 3), (None, 1), ("d2", 1), ("fliplr", 2)):
    eng = TiledInferenceEngine((700, 900, 3), 128, 64, channels=K, weight="pyramid", tta=tta_name)
    pred = torch.randn(eng.views * eng.num_tiles, K, 128, 128, generator=torch.Generator().manual_seed(5)).cuda()
    for rep in range(3):
        fast = eng.merge(pred)
        torch.cuda.synchronize()
        os.environ["TB200_FORCE_GENERIC_MERGE"] = "1"
        slow = eng.merge(pred)
        torch.cuda.synchronize()
        del os.environ["TB200_FORCE_GENERIC_MERGE"]
        d = (fast != slow)
        nz = d.nonzero()
        print(tta_name, K, "rep", rep, "mismatch", int(d.sum()), "nan fast", int(torch.isnan(fast).sum()), "nan slow", int(torch.isnan(slow).sum()),
              "maxdiff", float((fast - slow).abs().nan_to_num().max()))
        if len(nz):
            print("  first", nz[:5].tolist(), "last", nz[-3:].tolist())
            ks = nz[:, 0].unique().tolist(); ys = nz[:, 1].unique(); xs = nz[:, 2].unique()
            print("  k", ks, "y range", int(ys.min()), int(ys.max()), len(ys), "x range", int(xs.min()), int(xs.max()), len(xs))
    if tta_name is None:
        # compare against the oracle merger
        m = TO.TorchMerger(eng.slicer.target_shape, K, eng.slicer.weight)
        m.integrate_batch(pred.cpu(), eng.slicer.crops)
        ref = m.merge()[:, eng.slicer.margin_top:eng.slicer.margin_top+700, eng.slicer.margin_left:eng.slicer.margin_left+900]
        print("  fast vs oracle", float((fast.cpu()-ref).abs().max()), " slow vs oracle", float((slow.cpu()-ref).abs().max()))
