import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pytorch_toolbelt_b200.inference.engine import TiledInferenceEngine
eng = TiledInferenceEngine((5000, 5000, 3), 512, 256, channels=8, weight="pyramid", tta="flips")
pred = torch.randn(eng.views * eng.num_tiles, 8, 512, 512, device="cuda")
out = torch.empty((8, 5000, 5000), device="cuda")
for _ in range(3):
    eng.merge(pred, out=out)
torch.cuda.synchronize()
