import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pytorch_toolbelt_b200.inference.engine import TiledInferenceEngine
from pytorch_toolbelt_b200.inference import tta
from pytorch_toolbelt_b200.inference.tiles import CudaTileMerger
from pytorch_toolbelt_b200._native import TB
variant = sys.argv[1] if len(sys.argv) > 1 else "1"
os.environ["TB200_MERGE_VARIANT"] = variant
eng = TiledInferenceEngine((700, 900, 3), 128, 64, channels=3, weight="pyramid", tta="d4")
img = torch.from_numpy(np.random.default_rng(0).integers(0, 256, (700, 900, 3)).astype(np.uint8)).cuda()
aug = eng.slice(img)
pred = torch.randn(eng.views * eng.num_tiles, 3, 128, 128, generator=torch.Generator().manual_seed(1)).cuda()
out = eng.merge(pred)
torch.cuda.synchronize()
print("merge ok variant", variant, float(out.abs().max()))
y = tta.d4_image_deaugment(pred[:64])
m = CudaTileMerger(eng.slicer.target_shape, 3, eng.slicer.weight)
m.integrate_batch(pred[:64], eng.slicer.crops[:8], tta=TB.TTA_D4)
torch.cuda.synchronize()
print("deaug + scatter ok")
