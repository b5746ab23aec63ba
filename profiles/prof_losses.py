"""ncu driver: focal + dice fwd+bwd at cfg3 size (one launch set after warm-up)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pytorch_toolbelt_b200 import losses as L
B, C, H, W = 32, 4, 1024, 1024
g = torch.Generator().manual_seed(0)
x = (torch.randn(B, C, H, W, generator=g) * 3).cuda()
t = (torch.rand(B, C, H, W, generator=g) < 0.3).float().cuda()
lab = torch.randint(0, C, (B, H, W), generator=g).cuda()
for _ in range(2):
    for fn in (lambda z: L.BinaryFocalLoss()(z, t), lambda z: L.DiceLoss("multilabel")(z, t), lambda z: L.DiceLoss("multiclass")(z, lab),
               lambda z: L.CrossEntropyFocalLoss()(z, lab), lambda z: L.WingLoss()(z, t)):
        xr = x.detach().requires_grad_(True)
        fn(xr).backward()
torch.cuda.synchronize()
print("done")
