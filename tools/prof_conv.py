import os, sys, torch
sys.path.insert(0, ".")
from hdrnet_b200 import models
os.environ["HDRNET_CONV_TCGEN05"] = "1"
B = 8
x = torch.rand(B, 16, 16, 64, device="cuda"); w = torch.rand(3, 3, 64, 64, device="cuda"); b = torch.rand(64, device="cuda")
packed = models.pack_conv_weights(w)
for _ in range(4):
    models._conv(x, (w, b, packed), stride=1)
torch.cuda.synchronize()
