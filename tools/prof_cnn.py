"""ncu target: the coefficient network at batch 1 (launch chain), 3 warm-up calls + 1."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hdrnet_b200 import models
p = dict(models.DEFAULT_PARAMS); p["weights"] = models.init_weights(p, seed=0)
B = int(os.environ.get("B", "1"))
low = torch.rand(B, 256, 256, 3, device="cuda")
models.CHAIN_CNN_MAX_BATCH = 64
for _ in range(4): models.HDRNetCurves._coefficients(low, p)
torch.cuda.synchronize()
