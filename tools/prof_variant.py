"""Run a few launches of one slice-apply variant at the headline shape (for ncu)."""
import sys
import torch
sys.path.insert(0, ".")
from hdrnet_b200 import hdrnet_ops
variant = int(sys.argv[1]) if len(sys.argv) > 1 else 0
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
gen = torch.Generator(device="cuda").manual_seed(1234)
grid = torch.rand(B, 16, 16, 8, 12, device="cuda", generator=gen)
guide = torch.rand(B, 2160, 3840, device="cuda", generator=gen)
inp = torch.rand(B, 2160, 3840, 3, device="cuda", generator=gen)
out = torch.empty_like(inp)
for _ in range(4):
    hdrnet_ops.bilateral_slice_apply(grid, guide, inp, True, out=out, variant=variant)
torch.cuda.synchronize()
