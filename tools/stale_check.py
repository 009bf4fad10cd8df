"""Does a kernel form read slab rows of an EARLIER call?  Runs every seeded case of tests/knob_runner.py
through the issuer-warp form, then again right after a call with another grid (whose rows stay in the
workspace block the caching allocator hands out next), and compares with the block-synchronous form."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from hdrnet_b200 import hdrnet_ops, _lib
from util import rand_case
import knob_runner
for k, (seed, B, H, W, gh, gw, gd, edge) in enumerate(knob_runner.APPLY_CASES):
    grid, guide, inp = rand_case(seed, B, H, W, gh, gw, gd, signed=True)
    g, u, i = (torch.from_numpy(a).cuda() for a in (grid, guide, inp))
    ref = hdrnet_ops.bilateral_slice_apply(g, u, i, True, variant=_lib.VARIANT_TEX)
    a1 = hdrnet_ops.bilateral_slice_apply(g, u, i, True, variant=_lib.VARIANT_TEX_ASYNC)
    other = torch.from_numpy(np.ascontiguousarray(grid[::-1, ::-1] * 1.5 + 0.25)).cuda()
    hdrnet_ops.bilateral_slice_apply(other, u, i, True, variant=_lib.VARIANT_TEX_ASYNC)
    a2 = hdrnet_ops.bilateral_slice_apply(g, u, i, True, variant=_lib.VARIANT_TEX_ASYNC)
    torch.cuda.synchronize()
    d1 = (a1 - ref).abs().max().item(); d2 = (a2 - ref).abs().max().item()
    bad = (a2 != ref).nonzero()
    print(f"case {k} {B}x{H}x{W} grid {gh}x{gw}x{gd}: fresh diff {d1:.3e}  after-other-grid diff {d2:.3e}  mismatches {bad.shape[0]}", bad[:3].tolist())
