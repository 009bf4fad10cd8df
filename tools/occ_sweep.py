import os, sys, torch
sys.path.insert(0, ".")
from hdrnet_b200 import hdrnet_ops, _lib
B=8
gen = torch.Generator(device="cuda").manual_seed(1234)
grid = torch.rand(B, 16, 16, 8, 12, device="cuda", generator=gen)
guide = torch.rand(B, 2160, 3840, device="cuda", generator=gen)
inp = torch.rand(B, 2160, 3840, 3, device="cuda", generator=gen)
out = torch.empty_like(inp)
def t(variant, iters=100):
    f = lambda: hdrnet_ops.bilateral_slice_apply(grid, guide, inp, True, out=out, variant=variant)
    for _ in range(5): f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): f()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters
ref = hdrnet_ops.bilateral_slice_apply(grid, guide, inp, True, variant=_lib.VARIANT_TMA).clone()
for occ in ("2", "3"):
    os.environ["HDRNET_TMA_OCC"] = occ
    print("occ", occ, "tma", round(t(_lib.VARIANT_TMA), 4), "ms")
    for c in (3, 4, 5):
        os.environ["HDRNET_TEX_CHUNKS"] = str(c)
        ms = t(_lib.VARIANT_TEX)
        ok = torch.equal(hdrnet_ops.bilateral_slice_apply(grid, guide, inp, True, variant=_lib.VARIANT_TEX), ref)
        print("occ", occ, "tex chunks", c, round(ms, 4), "ms", round(8*2160*3840*28/ms/1e6/6577.4, 4), "frac", "bitwise-ok" if ok else "MISMATCH")
