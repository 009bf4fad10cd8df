import os, sys, json, torch
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
print("tma", round(t(_lib.VARIANT_TMA), 4))
print("defaults: tex", round(t(_lib.VARIANT_TEX), 4), "tex_ws", round(t(_lib.VARIANT_TEX_WS), 4), "auto", round(t(_lib.VARIANT_AUTO), 4))
for thr in ("256", "512"):
    os.environ["HDRNET_TMA_THREADS"] = thr
    os.environ["HDRNET_TEX_CHUNKS"] = "4"
    print("threads", thr, "tma", round(t(_lib.VARIANT_TMA), 4), "tex", round(t(_lib.VARIANT_TEX), 4),
          "tex_ws", round(t(_lib.VARIANT_TEX_WS), 4))
    for c in (5, 6):
        os.environ["HDRNET_TEX_CHUNKS"] = str(c)
        print("   threads", thr, "tex chunks", c, round(t(_lib.VARIANT_TEX), 4))
os.environ["HDRNET_TMA_THREADS"] = "256"
for c in ():
    os.environ["HDRNET_TEX_CHUNKS"] = str(c)
    ms = t(_lib.VARIANT_TEX)
    print("tex chunks", c, round(ms, 4), "ms", round(8*2160*3840*28/ms/1e6/6577.4, 4), "frac")
