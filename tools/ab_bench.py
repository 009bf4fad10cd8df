"""Interleaved A/B timing of slice-apply kernel configurations at the headline shape: every
configuration is timed in short bursts, round-robin, and the median burst is reported (guards
against clock / thermal drift across a long sweep)."""
import os, sys, statistics, torch
sys.path.insert(0, ".")
from hdrnet_b200 import hdrnet_ops, _lib
B = 8
gen = torch.Generator(device="cuda").manual_seed(1234)
grid = torch.rand(B, 16, 16, 8, 12, device="cuda", generator=gen)
guide = torch.rand(B, 2160, 3840, device="cuda", generator=gen)
inp = torch.rand(B, 2160, 3840, 3, device="cuda", generator=gen)
out = torch.empty_like(inp)
ALL_CONFIGS = {
    "auto": (dict(), _lib.VARIANT_AUTO),
    "tex t512 c4": (dict(HDRNET_TMA_THREADS="512", HDRNET_TEX_CHUNKS="4"), _lib.VARIANT_TEX),
    "tex t512 c5": (dict(HDRNET_TMA_THREADS="512", HDRNET_TEX_CHUNKS="5"), _lib.VARIANT_TEX),
    "tex t512 c3": (dict(HDRNET_TMA_THREADS="512", HDRNET_TEX_CHUNKS="3"), _lib.VARIANT_TEX),
    "tex t256 c4": (dict(HDRNET_TMA_THREADS="256", HDRNET_TEX_CHUNKS="4"), _lib.VARIANT_TEX),
    "async c4": (dict(HDRNET_TEX_CHUNKS="4", HDRNET_ASYNC_STORE="0", HDRNET_ASYNC_SLAB="0"), _lib.VARIANT_TEX_ASYNC),
    "async c5": (dict(HDRNET_TEX_CHUNKS="5", HDRNET_ASYNC_STORE="0", HDRNET_ASYNC_SLAB="0"), _lib.VARIANT_TEX_ASYNC),
    "async c4 stg": (dict(HDRNET_TEX_CHUNKS="4", HDRNET_ASYNC_STORE="1", HDRNET_ASYNC_SLAB="0"), _lib.VARIANT_TEX_ASYNC),
    "async c5 stg": (dict(HDRNET_TEX_CHUNKS="5", HDRNET_ASYNC_STORE="1", HDRNET_ASYNC_SLAB="0"), _lib.VARIANT_TEX_ASYNC),
    "async c6 stg": (dict(HDRNET_TEX_CHUNKS="6", HDRNET_ASYNC_STORE="1", HDRNET_ASYNC_SLAB="0"), _lib.VARIANT_TEX_ASYNC),
    "async c4 slab": (dict(HDRNET_TEX_CHUNKS="4", HDRNET_ASYNC_STORE="0", HDRNET_ASYNC_SLAB="1"), _lib.VARIANT_TEX_ASYNC),
    "async c5 slab": (dict(HDRNET_TEX_CHUNKS="5", HDRNET_ASYNC_STORE="0", HDRNET_ASYNC_SLAB="1"), _lib.VARIANT_TEX_ASYNC),
    "async c4 stg slab": (dict(HDRNET_TEX_CHUNKS="4", HDRNET_ASYNC_STORE="1", HDRNET_ASYNC_SLAB="1"), _lib.VARIANT_TEX_ASYNC),
    "async c5 stg slab": (dict(HDRNET_TEX_CHUNKS="5", HDRNET_ASYNC_STORE="1", HDRNET_ASYNC_SLAB="1"), _lib.VARIANT_TEX_ASYNC),
    "async c6 stg slab": (dict(HDRNET_TEX_CHUNKS="6", HDRNET_ASYNC_STORE="1", HDRNET_ASYNC_SLAB="1"), _lib.VARIANT_TEX_ASYNC),
    "async c3": (dict(HDRNET_TEX_CHUNKS="3"), _lib.VARIANT_TEX_ASYNC),
    "async px c5": (dict(HDRNET_TEX_CHUNKS="5", HDRNET_ASYNC_LEAN="0"), _lib.VARIANT_TEX_ASYNC),
    "ws  t256": (dict(HDRNET_TMA_THREADS="256"), _lib.VARIANT_TEX_WS),
    "ws  t512": (dict(HDRNET_TMA_THREADS="512"), _lib.VARIANT_TEX_WS),
    "tex t320x3 c4": (dict(HDRNET_TMA_THREADS="320", HDRNET_TEX_CHUNKS="4"), _lib.VARIANT_TEX),
    "texin t512 c4": (dict(HDRNET_TMA_THREADS="512", HDRNET_TEX_CHUNKS="4"), _lib.VARIANT_TEX_IN),
    "texin t256 c4": (dict(HDRNET_TMA_THREADS="256", HDRNET_TEX_CHUNKS="4"), _lib.VARIANT_TEX_IN),
    "texin t256x3": (dict(HDRNET_TMA_THREADS="256", HDRNET_TEXIN_OCC="3"), _lib.VARIANT_TEX_IN),
    "texin t256x4": (dict(HDRNET_TMA_THREADS="256", HDRNET_TEXIN_OCC="4"), _lib.VARIANT_TEX_IN),
    "tma t256": (dict(HDRNET_TMA_THREADS="256"), _lib.VARIANT_TMA),
    "tma t512": (dict(HDRNET_TMA_THREADS="512"), _lib.VARIANT_TMA),
}
# python tools/ab_bench.py [substring ...]: only the configurations whose name contains one
sel = sys.argv[1:]
CONFIGS = {k: v for k, v in ALL_CONFIGS.items() if not sel or any(t in k for t in sel)}
KEYS = ("HDRNET_TMA_THREADS", "HDRNET_TEX_CHUNKS", "HDRNET_TMA_STAGES", "HDRNET_TMA_OCC", "HDRNET_TEXIN_OCC",
        "HDRNET_ASYNC_LEAN", "HDRNET_ASYNC_STORE", "HDRNET_ASYNC_SLAB")
def burst(env, variant, iters=40):
    for k in KEYS: os.environ.pop(k, None)
    os.environ.update(env)
    f = lambda: hdrnet_ops.bilateral_slice_apply(grid, guide, inp, True, out=out, variant=variant)
    for _ in range(3): f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): f()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters
res = {k: [] for k in CONFIGS}
for r in range(7):
    for k, (env, v) in CONFIGS.items():
        res[k].append(burst(env, v))
lines = []
for k, v in res.items():
    med = statistics.median(v)
    lines.append(f"{k:18s} median {med:.4f} ms  min {min(v):.4f}  max {max(v):.4f}  frac {8*2160*3840*28/med/1e6/6577.4:.4f}")
print("\n".join(lines))
os.makedirs("gpurun_out", exist_ok=True)
with open("gpurun_out/ab_bench.txt", "w") as f:
    f.write("# tools/ab_bench.py: 4K x 8, grid 16x16x8, 7 interleaved bursts of 40 launches per configuration\n")
    f.write("\n".join(lines) + "\n")
