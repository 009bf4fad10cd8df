"""Interleaved A/B of the CTA shapes of the fused-guide (model path) row kernel at 4K x 8:
256 threads x 2 CTAs/SM (default), 256 x 3 (HDRNET_TMA_OCC=3), 512 x 2 (HDRNET_TMA_THREADS=512),
and the issuer-warp control flow (HDRNET_FUSED_ASYNC=1, 8 math warps + issuer),
for the curves / pointwise-NN guides and float32 / uint8 pixels."""
import os, sys, statistics, torch
sys.path.insert(0, ".")
from hdrnet_b200 import models
B, H, W = 8, 2160, 3840
gen = torch.Generator(device="cuda").manual_seed(1)
im8 = torch.randint(0, 256, (B, H, W, 3), device="cuda", generator=gen, dtype=torch.uint8)
imf = models.image_to_float(im8)
ENVS = {"256x2": {}, "256x3": {"HDRNET_TMA_OCC": "3"}, "512x2": {"HDRNET_TMA_THREADS": "512"},
        "issuer-warp 288x2": {"HDRNET_FUSED_ASYNC": "1"}}
KEYS = ("HDRNET_TMA_OCC", "HDRNET_TMA_THREADS", "HDRNET_FUSED_ASYNC")
cases = {}
for kind, name in (("curves", "HDRNetCurves"), ("nn", "HDRNetPointwiseNNGuide")):
    p = dict(models.DEFAULT_PARAMS, model_name=name)
    p["weights"] = models.init_weights(p, seed=0, model_name=name)
    cls = getattr(models, name)
    coeffs = cls._coefficients(models.lowres_from_image(im8, 256), p)
    for px, im, dt in (("f32", imf, torch.float32), ("u8", im8, torch.uint8)):
        cases[f"{kind} {px}"] = (cls, coeffs, im, p, dt)

def burst(case, env, iters=20):
    cls, coeffs, im, p, dt = cases[case]
    for k in KEYS: os.environ.pop(k, None)
    os.environ.update(env)
    f = lambda: cls._fullres(coeffs, im, p, dt)
    for _ in range(2): f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): f()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters

res = {(c, e): [] for c in cases for e in ENVS}
for r in range(5):
    for c in cases:
        for e, env in ENVS.items():
            res[(c, e)].append(burst(c, env))
for c in cases:
    print(c.ljust(12), "  ".join(f"{e}: {statistics.median(res[(c, e)]):.4f} ms" for e in ENVS))
