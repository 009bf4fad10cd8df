"""Helper of the tuning-knob tests.  libhdrnet_b200.so reads its tuning record (HDRNET_ASYNC_THREADS,
HDRNET_TEX_CHUNKS, HDRNET_FUSED_ASYNC) ONCE per process, so a knob setting needs a process of its
own: this script runs one seeded case under the environment it was started with and prints the
SHA-256 of the result bytes.
    python tests/knob_runner.py apply <variant>     # the APPLY_CASES below, one "SHA256 <case> <hex>" line each
    python tests/knob_runner.py fused               # curves / nn x float32 / uint8 / uint16
One process per knob setting (not per case): a Python + torch start-up costs ~15 s on the GPU box.
"""
import hashlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))


# (seed, B, H, W, gh, gw, gd, out-of-range guides)
APPLY_CASES = [(5, 2, 64, 3840, 16, 16, 8, False),     # headline row shape
               (6, 1, 700, 1028, 5, 7, 3, True),       # many rows per CTA, ragged last segment, edge guides
               (7, 3, 9, 128, 8, 64, 4, False),        # W < 4 gw: per-pixel indices
               (8, 2, 300, 2048, 8, 8, 4, True)]       # small grid, 1536-byte slab rows, several rows per grid row


def apply_case(variant, seed, B, H, W, gh, gw, gd, edge):
    from hdrnet_b200 import hdrnet_ops
    from util import rand_case
    grid, guide, inp = rand_case(seed, B, H, W, gh, gw, gd, signed=True)
    if edge:
        guide[0, :, ::5] = 1.75
        guide[0, :, 1::5] = -0.6
    # a call with ANOTHER grid first: it leaves its slab rows in the workspace block the caching allocator
    # hands to the next call -- a kernel that read stale rows (e.g. through a cache) would show it
    other = torch.from_numpy(np.ascontiguousarray(grid[::-1, ::-1] * 1.5 + 0.25)).cuda()
    hdrnet_ops.bilateral_slice_apply(other, torch.from_numpy(guide).cuda(), torch.from_numpy(inp).cuda(), True,
                                     variant=variant)
    out = hdrnet_ops.bilateral_slice_apply(torch.from_numpy(grid).cuda(), torch.from_numpy(guide).cuda(),
                                           torch.from_numpy(inp).cuda(), True, variant=variant)
    torch.cuda.synchronize()
    return out.cpu().numpy()


def fused_case(kind, dtype):
    from hdrnet_b200 import models
    from oracle import model_np as M
    cls = models.HDRNetCurves if kind == "curves" else models.HDRNetPointwiseNNGuide
    p = dict(M.DEFAULT_PARAMS, net_input_size=64, spatial_bin=8, luma_bins=8)
    if kind == "nn":
        p.update(model_name="HDRNetPointwiseNNGuide", batch_norm=True)
    p["weights"] = models.init_weights(p, seed=3)
    rng = np.random.RandomState(5)
    B, H, W = 2, 300, 3840           # >= 2 Mi pixels: the texture-assisted fused kernels
    if dtype == "float32":
        im = torch.from_numpy(rng.rand(B, H, W, 3).astype(np.float32)).cuda()
        low = torch.from_numpy(rng.rand(B, 64, 64, 3).astype(np.float32)).cuda()
        out_dtype = torch.float32
    else:
        hi = 256 if dtype == "uint8" else 65536
        im = torch.from_numpy(rng.randint(0, hi, size=(B, H, W, 3)).astype(dtype)).cuda()
        low = models.lowres_from_image(im, 64)
        out_dtype = torch.uint8
    coeffs = cls._coefficients(low, p)
    out = cls._fullres(coeffs, im, p, out_dtype)
    torch.cuda.synchronize()
    return out.cpu().numpy()


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


if __name__ == "__main__":
    if sys.argv[1] == "apply":
        for k, case in enumerate(APPLY_CASES):
            print("SHA256", f"apply{k}", sha(apply_case(int(sys.argv[2]), *case)), flush=True)
    else:
        for kind in ("curves", "nn"):
            for dtype in ("float32", "uint8", "uint16"):
                print("SHA256", f"{kind}-{dtype}", sha(fused_case(kind, dtype)), flush=True)
