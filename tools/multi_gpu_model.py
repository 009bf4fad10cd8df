#!/usr/bin/env python
"""Batch-sharded HDRNetCurves inference across the GPUs of one box:
    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/multi_gpu_model.py
Rank 0 owns the weights; ONE NCCL broadcast distributes them at init; every rank then processes
its contiguous shard of the image batch with no further communication.  Prints one JSON line."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, ROOT)
from hdrnet_b200 import models, parallel  # noqa: E402


def main():
    rank, world, local_rank = parallel.init_distributed()
    torch.cuda.set_device(local_rank)
    params = dict(models.DEFAULT_PARAMS)
    weights = models.init_weights(params, seed=0) if rank == 0 else None
    weights = parallel.broadcast_weights(weights, src=0)          # the only collective
    checksum = float(sum(np.asarray(v, np.float64).sum() for v in weights.values()))
    params["weights"] = weights
    total_images, H, W = 8 * world, 2160, 3840
    lo, hi = parallel.shard_batch(total_images, rank, world)
    gen = torch.Generator(device="cuda").manual_seed(100 + rank)
    low = torch.rand(hi - lo, 256, 256, 3, device="cuda", generator=gen)
    full = torch.rand(hi - lo, H, W, 3, device="cuda", generator=gen)
    for _ in range(3):
        out = models.HDRNetCurves.inference(low, full, params)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 20
    a.record()
    for _ in range(iters):
        out = models.HDRNetCurves.inference(low, full, params)
    b.record()
    torch.cuda.synchronize()
    ms = parallel.max_over_ranks(a.elapsed_time(b) / iters)
    sums = [None] * world
    torch.distributed.all_gather_object(sums, checksum) if world > 1 else sums.__setitem__(0, checksum)
    if rank == 0:
        print(json.dumps({"case": "HDRNetCurves model, batch-sharded", "n_gpus": world,
                          "images": total_images, "shape": [H, W], "ms_per_batch": round(ms, 4),
                          "MP/s": round(total_images * H * W / ms / 1e3, 1),
                          "weights_identical_on_all_ranks": len(set(sums)) == 1,
                          "finite": bool(torch.isfinite(out).all())}), flush=True)
    parallel.finalize()


if __name__ == "__main__":
    main()
