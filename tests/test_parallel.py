"""world_size-2 gloo tests (CPU) of the N>1 host logic: batch sharding, the single weight
broadcast, and the max-over-ranks timing reduction bench.py uses."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from hdrnet_b200 import parallel


def test_shard_batch_partitions_exactly():
    for n in (0, 1, 7, 8, 64, 65):
        for world in (1, 2, 3, 8):
            spans = [parallel.shard_batch(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
                assert a1 == b0 and a1 >= a0 and b1 >= b0
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    assert parallel.shard_rows(2160, 1, 8) == (270, 540)


def test_shard_plan_tiles_the_job_exactly():
    """whole images per rank while there are enough of them, else row bands of
    every image; either way the parts of all ranks tile the job with nothing shared or left out."""
    for n, H, world in ((8, 2160, 8), (64, 3024, 8), (9, 100, 4), (1, 2160, 8), (3, 37, 8), (2, 5, 8), (1, 1, 2)):
        plans = [parallel.shard_plan(n, H, r, world) for r in range(world)]
        kinds = {p[0] for p in plans}
        assert kinds == ({"batch"} if n >= world else {"rows"})
        extent = n if n >= world else H
        assert plans[0][1] == 0 and plans[-1][2] == extent
        for a, b in zip(plans, plans[1:]):
            assert a[2] == b[1] and a[2] >= a[1]
        sizes = [p[2] - p[1] for p in plans]
        assert sum(sizes) == extent and max(sizes) - min(sizes) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                      MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    r, w, _ = parallel.init_distributed("gloo")
    weights = None
    if r == 0:
        rng = np.random.RandomState(0)
        weights = {"inference/coefficients/splat/conv1/weights": rng.randn(3, 3, 3, 8).astype(np.float32),
                   "inference/guide/ccm": np.eye(3, dtype=np.float32),
                   "inference/coefficients/global/fc1/weights": rng.randn(16, 4).astype(np.float32)}
    got = parallel.broadcast_weights(weights)
    checksum = float(sum(np.asarray(v, np.float64).sum() for v in got.values()))
    shapes = {k: v.shape for k, v in got.items()}
    slowest = parallel.max_over_ranks(10.0 + rank)
    span = parallel.shard_batch(9, r, w)
    parallel.finalize()
    q.put((rank, checksum, shapes, slowest, span))


def test_weight_broadcast_and_timing_reduction_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, c0, s0, t0, span0), (r1, c1, s1, t1, span1) = results
    assert c0 == c1 and s0 == s1
    assert s0["inference/coefficients/splat/conv1/weights"] == (3, 3, 3, 8)
    assert t0 == t1 == 11.0                       # max over ranks
    assert span0 == (0, 5) and span1 == (5, 9)    # batch shard covers all 9 images


def test_numa_binding_is_safe_without_nvml_and_never_widens_the_mask():
    """bind_to_gpu_numa must be a no-op when NVML / the GPU is absent (this container), and when it
    does bind, the new mask is a subset of the CPUs the process was allowed before."""
    import os
    from hdrnet_b200 import parallel
    before = os.sched_getaffinity(0)
    try:
        cpus = parallel.bind_to_gpu_numa(0)
        after = os.sched_getaffinity(0)
        assert set(cpus) <= before and after <= before
        assert (not cpus and after == before) or (cpus and after == set(cpus))
    finally:
        os.sched_setaffinity(0, before)
