"""Shared-memory wavefronts of the row kernels' corner loads, from the access pattern alone.

Model (B200_PROFILING / microbenchmarks in tools/ubench): an LDS.128 is served per QUARTER-WARP
(8 lanes x 16 B = one 128-byte wavefront); the 32 banks form 8 groups of 16 bytes; lanes reading the
SAME 16-byte address share a wavefront slot, lanes reading DIFFERENT addresses of one group serialise.
A quarter-warp therefore costs max over groups of the number of distinct addresses in the group.

Access pattern of slice_apply_rows_async_kernel (csrc/slice_apply_async.cu, process_quad_lean): lane q
of a warp owns the 4 consecutive pixels x0 + 4 q .. + 3; for pixel i of the quad, corner (dx, dz),
part p it reads   slab + (xcell + dx) * gd * 48 + clamp(zcell + dz) * 48 + 16 p
with xcell = floor((x + 0.5) * gw / W - 0.5) and zcell = floor(guide * gd - 0.5).

    python tools/sim_bank_conflicts.py            # table for the grids of BASELINE.json config 5
Prints wavefronts per warp-wide LDS.128 (ideal 4.00) for a uniformly random guide (the benchmark's)
and for a smooth guide (neighbouring pixels in one depth cell).  ncu measures 4.40 at 16x16x8 with
the random guide (profiles/r02_slabwarp_ncu_summary.md)."""
import numpy as np

W = 3840


def wavefronts(gw, gd, guide_row, rng):
    """mean wavefronts per warp instruction over all corner loads of one image row"""
    x = np.arange(W)
    tx = (x + 0.5) * (gw / W) - 0.5
    xcell = np.floor(tx).astype(np.int64)
    zcell = np.floor(guide_row * gd - 0.5).astype(np.int64)
    total, count = 0, 0
    quads = W // 4
    for i in range(4):                                   # pixel of the quad: one instruction stream each
        px = 4 * np.arange(quads) + i
        for dx in (0, 1):
            xc = np.clip(xcell[px] + dx, 0, gw - 1)
            for dz in (0, 1):
                zc = np.clip(zcell[px] + dz, 0, gd - 1)
                addr16 = xc * gd * 3 + zc * 3            # 16-byte units; part p adds p to every lane alike
                a = addr16[: quads // 8 * 8].reshape(-1, 8)          # quarter-warps of 8 consecutive quads
                grp = a % 8
                wf = np.zeros(len(a), np.int64)
                for g in range(8):
                    m = grp == g
                    # distinct addresses within the group, per quarter-warp
                    vals = np.where(m, a, -1)
                    vals.sort(axis=1)
                    distinct = (np.diff(vals, axis=1) != 0).sum(1) + 1 - (vals[:, 0] == -1)
                    wf = np.maximum(wf, distinct)
                total += wf.sum() * 4 / len(a) * 1.0     # 4 quarter-warps per warp
                count += 1
    return total / count


def tex_lines(gw, gd, guide_row):
    """mean number of distinct 128-byte lines one warp-wide texture fetch of a corner chunk touches
    (32 lanes; the texture path has no banks, its cost grows with the lines a request spreads over)"""
    x = np.arange(W)
    xcell = np.floor((x + 0.5) * (gw / W) - 0.5).astype(np.int64)
    zcell = np.floor(guide_row * gd - 0.5).astype(np.int64)
    quads = W // 4
    total, count = 0.0, 0
    for i in range(4):
        px = 4 * np.arange(quads) + i
        for dx in (0, 1):
            xc = np.clip(xcell[px] + dx, 0, gw - 1)
            for dz in (0, 1):
                zc = np.clip(zcell[px] + dz, 0, gd - 1)
                line = (xc * gd * 48 + zc * 48) // 128          # part p shifts all lanes alike
                a = np.sort(line[: quads // 32 * 32].reshape(-1, 32), axis=1)
                total += ((np.diff(a, axis=1) != 0).sum(1) + 1).mean()
                count += 1
    return total / count


def main():
    rng = np.random.RandomState(0)
    rows = 64
    xx = np.linspace(0, 1, W)
    smooth_rows = [np.clip(0.5 + 0.45 * np.sin(6.2831853 * (3 * xx + ph)) + 0.01 * (rng.rand(W) - 0.5), 0, 1)
                   for ph in np.linspace(0, 1, 16, endpoint=False)]
    print("| grid | random guide: wavefronts per LDS.128 | smooth guide | ideal | random guide: 128-byte lines per warp-wide texture fetch | smooth guide |")
    print("|---|---|---|---|---|---|")
    for gh, gw, gd in ((8, 8, 4), (16, 16, 4), (16, 16, 8), (32, 32, 8), (16, 16, 16), (32, 32, 16)):
        rand_rows = [rng.rand(W) for _ in range(rows)]
        rnd = np.mean([wavefronts(gw, gd, r, rng) for r in rand_rows])
        smooth = np.mean([wavefronts(gw, gd, r, rng) for r in smooth_rows])
        tl_r = np.mean([tex_lines(gw, gd, r) for r in rand_rows])
        tl_s = np.mean([tex_lines(gw, gd, r) for r in smooth_rows])
        print(f"| {gh}x{gw}x{gd} | {rnd:.2f} | {smooth:.2f} | 4.00 | {tl_r:.2f} | {tl_s:.2f} |")


if __name__ == "__main__":
    main()
