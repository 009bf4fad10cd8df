"""CPU emulation of the arithmetic of the experimental tensor-core form (csrc/slice_apply_tc.cu):
the depth interpolation as D = A x B per 128-pixel tile with 3xTF32 operand splitting
(hi = the 19 bits the tensor core reads, lo = the exact remainder; A_hi B_hi + A_lo B_hi + A_hi B_lo),
the tile's three-x-cell window and the per-pixel choice of two of them -- against the oracle.
The kernel itself has not run on a GPU yet (its GPU tests are gated behind
HDRNET_TEST_EXPERIMENTAL=1); this pins the ALGORITHM to the 1e-5 bar before that first run."""
import numpy as np
import pytest

import oracle
from util import RTOL, rand_case, rel_err

F = np.float32


def tf32_split(v):
    v = np.asarray(v, dtype=F)
    hi = (v.view(np.uint32) & np.uint32(0xFFFFE000)).view(F)
    return hi, (v - hi).astype(F)


def axis(p, scale):
    """common.cuh spatial_axis: explicitly rounded float32 steps."""
    t = (((p.astype(F) + F(0.5)) * F(scale)).astype(F) - F(0.5)).astype(F)
    fl = np.floor(t)
    return fl.astype(np.int64), (t - fl).astype(F)


def emulate(grid, guide, inp, tile=128):
    B, gh, gw, gd, _ = grid.shape
    _, H, W = guide.shape
    assert gd == 8 and gw >= 3 and W >= 128 * gw
    sx, sy = F(gw) / F(W), F(gh) / F(H)
    out = np.empty((B, H, W, 3), dtype=F)
    xs = np.arange(W)
    ix, fx = axis(xs, sx)
    for b in range(B):
        for y in range(H):
            iy, fy = axis(np.array([y]), sy)
            g0 = grid[b, np.clip(iy[0], 0, gh - 1)].astype(F)          # [gw, 8, 12]
            g1 = grid[b, np.clip(iy[0] + 1, 0, gh - 1)].astype(F)
            w1, w0 = fy[0], F(1.0) - fy[0]
            slab = (np.float64(w1) * g1 + np.float64((w0 * g0).astype(F))).astype(F)   # lerp4: fma(w1, b, w0 * a)
            s_hi, s_lo = tf32_split(slab)
            # depth axis per pixel (range_axis + smoothed_weights)
            tz = ((guide[b, y].astype(F) * F(8.0)).astype(F) - F(0.5)).astype(F)
            iz = np.floor(tz).astype(np.int64)
            fz = (tz - iz.astype(F)).astype(F)
            u = (F(1.0) - fz).astype(F)
            wz0 = np.maximum(F(1.0) - np.sqrt((fz * fz + F(1e-8)).astype(F)), F(0.0)).astype(F)
            wz1 = np.maximum(F(1.0) - np.sqrt((u * u + F(1e-8)).astype(F)), F(0.0)).astype(F)
            zc0, zc1 = np.clip(iz, 0, 7), np.clip(iz + 1, 0, 7)
            A = np.zeros((W, 8), dtype=F)
            np.add.at(A, (xs, zc0), wz0)
            np.add.at(A, (xs, zc1), wz1)
            a_hi, a_lo = tf32_split(A)
            for x0 in range(0, W, tile):
                sl = slice(x0, min(x0 + tile, W))
                cb = int(np.clip(ix[x0], 0, gw - 3))
                assert ix[sl].max() <= ix[x0] + 1                      # a tile spans at most two floors
                bh = s_hi[cb:cb + 3].astype(np.float64)                # [3, 8, 12]
                bl = s_lo[cb:cb + 3].astype(np.float64)
                ah, al = a_hi[sl].astype(np.float64), a_lo[sl].astype(np.float64)
                D = (np.einsum("pk,ckj->pcj", ah, bh) + np.einsum("pk,ckj->pcj", al, bh)
                     + np.einsum("pk,ckj->pcj", ah, bl)).astype(F)     # [p, 3 cells, 12]
                l0 = np.clip(ix[sl], 0, gw - 1) - cb
                l1 = np.clip(ix[sl] + 1, 0, gw - 1) - cb
                assert l0.min() >= 0 and l1.max() <= 2
                n = np.arange(D.shape[0])
                f = fx[sl][:, None]
                v = (f * D[n, l1] + ((F(1.0) - f) * D[n, l0]).astype(F)).astype(F)   # [p, 12]
                rgb1 = np.concatenate([inp[b, y, sl].astype(F), np.ones((D.shape[0], 1), F)], axis=1)
                out[b, y, sl] = np.einsum("pij,pj->pi", v.reshape(-1, 3, 4).astype(np.float64),
                                          rgb1.astype(np.float64)).astype(F)
    return out


@pytest.mark.parametrize("shape", [(1, 6, 512, 4, 4, 8), (2, 5, 1152, 3, 9, 8), (1, 3, 3840, 16, 16, 8)],
                         ids=lambda s: "x".join(map(str, s)))
def test_tensor_core_algorithm_meets_the_parity_bar(shape):
    B, H, W, gh, gw, gd = shape
    grid, guide, inp = rand_case(4242, B, H, W, gh, gw, gd, signed=True)
    guide[0, 0, :6] = [0.0, 1.0, -0.3, 1.7, 0.0625, 0.9375]    # clamped depth cells, cell centres
    want = oracle.best().bilateral_slice_apply(grid, guide, inp, True)
    got = emulate(grid, guide, inp)
    assert rel_err(got, want) <= RTOL, rel_err(got, want)


def test_plain_tf32_would_miss_the_bar():
    """Why the operands are split: a single TF32 product is ~1e-4 off."""
    grid, guide, inp = rand_case(7, 1, 4, 512, 4, 4, 8, signed=True)
    want = oracle.best().bilateral_slice_apply(grid, guide, inp, True)
    hi = (grid.view(np.uint32) & np.uint32(0xFFFFE000)).view(F)
    got = oracle.best().bilateral_slice_apply(hi, guide, inp, True)
    assert rel_err(got, want) > RTOL


def emulate_gather(grid, guide, inp, tile=128):
    """The second form (slice_apply_rows_tcg_kernel): one-hot A (exact), B' = both depth rows of
    every x cell split hi + lo, D = A B'_hi + A B'_lo -- a copy of the pixel's two depth rows to
    ~2^-22 -- and the 4-corner blend in float32 as the row kernels do it."""
    B, gh, gw, gd, _ = grid.shape
    _, H, W = guide.shape
    sx, sy = F(gw) / F(W), F(gh) / F(H)
    out = np.empty((B, H, W, 3), dtype=F)
    xs = np.arange(W)
    ix, fx = axis(xs, sx)
    for b in range(B):
        for y in range(H):
            iy, fy = axis(np.array([y]), sy)
            g0 = grid[b, np.clip(iy[0], 0, gh - 1)].astype(F)
            g1 = grid[b, np.clip(iy[0] + 1, 0, gh - 1)].astype(F)
            w1, w0 = fy[0], F(1.0) - fy[0]
            slab = (np.float64(w1) * g1 + np.float64((w0 * g0).astype(F))).astype(F)   # [gw, 8, 12]
            s_hi, s_lo = tf32_split(slab)
            copy = (s_hi.astype(np.float64) + s_lo.astype(np.float64)).astype(F)       # what D holds
            zup = np.minimum(np.arange(8) + 1, 7)
            tz = ((guide[b, y].astype(F) * F(8.0)).astype(F) - F(0.5)).astype(F)
            iz = np.floor(tz).astype(np.int64)
            fz = (tz - iz.astype(F)).astype(F)
            u = (F(1.0) - fz).astype(F)
            wz0 = np.maximum(F(1.0) - np.sqrt((fz * fz + F(1e-8)).astype(F)), F(0.0)).astype(F)
            wz1 = np.maximum(F(1.0) - np.sqrt((u * u + F(1e-8)).astype(F)), F(0.0)).astype(F)
            zc0, zc1 = np.clip(iz, 0, 7), np.clip(iz + 1, 0, 7)
            same = zc0 == zc1
            wz0 = np.where(same, (wz0 + wz1).astype(F), wz0)
            wz1 = np.where(same, F(0.0), wz1)
            c0, c1 = np.clip(ix, 0, gw - 1), np.clip(ix + 1, 0, gw - 1)
            wx1, wx0 = fx, (F(1.0) - fx).astype(F)
            d00, d01 = copy[c0, zc0], copy[c0, zup[zc0]]       # half 0 = row z0, half 1 = row min(z0 + 1, 7)
            d10, d11 = copy[c1, zc0], copy[c1, zup[zc0]]
            w00, w01 = (wx0 * wz0).astype(F)[:, None], (wx0 * wz1).astype(F)[:, None]
            w10, w11 = (wx1 * wz0).astype(F)[:, None], (wx1 * wz1).astype(F)[:, None]
            v = (w11 * d11 + (w10 * d10 + (w01 * d01 + (w00 * d00).astype(F)).astype(F)).astype(F)).astype(F)
            rgb1 = np.concatenate([inp[b, y].astype(F), np.ones((W, 1), F)], axis=1)
            out[b, y] = np.einsum("pij,pj->pi", v.reshape(-1, 3, 4).astype(np.float64),
                                  rgb1.astype(np.float64)).astype(F)
    return out


@pytest.mark.parametrize("shape", [(1, 6, 512, 4, 4, 8), (1, 3, 3840, 16, 16, 8)],
                         ids=lambda s: "x".join(map(str, s)))
def test_gather_form_algorithm_meets_the_parity_bar(shape):
    B, H, W, gh, gw, gd = shape
    grid, guide, inp = rand_case(4243, B, H, W, gh, gw, gd, signed=True)
    guide[0, 0, :8] = [0.0, 1.0, -0.3, 1.7, 0.0625, 0.9375, 0.99, 0.01]   # both clamps, cell centres
    want = oracle.best().bilateral_slice_apply(grid, guide, inp, True)
    assert rel_err(emulate_gather(grid, guide, inp), want) <= RTOL
