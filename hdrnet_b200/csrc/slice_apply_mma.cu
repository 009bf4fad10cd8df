// slice_apply_mma.cu -- fused BilateralSliceApply with the per-pixel coefficient GATHER on the
// 5th-generation tensor cores (tcgen05.mma kind::tf32, accumulator and A operand in tensor memory).
//
// Replaces the hot loop hdrnet/ops/bilateral_slice_apply.cu.cc:36-126 (and, in its fused-guide
// forms, HDRNetCurves / HDRNetPointwiseNNGuide `_guide` + `_output`, hdrnet/models.py:145-210).
//
// Why: every earlier form of this kernel is bound by the SM's gather pipes -- a pixel needs 4
// corners x 12 coefficients = 192 B out of the row's y-pre-blended slab, and the shared-memory
// crossbar (128 B/clk/SM) plus the texture pipe (64 B/clk/SM) deliver that in ~1.8 clk/px/SM where
// the HBM roofline allows 1.24 (profiles/r01_async_ncu_full_summary.txt).  Tensor memory is read
// at ~750-900 B/clk/SM (tools/ubench/tmem_paths.cu), so the gather is phrased as a matrix product
// whose result lands there:
//
//     D[128 px][48] = A[128 px][K] x B'[K][48]                 K = 8 (gd <= 8) or 16 (gd <= 16)
//
//   A[p][k]  = 1.0 where k is pixel p's lower depth cell, else 0 -- EXACT in TF32, no operand
//              split; written to tensor memory by the pixel's own thread (tcgen05.st);
//   B'[k][n] = for the tile's two x cells (cx = x0, x0 + 1) and both depth rows a pixel of depth
//              cell k needs (h = 0, 1):  n = cx * 24 + h * 12 + c  ->  slab[cx][min(k + h, gd-1)][c],
//              split hi + lo into two TF32 operands (hi = top 19 bits, lo = v - hi), shared memory,
//              UMMA K-major no-swizzle layout;
//   D        = A B'_hi + A B'_lo: a copy (to 2^-22 relative) of the pixel's 4 corner vectors, which
//              the thread reads back with tcgen05.ld and blends in registers (24 FFMA2) before the
//              3x4 affine apply -- the arithmetic of the other row kernels.
//
// Tiles never straddle an x-cell boundary: a row is cut into RUNS of pixels that share the lower
// x cell x0 (floor((x + .5) gw / W - .5), found with the reference's own float arithmetic), runs
// into tiles of <= 128 pixels, so ONE cell pair (N = 48) serves a whole tile and every thread reads
// the same 48 TMEM columns (at 4K: 32 tiles of 120 / 128 / 112 pixels per row, 94 % of the lanes).
//
// No pre-pass, no workspace, no texture: a SLAB warp TMA-loads the image row's two grid rows,
// blends them in y (wy is constant along a row), splits and writes B' two rows ahead.
//
// CTA (one per SM, persistent over a contiguous range of row segments):
//   warps 0..15  four MATH warpgroups; a tile = one pixel per thread of a warpgroup.  Two TMEM
//                slots per warpgroup (D[48] | A[16]): the set-up of tile j+1 (guide -> depth cell
//                -> one-hot row -> tcgen05.st) is issued before the epilogue of tile j.  The LAST
//                of a warpgroup's four warps to finish its part of A (shared-memory counter,
//                acq_rel) issues the tile's MMAs and commits them to the slot's mbarrier: nobody
//                waits at a barrier.
//   warp 16      ISSUER: one lane issues every bulk copy of pixel segments (TMA ring in, one bulk
//                store per segment out), as in the issuer-warp row kernel.
//   warp 17      SLAB warp (above).
#include <cuda_runtime.h>

#include <algorithm>
#include <climits>
#include <cstdint>

#include "slice_rows.cuh"

namespace hdrnet_b200 {

constexpr int kMmWgs = 4;
constexpr int kMmMathWarps = kMmWgs * 4;
constexpr int kMmSlabWarps = 2;
constexpr int kMmThreads = (kMmMathWarps + 1 + kMmSlabWarps) * 32;   // + issuer warp + slab warps = 608
constexpr int kMmMaxStages = 4;
constexpr int kMmTile = 128;
constexpr int kMmWgCols = 128;      // per warpgroup: D of two tiles [0, 96), A operands [96, 128)
constexpr int kMmN = 48;            // 2 x cells x 2 depth rows x 12 coefficients
constexpr int kMmTmemCols = kMmWgs * kMmWgCols;   // 512: the whole tensor memory of the SM

struct MmArgs {
  const float* grid;
  const float* guide;       // GuideFromInput form
  float* guide_out;         // optional guide dump of the fused forms
  const unsigned char* input;
  unsigned char* out;
  SliceGeom g;
  int nseg, seg_px, stages, stage_bytes, off_guide, off_out;
  int in_bpp, out_bpp;
  int row_floats;           // gw * gd * 12
  int bsplit_bytes;         // one operand split of one row's B'
  int max_tiles;            // capacity of a segment's tile list
  int off_tab, off_raw, off_b, off_stage, smem_bytes, ctas;
};

// instruction descriptor: D f32, A / B tf32, both K-major, N = 48, M = 128
constexpr uint32_t kMmIdesc = (1u << 4) | (2u << 7) | (2u << 10) |
                              (static_cast<uint32_t>(kMmN >> 3) << 17) |
                              (static_cast<uint32_t>(kMmTile >> 4) << 24);

__device__ __forceinline__ uint64_t mm_kmajor_desc(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3fffu);
  d |= static_cast<uint64_t>((lbo >> 4) & 0x3fffu) << 16;
  d |= static_cast<uint64_t>((sbo >> 4) & 0x3fffu) << 32;
  d |= static_cast<uint64_t>(1) << 46;   // descriptor version (sm_100), no swizzle
  return d;
}
__device__ __forceinline__ void mm_mma(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(kMmIdesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void mm_commit(uint32_t bar_smem_addr) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_smem_addr)
               : "memory");
}
__device__ __forceinline__ void mm_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void mm_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void mm_st8(uint32_t taddr, const float (&a)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr),
               "f"(a[0]), "f"(a[1]), "f"(a[2]), "f"(a[3]), "f"(a[4]), "f"(a[5]), "f"(a[6]), "f"(a[7])
               : "memory");
}
// 16 consecutive columns of this thread's TMEM lane
__device__ __forceinline__ void mm_ld16(uint32_t taddr, float* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7]),
        "=f"(r[8]), "=f"(r[9]), "=f"(r[10]), "=f"(r[11]), "=f"(r[12]), "=f"(r[13]), "=f"(r[14]), "=f"(r[15])
      : "r"(taddr));
}
// Relaxed on purpose: what the counter orders are tensor-memory stores, which each warp has
// completed (tcgen05.wait::st) and fenced (tcgen05.fence::before_thread_sync) before it counts
// itself; an acq_rel atomic costs a MEMBAR.ALL.CTA per tile and warp.
__device__ __forceinline__ uint32_t mm_atom_inc(uint32_t smem_addr) {
  uint32_t old;
  asm volatile("atom.relaxed.cta.shared::cta.add.u32 %0, [%1], 1;" : "=r"(old) : "r"(smem_addr) : "memory");
  return old;
}
__device__ __forceinline__ void mm_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mm_elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0u;
}
// tells the compiler a value is the same in every lane (it then lives in a uniform register)
__device__ __forceinline__ uint32_t mm_uniform(uint32_t v) { return __shfl_sync(0xffffffffu, v, 0); }
// mbarrier wait of a SERVICE warp: backs off between polls so that it does not compete with the
// math warps for issue slots (the plain try_wait loop re-polls every few dozen cycles)
__device__ __forceinline__ void mm_wait_sleepy(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  for (;;) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
    if (ok) break;
    __nanosleep(200);
  }
}
// 1.0f / 0.0f without a predicate + select pair (SASS FSET.BF)
__device__ __forceinline__ float mm_eq_one(float a, float b) {
  float r;
  asm("set.eq.f32.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
  return r;
}

// Pixel accessors of a staged tile, one pixel per thread.
template <int kFmt>
__device__ __forceinline__ void mm_load_px(const unsigned char* tile, int p, float& r, float& g, float& b) {
  if constexpr (kFmt == kPxF32) {
    const float* t = reinterpret_cast<const float*>(tile) + 3 * p;
    r = t[0]; g = t[1]; b = t[2];
  } else if constexpr (kFmt == kPxU8) {
    const unsigned char* t = tile + 3 * p;
    r = px_to_float<kPxU8>(t[0]); g = px_to_float<kPxU8>(t[1]); b = px_to_float<kPxU8>(t[2]);
  } else {
    const unsigned short* t = reinterpret_cast<const unsigned short*>(tile) + 3 * p;
    r = px_to_float<kPxU16>(t[0]); g = px_to_float<kPxU16>(t[1]); b = px_to_float<kPxU16>(t[2]);
  }
}
template <int kFmt>
__device__ __forceinline__ void mm_store_px(unsigned char* tile, int p, float r, float g, float b) {
  if constexpr (kFmt == kPxF32) {
    float* t = reinterpret_cast<float*>(tile) + 3 * p;
    t[0] = r; t[1] = g; t[2] = b;
  } else {
    static_assert(kFmt == kPxU8, "results leave as float32 or uint8");
    unsigned char* t = tile + 3 * p;
    t[0] = static_cast<unsigned char>(float_to_u8(r));
    t[1] = static_cast<unsigned char>(float_to_u8(g));
    t[2] = static_cast<unsigned char>(float_to_u8(b));
  }
}

// A tile of a segment's list: start pixel inside the segment, pixel count, padded cell index
// (x0 + 1; the tile's B' window starts at that padded cell).
__host__ __device__ constexpr uint32_t mm_pack_tile(int start, int n, int cellp) {
  return static_cast<uint32_t>(start) | (static_cast<uint32_t>(n) << 16) | (static_cast<uint32_t>(cellp) << 24);
}

template <class GuideFn, int kIn, int kOut, int kNSplit, int kKSteps>
__global__ void __launch_bounds__(kMmThreads, 1)
slice_apply_rows_mma_kernel(const MmArgs args, const GuideFn guide_fn) {
  extern __shared__ __align__(1024) unsigned char smem[];
  constexpr bool kGuideIn = GuideFn::kFromInput;
  constexpr int kK = 8 * kKSteps;            // depth rows of the product
  constexpr int kKC = kK / 4;                // 16-byte K chunks
  constexpr uint32_t kSbo = kKC * 128u;      // bytes between groups of eight n
  constexpr uint32_t kCellBytes = 3u * kSbo; // a cell = 24 n = three groups
  const SliceGeom& g = args.g;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // ---- shared-memory map -------------------------------------------------------------------
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);     // [4]  pixel segment landed
  uint64_t* done = full + kMmMaxStages;                    // [4]  every math warp is through it
  uint64_t* raw_full = done + kMmMaxStages;                // [2]  grid row landed
  uint64_t* b_full = raw_full + 2;                         // [2]  B' of a row ready
  uint64_t* b_free = b_full + 2;                           // [2]  B' of a row no longer read
  uint64_t* d_ready = b_free + 2;                          // [8]  a tile's MMAs complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 192);
  uint32_t* a_cnt = reinterpret_cast<uint32_t*>(smem + 200);   // [8]  warps done with a tile's A rows
  int* bnd = reinterpret_cast<int*>(smem + args.off_tab);      // [gw + 2] run boundaries
  int* seg_nt = bnd + 40;                                      // [nseg]   tiles per segment
  uint32_t* tiles = reinterpret_cast<uint32_t*>(seg_nt + 24);  // [nseg][max_tiles]
  float* raw = reinterpret_cast<float*>(smem + args.off_raw);  // two grid rows
  unsigned char* bt = smem + args.off_b;                       // [2 rows][kNSplit][bsplit_bytes]
  unsigned char* stage_base = smem + args.off_stage;

  const long long total_items = static_cast<long long>(g.B) * g.rows * args.nseg;
  const long long i_begin = total_items * blockIdx.x / gridDim.x;
  const long long i_end = total_items * (blockIdx.x + 1) / gridDim.x;
  if (i_end <= i_begin) return;
  const long long r_begin = i_begin / args.nseg, r_end = (i_end - 1) / args.nseg + 1;
  const int seg_first = static_cast<int>(i_begin - r_begin * args.nseg);
  const int seg_last = static_cast<int>((i_end - 1) - (r_end - 1) * args.nseg);   // inclusive
  auto row_seg0 = [&](long long row) { return row == r_begin ? seg_first : 0; };
  auto row_seg1 = [&](long long row) { return row == r_end - 1 ? seg_last + 1 : args.nseg; };

  // ---- start-up ---------------------------------------------------------------------------
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(static_cast<uint32_t>(kMmTmemCols)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 32) {
    for (int s = 0; s < kMmMaxStages; ++s) { mbar_init(&full[s], 1); mbar_init(&done[s], kMmMathWarps); }
    for (int i = 0; i < 2; ++i) { mbar_init(&raw_full[i], 1); mbar_init(&b_full[i], kMmSlabWarps); mbar_init(&b_free[i], kMmMathWarps); }
    for (int i = 0; i < kMmWgs; ++i) { mbar_init(&d_ready[i], 1); a_cnt[i] = 0u; }
    fence_mbar_init();
  }
  // run boundaries: bnd[k] = first pixel whose lower x cell is >= k - 1 (k = 0: 0; k = gw + 1: W),
  // found with the kernels' own coordinate arithmetic (spatial_axis is monotone in x)
  if (tid >= 64 && tid < 64 + g.gw + 2) {
    const int k = tid - 64;
    int lo = 0, hi = g.W;
    if (k == 0) hi = 0;
    if (k == g.gw + 1) lo = g.W;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (spatial_axis(mid, g.scale_x).i0 >= k - 1) hi = mid; else lo = mid + 1;
    }
    bnd[k] = lo;
  }
  __syncthreads();
  // tile lists: (run ^ segment) cut into pieces of <= 128 pixels
  if (tid >= 64 && tid < 64 + args.nseg) {
    const int sg = tid - 64;
    const int s_lo = sg * args.seg_px, s_hi = min(g.W, s_lo + args.seg_px);
    uint32_t* tl = tiles + sg * args.max_tiles;
    int nt = 0;
    for (int k = 0; k <= g.gw; ++k) {
      int lo = max(bnd[k], s_lo);
      const int hi = min(bnd[k + 1], s_hi);
      while (lo < hi) {
        const int n = min(kMmTile, hi - lo);
        tl[nt++] = mm_pack_tile(lo - s_lo, n, k);
        lo += n;
      }
    }
    if (nt & 1) tl[nt] = mm_pack_tile(0, 0, 0);   // rounds take tiles in pairs: an empty partner
    seg_nt[sg] = nt;
  }
  mm_fence_before();
  __syncthreads();
  mm_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int NS = args.stages;

  if (warp == kMmMathWarps) {
    // =============================== issuer warp =============================================
    if (lane == 0) {
      long long l_row = r_begin;
      int l_seg = seg_first, l_s = 0;
      auto issue_next_load = [&]() {
        if (l_row >= r_end) return;
        const int x0 = l_seg * args.seg_px;
        const int npx = min(args.seg_px, g.W - x0);
        unsigned char* st = stage_base + static_cast<size_t>(l_s) * args.stage_bytes;
        const size_t pix = static_cast<size_t>(l_row) * g.W + x0;
        const uint32_t in_bytes = static_cast<uint32_t>(npx) * args.in_bpp;
        mbar_expect_tx(&full[l_s], in_bytes + (kGuideIn ? static_cast<uint32_t>(npx) * 4u : 0u));
        tma_load_1d(st, args.input + pix * args.in_bpp, in_bytes, &full[l_s]);
        if (kGuideIn) tma_load_1d(st + args.off_guide, args.guide + pix, static_cast<uint32_t>(npx) * 4u, &full[l_s]);
        if (++l_s == NS) l_s = 0;
        if (++l_seg >= row_seg1(l_row)) { l_seg = 0; ++l_row; }
      };
      for (int i = 0; i < NS - 1; ++i) issue_next_load();
      int s = 0;
      uint32_t ph = 0;
      for (long long row = r_begin; row < r_end; ++row) {
        const int sg1 = row_seg1(row);
        for (int sg = row_seg0(row); sg < sg1; ++sg) {
          mm_wait_sleepy(&done[s], ph);   // every math warp has written (and proxy-fenced) its results
          const int x0 = sg * args.seg_px;
          const int npx = min(args.seg_px, g.W - x0);
          unsigned char* st = stage_base + static_cast<size_t>(s) * args.stage_bytes;
          const size_t pix = static_cast<size_t>(row) * g.W + x0;
          tma_store_1d(args.out + pix * args.out_bpp, st + args.off_out, static_cast<uint32_t>(npx) * args.out_bpp);
          tma_store_commit();
          if (l_row < r_end) {
            tma_store_wait_read<1>();   // the stage stored one item ago is free again
            issue_next_load();
          }
          if (++s == NS) { s = 0; ph ^= 1u; }
        }
      }
      tma_store_wait_all<0>();
    }
  } else if (warp > kMmMathWarps) {
    // ================================ slab warps =============================================
    // Both warps walk the rows together (a named barrier keeps the shared grid-row slots
    // consistent); slab warp 0 issues the grid-row loads, each warp builds half of B'.
    const int sw = warp - kMmMathWarps - 1;
    // raw[slot] holds grid row key (b * gh + gy); -1 = empty
    int key0 = -1, key1 = -1;          // grid row held by raw slot 0 / 1
    uint32_t rpar0 = 0u, rpar1 = 0u;   // parity of the next completion of raw_full[0 / 1]
    const uint32_t raw_bytes = static_cast<uint32_t>(args.row_floats) * 4u;
    const int tasks = g.gw * kGc * kKC;      // (cell, coefficient, K chunk)
    for (long long row = r_begin; row < r_end; ++row) {
      const int rowk = static_cast<int>(row - r_begin), rb = rowk & 1;
      const int b = static_cast<int>(row / g.rows);
      const int y = g.y_off + static_cast<int>(row - static_cast<long long>(b) * g.rows);
      const Axis ay = spatial_axis(y, g.scale_y);
      const int k0 = b * g.gh + clampi(ay.i0, 0, g.gh - 1);
      const int k1 = b * g.gh + clampi(ay.i0 + 1, 0, g.gh - 1);
      // which slots hold k0 / k1; load what is missing into the slot the other does not use
      // nobody still reads the grid-row slot a load below may overwrite
      asm volatile("bar.sync 1, %0;" ::"n"(kMmSlabWarps * 32) : "memory");
      int s0 = (key0 == k0) ? 0 : ((key1 == k0) ? 1 : -1);
      int s1 = (key0 == k1) ? 0 : ((key1 == k1) ? 1 : -1);
      auto fetch = [&](int slot, int k) {   // grid row k -> raw slot (warp-uniform)
        if (lane == 0 && sw == 0) {
          mbar_expect_tx(&raw_full[slot], raw_bytes);
          tma_load_1d(raw + static_cast<size_t>(slot) * args.row_floats,
                      args.grid + static_cast<size_t>(k) * args.row_floats, raw_bytes, &raw_full[slot]);
        }
        if (slot == 0) { key0 = k; mbar_wait(&raw_full[0], rpar0); rpar0 ^= 1u; }
        else { key1 = k; mbar_wait(&raw_full[1], rpar1); rpar1 ^= 1u; }
      };
      if (s0 < 0) { s0 = (s1 == 0) ? 1 : 0; fetch(s0, k0); if (k1 == k0) s1 = s0; }
      if (s1 < 0) { s1 = s0 ^ 1; fetch(s1, k1); }
      // the row buffer is free once every math warp is through row - 2
      if (rowk >= 2) mm_wait_sleepy(&b_free[rb], static_cast<uint32_t>((rowk >> 1) - 1) & 1u);
      const float wy1 = ay.f, wy0 = 1.0f - ay.f;
      const float* g0 = raw + static_cast<size_t>(s0) * args.row_floats;
      const float* g1 = raw + static_cast<size_t>(s1) * args.row_floats;
      unsigned char* brow = bt + static_cast<size_t>(rb) * kNSplit * args.bsplit_bytes;
      for (int e = sw * 32 + lane; e < tasks; e += 32 * kMmSlabWarps) {
        const int kc = e % kKC, ec = e / kKC;
        const int cell = ec / kGc, c = ec - cell * kGc;
        // depth cells 4 kc .. 4 kc + 4 of this (cell, coefficient), y-blended (lerp4's order)
        float v[5];
#pragma unroll
        for (int i = 0; i < 5; ++i) {
          const int z = min(4 * kc + i, g.gd - 1);
          const int o = (cell * g.gd + z) * kGc + c;
          v[i] = fmaf(wy1, g1[o], wy0 * g0[o]);
        }
        // rows k >= gd of the product are never selected; they must still be finite
        float h0[4], h1[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const bool live = 4 * kc + i < g.gd;
          h0[i] = live ? v[i] : 0.0f;
          h1[i] = live ? v[i + 1] : 0.0f;
        }
        float s_h0[kNSplit][4], s_h1[kNSplit][4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float r0 = h0[i], r1 = h1[i];
#pragma unroll
          for (int sp = 0; sp < kNSplit; ++sp) {
            if (sp == kNSplit - 1 && kNSplit == 3) { s_h0[sp][i] = r0; s_h1[sp][i] = r1; break; }
            const float t0 = __uint_as_float(__float_as_uint(r0) & 0xffffe000u);
            const float t1 = __uint_as_float(__float_as_uint(r1) & 0xffffe000u);
            s_h0[sp][i] = t0; s_h1[sp][i] = t1;
            r0 -= t0; r1 -= t1;
          }
        }
        // padded cell index cell + 1; the border cells are stored twice (clamped neighbours)
        const int ndup = 1 + (cell == 0 ? 1 : 0) + (cell == g.gw - 1 ? 1 : 0);
        for (int d = 0; d < ndup; ++d) {
          int cellp = cell + 1;
          if (d >= 1) cellp = (cell == 0 && d == 1) ? 0 : g.gw + 1;
          const int n0 = cellp * 24 + c, n1 = n0 + kGc;
          const uint32_t o0 = static_cast<uint32_t>(n0 >> 3) * kSbo + static_cast<uint32_t>(kc) * 128u + static_cast<uint32_t>(n0 & 7) * 16u;
          const uint32_t o1 = static_cast<uint32_t>(n1 >> 3) * kSbo + static_cast<uint32_t>(kc) * 128u + static_cast<uint32_t>(n1 & 7) * 16u;
#pragma unroll
          for (int sp = 0; sp < kNSplit; ++sp) {
            unsigned char* bs = brow + static_cast<size_t>(sp) * args.bsplit_bytes;
            *reinterpret_cast<float4*>(bs + o0) = make_float4(s_h0[sp][0], s_h0[sp][1], s_h0[sp][2], s_h0[sp][3]);
            *reinterpret_cast<float4*>(bs + o1) = make_float4(s_h1[sp][0], s_h1[sp][1], s_h1[sp][2], s_h1[sp][3]);
          }
        }
      }
      fence_proxy_async_smem();   // generic writes -> the tensor core's (async-proxy) operand reads
      __syncwarp();
      if (lane == 0) mm_arrive(&b_full[rb]);
    }
  } else {
    // ================================ math warpgroups ========================================
    const int wg = static_cast<int>(mm_uniform(static_cast<uint32_t>(warp >> 2))), t = tid & 127;
    const uint32_t lane_sel = static_cast<uint32_t>((warp & 3) * 32) << 16;   // this warp's TMEM lanes
    const uint32_t tm_wg = tmem_base + static_cast<uint32_t>(wg) * kMmWgCols;
    const uint32_t bt_addr = smem_u32(bt);
    const float gd_f = static_cast<float>(g.gd);

    const uint32_t stage_addr = smem_u32(stage_base);
    const uint32_t tiles_addr = smem_u32(tiles);
    const uint32_t dbar = smem_u32(&d_ready[wg]);
    const uint32_t acnt = smem_u32(&a_cnt[wg]);
    const uint32_t done_addr = smem_u32(done);
    const uint32_t bfree_addr = smem_u32(b_free);
    // tensor memory of this warpgroup (128 columns): D of the round's two tiles at [0, 48) and
    // [48, 96); A operands at [96, 128): kKSteps == 1 -- two BUFFERS of two tiles x 8 columns, so
    // that the one-hot rows of round r + 2 are written while the MMAs of round r + 1 run;
    // kKSteps == 2 -- one buffer of two tiles x 16 columns (the set-up then waits for the MMAs).
    constexpr bool kPipeA = (kKSteps == 1);
    const uint32_t tD0 = tm_wg, tD1 = tm_wg + 48u, tA0 = tm_wg + 96u;
    uint32_t dpar = 0u;

    // One pixel's depth axis: guide -> lower depth cell (as a float, for the one-hot compare) and
    // the two smoothed weights (both on the first row when the two cells clamp to cell 0, whose
    // row of B' is (s0, s1)).
    auto depth_axis = [&](float gv, float& zf, float& wz0, float& wz1) {
      const float tz = __fsub_rn(__fmul_rn(gv, gd_f), 0.5f);
      const int iz = __float2int_rd(tz);
      const float fz = tz - static_cast<float>(iz);
      smoothed_weights(fz, wz0, wz1);
      if (iz < 0) { wz0 += wz1; wz1 = 0.0f; }
      zf = static_cast<float>(clampi(iz, 0, g.gd - 1));
    };
    auto put_onehot = [&](uint32_t ta, float zf) {
#pragma unroll
      for (int ks = 0; ks < kKSteps; ++ks) {
        float a[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) a[k] = mm_eq_one(zf, static_cast<float>(8 * ks + k));
        mm_st8(ta + 8 * ks + lane_sel, a);
      }
    };
    auto issue_mmas = [&](uint32_t td, uint32_t ta, uint32_t b0) {
#pragma unroll
      for (int sp = 0; sp < kNSplit; ++sp) {
#pragma unroll
        for (int ks = 0; ks < kKSteps; ++ks) {
          const uint64_t bd = mm_kmajor_desc(b0 + static_cast<uint32_t>(sp) * static_cast<uint32_t>(args.bsplit_bytes) + ks * 256u, 128, kSbo);
          mm_mma(td, ta + 8 * ks, bd, (sp | ks) ? 1u : 0u);
        }
      }
    };
    // read a pixel's 4 corner vectors back from tensor memory, blend, apply, store in place
    auto finish_px = [&](uint32_t td, unsigned char* st, int x, uint32_t tw, float wz0, float wz1) {
      float d[48];
      const uint32_t taddr = td + lane_sel;
      mm_ld16(taddr, d);
      mm_ld16(taddr + 16, d + 16);
      mm_ld16(taddr + 32, d + 32);
      const int n = static_cast<int>((tw >> 16) & 0xffu), cellp = static_cast<int>(tw >> 24);
      const bool valid = t < n;
      const int p = static_cast<int>(tw & 0xffffu) + (valid ? t : 0);
      float r, gg, bb;
      mm_load_px<kIn>(st, p, r, gg, bb);
      // x fraction with the reference's roundings; floor(tx) is the tile's x0 by construction
      const float tx = __fsub_rn(__fmul_rn(__fadd_rn(static_cast<float>(x + p), 0.5f), g.scale_x), 0.5f);
      const float wx1 = tx - static_cast<float>(cellp - 1), wx0 = 1.0f - wx1;
      const float w00 = wx0 * wz0, w01 = wx0 * wz1, w10 = wx1 * wz0, w11 = wx1 * wz1;
      const unsigned long long W00 = pack2(w00, w00), W01 = pack2(w01, w01);
      const unsigned long long W10 = pack2(w10, w10), W11 = pack2(w11, w11);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      float v[12];
#pragma unroll
      for (int c = 0; c < 12; c += 2) {
        const unsigned long long acc =
            fma2(W11, pack2(d[36 + c], d[37 + c]),
                 fma2(W10, pack2(d[24 + c], d[25 + c]),
                      fma2(W01, pack2(d[12 + c], d[13 + c]), mul2(W00, pack2(d[c], d[c + 1])))));
        unpack2(acc, v[c], v[c + 1]);
      }
      const float o_r = fmaf(v[2], bb, fmaf(v[1], gg, fmaf(v[0], r, v[3])));
      const float o_g = fmaf(v[6], bb, fmaf(v[5], gg, fmaf(v[4], r, v[7])));
      const float o_b = fmaf(v[10], bb, fmaf(v[9], gg, fmaf(v[8], r, v[11])));
      if (valid) mm_store_px<kOut>(st + args.off_out, p, o_r, o_g, o_b);
    };

    // A ROUND = a pair of tiles (one pixel of each per thread).  What is kept of it between its
    // set-up (A written), the issue of its MMAs and its epilogue:
    struct Round {
      float wzA0, wzA1, wzB0, wzB1;   // per thread
      uint32_t twA, twB;              // the tile words
      unsigned char* st;              // stage of its segment
      int xs;                         // first pixel of the segment
      uint32_t b_row;                 // B' of its row
      uint32_t abuf;                  // TMEM column of its A operands
      uint32_t done_bar, free_bar;    // != 0: arrive there after the epilogue
      int segc, rowk;                 // segment / row counters (look-ahead bounds)
    };
    Round cur, nxt;                   // cur: MMAs issued; nxt: A written
    bool have_cur = false, have_nxt = false;

    // ---- cursor over the tile pairs of this CTA's segments; every fourth pair is this warpgroup's.
    // A warp observes full[] / b_full[] of every segment / row it passes, also those it has no
    // tile in: that bounds how far it can run ahead of the others (it must never arrive on
    // done[] / b_free[] for a later use of the same barrier).
    long long c_row = r_begin;
    int c_sg = seg_first, c_s = 0, c_gp = 0, c_segc = 0;
    uint32_t c_fph = 0;
    bool c_row_open = false, c_seg_open = false;   // b_full[] / full[] of the current row / segment observed
    int c_np = 0, c_j = 0;
    int nfetch = 0;
    // marks on the most recently fetched round that is still pending (nxt, else cur)
    auto newest = [&]() -> Round* { return have_nxt ? &nxt : (have_cur ? &cur : nullptr); };

    for (;;) {
      // ---- 1. epilogue of the round whose MMAs were issued one iteration ago ------------------
      if (have_cur) {
        mbar_wait_addr(dbar, dpar);
        dpar ^= 1u;
        mm_fence_after();
        finish_px(tD0, cur.st, cur.xs, cur.twA, cur.wzA0, cur.wzA1);
        finish_px(tD1, cur.st, cur.xs, cur.twB, cur.wzB0, cur.wzB1);
        mm_fence_before();   // D is read: order it before the MMAs of the next round
        if (cur.done_bar != 0u) {
          fence_proxy_async_smem();   // results (generic writes) -> the issuer's bulk store
          __syncwarp();
          if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(cur.done_bar) : "memory");
        }
        if (cur.free_bar != 0u) {
          __syncwarp();
          if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(cur.free_bar) : "memory");
        }
        have_cur = false;
      }
      // ---- 2. MMAs of the next round: its A rows are written, D is free ----------------------
      if (have_nxt) {
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        mm_fence_before();
        uint32_t old = 0u;
        if (lane == 0) old = mm_atom_inc(acnt);
        old = mm_uniform(old);
        if ((old & 3u) == 3u) {   // the last of the four warps: a warp-uniform branch
          mm_fence_after();
          const uint32_t bA = mm_uniform(nxt.b_row + (nxt.twA >> 24) * kCellBytes);
          const uint32_t bB = mm_uniform(nxt.b_row + (nxt.twB >> 24) * kCellBytes);
          const uint32_t ab = mm_uniform(nxt.abuf);
          if (mm_elect_one()) {
            issue_mmas(tD0, ab, bA);
            issue_mmas(tD1, ab + 8u * kKSteps, bB);
            mm_commit(dbar);
          }
          __syncwarp();
        }
        cur = nxt;
        have_cur = true;
        have_nxt = false;
      }
      // ---- 3. fetch + set-up of one more round (overlaps the MMAs just issued) ---------------
      bool fetched = false;
      while (c_row < r_end) {
        const int rowk = static_cast<int>(c_row - r_begin), rb = rowk & 1;
        // look-ahead bounds: B' has two row buffers, the ring NS stages
        if (have_cur && (rowk - cur.rowk > 1 || c_segc - cur.segc > NS - 2)) break;
        if (!c_row_open) { mbar_wait(&b_full[rb], static_cast<uint32_t>(rowk >> 1) & 1u); c_row_open = true; }
        if (!c_seg_open) {
          mbar_wait(&full[c_s], c_fph);
          c_seg_open = true;
          c_np = (seg_nt[c_sg] + 1) >> 1;
          c_j = (wg - c_gp) & 3;
          if (c_j >= c_np) { __syncwarp(); if (lane == 0) mm_arrive(&done[c_s]); }   // no pair of ours here
        }
        if (c_j < c_np) {
          // ---- set-up: guide -> depth cell, weights, one-hot rows of A ------------------------
          unsigned char* st = stage_base + static_cast<size_t>(c_s) * args.stage_bytes;
          const uint32_t st_a = stage_addr + static_cast<uint32_t>(c_s) * static_cast<uint32_t>(args.stage_bytes);
          uint32_t twA, twB;
          asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(twA), "=r"(twB)
                       : "r"(tiles_addr + 4u * static_cast<uint32_t>(c_sg * args.max_tiles) + 8u * c_j));
          const int nA = static_cast<int>((twA >> 16) & 0xffu), nB = static_cast<int>((twB >> 16) & 0xffu);
          const int pA = static_cast<int>(twA & 0xffffu) + (t < nA ? t : 0);
          const int pB = static_cast<int>(twB & 0xffffu) + (t < nB ? t : 0);
          const int xs = c_sg * args.seg_px;
          float gvA, gvB;
          if constexpr (kGuideIn) {
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(gvA) : "r"(st_a + args.off_guide + 4u * pA));
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(gvB) : "r"(st_a + args.off_guide + 4u * pB));
          } else {
            float rA, gA, bA, rB, gB, bB;
            mm_load_px<kIn>(st, pA, rA, gA, bA);
            mm_load_px<kIn>(st, pB, rB, gB, bB);
            gvA = guide_fn(rA, gA, bA);
            gvB = guide_fn(rB, gB, bB);
            if (args.guide_out != nullptr) {
              float* go = args.guide_out + static_cast<size_t>(c_row) * g.W + xs;
              if (t < nA) go[pA] = gvA;
              if (t < nB) go[pB] = gvB;
            }
          }
          float zfA, zfB;
          depth_axis(gvA, zfA, nxt.wzA0, nxt.wzA1);
          depth_axis(gvB, zfB, nxt.wzB0, nxt.wzB1);
          const uint32_t abuf = tA0 + (kPipeA ? (static_cast<uint32_t>(nfetch) & 1u) * 16u : 0u);
          if constexpr (!kPipeA) {
            // one A buffer: the MMAs of `cur` (issued in step 2) read it -- wait for them first
            if (have_cur) mbar_wait_addr(dbar, dpar);
          }
          put_onehot(abuf, zfA);
          put_onehot(abuf + 8u * kKSteps, zfB);
          ++nfetch;
          nxt.twA = twA; nxt.twB = twB;
          nxt.st = st; nxt.xs = xs;
          nxt.b_row = bt_addr + static_cast<uint32_t>(rb) * kNSplit * static_cast<uint32_t>(args.bsplit_bytes);
          nxt.abuf = abuf;
          nxt.done_bar = 0u; nxt.free_bar = 0u;
          nxt.segc = c_segc; nxt.rowk = rowk;
          have_nxt = true;
          fetched = true;
          c_j += 4;
          if (c_j < c_np) break;       // more pairs of ours in this segment: come back next iteration
        }
        // ---- leave the segment (all our pairs of it are fetched) -------------------------------
        {
          Round* nw = newest();
          if (nw != nullptr && nw->segc == c_segc) nw->done_bar = done_addr + 8u * c_s;
          // (a segment without a pair of ours was acknowledged when it was opened)
        }
        c_gp = (c_gp + c_np) & 3;
        ++c_segc;
        if (++c_s == NS) { c_s = 0; c_fph ^= 1u; }
        c_seg_open = false;
        if (++c_sg >= row_seg1(c_row)) {
          Round* nw = newest();
          if (nw != nullptr && nw->rowk == rowk) nw->free_bar = bfree_addr + 8u * rb;
          else { __syncwarp(); if (lane == 0) mm_arrive(&b_free[rb]); }
          c_sg = 0;
          ++c_row;
          c_row_open = false;
        }
        if (fetched) break;
      }
      if (!have_cur && !have_nxt) break;
    }
  }
  mm_fence_before();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"(static_cast<uint32_t>(kMmTmemCols)));
  }
}

// ---- host side ------------------------------------------------------------------------------

static inline int mm_round_up(int v, int m) { return (v + m - 1) / m * m; }

// Plans the launch; false when the shapes do not suit this form (the callers fall back to the
// shared-memory row kernel / the generic kernel).
bool make_mma_plan(const SliceGeom& g, int in_fmt, int out_fmt, bool guide_from_input, int nsplit,
                   int max_smem, int sms, MmArgs* out) {
  if (g.gd > 16 || g.gw > 32 || g.gw < 1) return false;
  const int in_bpp = 3 * px_bytes_per_channel(in_fmt), out_bpp = 3 * px_bytes_per_channel(out_fmt);
  const int gran = std::max(4, std::max(16 / std::__gcd(16, in_bpp), 16 / std::__gcd(16, out_bpp)));
  if (g.W < gran || (g.W % gran) != 0) return false;
  // runs of one x cell must fill a useful part of a 128-pixel tile
  if (static_cast<long long>(g.W) < 48LL * g.gw) return false;
  MmArgs a = {};
  a.g = g;
  a.in_bpp = in_bpp;
  a.out_bpp = out_bpp;
  a.row_floats = g.gw * g.gd * kGc;
  const int ksteps = g.gd > 8 ? 2 : 1;
  a.bsplit_bytes = (g.gw + 2) * 3 * (ksteps * 2) * 128;
  const long long rows_total = static_cast<long long>(g.B) * g.rows;
  const int ctas = static_cast<int>(std::min<long long>(rows_total, sms));
  // segments: <= 2048 pixels; more (shorter) segments while a CTA would get fewer than 16 items
  int nseg = (g.W + 2047) / 2048;
  while (rows_total * nseg < 16LL * ctas && g.W / (nseg * 2) >= 512 && nseg * 2 <= 16) nseg *= 2;
  a.seg_px = mm_round_up((g.W + nseg - 1) / nseg, std::max(gran, 16));
  a.nseg = (g.W + a.seg_px - 1) / a.seg_px;
  if (a.nseg > 24 || a.seg_px > 65535) return false;
  a.max_tiles = (a.seg_px / kMmTile + g.gw + 5) & ~1;
  a.off_guide = mm_round_up(a.seg_px * in_bpp, 16);
  const int after_in = guide_from_input ? a.off_guide + a.seg_px * 4 : a.off_guide;
  a.off_out = (out_bpp == in_bpp) ? 0 : mm_round_up(after_in, 16);
  a.stage_bytes = mm_round_up(a.off_out ? a.off_out + a.seg_px * out_bpp : after_in, 128);
  a.off_tab = 256;
  a.off_raw = mm_round_up(a.off_tab + (40 + 24 + a.nseg * a.max_tiles) * 4, 128);
  a.off_b = mm_round_up(a.off_raw + 2 * a.row_floats * 4, 1024);
  a.off_stage = mm_round_up(a.off_b + 2 * nsplit * a.bsplit_bytes, 128);
  a.stages = 0;
  // >= 3 stages: a math warp sets up the first tile of segment k + 1 before the epilogue of its
  // last tile of segment k, and the load of k + 1 must not wait for done[k]
  for (int ns = kMmMaxStages; ns >= 3; --ns)
    if (a.off_stage + ns * a.stage_bytes <= max_smem) { a.stages = ns; break; }
  if (a.stages == 0) return false;
  a.smem_bytes = a.off_stage + a.stages * a.stage_bytes;
  a.ctas = static_cast<int>(std::min<long long>(rows_total * a.nseg, sms));
  *out = a;
  return true;
}

template <class GuideFn, int kIn, int kOut, int kNSplit>
static int launch_mma_k(const MmArgs& a, const GuideFn& fn, cudaStream_t stream) {
  if (a.g.gd > 8) {
    auto kern = slice_apply_rows_mma_kernel<GuideFn, kIn, kOut, kNSplit, 2>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, a.smem_bytes);
    if (e != cudaSuccess) return static_cast<int>(e);
    kern<<<a.ctas, kMmThreads, a.smem_bytes, stream>>>(a, fn);
  } else {
    auto kern = slice_apply_rows_mma_kernel<GuideFn, kIn, kOut, kNSplit, 1>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, a.smem_bytes);
    if (e != cudaSuccess) return static_cast<int>(e);
    kern<<<a.ctas, kMmThreads, a.smem_bytes, stream>>>(a, fn);
  }
  return static_cast<int>(cudaGetLastError());
}

// Op-API form: float32 pixels, guide as an input tensor.
int launch_slice_apply_mma(const float* grid, const float* guide, const float* input, float* out,
                           const SliceGeom& g, int nsplit, int max_smem, int sms, cudaStream_t stream) {
  MmArgs a;
  if (!make_mma_plan(g, kPxF32, kPxF32, true, nsplit, max_smem, sms, &a)) return HDRNET_E_UNSUPPORTED;
  a.grid = grid; a.guide = guide; a.guide_out = nullptr;
  a.input = reinterpret_cast<const unsigned char*>(input);
  a.out = reinterpret_cast<unsigned char*>(out);
  if (nsplit == 3) return launch_mma_k<GuideFromInput, kPxF32, kPxF32, 3>(a, GuideFromInput{}, stream);
  return launch_mma_k<GuideFromInput, kPxF32, kPxF32, 2>(a, GuideFromInput{}, stream);
}

}  // namespace hdrnet_b200
