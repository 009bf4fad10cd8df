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
constexpr int kMmThreads = (kMmMathWarps + 2 + kMmSlabWarps) * 32;   // + issuer, scheduler, slab warps = 640
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
  int off_wx, off_raw, off_b, off_stage, smem_bytes, ctas;
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

// ---- shared-memory header of the kernel (byte offsets) -------------------------------------------
constexpr uint32_t kOffFull = 0;         // [4]   pixel segment landed (TMA)
constexpr uint32_t kOffDone = 32;        // [4]   every math warp is through the segment
constexpr uint32_t kOffRawFull = 64;     // [2]   grid row landed
constexpr uint32_t kOffBFull = 80;       // [2]   B' of a row built
constexpr uint32_t kOffBFree = 96;       // [2]   B' of a row no longer read by any MMA
constexpr uint32_t kOffDReady = 112;     // [4]   a warpgroup's MMAs complete
constexpr uint32_t kOffDescReady = 144;  // [4][4] round descriptor published
constexpr uint32_t kOffDescFree = 272;   // [4][4] round descriptor consumed
constexpr uint32_t kOffTmem = 400;
constexpr uint32_t kOffACnt = 416;       // [4]   warps done with a round's A rows
constexpr uint32_t kOffDesc = 512;       // [4][4][8 words] round descriptors
constexpr uint32_t kOffTab = 1024;       // run boundaries, tile lists, x weights (MmArgs::off_tab)

// A ROUND descriptor (8 words), written by the scheduler warp, read by a warpgroup's math warps:
//   0 stage address | 1 tile A (start | n << 16) | 2 tile B | 3 B' window of A | 4 B' window of B |
//   5 first pixel of the segment in the row | 6 flags | 7 first pixel of the segment in the image
constexpr uint32_t kDescEnd = 1u, kDescDone = 2u, kDescFree = 16u;   // flags: | stage << 2 | rb << 5

__device__ __forceinline__ uint32_t lds_u32(uint32_t a) {
  uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v;
}
__device__ __forceinline__ float lds_f32(uint32_t a) {
  float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); return v;
}
__device__ __forceinline__ uint4 lds_v4(uint32_t a) {
  uint4 v; asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts_v4(uint32_t a, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
__device__ __forceinline__ void mm_arrive_a(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mm_arrive_n(uint32_t bar, uint32_t n) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(n) : "memory");
}
__device__ __forceinline__ bool mm_test(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0u;
}
__device__ __forceinline__ void mm_wait_sleepy_a(uint32_t bar, uint32_t parity) {
  while (!mm_test(bar, parity)) __nanosleep(200);
}
// one pixel of a staged tile, by shared-memory address
template <int kFmt>
__device__ __forceinline__ void mm_load_px_a(uint32_t tile, int p, float& r, float& g, float& b) {
  if constexpr (kFmt == kPxF32) {
    const uint32_t a = tile + 12u * p;
    r = lds_f32(a); g = lds_f32(a + 4); b = lds_f32(a + 8);
  } else if constexpr (kFmt == kPxU8) {
    const uint32_t a = tile + 3u * p;
    uint32_t x, y, z;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(x) : "r"(a));
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(y) : "r"(a + 1));
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(z) : "r"(a + 2));
    r = px_to_float<kPxU8>(x); g = px_to_float<kPxU8>(y); b = px_to_float<kPxU8>(z);
  } else {
    const uint32_t a = tile + 6u * p;
    uint32_t x, y, z;
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(x) : "r"(a));
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(y) : "r"(a + 2));
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(z) : "r"(a + 4));
    r = px_to_float<kPxU16>(x); g = px_to_float<kPxU16>(y); b = px_to_float<kPxU16>(z);
  }
}
template <int kFmt>
__device__ __forceinline__ void mm_store_px_a(uint32_t tile, int p, float r, float g, float b) {
  if constexpr (kFmt == kPxF32) {
    const uint32_t a = tile + 12u * p;
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(r) : "memory");
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(a + 4), "f"(g) : "memory");
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(a + 8), "f"(b) : "memory");
  } else {
    static_assert(kFmt == kPxU8, "results leave as float32 or uint8");
    const uint32_t a = tile + 3u * p;
    asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"(float_to_u8(r)) : "memory");
    asm volatile("st.shared.u8 [%0], %1;" ::"r"(a + 1), "r"(float_to_u8(g)) : "memory");
    asm volatile("st.shared.u8 [%0], %1;" ::"r"(a + 2), "r"(float_to_u8(b)) : "memory");
  }
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
  const uint32_t sm0 = smem_u32(smem);

  int* bnd = reinterpret_cast<int*>(smem + kOffTab);          // [gw + 2] run boundaries
  int* seg_nt = bnd + 40;                                      // [nseg]   tiles per segment
  uint32_t* tiles = reinterpret_cast<uint32_t*>(seg_nt + 24);  // [nseg][max_tiles]
  float* wx = reinterpret_cast<float*>(smem + args.off_wx);    // [W] x fraction of every pixel column
  float* raw = reinterpret_cast<float*>(smem + args.off_raw);  // two grid rows
  unsigned char* bt = smem + args.off_b;                       // [2 rows][kNSplit][bsplit_bytes]
  const uint32_t bt_addr = sm0 + args.off_b;
  const uint32_t stage_addr = sm0 + args.off_stage;
  const uint32_t tiles_addr = smem_u32(tiles);
  const uint32_t wx_addr = sm0 + args.off_wx;

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
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sm0 + kOffTmem),
                 "r"(static_cast<uint32_t>(kMmTmemCols)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 32) {
    auto init = [&](uint32_t off, int n, uint32_t count) {
      for (int i = 0; i < n; ++i)
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sm0 + off + 8u * i), "r"(count));
    };
    init(kOffFull, kMmMaxStages, 1); init(kOffDone, kMmMaxStages, kMmMathWarps);
    init(kOffRawFull, 2, 1); init(kOffBFull, 2, kMmSlabWarps); init(kOffBFree, 2, kMmMathWarps);
    init(kOffDReady, kMmWgs, 1); init(kOffDescReady, 16, 1); init(kOffDescFree, 16, 4);
    for (int i = 0; i < kMmWgs; ++i) reinterpret_cast<uint32_t*>(smem + kOffACnt)[i] = 0u;
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
  // x fractions: the same for every row (the reference's roundings; wx1 = f, wx0 = 1 - f)
  for (int x = tid; x < g.W; x += kMmThreads) wx[x] = spatial_axis(x, g.scale_x).f;
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
  const uint32_t tmem_base = *reinterpret_cast<uint32_t*>(smem + kOffTmem);
  const int NS = args.stages;

  if (warp == kMmMathWarps) {
    // =============================== issuer warp =============================================
    // one lane issues every bulk copy of pixel segments: TMA ring in, one bulk store per segment out
    if (lane == 0) {
      long long l_row = r_begin;
      int l_seg = seg_first, l_s = 0;
      auto issue_next_load = [&]() {
        if (l_row >= r_end) return;
        const int x0 = l_seg * args.seg_px;
        const int npx = min(args.seg_px, g.W - x0);
        unsigned char* st = smem + args.off_stage + static_cast<size_t>(l_s) * args.stage_bytes;
        const size_t pix = static_cast<size_t>(l_row) * g.W + x0;
        const uint32_t in_bytes = static_cast<uint32_t>(npx) * args.in_bpp;
        uint64_t* fb = reinterpret_cast<uint64_t*>(smem + kOffFull) + l_s;
        mbar_expect_tx(fb, in_bytes + (kGuideIn ? static_cast<uint32_t>(npx) * 4u : 0u));
        tma_load_1d(st, args.input + pix * args.in_bpp, in_bytes, fb);
        if (kGuideIn) tma_load_1d(st + args.off_guide, args.guide + pix, static_cast<uint32_t>(npx) * 4u, fb);
        if (++l_s == NS) l_s = 0;
        if (++l_seg >= row_seg1(l_row)) { l_seg = 0; ++l_row; }
      };
      for (int i = 0; i < NS - 1; ++i) issue_next_load();
      int s = 0;
      uint32_t ph = 0;
      for (long long row = r_begin; row < r_end; ++row) {
        const int sg1 = row_seg1(row);
        for (int sg = row_seg0(row); sg < sg1; ++sg) {
          mm_wait_sleepy_a(sm0 + kOffDone + 8u * s, ph);   // every math warp has written (and proxy-fenced) its results
          const int x0 = sg * args.seg_px;
          const int npx = min(args.seg_px, g.W - x0);
          unsigned char* st = smem + args.off_stage + static_cast<size_t>(s) * args.stage_bytes;
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
  } else if (warp == kMmMathWarps + 1) {
    // ============================== scheduler warp ===========================================
    // One lane walks the CTA's rows / segments / tile pairs in order and publishes ROUND
    // descriptors, round-robin to the four warpgroups (round i belongs to warpgroup i & 3 and is
    // that warpgroup's (i >> 2)-th): the math warps never touch full[] / b_full[] themselves, a
    // descriptor only appears once its segment has landed and its row's B' is built.
    if (lane == 0) {
      int s = 0;
      uint32_t fph = 0u, gp = 0u;
      auto publish = [&](uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3, uint32_t w4, uint32_t w5,
                         uint32_t w6, uint32_t w7, uint32_t wg_, uint32_t n) {
        const uint32_t slot = n & 3u, idx = wg_ * 4u + slot;
        if (n >= 4u) mm_wait_sleepy_a(sm0 + kOffDescFree + 8u * idx, ((n >> 2) - 1u) & 1u);
        const uint32_t da = sm0 + kOffDesc + 32u * idx;
        sts_v4(da, w0, w1, w2, w3);
        sts_v4(da + 16, w4, w5, w6, w7);
        mm_arrive_a(sm0 + kOffDescReady + 8u * idx);   // release: the descriptor is visible to the waiters
      };
      for (long long row = r_begin; row < r_end; ++row) {
        const int rowk = static_cast<int>(row - r_begin), rb = rowk & 1;
        mm_wait_sleepy_a(sm0 + kOffBFull + 8u * rb, static_cast<uint32_t>(rowk >> 1) & 1u);
        const uint32_t b_row = bt_addr + static_cast<uint32_t>(rb) * kNSplit * static_cast<uint32_t>(args.bsplit_bytes);
        const int sg0 = row_seg0(row), sg1 = row_seg1(row);
        int row_left = 0;   // rounds of this row not yet published
        for (int sg = sg0; sg < sg1; ++sg) row_left += (seg_nt[sg] + 1) >> 1;
        const int row_rounds = row_left;
        for (int sg = sg0; sg < sg1; ++sg) {
          mm_wait_sleepy_a(sm0 + kOffFull + 8u * s, fph);
          const int np = (seg_nt[sg] + 1) >> 1;
          const uint32_t st = stage_addr + static_cast<uint32_t>(s) * static_cast<uint32_t>(args.stage_bytes);
          const uint32_t tl = tiles_addr + 4u * static_cast<uint32_t>(sg * args.max_tiles);
          const uint32_t xs = static_cast<uint32_t>(sg * args.seg_px);
          const uint32_t pixbase = static_cast<uint32_t>(static_cast<unsigned long long>(row) * g.W + xs);
          for (int j = 0; j < np; ++j) {
            const uint32_t twA = lds_u32(tl + 8u * j), twB = lds_u32(tl + 8u * j + 4u);
            --row_left;
            uint32_t flags = (static_cast<uint32_t>(s) << 2) | (static_cast<uint32_t>(rb) << 5);
            if (j + 4 >= np) flags |= kDescDone;        // this warpgroup's last round of the segment
            if (row_left < 4) flags |= kDescFree;       // ... and of the row
            publish(st, twA & 0xffffffu, twB & 0xffffffu, b_row + (twA >> 24) * kCellBytes,
                    b_row + (twB >> 24) * kCellBytes, xs, flags, pixbase, gp & 3u, gp >> 2);
            ++gp;
          }
          // warpgroups without a round in this segment / row: acknowledge on their behalf
          if (np < 4) mm_arrive_n(sm0 + kOffDone + 8u * s, 4u * static_cast<uint32_t>(4 - np));
          if (++s == NS) { s = 0; fph ^= 1u; }
        }
        if (row_rounds < 4) mm_arrive_n(sm0 + kOffBFree + 8u * rb, 4u * static_cast<uint32_t>(4 - row_rounds));
      }
      for (uint32_t w = 0; w < 4u; ++w) publish(0u, 0u, 0u, 0u, 0u, 0u, kDescEnd, 0u, w, (gp + 3u - w) >> 2);
    }
  } else if (warp > kMmMathWarps + 1) {
    // ================================ slab warps =============================================
    // Both warps walk the rows together (a named barrier keeps the shared grid-row slots
    // consistent); slab warp 0 issues the grid-row loads, each warp builds half of B'.
    const int sw = warp - kMmMathWarps - 2;
    uint64_t* raw_full = reinterpret_cast<uint64_t*>(smem + kOffRawFull);
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
      // nobody still reads the grid-row slot a load below may overwrite
      asm volatile("bar.sync 1, %0;" ::"n"(kMmSlabWarps * 32) : "memory");
      // which slots hold k0 / k1; load what is missing into the slot the other does not use
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
      if (rowk >= 2) mm_wait_sleepy_a(sm0 + kOffBFree + 8u * rb, static_cast<uint32_t>((rowk >> 1) - 1) & 1u);
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
      if (lane == 0) mm_arrive_a(sm0 + kOffBFull + 8u * rb);
    }
  } else {
    // ================================ math warpgroups ========================================
    const uint32_t wg = mm_uniform(static_cast<uint32_t>(warp >> 2));
    const int t = tid & 127;
    const uint32_t lane_sel = static_cast<uint32_t>((warp & 3) * 32) << 16;   // this warp's TMEM lanes
    const uint32_t tm_wg = tmem_base + wg * kMmWgCols;
    const uint32_t dbar = sm0 + kOffDReady + 8u * wg;
    const uint32_t acnt = sm0 + kOffACnt + 4u * wg;
    const uint32_t descr = sm0 + kOffDescReady + 32u * wg, descf = sm0 + kOffDescFree + 32u * wg;
    const uint32_t descb = sm0 + kOffDesc + 128u * wg;
    const float gd_f = static_cast<float>(g.gd);
    // tensor memory of this warpgroup (128 columns): D of the round's two tiles at [0, 48) and
    // [48, 96); A operands at [96, 128): kKSteps == 1 -- two BUFFERS of two tiles x 8 columns, so
    // that the one-hot rows of round r + 2 are written while the MMAs of round r + 1 run;
    // kKSteps == 2 -- one buffer of two tiles x 16 columns (the set-up then waits for the MMAs).
    constexpr bool kPipeA = (kKSteps == 1);
    const uint32_t tD0 = tm_wg, tD1 = tm_wg + 48u, tA0 = tm_wg + 96u;

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
    // a pixel's 4 corner vectors back from tensor memory, blend, affine apply, store in place
    auto finish_px = [&](uint32_t td, uint32_t st, uint32_t tw, uint32_t wxs, float wz0, float wz1) {
      float d[48];
      const uint32_t taddr = td + lane_sel;
      mm_ld16(taddr, d);
      mm_ld16(taddr + 16, d + 16);
      mm_ld16(taddr + 32, d + 32);
      const bool valid = t < static_cast<int>(tw >> 16);
      const int p = static_cast<int>(tw & 0xffffu) + (valid ? t : 0);
      float r, gg, bb;
      mm_load_px_a<kIn>(st, p, r, gg, bb);
      const float wx1 = lds_f32(wxs + 4u * p), wx0 = 1.0f - wx1;
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
      if (valid) mm_store_px_a<kOut>(st + args.off_out, p, o_r, o_g, o_b);
    };

    // Per iteration: epilogue of round r (its MMAs were issued one iteration ago) -> MMAs of round
    // r + 1 (its A rows were written one iteration ago; D is free now) -> set-up of round r + 2
    // (overlaps those MMAs).  `cur` = MMAs issued, `nxt` = A written.
    float cw0 = 0.f, cw1 = 0.f, cw2 = 0.f, cw3 = 0.f, nw0 = 0.f, nw1 = 0.f, nw2 = 0.f, nw3 = 0.f;
    uint32_t cslot = 0u, nslot = 0u, nab = tA0;
    bool have_cur = false, have_nxt = false, ended = false;
    uint32_t nfetch = 0u, nsetup = 0u, dpar = 0u;
    for (;;) {
      // ---- 1. epilogue ---------------------------------------------------------------------
      if (have_cur) {
        mbar_wait_addr(dbar, dpar);
        dpar ^= 1u;
        mm_fence_after();
        const uint32_t da = descb + 32u * cslot;
        const uint4 d0 = lds_v4(da);
        const uint32_t xs = lds_u32(da + 20u), flags = lds_u32(da + 24u);
        const uint32_t wxs = wx_addr + 4u * xs;
        finish_px(tD0, d0.x, d0.y, wxs, cw0, cw1);
        finish_px(tD1, d0.x, d0.z, wxs, cw2, cw3);
        mm_fence_before();   // D is read: order it before the MMAs of the next round
        if (flags & kDescDone) fence_proxy_async_smem();   // results (generic writes) -> the issuer's bulk store
        __syncwarp();
        if (lane == 0) {
          mm_arrive_a(descf + 8u * cslot);
          if (flags & kDescDone) mm_arrive_a(sm0 + kOffDone + 8u * ((flags >> 2) & 3u));
          if (flags & kDescFree) mm_arrive_a(sm0 + kOffBFree + 8u * ((flags >> 5) & 1u));
        }
        have_cur = false;
      }
      // ---- 2. MMAs of the next round --------------------------------------------------------
      if (have_nxt) {
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        mm_fence_before();
        uint32_t old = 0u;
        if (lane == 0) old = mm_atom_inc(acnt);
        old = mm_uniform(old);
        if ((old & 3u) == 3u) {   // the last of the four warps: a warp-uniform branch
          mm_fence_after();
          const uint32_t da = descb + 32u * nslot;
          const uint32_t bA = mm_uniform(lds_u32(da + 12u)), bB = mm_uniform(lds_u32(da + 16u));
          const uint32_t ab = mm_uniform(nab);
          if (mm_elect_one()) {
            issue_mmas(tD0, ab, bA);
            issue_mmas(tD1, ab + 8u * kKSteps, bB);
            mm_commit(dbar);
          }
          __syncwarp();
        }
        cw0 = nw0; cw1 = nw1; cw2 = nw2; cw3 = nw3;
        cslot = nslot;
        have_cur = true;
        have_nxt = false;
      }
      // ---- 3. fetch + set-up of one more round ---------------------------------------------
      if (!ended) {
        const uint32_t slot = nfetch & 3u, par = (nfetch >> 2) & 1u;
        bool ready = true;
        // with MMAs in flight do not block here: their epilogue may be what the next descriptor waits for
        if (have_cur) ready = mm_test(descr + 8u * slot, par);
        else mbar_wait_addr(descr + 8u * slot, par);
        if (ready) {
          ++nfetch;
          const uint32_t da = descb + 32u * slot;
          const uint4 d0 = lds_v4(da);
          const uint32_t flags = lds_u32(da + 24u);
          if (flags & kDescEnd) {
            ended = true;
          } else {
            const uint32_t st = d0.x, twA = d0.y, twB = d0.z;
            const int pA = static_cast<int>(twA & 0xffffu) + (t < static_cast<int>(twA >> 16) ? t : 0);
            const int pB = static_cast<int>(twB & 0xffffu) + (t < static_cast<int>(twB >> 16) ? t : 0);
            float gvA, gvB;
            if constexpr (kGuideIn) {
              gvA = lds_f32(st + args.off_guide + 4u * pA);
              gvB = lds_f32(st + args.off_guide + 4u * pB);
            } else {
              float rA, gA, bA, rB, gB, bB;
              mm_load_px_a<kIn>(st, pA, rA, gA, bA);
              mm_load_px_a<kIn>(st, pB, rB, gB, bB);
              gvA = guide_fn(rA, gA, bA);
              gvB = guide_fn(rB, gB, bB);
              if (args.guide_out != nullptr) {
                float* go = args.guide_out + lds_u32(da + 28u);
                if (t < static_cast<int>(twA >> 16)) go[pA] = gvA;
                if (t < static_cast<int>(twB >> 16)) go[pB] = gvB;
              }
            }
            float zfA, zfB;
            depth_axis(gvA, zfA, nw0, nw1);
            depth_axis(gvB, zfB, nw2, nw3);
            nab = tA0 + (kPipeA ? (nsetup & 1u) * 16u : 0u);
            if constexpr (!kPipeA) {
              // one A buffer: the MMAs of `cur` (issued in step 2) read it -- wait for them first
              if (have_cur) mbar_wait_addr(dbar, dpar);
            }
            put_onehot(nab, zfA);
            put_onehot(nab + 8u * kKSteps, zfB);
            ++nsetup;
            nslot = slot;
            have_nxt = true;
          }
        }
      }
      if (ended && !have_cur && !have_nxt) break;
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
  a.off_wx = mm_round_up(1024 + (40 + 24 + a.nseg * a.max_tiles) * 4, 16);   // after the header and the tables (kOffTab)
  a.off_raw = mm_round_up(a.off_wx + g.W * 4, 128);
  if (static_cast<long long>(g.B) * g.rows * g.W >= (1LL << 32)) return false;   // 32-bit pixel offsets in descriptors
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
