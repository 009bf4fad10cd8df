// slice_apply_tc.cu -- EXPERIMENTAL (HDRNET_VARIANT_TC): fused BilateralSliceApply with the depth
// interpolation on the 5th-generation tensor cores.  Written at the end of round 1 from the
// corrected tensor-memory measurement (tools/ubench/tmem_paths.cu section 3: tcgen05.ld sustains
// ~900 B/clk/SM, seven times the shared-memory crossbar that bounds every other form).  One GPU
// run so far (tools/cabi_check.cu, 1 x 16 x 3840, grid 16x16x8: 6.8e-7 from the generic kernel, two
// launches; profiles/r01_tc_first_run.txt); no timing at scale; its pytest cases are gated behind
// HDRNET_TEST_EXPERIMENTAL=1 and AUTO never selects it.  Known cost of this first version: ~430 SASS
// instructions per pixel (A-row construction ~60, three-cell selection ~50, per-tile barrier /
// fence / wait overhead paid per pixel because a thread owns one pixel per tile).
//
// Replaces the per-pixel gather of hdrnet/ops/bilateral_slice_apply.cu.cc:36-126 by a tiny matrix
// product per 128-pixel tile (gd == 8):
//     D[128 px][48] = A[128 px][8] x B[8][48]          tcgen05.mma kind::tf32, M128 N48 K8
//   A[p][k] = smoothed depth weight of pixel p for depth cell k (two non-zeros; numerics.h:108-113),
//             written to TENSOR MEMORY by the pixel's own thread (tcgen05.st), split hi + lo;
//   B[k][n] = the image row's y-pre-blended slab for THREE consecutive x cells (n = cell * 16 +
//             coefficient, 12 used), shared memory, UMMA K-major no-swizzle layout, split hi + lo
//             once per image row by the issuer warp;
//   D       = for each of the three x cells the 12 coefficients already blended along depth:
//             A_hi B_hi + A_lo B_hi + A_hi B_lo (3xTF32: float32-grade accuracy), accumulated in
//             TMEM and read back with tcgen05.ld; the thread finishes with the x blend (two of the
//             three cells) and the 3x4 affine apply in registers.
// The shared-memory crossbar then carries only the pixel tiles (~0.45 wavefronts per pixel instead
// of 1.42), the texture pipe is not used at all.
//
// CTA = 2 math warpgroups (each owns 128 TMEM lanes x 128 columns = two tile buffers of
// D[48] | A_hi[8] | A_lo[8]) + the issuer warp of the issuer-warp form (slice_apply.cu): TMA ring
// for the pixel tiles, one bulk store per segment, slab rows from the pre-pass workspace.  Two
// CTAs per SM use all 512 TMEM columns.
#include <cuda_runtime.h>

#include <algorithm>
#include <climits>
#include <cstdint>
#include <cstdlib>

#include "common.cuh"

namespace hdrnet_b200 {

constexpr int kTcWgs = 2;                       // math warpgroups per CTA
constexpr int kTcMathWarps = kTcWgs * 4;
constexpr int kTcThreads = kTcWgs * 128 + 32;   // + the issuer warp
constexpr int kTcMaxStages = 4;
constexpr int kTcTile = 128;                    // pixels per MMA tile (M)
constexpr int kTcBufCols = 64;                  // TMEM columns per tile buffer
constexpr int kTcColA = 48;                     // A_hi at [48, 56), A_lo at [56, 64); D at [0, 48)
constexpr int kTcN = 48;                        // three x cells x 16 coefficient slots
constexpr int kTcTmemCols = kTcWgs * 2 * kTcBufCols;   // 256 per CTA

struct TcArgs {
  const float* guide;
  const unsigned char* input;
  unsigned char* out;
  const float* yslab;   // [B * rows][gw * 8 * 12] y-pre-blended slab rows (pre-pass workspace)
  SliceGeom g;
  int nseg, seg_px, stages, stage_bytes, off_guide;
  int off_raw, off_b, off_stage, raw_bytes, b_bytes, smem_bytes, ctas;
};

// instruction descriptor: D f32, A/B tf32, K-major both, N = 48, M = 128 (tools/ubench/tmem_paths.cu)
constexpr uint32_t kTcIdesc = (1u << 4) | (2u << 7) | (2u << 10) |
                              (static_cast<uint32_t>(kTcN >> 3) << 17) |
                              (static_cast<uint32_t>(kTcTile >> 4) << 24);

__device__ __forceinline__ uint64_t tc_kmajor_desc(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3fffu);
  d |= static_cast<uint64_t>((lbo >> 4) & 0x3fffu) << 16;
  d |= static_cast<uint64_t>((sbo >> 4) & 0x3fffu) << 32;
  d |= static_cast<uint64_t>(1) << 46;   // descriptor version (sm_100), no swizzle
  return d;
}
// B element (n, k), k < 8: 8-row x 16-byte core matrices, K chunks 128 B apart (LBO), groups of
// eight n 256 B apart (SBO) -- the layout the semantics probe validated.
__host__ __device__ constexpr int tc_b_off(int n, int k) {
  return (n >> 3) * 64 + (k >> 2) * 32 + (n & 7) * 4 + (k & 3);
}
__device__ __forceinline__ void tc_split(float v, float& hi, float& lo) {
  hi = __uint_as_float(__float_as_uint(v) & 0xffffe000u);   // what the tensor core reads of v
  lo = v - hi;
}
__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(kTcIdesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tc_commit(uint32_t bar_smem_addr) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_smem_addr)
               : "memory");
}
__device__ __forceinline__ void tc_st16(uint32_t taddr, const float (&hi)[8], const float (&lo)[8]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "f"(hi[0]), "f"(hi[1]), "f"(hi[2]), "f"(hi[3]), "f"(hi[4]), "f"(hi[5]), "f"(hi[6]), "f"(hi[7]),
      "f"(lo[0]), "f"(lo[1]), "f"(lo[2]), "f"(lo[3]), "f"(lo[4]), "f"(lo[5]), "f"(lo[6]), "f"(lo[7])
      : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, float (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7]),
        "=f"(r[8]), "=f"(r[9]), "=f"(r[10]), "=f"(r[11]), "=f"(r[12]), "=f"(r[13]), "=f"(r[14]), "=f"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_named_barrier(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// What a thread keeps of a pixel between the tile's set-up (A row, MMA issue) and its epilogue.
struct TcPixel {
  float r, g, b;     // input
  float fx;          // x fraction: weights (1 - fx, fx)
  int l0, l1;        // its two x cells, local to the tile's three (0..2)
  int p;             // pixel inside the segment, -1 = past the end
};

__global__ void __launch_bounds__(kTcThreads, 2)
slice_apply_rows_tc_kernel(const TcArgs args) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const SliceGeom& g = args.g;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  uint64_t* full = reinterpret_cast<uint64_t*>(smem);   // [kTcMaxStages]  pixel tile landed
  uint64_t* done = full + kTcMaxStages;                  // [kTcMaxStages]  every math warp is through
  uint64_t* raw_full = done + kTcMaxStages;              // [2]  fp32 slab row landed
  uint64_t* b_full = raw_full + 2;                       // [2]  hi / lo operand tiles of a row ready
  uint64_t* dbar = b_full + 2;                           // [kTcWgs][2]  a tile's MMAs complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 240);
  unsigned char* raw0 = smem + args.off_raw;             // two fp32 slab rows
  float* bt = reinterpret_cast<float*>(smem + args.off_b);   // [2 rows][hi, lo][gw * 16 * 8] operand tiles
  unsigned char* stage_base = smem + args.off_stage;
  const int b_floats = args.b_bytes / 4;

  const long long total_items = static_cast<long long>(g.B) * g.rows * args.nseg;
  const long long i_begin = total_items * blockIdx.x / gridDim.x;
  const long long i_end = total_items * (blockIdx.x + 1) / gridDim.x;
  if (i_end <= i_begin) return;
  const long long r_begin = i_begin / args.nseg, r_end = (i_end - 1) / args.nseg + 1;
  const int x_first = static_cast<int>(i_begin - r_begin * args.nseg) * args.seg_px;
  const int x_last = min(g.W, (static_cast<int>((i_end - 1) - (r_end - 1) * args.nseg) + 1) * args.seg_px);
  auto row_x0 = [&](long long row) { return row == r_begin ? x_first : 0; };
  auto row_x1 = [&](long long row) { return row == r_end - 1 ? x_last : g.W; };

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(static_cast<uint32_t>(kTcTmemCols)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    for (int s = 0; s < args.stages; ++s) { mbar_init(&full[s], 1); mbar_init(&done[s], kTcMathWarps); }
    for (int i = 0; i < 2; ++i) { mbar_init(&raw_full[i], 1); mbar_init(&b_full[i], 1); }
    for (int i = 0; i < kTcWgs * 2; ++i) mbar_init(&dbar[i], 1);
    fence_mbar_init();
  }
  // the coefficient slots 12..15 of every cell are padding: zero all operand tiles once
  for (int e = tid; e < 4 * b_floats; e += kTcThreads) bt[e] = 0.0f;
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int NS = args.stages;
  const uint32_t raw_bytes = static_cast<uint32_t>(args.raw_bytes);
  auto arrive = [&](uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
  };

  if (warp == kTcMathWarps) {
    // ------------------------------- issuer warp --------------------------------------------
    auto load_raw = [&](long long row) {   // lane 0
      const int rb = static_cast<int>(row - r_begin) & 1;
      mbar_expect_tx(&raw_full[rb], raw_bytes);
      tma_load_1d(raw0 + static_cast<size_t>(rb) * raw_bytes,
                  args.yslab + static_cast<size_t>(row) * (raw_bytes / 4), raw_bytes, &raw_full[rb]);
    };
    // fp32 slab row [cell][z][12] -> hi / lo operand tiles in the UMMA K-major layout (whole warp)
    auto make_b = [&](long long row) {
      const int rowk = static_cast<int>(row - r_begin), rb = rowk & 1;
      mbar_wait(&raw_full[rb], static_cast<uint32_t>(rowk >> 1) & 1u);
      const float* src = reinterpret_cast<const float*>(raw0 + static_cast<size_t>(rb) * raw_bytes);
      float* hi_t = bt + static_cast<size_t>(rb) * 2 * b_floats;
      float* lo_t = hi_t + b_floats;
      const int n_src = g.gw * 96;
      for (int e = lane; e < n_src; e += 32) {
        const int cell = e / 96, rem = e - cell * 96;
        const int z = rem / 12, j = rem - z * 12;
        float hi, lo;
        tc_split(src[e], hi, lo);
        const int off = tc_b_off(cell * 16 + j, z);
        hi_t[off] = hi;
        lo_t[off] = lo;
      }
      fence_proxy_async_smem();   // generic writes -> the tensor core's (async-proxy) operand reads
      __syncwarp();
      if (lane == 0) {
        arrive(&b_full[rb]);
        if (row + 2 < r_end) load_raw(row + 2);   // the fp32 row buffer is free again
      }
    };
    long long l_row = r_begin;
    int l_x0 = x_first, l_s = 0;
    auto issue_next_load = [&]() {  // lane 0
      if (l_row >= r_end) return;
      const int npx = min(args.seg_px, g.W - l_x0);
      unsigned char* st = stage_base + static_cast<size_t>(l_s) * args.stage_bytes;
      const size_t pix = static_cast<size_t>(l_row) * g.W + l_x0;
      mbar_expect_tx(&full[l_s], static_cast<uint32_t>(npx) * 16u);
      tma_load_1d(st, args.input + pix * 12, static_cast<uint32_t>(npx) * 12u, &full[l_s]);
      tma_load_1d(st + args.off_guide, args.guide + pix, static_cast<uint32_t>(npx) * 4u, &full[l_s]);
      if (++l_s == NS) l_s = 0;
      l_x0 += args.seg_px;
      if (l_x0 >= row_x1(l_row)) { l_x0 = 0; ++l_row; }
    };
    if (lane == 0) {
      for (int i = 0; i < NS - 1; ++i) issue_next_load();
      load_raw(r_begin);
      if (r_begin + 1 < r_end) load_raw(r_begin + 1);
    }
    __syncwarp();
    make_b(r_begin);
    if (r_begin + 1 < r_end) make_b(r_begin + 1);

    int s = 0;
    uint32_t ph = 0;
    for (long long row = r_begin; row < r_end; ++row) {
      if (lane == 0) {
        const int x_end = row_x1(row);
        for (int x0 = row_x0(row); x0 < x_end; x0 += args.seg_px) {
          mbar_wait(&done[s], ph);  // every math warp has written (and proxy-fenced) its results
          const int npx = min(args.seg_px, g.W - x0);
          unsigned char* st = stage_base + static_cast<size_t>(s) * args.stage_bytes;
          const size_t pix = static_cast<size_t>(row) * g.W + x0;
          tma_store_1d(args.out + pix * 12, st, static_cast<uint32_t>(npx) * 12u);
          tma_store_commit();
          if (l_row < r_end) {
            tma_store_wait_read<1>();
            issue_next_load();
          }
          if (++s == NS) { s = 0; ph ^= 1u; }
        }
      }
      __syncwarp();
      // the row's operand tiles are free: every math warp arrived after its last MMA on them
      if (row + 2 < r_end) make_b(row + 2);
    }
    if (lane == 0) tma_store_wait_all<0>();
  } else {
    // --------------------------------- math warpgroups ---------------------------------------
    const int wg = warp >> 2, t = tid & 127;
    const uint32_t lane_sel = static_cast<uint32_t>((warp & 3) * 32) << 16;   // this warp's TMEM lanes
    const uint32_t tm_wg = tmem_base + static_cast<uint32_t>(wg) * (2 * kTcBufCols);
    const float gd_f = 8.0f;
    uint32_t dpar = 0u;            // bit b: phase parity of dbar[wg][b]
    const uint32_t dbar0 = smem_u32(&dbar[wg * 2]);
    const uint32_t bt_addr = smem_u32(bt);

    int s = 0;
    uint32_t ph = 0;
    for (long long row = r_begin; row < r_end; ++row) {
      const int rowk = static_cast<int>(row - r_begin), rb = rowk & 1;
      mbar_wait(&b_full[rb], static_cast<uint32_t>(rowk >> 1) & 1u);
      const uint32_t b_hi = bt_addr + static_cast<uint32_t>(rb) * 2u * static_cast<uint32_t>(args.b_bytes);
      const uint32_t b_lo = b_hi + static_cast<uint32_t>(args.b_bytes);
      const int x_end = row_x1(row);
      for (int x0 = row_x0(row); x0 < x_end; x0 += args.seg_px) {
        const int npx = min(args.seg_px, g.W - x0);
        unsigned char* st = stage_base + static_cast<size_t>(s) * args.stage_bytes;
        const float* rgb = reinterpret_cast<const float*>(st);
        const float* gui = reinterpret_cast<const float*>(st + args.off_guide);
        mbar_wait(&full[s], ph);

        const int ntiles = (npx + kTcTile - 1) / kTcTile;
        TcPixel prev;
        prev.p = INT_MIN;            // no tile in flight
        int prev_buf = 0, buf = 0;

        // Epilogue of a tile: read D, blend the pixel's two x cells, affine apply, store in place.
        auto epilogue = [&](const TcPixel& px, int b) {
          mbar_wait_addr(dbar0 + 8u * b, (dpar >> b) & 1u);
          dpar ^= 1u << b;
          tc_fence_after();
          float d0[16], d1[16], d2[16];
          const uint32_t taddr = tm_wg + static_cast<uint32_t>(b) * kTcBufCols + lane_sel;
          tc_ld16(taddr, d0);
          tc_ld16(taddr + 16, d1);
          tc_ld16(taddr + 32, d2);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          if (px.p >= 0) {
            const float w1 = px.fx, w0 = 1.0f - px.fx;
            float v[12];
#pragma unroll
            for (int c = 0; c < 12; ++c) {
              const float a = px.l0 == 0 ? d0[c] : (px.l0 == 1 ? d1[c] : d2[c]);
              const float bb = px.l1 == 0 ? d0[c] : (px.l1 == 1 ? d1[c] : d2[c]);
              v[c] = fmaf(w1, bb, w0 * a);
            }
            float* o = reinterpret_cast<float*>(st) + 3 * px.p;
            o[0] = fmaf(v[2], px.b, fmaf(v[1], px.g, fmaf(v[0], px.r, v[3])));
            o[1] = fmaf(v[6], px.b, fmaf(v[5], px.g, fmaf(v[4], px.r, v[7])));
            o[2] = fmaf(v[10], px.b, fmaf(v[9], px.g, fmaf(v[8], px.r, v[11])));
          }
        };

        for (int j = wg; j < ntiles; j += kTcWgs) {
          // ---- set-up of tile j: this thread's pixel, its A row, the tile's MMAs --------------
          const int p = j * kTcTile + t;
          TcPixel cur;
          cur.p = p < npx ? p : -1;
          const int pc = p < npx ? p : npx - 1;          // clamp reads of idle lanes
          cur.r = rgb[3 * pc]; cur.g = rgb[3 * pc + 1]; cur.b = rgb[3 * pc + 2];
          const float gv = gui[pc];
          // depth axis: range_axis / smoothed_weights, bit-exact cell index
          const float tz = __fsub_rn(__fmul_rn(gv, gd_f), 0.5f);
          const int iz = __float2int_rd(tz);
          const float fz = tz - static_cast<float>(iz);
          const int zc0 = clampi(iz, 0, 7), zc1 = clampi(iz + 1, 0, 7);
          float wz0, wz1;
          smoothed_weights(fz, wz0, wz1);
          if (cur.p < 0) { wz0 = 0.0f; wz1 = 0.0f; }
          float a_hi[8], a_lo[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const float a = (k == zc0 ? wz0 : 0.0f) + (k == zc1 ? wz1 : 0.0f);
            tc_split(a, a_hi[k], a_lo[k]);
          }
          const uint32_t tbuf = tm_wg + static_cast<uint32_t>(buf) * kTcBufCols;
          tc_st16(tbuf + kTcColA + lane_sel, a_hi, a_lo);
          asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
          // x axis: the tile's three cells start at cb (computed identically by every thread)
          const Axis a_first = spatial_axis(x0 + j * kTcTile, g.scale_x);
          const int cb = clampi(a_first.i0, 0, g.gw - 3);
          const Axis ax = spatial_axis(x0 + pc, g.scale_x);
          cur.fx = ax.f;
          cur.l0 = clampi(ax.i0, 0, g.gw - 1) - cb;
          cur.l1 = clampi(ax.i0 + 1, 0, g.gw - 1) - cb;
          tc_fence_before();
          tc_named_barrier(1 + wg, 128);
          if (t == 0) {
            tc_fence_after();
            // cell cb's columns start 2 groups of eight n (2 x 256 B) per cell into the row's tiles
            const uint64_t dh = tc_kmajor_desc(b_hi + static_cast<uint32_t>(cb) * 512u, 128, 256);
            const uint64_t dl = tc_kmajor_desc(b_lo + static_cast<uint32_t>(cb) * 512u, 128, 256);
            tc_mma(tbuf, tbuf + kTcColA, dh, 0u);        // A_hi B_hi
            tc_mma(tbuf, tbuf + kTcColA + 8, dh, 1u);    // A_lo B_hi
            tc_mma(tbuf, tbuf + kTcColA, dl, 1u);        // A_hi B_lo
            tc_commit(dbar0 + 8u * buf);
          }
          // ---- epilogue of the previous tile while this one's MMAs run ------------------------
          if (prev.p != INT_MIN) epilogue(prev, prev_buf);
          prev = cur;
          prev_buf = buf;
          buf ^= 1;
        }
        if (prev.p != INT_MIN) epilogue(prev, prev_buf);
        fence_proxy_async_smem();   // results (generic writes) -> the issuer's bulk store
        __syncwarp();
        if (lane == 0) arrive(&done[s]);
        if (++s == NS) { s = 0; ph ^= 1u; }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"(static_cast<uint32_t>(kTcTmemCols)));
  }
}


// =========================================================================================
// Second form (HDRNET_VARIANT_TC_GATHER): the tensor core as an exact GATHER engine.
// =========================================================================================
// The first form spends ~60 instructions per pixel on a weighted one-hot A row (hi + lo halves) and
// leaves the x blend to the thread anyway.  Here A is a plain one-hot row -- 1.0 in the column of
// the pixel's lower depth cell, exact in TF32, no split, 16 instructions -- and B carries, per x
// cell, BOTH depth rows a pixel of that cell needs:
//     B'[k][cell * 32 + h * 16 + j] = slab[cell][min(k + h, 7)][j]      h = 0, 1;  j < 12
// so that D[p][.] = A B'_hi + A B'_lo (2 MMAs, M128 N96 K8) is an fp32-exact COPY (to 2^-22) of
// the pixel's depth rows z0 and z0 + 1 for the tile's three x cells, and the thread does the
// trilinear blend of its 4 corners x 12 coefficients in registers exactly as the row kernels do
// (24 FFMA2 + 9 FMA).  Pixels whose two depth cells clamp to the same cell fold both weights onto
// the first row.  Warps whose pixels share one x-cell pair (7 of 8 at 4K) read 64 contiguous TMEM
// columns; the others read all 96 and weight the three cells.
// CTA = 4 math warpgroups (one 128-column TMEM slot each: D[96] | A[8]) + the issuer warp, ONE CTA
// per SM (512 TMEM columns, ~157 KB of shared memory: B' for two image rows is 64 KB at gw = 16).
// Tiles are dealt to the warpgroups round-robin across segment boundaries.
// NEVER RUN ON A GPU (written after the round's GPU budget was spent); arithmetic as in
// tests/test_tc_emulation.py::test_gather_form_*.
constexpr int kTgWgs = 4;
constexpr int kTgMathWarps = kTgWgs * 4;
constexpr int kTgThreads = kTgWgs * 128 + 32;
constexpr int kTgSlotCols = 128;            // D at [0, 96), A at [96, 104)
constexpr int kTgColA = 96;
constexpr int kTgN = 96;
constexpr int kTgMaxStages = 4;
constexpr uint32_t kTgIdesc = (1u << 4) | (2u << 7) | (2u << 10) |
                              (static_cast<uint32_t>(kTgN >> 3) << 17) |
                              (static_cast<uint32_t>(kTcTile >> 4) << 24);

__device__ __forceinline__ void tg_mma(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(kTgIdesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tg_st8(uint32_t taddr, const float (&a)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr),
               "f"(a[0]), "f"(a[1]), "f"(a[2]), "f"(a[3]), "f"(a[4]), "f"(a[5]), "f"(a[6]), "f"(a[7])
               : "memory");
}

__global__ void __launch_bounds__(kTgThreads, 1)
slice_apply_rows_tcg_kernel(const TcArgs args) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const SliceGeom& g = args.g;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  uint64_t* full = reinterpret_cast<uint64_t*>(smem);   // [kTgMaxStages]
  uint64_t* done = full + kTgMaxStages;                  // [kTgMaxStages]
  uint64_t* raw_full = done + kTgMaxStages;              // [2]
  uint64_t* b_full = raw_full + 2;                       // [2]
  uint64_t* dbar = b_full + 2;                           // [kTgWgs]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 240);
  unsigned char* raw0 = smem + args.off_raw;
  float* bt = reinterpret_cast<float*>(smem + args.off_b);   // [2 rows][hi, lo][gw * 32 * 8]
  unsigned char* stage_base = smem + args.off_stage;
  const int b_floats = args.b_bytes / 4;

  const long long total_items = static_cast<long long>(g.B) * g.rows * args.nseg;
  const long long i_begin = total_items * blockIdx.x / gridDim.x;
  const long long i_end = total_items * (blockIdx.x + 1) / gridDim.x;
  if (i_end <= i_begin) return;
  const long long r_begin = i_begin / args.nseg, r_end = (i_end - 1) / args.nseg + 1;
  const int x_first = static_cast<int>(i_begin - r_begin * args.nseg) * args.seg_px;
  const int x_last = min(g.W, (static_cast<int>((i_end - 1) - (r_end - 1) * args.nseg) + 1) * args.seg_px);
  auto row_x0 = [&](long long row) { return row == r_begin ? x_first : 0; };
  auto row_x1 = [&](long long row) { return row == r_end - 1 ? x_last : g.W; };

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(static_cast<uint32_t>(kTgWgs * kTgSlotCols)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    for (int s = 0; s < args.stages; ++s) { mbar_init(&full[s], 1); mbar_init(&done[s], kTgMathWarps); }
    for (int i = 0; i < 2; ++i) { mbar_init(&raw_full[i], 1); mbar_init(&b_full[i], 1); }
    for (int i = 0; i < kTgWgs; ++i) mbar_init(&dbar[i], 1);
    fence_mbar_init();
  }
  for (int e = tid; e < 4 * b_floats; e += kTgThreads) bt[e] = 0.0f;   // padding slots 12..15 stay zero
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int NS = args.stages;
  const uint32_t raw_bytes = static_cast<uint32_t>(args.raw_bytes);
  auto arrive = [&](uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
  };

  if (warp == kTgMathWarps) {
    // ------------------------------- issuer warp --------------------------------------------
    auto load_raw = [&](long long row) {   // lane 0
      const int rb = static_cast<int>(row - r_begin) & 1;
      mbar_expect_tx(&raw_full[rb], raw_bytes);
      tma_load_1d(raw0 + static_cast<size_t>(rb) * raw_bytes,
                  args.yslab + static_cast<size_t>(row) * (raw_bytes / 4), raw_bytes, &raw_full[rb]);
    };
    // fp32 slab row [cell][z][12] -> B' hi / lo: every value lands twice, as depth row z of half 0
    // and as depth row z - 1 of half 1 (and row 7 of half 1 repeats depth cell 7)
    auto make_b = [&](long long row) {
      const int rowk = static_cast<int>(row - r_begin), rb = rowk & 1;
      mbar_wait(&raw_full[rb], static_cast<uint32_t>(rowk >> 1) & 1u);
      const float* src = reinterpret_cast<const float*>(raw0 + static_cast<size_t>(rb) * raw_bytes);
      float* hi_t = bt + static_cast<size_t>(rb) * 2 * b_floats;
      float* lo_t = hi_t + b_floats;
      const int n_src = g.gw * 96;
      for (int e = lane; e < n_src; e += 32) {
        const int cell = e / 96, rem = e - cell * 96;
        const int z = rem / 12, j = rem - z * 12;
        float hi, lo;
        tc_split(src[e], hi, lo);
        const int n0 = cell * 32 + j;
        const int o0 = tc_b_off(n0, z);                 // half 0: k = z
        hi_t[o0] = hi; lo_t[o0] = lo;
        if (z >= 1) {                                   // half 1: k = z - 1
          const int o1 = tc_b_off(n0 + 16, z - 1);
          hi_t[o1] = hi; lo_t[o1] = lo;
        }
        if (z == 7) {                                   // half 1, k = 7: min(7 + 1, 7)
          const int o2 = tc_b_off(n0 + 16, 7);
          hi_t[o2] = hi; lo_t[o2] = lo;
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        arrive(&b_full[rb]);
        if (row + 2 < r_end) load_raw(row + 2);
      }
    };
    long long l_row = r_begin;
    int l_x0 = x_first, l_s = 0;
    auto issue_next_load = [&]() {  // lane 0
      if (l_row >= r_end) return;
      const int npx = min(args.seg_px, g.W - l_x0);
      unsigned char* st = stage_base + static_cast<size_t>(l_s) * args.stage_bytes;
      const size_t pix = static_cast<size_t>(l_row) * g.W + l_x0;
      mbar_expect_tx(&full[l_s], static_cast<uint32_t>(npx) * 16u);
      tma_load_1d(st, args.input + pix * 12, static_cast<uint32_t>(npx) * 12u, &full[l_s]);
      tma_load_1d(st + args.off_guide, args.guide + pix, static_cast<uint32_t>(npx) * 4u, &full[l_s]);
      if (++l_s == NS) l_s = 0;
      l_x0 += args.seg_px;
      if (l_x0 >= row_x1(l_row)) { l_x0 = 0; ++l_row; }
    };
    if (lane == 0) {
      for (int i = 0; i < NS - 1; ++i) issue_next_load();
      load_raw(r_begin);
      if (r_begin + 1 < r_end) load_raw(r_begin + 1);
    }
    __syncwarp();
    make_b(r_begin);
    if (r_begin + 1 < r_end) make_b(r_begin + 1);

    int s = 0;
    uint32_t ph = 0;
    for (long long row = r_begin; row < r_end; ++row) {
      if (lane == 0) {
        const int x_end = row_x1(row);
        for (int x0 = row_x0(row); x0 < x_end; x0 += args.seg_px) {
          mbar_wait(&done[s], ph);
          const int npx = min(args.seg_px, g.W - x0);
          unsigned char* st = stage_base + static_cast<size_t>(s) * args.stage_bytes;
          const size_t pix = static_cast<size_t>(row) * g.W + x0;
          tma_store_1d(args.out + pix * 12, st, static_cast<uint32_t>(npx) * 12u);
          tma_store_commit();
          if (l_row < r_end) {
            tma_store_wait_read<1>();
            issue_next_load();
          }
          if (++s == NS) { s = 0; ph ^= 1u; }
        }
      }
      __syncwarp();
      if (row + 2 < r_end) make_b(row + 2);
    }
    if (lane == 0) tma_store_wait_all<0>();
  } else {
    // --------------------------------- math warpgroups ---------------------------------------
    const int wg = warp >> 2, t = tid & 127;
    const uint32_t lane_sel = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const uint32_t tslot = tmem_base + static_cast<uint32_t>(wg) * kTgSlotCols;
    uint32_t dpar = 0u;
    const uint32_t dbar_addr = smem_u32(&dbar[wg]);
    const uint32_t bt_addr = smem_u32(bt);
    int tile_base = 0;   // tiles of this CTA's earlier segments, mod kTgWgs: deals tiles round-robin

    int s = 0;
    uint32_t ph = 0;
    for (long long row = r_begin; row < r_end; ++row) {
      const int rowk = static_cast<int>(row - r_begin), rb = rowk & 1;
      mbar_wait(&b_full[rb], static_cast<uint32_t>(rowk >> 1) & 1u);
      const uint32_t b_hi = bt_addr + static_cast<uint32_t>(rb) * 2u * static_cast<uint32_t>(args.b_bytes);
      const uint32_t b_lo = b_hi + static_cast<uint32_t>(args.b_bytes);
      const int x_end = row_x1(row);
      for (int x0 = row_x0(row); x0 < x_end; x0 += args.seg_px) {
        const int npx = min(args.seg_px, g.W - x0);
        unsigned char* st = stage_base + static_cast<size_t>(s) * args.stage_bytes;
        const float* rgb = reinterpret_cast<const float*>(st);
        const float* gui = reinterpret_cast<const float*>(st + args.off_guide);
        mbar_wait(&full[s], ph);
        const int ntiles = (npx + kTcTile - 1) / kTcTile;
        // first tile of this segment that belongs to this warpgroup
        for (int j = (wg - tile_base) & (kTgWgs - 1); j < ntiles; j += kTgWgs) {
          const int p = j * kTcTile + t;
          const bool valid = p < npx;
          const int pc = valid ? p : npx - 1;
          const float pr = rgb[3 * pc], pg = rgb[3 * pc + 1], pb = rgb[3 * pc + 2];
          const float gv = gui[pc];
          // depth axis (range_axis / smoothed_weights); both weights on the first row when the
          // two cells clamp to the same one
          const float tz = __fsub_rn(__fmul_rn(gv, 8.0f), 0.5f);
          const int iz = __float2int_rd(tz);
          const float fz = tz - static_cast<float>(iz);
          const int zc0 = clampi(iz, 0, 7), zc1 = clampi(iz + 1, 0, 7);
          float wz0, wz1;
          smoothed_weights(fz, wz0, wz1);
          if (zc1 == zc0) { wz0 += wz1; wz1 = 0.0f; }
          float a[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) a[k] = (valid && k == zc0) ? 1.0f : 0.0f;
          tg_st8(tslot + kTgColA + lane_sel, a);
          asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
          // x axis
          const Axis a_first = spatial_axis(x0 + j * kTcTile, g.scale_x);
          const int cb = clampi(a_first.i0, 0, g.gw - 3);
          const Axis ax = spatial_axis(x0 + pc, g.scale_x);
          const int l0 = clampi(ax.i0, 0, g.gw - 1) - cb;
          const int l1 = clampi(ax.i0 + 1, 0, g.gw - 1) - cb;
          const float wx1 = ax.f, wx0 = 1.0f - ax.f;
          tc_fence_before();
          tc_named_barrier(1 + wg, 128);
          if (t == 0) {
            tc_fence_after();
            // a cell is 4 groups of eight n (4 x 256 B) in the row's B' tiles
            const uint64_t dh = tc_kmajor_desc(b_hi + static_cast<uint32_t>(cb) * 1024u, 128, 256);
            const uint64_t dl = tc_kmajor_desc(b_lo + static_cast<uint32_t>(cb) * 1024u, 128, 256);
            tg_mma(tslot, tslot + kTgColA, dh, 0u);
            tg_mma(tslot, tslot + kTgColA, dl, 1u);
            tc_commit(dbar_addr);
          }
          mbar_wait_addr(dbar_addr, dpar);
          dpar ^= 1u;
          tc_fence_after();
          const uint32_t taddr = tslot + lane_sel;
          float v[12];
          const int l0u = __shfl_sync(0xffffffffu, l0, 0);
          const bool uni = __all_sync(0xffffffffu, l0 == l0u && l1 == l0u + 1) != 0;
          if (uni) {
            // columns l0*32 ..: [cell l0: z0 | z1][cell l0+1: z0 | z1], 16 each
            float d00[16], d01[16], d10[16], d11[16];
            const uint32_t c0 = taddr + static_cast<uint32_t>(l0u) * 32u;
            tc_ld16(c0, d00);
            tc_ld16(c0 + 16, d01);
            tc_ld16(c0 + 32, d10);
            tc_ld16(c0 + 48, d11);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            const float w00 = wx0 * wz0, w01 = wx0 * wz1, w10 = wx1 * wz0, w11 = wx1 * wz1;
            // packed fp32x2 (FFMA2), the row kernels' order of operations
            const unsigned long long W00 = pack2(w00, w00), W01 = pack2(w01, w01);
            const unsigned long long W10 = pack2(w10, w10), W11 = pack2(w11, w11);
#pragma unroll
            for (int c = 0; c < 12; c += 2) {
              const unsigned long long acc =
                  fma2(W11, pack2(d11[c], d11[c + 1]),
                       fma2(W10, pack2(d10[c], d10[c + 1]),
                            fma2(W01, pack2(d01[c], d01[c + 1]), mul2(W00, pack2(d00[c], d00[c + 1])))));
              unpack2(acc, v[c], v[c + 1]);
            }
          } else {
            // the tile's three cells, weighted: a cell that is not one of the pixel's two gets 0
            float wc[3];
#pragma unroll
            for (int c3 = 0; c3 < 3; ++c3) wc[c3] = (l0 == c3 ? wx0 : 0.0f) + (l1 == c3 ? wx1 : 0.0f);
#pragma unroll
            for (int c = 0; c < 12; ++c) v[c] = 0.0f;
#pragma unroll
            for (int c3 = 0; c3 < 3; ++c3) {
              float dz0[16], dz1[16];
              tc_ld16(taddr + static_cast<uint32_t>(c3) * 32u, dz0);
              tc_ld16(taddr + static_cast<uint32_t>(c3) * 32u + 16u, dz1);
              asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
              const float u0 = wc[c3] * wz0, u1 = wc[c3] * wz1;
#pragma unroll
              for (int c = 0; c < 12; ++c) v[c] = fmaf(u1, dz1[c], fmaf(u0, dz0[c], v[c]));
            }
          }
          if (valid) {
            float* o = reinterpret_cast<float*>(st) + 3 * p;
            o[0] = fmaf(v[2], pb, fmaf(v[1], pg, fmaf(v[0], pr, v[3])));
            o[1] = fmaf(v[6], pb, fmaf(v[5], pg, fmaf(v[4], pr, v[7])));
            o[2] = fmaf(v[10], pb, fmaf(v[9], pg, fmaf(v[8], pr, v[11])));
          }
          // the slot is reused by this warpgroup's next tile: its reads of D are complete
          // (wait::ld above); order them before the next tile's MMAs
          tc_fence_before();
        }
        tile_base = (tile_base + ntiles) & (kTgWgs - 1);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) arrive(&done[s]);
        if (++s == NS) { s = 0; ph ^= 1u; }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"(static_cast<uint32_t>(kTgWgs * kTgSlotCols)));
  }
}

static inline int tc_round_up(int v, int m) { return (v + m - 1) / m * m; }

// Plans and launches the tensor-core form; HDRNET_E_UNSUPPORTED for shapes it does not take
// (gd != 8, fewer than three x cells, x cells narrower than a 128-pixel tile, ...).
int launch_slice_apply_tc(const float* guide, const float* input, float* out, const float* yslab,
                          const SliceGeom& g, int max_smem, int sms, cudaStream_t stream) {
  if (g.gd != 8 || g.gw < 3 || g.W % 4 != 0 || static_cast<long long>(g.W) < 128LL * g.gw)
    return HDRNET_E_UNSUPPORTED;
  TcArgs a;
  a.guide = guide;
  a.input = reinterpret_cast<const unsigned char*>(input);
  a.out = reinterpret_cast<unsigned char*>(out);
  a.yslab = yslab;
  a.g = g;
  // segments of whole tiles, at most 1280 pixels (ten tiles: five per warpgroup)
  const int max_seg = 1280;
  a.nseg = (g.W + max_seg - 1) / max_seg;
  a.seg_px = tc_round_up((g.W + a.nseg - 1) / a.nseg, kTcTile);
  a.nseg = (g.W + a.seg_px - 1) / a.seg_px;
  a.off_guide = tc_round_up(a.seg_px * 12, 16);
  a.stage_bytes = tc_round_up(a.off_guide + a.seg_px * 4, 128);
  a.raw_bytes = g.gw * 8 * 12 * 4;
  a.b_bytes = g.gw * 16 * 8 * 4;
  a.off_raw = 256;
  a.off_b = tc_round_up(a.off_raw + 2 * a.raw_bytes, 1024);
  a.off_stage = tc_round_up(a.off_b + 4 * a.b_bytes, 128);
  const int per_cta_2 = (max_smem + 1024) / 2 - 1024;
  a.stages = 0;
  for (int ns = kTcMaxStages; ns >= 2; --ns)
    if (a.off_stage + ns * a.stage_bytes <= per_cta_2) { a.stages = ns; break; }
  if (a.stages == 0) return HDRNET_E_UNSUPPORTED;
  a.smem_bytes = a.off_stage + a.stages * a.stage_bytes;
  const long long total_items = static_cast<long long>(g.B) * g.rows * a.nseg;
  a.ctas = static_cast<int>(std::min<long long>(total_items, static_cast<long long>(sms) * 2));
  cudaError_t e = cudaFuncSetAttribute(slice_apply_rows_tc_kernel,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, a.smem_bytes);
  if (e != cudaSuccess) return static_cast<int>(e);
  slice_apply_rows_tc_kernel<<<a.ctas, kTcThreads, a.smem_bytes, stream>>>(a);
  return static_cast<int>(cudaGetLastError());
}

int launch_slice_apply_tcg(const float* guide, const float* input, float* out, const float* yslab,
                           const SliceGeom& g, int max_smem, int sms, cudaStream_t stream) {
  if (g.gd != 8 || g.gw < 3 || g.W % 4 != 0 || static_cast<long long>(g.W) < 128LL * g.gw)
    return HDRNET_E_UNSUPPORTED;
  TcArgs a;
  a.guide = guide;
  a.input = reinterpret_cast<const unsigned char*>(input);
  a.out = reinterpret_cast<unsigned char*>(out);
  a.yslab = yslab;
  a.g = g;
  const int max_seg = 1280;
  a.nseg = (g.W + max_seg - 1) / max_seg;
  a.seg_px = tc_round_up((g.W + a.nseg - 1) / a.nseg, kTcTile);
  a.nseg = (g.W + a.seg_px - 1) / a.seg_px;
  a.off_guide = tc_round_up(a.seg_px * 12, 16);
  a.stage_bytes = tc_round_up(a.off_guide + a.seg_px * 4, 128);
  a.raw_bytes = g.gw * 8 * 12 * 4;
  a.b_bytes = g.gw * 32 * 8 * 4;          // B': two depth rows per cell
  a.off_raw = 256;
  a.off_b = tc_round_up(a.off_raw + 2 * a.raw_bytes, 1024);
  a.off_stage = tc_round_up(a.off_b + 4 * a.b_bytes, 128);
  a.stages = 0;
  for (int ns = kTgMaxStages; ns >= 2; --ns)
    if (a.off_stage + ns * a.stage_bytes <= max_smem) { a.stages = ns; break; }
  if (a.stages == 0) return HDRNET_E_UNSUPPORTED;
  a.smem_bytes = a.off_stage + a.stages * a.stage_bytes;
  const long long total_items = static_cast<long long>(g.B) * g.rows * a.nseg;
  a.ctas = static_cast<int>(std::min<long long>(total_items, static_cast<long long>(sms)));   // one CTA per SM
  cudaError_t e = cudaFuncSetAttribute(slice_apply_rows_tcg_kernel,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, a.smem_bytes);
  if (e != cudaSuccess) return static_cast<int>(e);
  slice_apply_rows_tcg_kernel<<<a.ctas, kTgThreads, a.smem_bytes, stream>>>(a);
  return static_cast<int>(cudaGetLastError());
}

}  // namespace hdrnet_b200
