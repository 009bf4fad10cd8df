// slice_apply_zsort.cu -- z-bucketed fused BilateralSliceApply (HDRNET_VARIANT_ZSORT).
//
// Why: ncu shows the row kernel in slice_apply.cu bound by the shared-memory data pipe, not by
// HBM: every pixel needs 4 corner vectors x 12 floats = 192 B delivered from the slab and the
// LSU delivers 128 B/clk/SM (1.5 clk/px/SM against the 1.21 clk/px/SM HBM allows; SHFL and
// L1-hit LDG share the same crossbar -- tools/ubench/gather_paths.cu).  The only way under
// that floor is to deliver fewer coefficient bytes per pixel, i.e. to let one thread reuse a
// set of corner vectors for several pixels.  Pixels that share (x cell, depth cell) share all
// four corner vectors, so each CTA counting-sorts the pixels of a row segment by that key
// (buckets padded to multiples of 4) and a thread then processes 4 consecutive SORTED pixels
// with ONE load of the 48 coefficients:  192 B/px -> 48 B/px of slab traffic, paid for with a
// warp-match ranking pass, an 8-byte record per pixel and a gather / scatter of the RGB.
//
// Everything else is as in slice_apply_rows_tma_kernel: persistent CTAs over contiguous image
// rows, TMA bulk copies in and out through an mbarrier ring, grid rows staged by TMA and
// pre-blended in y once per image row, bit-exact cell indices, FFMA2 blends.  The per-pixel
// arithmetic is identical (same functions from common.cuh), only the order differs, so the
// result is bitwise equal to the row kernel's.
#include <cuda_runtime.h>

#include <climits>
#include <cstdint>

#include "common.cuh"

namespace hdrnet_b200 {

constexpr int kZsThreads = 288;      // 9 warps: 1152 sorted slots >= 960 px + 3 * 64 padding
constexpr int kZsWarps = kZsThreads / 32;
constexpr int kZsMaxBuckets = 64;    // (x cells in a segment) * (gd + 1)
constexpr int kZsMaxSegPx = 960;
constexpr int kZsSlots = 4 * kZsThreads;
constexpr int kZsStagesMax = 8;
constexpr int kZsGc = 12;

struct ZsPlan {
  int ctas, stages, nseg, seg_px, row_floats, smem_bytes;
  int pieces;  // max distinct x cells (gx0 values) inside one segment
  int off_raw, off_slab, off_rec, off_cnt, off_stage, stage_bytes;
};

struct ZsArgs {
  const float* grid;
  const float* guide;
  const float* input;
  float* out;
  SliceGeom g;
  ZsPlan p;
};

__device__ __forceinline__ float4 zs_lerp4(float w0, float4 a, float w1, float4 b) {
  return make_float4(fmaf(w1, b.x, w0 * a.x), fmaf(w1, b.y, w0 * a.y), fmaf(w1, b.z, w0 * a.z),
                     fmaf(w1, b.w, w0 * a.w));
}

__global__ void __launch_bounds__(kZsThreads, 2)
slice_apply_zsort_kernel(const ZsArgs args) {
  extern __shared__ __align__(128) unsigned char smem[];
  const SliceGeom& g = args.g;
  const ZsPlan& pl = args.p;
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const unsigned full_mask = 0xffffffffu;
  const unsigned lt_mask = (1u << lane) - 1u;

  uint64_t* full = reinterpret_cast<uint64_t*>(smem);  // [kZsStagesMax]
  uint64_t* gridbar = full + kZsStagesMax;
  float* raw0 = reinterpret_cast<float*>(smem + pl.off_raw);
  float* raw1 = raw0 + pl.row_floats;
  float* slab = reinterpret_cast<float*>(smem + pl.off_slab);
  int2* rec = reinterpret_cast<int2*>(smem + pl.off_rec);  // [kZsSlots] {fz bits, idx | key<<16}
  int* cnt = reinterpret_cast<int*>(smem + pl.off_cnt);    // [kZsWarps][kZsMaxBuckets]
  unsigned char* stage_base = smem + pl.off_stage;

  const long long total_rows = static_cast<long long>(g.B) * g.rows;
  const long long r_begin = total_rows * blockIdx.x / gridDim.x;
  const long long r_end = total_rows * (blockIdx.x + 1) / gridDim.x;
  const int nitems = static_cast<int>(r_end - r_begin) * pl.nseg;
  if (nitems <= 0) return;

  if (tid == 0) {
    for (int s = 0; s < pl.stages; ++s) mbar_init(&full[s], 1);
    mbar_init(gridbar, 1);
    fence_mbar_init();
  }
  __syncthreads();

  const int NS = pl.stages;
  const int seg_rgb_bytes_max = pl.seg_px * 12;
  auto stage_rgb = [&](int s) { return stage_base + static_cast<size_t>(s) * pl.stage_bytes; };
  auto stage_guide = [&](int s) {
    return stage_base + static_cast<size_t>(s) * pl.stage_bytes + seg_rgb_bytes_max;
  };
  auto item_span = [&](int item, long long& row, int& x0, int& npx) {
    const int rr = item / pl.nseg;
    const int seg = item - rr * pl.nseg;
    row = r_begin + rr;
    x0 = seg * pl.seg_px;
    npx = min(pl.seg_px, g.W - x0);
  };
  auto issue_load = [&](int item) {  // thread 0 only
    long long row; int x0, npx;
    item_span(item, row, x0, npx);
    const int s = item % NS;
    const size_t pix = static_cast<size_t>(row) * g.W + x0;
    mbar_expect_tx(&full[s], static_cast<uint32_t>(npx) * 16u);
    tma_load_1d(stage_rgb(s), args.input + pix * 3, static_cast<uint32_t>(npx) * 12u, &full[s]);
    tma_load_1d(stage_guide(s), args.guide + pix, static_cast<uint32_t>(npx) * 4u, &full[s]);
  };
  if (tid == 0) {
    const int pre = min(NS - 1, nitems);
    for (int it = 0; it < pre; ++it) issue_load(it);
  }

  const float gd_f = static_cast<float>(g.gd);
  const int x_stride = g.gd * kZsGc;
  const int nzb = g.gd + 1;                 // depth buckets: clamp(z0, -1, gd-1) + 1
  int cur_b = -1, cur_gy0 = INT_MIN;
  uint32_t grid_phase = 0;

  for (int item = 0; item < nitems; ++item) {
    long long row; int x0, npx;
    item_span(item, row, x0, npx);

    if (x0 == 0) {  // new image row: grid rows + y pre-blend (identical to the row kernel)
      const int b = static_cast<int>(row / g.rows);
      const int y = g.y_off + static_cast<int>(row - static_cast<long long>(b) * g.rows);
      const Axis ay = spatial_axis(y, g.scale_y);
      if (b != cur_b || ay.i0 != cur_gy0) {
        if (tid == 0) {
          const int gy0c = clampi(ay.i0, 0, g.gh - 1);
          const int gy1c = clampi(ay.i0 + 1, 0, g.gh - 1);
          const float* gb = args.grid + static_cast<size_t>(b) * g.gh * pl.row_floats;
          const uint32_t bytes = static_cast<uint32_t>(pl.row_floats) * 4u;
          mbar_expect_tx(gridbar, 2u * bytes);
          tma_load_1d(raw0, gb + static_cast<size_t>(gy0c) * pl.row_floats, bytes, gridbar);
          tma_load_1d(raw1, gb + static_cast<size_t>(gy1c) * pl.row_floats, bytes, gridbar);
        }
        mbar_wait(gridbar, grid_phase);
        grid_phase ^= 1u;
        cur_b = b;
        cur_gy0 = ay.i0;
      }
      const float wy1 = ay.f, wy0 = 1.0f - ay.f;
      const float4* a4 = reinterpret_cast<const float4*>(raw0);
      const float4* b4 = reinterpret_cast<const float4*>(raw1);
      float4* s4 = reinterpret_cast<float4*>(slab);
      for (int e = tid; e < pl.row_floats / 4; e += kZsThreads) s4[e] = zs_lerp4(wy0, a4[e], wy1, b4[e]);
      // (the barrier after phase 1 below orders these writes before any slab read)
    }

    const int s = item % NS;
    // First x cell of the segment: every pixel's gx0 lies in [gx_first, gx_first + pieces).
    const int gx_first = spatial_axis(x0, g.scale_x).i0;

    // Clear this warp's bucket counters and the sorted-slot index words (padding = -1).
    cnt[warp * kZsMaxBuckets + lane] = 0;
    cnt[warp * kZsMaxBuckets + lane + 32] = 0;
    reinterpret_cast<int4*>(rec)[2 * tid] = make_int4(0, -1, 0, -1);
    reinterpret_cast<int4*>(rec)[2 * tid + 1] = make_int4(0, -1, 0, -1);

    mbar_wait(&full[s], static_cast<uint32_t>(item / NS) & 1u);
    __syncwarp();

    // ---- phase 1: key + rank of this thread's 4 original pixels -------------------------
    const bool have = tid * 4 < npx;
    int key[4], off[4];
    float fzv[4];
    {
      float4 gq = make_float4(0.f, 0.f, 0.f, 0.f);
      if (have) gq = reinterpret_cast<const float4*>(stage_guide(s))[tid];
      const float gv[4] = {gq.x, gq.y, gq.z, gq.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const Axis az = range_axis(gv[k], gd_f);
        const Axis ax = spatial_axis(x0 + 4 * tid + k, g.scale_x);
        fzv[k] = az.f;
        const int kz = clampi(az.i0, -1, g.gd - 1) + 1;
        key[k] = have ? (ax.i0 - gx_first) * nzb + kz : 0x7fff;
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const unsigned m = __match_any_sync(full_mask, key[k]);
        const int r = __popc(m & lt_mask);
        const int leader = __ffs(m) - 1;
        int old = 0;
        if (r == 0 && have) {
          int* c = cnt + warp * kZsMaxBuckets + key[k];
          old = *c;
          *c = old + __popc(m);
        }
        old = __shfl_sync(full_mask, old, leader);
        off[k] = old + r;
        __syncwarp();
      }
    }
    __syncthreads();  // counters of all warps complete; slab + cleared records visible

    // ---- bucket bases: padded exclusive scan over buckets, plus earlier warps' counts -----
    int base_lo, base_hi, nslots;
    {
      int tot_lo = 0, tot_hi = 0, pre_lo = 0, pre_hi = 0;
#pragma unroll
      for (int w = 0; w < kZsWarps; ++w) {
        const int c0 = cnt[w * kZsMaxBuckets + lane];
        const int c1 = cnt[w * kZsMaxBuckets + lane + 32];
        if (w < warp) { pre_lo += c0; pre_hi += c1; }
        tot_lo += c0; tot_hi += c1;
      }
      const int pad_lo = (tot_lo + 3) & ~3, pad_hi = (tot_hi + 3) & ~3;
      int inc_lo = pad_lo, inc_hi = pad_hi;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int a = __shfl_up_sync(full_mask, inc_lo, d);
        const int b2 = __shfl_up_sync(full_mask, inc_hi, d);
        if (lane >= d) { inc_lo += a; inc_hi += b2; }
      }
      const int sum_lo = __shfl_sync(full_mask, inc_lo, 31);
      nslots = sum_lo + __shfl_sync(full_mask, inc_hi, 31);
      base_lo = inc_lo - pad_lo + pre_lo;
      base_hi = sum_lo + inc_hi - pad_hi + pre_hi;
    }
    // ---- scatter the 8-byte records into sorted order ------------------------------------
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int kk = key[k] & 63;
      const int blo = __shfl_sync(full_mask, base_lo, kk & 31);
      const int bhi = __shfl_sync(full_mask, base_hi, kk & 31);
      if (have) {
        const int dest = ((kk < 32) ? blo : bhi) + off[k];
        rec[dest] = make_int2(__float_as_int(fzv[k]), (4 * tid + k) | (key[k] << 16));
      }
    }
    __syncthreads();  // sorted records complete

    // ---- phase 2: 4 consecutive sorted pixels share one set of corner vectors ------------
    if (tid * 4 < nslots) {
      const int4 ra = reinterpret_cast<const int4*>(rec)[2 * tid];      // slots 4t, 4t+1
      const int4 rb = reinterpret_cast<const int4*>(rec)[2 * tid + 1];  // slots 4t+2, 4t+3
      const int ridx[4] = {ra.y, ra.w, rb.y, rb.w};
      const float rfz[4] = {__int_as_float(ra.x), __int_as_float(ra.z), __int_as_float(rb.x),
                            __int_as_float(rb.z)};
      // slot 4t is always a real pixel (buckets are padded at their END)
      const int ckey = ridx[0] >> 16;
      const int piece = ckey / nzb;
      const int kz = ckey - piece * nzb;
      const int gx0 = gx_first + piece;
      const int xo0 = clampi(gx0, 0, g.gw - 1) * x_stride;
      const int xo1 = clampi(gx0 + 1, 0, g.gw - 1) * x_stride;
      const int zo0 = clampi(kz - 1, 0, g.gd - 1) * kZsGc;
      const int zo1 = clampi(kz, 0, g.gd - 1) * kZsGc;
      unsigned long long v00[6], v01[6], v10[6], v11[6];
      {
        const ulonglong2* p00 = reinterpret_cast<const ulonglong2*>(slab + xo0 + zo0);
        const ulonglong2* p01 = reinterpret_cast<const ulonglong2*>(slab + xo0 + zo1);
        const ulonglong2* p10 = reinterpret_cast<const ulonglong2*>(slab + xo1 + zo0);
        const ulonglong2* p11 = reinterpret_cast<const ulonglong2*>(slab + xo1 + zo1);
#pragma unroll
        for (int q = 0; q < 3; ++q) {
          const ulonglong2 a = p00[q], b2 = p01[q], c = p10[q], d = p11[q];
          v00[2 * q] = a.x; v00[2 * q + 1] = a.y;
          v01[2 * q] = b2.x; v01[2 * q + 1] = b2.y;
          v10[2 * q] = c.x; v10[2 * q + 1] = c.y;
          v11[2 * q] = d.x; v11[2 * q + 1] = d.y;
        }
      }
      float* rgb = reinterpret_cast<float*>(stage_rgb(s));
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (ridx[k] < 0) continue;  // bucket padding
        const int idx = ridx[k] & 0xffff;
        float* px = rgb + 3 * idx;
        const float r = px[0], gg = px[1], bb = px[2];
        const Axis ax = spatial_axis(x0 + idx, g.scale_x);
        float wz0, wz1;
        smoothed_weights(rfz[k], wz0, wz1);
        const float wx1 = ax.f, wx0 = 1.0f - ax.f;
        const float w00 = wx0 * wz0, w01 = wx0 * wz1, w10 = wx1 * wz0, w11 = wx1 * wz1;
        const unsigned long long W00 = pack2(w00, w00), W01 = pack2(w01, w01);
        const unsigned long long W10 = pack2(w10, w10), W11 = pack2(w11, w11);
        unsigned long long acc[6];
#pragma unroll
        for (int q = 0; q < 6; ++q)
          acc[q] = fma2(W11, v11[q], fma2(W10, v10[q], fma2(W01, v01[q], mul2(W00, v00[q]))));
        float a0, a1, a2, a3;
        unpack2(acc[0], a0, a1);
        unpack2(acc[1], a2, a3);
        const float o_r = fmaf(a2, bb, fmaf(a1, gg, fmaf(a0, r, a3)));
        unpack2(acc[2], a0, a1);
        unpack2(acc[3], a2, a3);
        const float o_g = fmaf(a2, bb, fmaf(a1, gg, fmaf(a0, r, a3)));
        unpack2(acc[4], a0, a1);
        unpack2(acc[5], a2, a3);
        const float o_b = fmaf(a2, bb, fmaf(a1, gg, fmaf(a0, r, a3)));
        px[0] = o_r; px[1] = o_g; px[2] = o_b;
      }
      fence_proxy_async_smem();
    }
    __syncthreads();

    if (tid == 0) {
      const size_t pix = static_cast<size_t>(row) * g.W + x0;
      tma_store_1d(args.out + pix * 3, stage_rgb(s), static_cast<uint32_t>(npx) * 12u);
      tma_store_commit();
      const int nxt = item + NS - 1;
      if (nxt < nitems) {
        tma_store_wait_read<1>();
        issue_load(nxt);
      }
    }
  }
  if (tid == 0) tma_store_wait_all<0>();
}

// ---- host side ------------------------------------------------------------------------------
static inline int zs_round_up(int v, int m) { return (v + m - 1) / m * m; }

// Returns false when the shapes do not suit the z-bucketed kernel (too many buckets per
// segment, W not a multiple of 4, ...).
bool make_zsort_plan(const SliceGeom& g, int max_smem, int sms, ZsPlan* out) {
  if (g.W < 4 || (g.W % 4) != 0) return false;
  const int nzb = g.gd + 1;
  if (nzb > 32) return false;
  const int max_pieces = kZsMaxBuckets / nzb;
  if (max_pieces < 3) return false;
  // A segment of S pixels spans at most floor(S * gw / W) + 2 distinct gx0 values.
  const double cell_px = static_cast<double>(g.W) / g.gw;
  int max_seg = static_cast<int>((max_pieces - 2) * cell_px) / 4 * 4;
  if (max_seg > kZsMaxSegPx) max_seg = kZsMaxSegPx;
  if (max_seg < 256) return false;  // sorting tiny segments does not pay
  ZsPlan p;
  const int quads = g.W / 4;
  const int max_quads = max_seg / 4;
  p.nseg = (quads + max_quads - 1) / max_quads;
  p.seg_px = 4 * ((quads + p.nseg - 1) / p.nseg);
  p.pieces = static_cast<int>(p.seg_px / cell_px) + 2;
  if (p.pieces > max_pieces) p.pieces = max_pieces;  // guaranteed by max_seg; keeps nbuckets <= 64
  p.row_floats = g.gw * g.gd * kZsGc;
  p.stage_bytes = zs_round_up(p.seg_px * 16, 128);
  p.off_raw = 128;
  p.off_slab = p.off_raw + zs_round_up(2 * p.row_floats * 4, 128);
  p.off_rec = p.off_slab + zs_round_up(p.row_floats * 4, 128);
  p.off_cnt = p.off_rec + kZsSlots * 8;
  p.off_stage = p.off_cnt + kZsWarps * kZsMaxBuckets * 4;
  const int per_cta_2 = (max_smem + 1024) / 2 - 1024;
  int stages = 0;
  for (int ns = 4; ns >= 2; --ns)
    if (p.off_stage + ns * p.stage_bytes <= per_cta_2) { stages = ns; break; }
  if (stages == 0) return false;  // only worth it at two CTAs per SM
  p.stages = stages;
  p.smem_bytes = p.off_stage + stages * p.stage_bytes;
  const long long total_rows = static_cast<long long>(g.B) * g.rows;
  long long ctas = static_cast<long long>(sms) * 2;
  if (ctas > total_rows) ctas = total_rows;
  p.ctas = static_cast<int>(ctas);
  *out = p;
  return true;
}

int launch_zsort(const float* grid, const float* guide, const float* input, float* out,
                 const SliceGeom& g, const ZsPlan& plan, cudaStream_t stream) {
  cudaError_t e = cudaFuncSetAttribute(slice_apply_zsort_kernel,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       plan.smem_bytes);
  if (e != cudaSuccess) return static_cast<int>(e);
  ZsArgs a;
  a.grid = grid; a.guide = guide; a.input = input; a.out = out; a.g = g; a.p = plan;
  slice_apply_zsort_kernel<<<plan.ctas, kZsThreads, plan.smem_bytes, stream>>>(a);
  return static_cast<int>(cudaGetLastError());
}

}  // namespace hdrnet_b200
