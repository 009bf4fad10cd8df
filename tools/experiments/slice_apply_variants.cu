// slice_apply_variants.cu -- opt-in forms of the texture-assisted row kernel that measured SLOWER
// than the default and are kept, with their parity tests, as documented negative results
// (DESIGN.md section 3): the texture-fed input form (HDRNET_VARIANT_TEX_IN) and the first
// warp-specialised form (HDRNET_VARIANT_TEX_WS).  Shared device code: slice_rows.cuh.
#include <cuda_runtime.h>

#include <climits>
#include <cstdint>
#include <cstdlib>

#include "slice_rows.cuh"

namespace hdrnet_b200 {

// =========================================================================================
// Texture-fed form of the texture-assisted row kernel (HDRNET_VARIANT_TEX_IN).
// =========================================================================================
// The block-synchronous kernel spends 0.25 of its 1.56 shared-memory wavefronts per pixel on
// staging the INPUT: the TMA engine writes 16 B/px into the ring and the threads read them back
// with LDS.128, while the texture pipe idles at 56 %.  Here a thread fetches its 4 pixels (3 RGB
// texels + 1 guide texel, float4 views of the caller's tensors) through the texture pipe
// straight into registers -- issued one item ahead, right before the block barrier, when no
// other value is live -- and shared memory only carries the slab rows and the OUTPUT tiles
// (3 x STS.128 per thread, one bulk store per segment).  Thread 0 asks the L2 for the segments two
// items ahead (cp.async.bulk.prefetch.L2) so that the texture fetches are L2 hits.
// Wavefronts per pixel: 1.29 (LSU) against 1.25 texture-pipe clocks -- the two pipes balanced.

template <int kTexChunks, int kThreads, int kMinBlocks = 2>
__global__ void __launch_bounds__(kThreads, kMinBlocks)
slice_apply_rows_texin_kernel(const TmaArgs args) {
  static_assert(kTexChunks > 0, "slab rows come from the pre-pass workspace");
  extern __shared__ __align__(128) unsigned char smem[];
  const SliceGeom& g = args.g;
  const TmaPlan& pl = args.p;
  const int tid = threadIdx.x;

  uint64_t* gridbar = reinterpret_cast<uint64_t*>(smem);  // [2] slab row landed
  float* raw0 = reinterpret_cast<float*>(smem + pl.off_raw);
  unsigned char* stage_base = smem + pl.off_stage;

  const long long total_rows = static_cast<long long>(g.B) * g.rows;
  const long long r_begin = total_rows * blockIdx.x / gridDim.x;
  const long long r_end = total_rows * (blockIdx.x + 1) / gridDim.x;
  const int nitems = static_cast<int>(r_end - r_begin) * pl.nseg;
  if (nitems <= 0) return;

  if (tid == 0) {
    mbar_init(&gridbar[0], 1);
    mbar_init(&gridbar[1], 1);
    fence_mbar_init();
  }
  __syncthreads();

  auto item_span = [&](int item, long long& row, int& x0, int& npx) {
    const int rr = item / pl.nseg;
    const int seg = item - rr * pl.nseg;
    row = r_begin + rr;
    x0 = seg * pl.seg_px;
    npx = min(pl.seg_px, g.W - x0);
  };
  auto prefetch_l2 = [&](int item) {  // thread 0 only
    long long row; int x0, npx;
    item_span(item, row, x0, npx);
    const size_t pix = static_cast<size_t>(row) * g.W + x0;
    l2_prefetch_bulk(args.input + pix * 12, static_cast<uint32_t>(npx) * 12u);
    l2_prefetch_bulk(args.guide + pix, static_cast<uint32_t>(npx) * 4u);
  };
  // This thread's quad of item `item`: three RGB texels and one guide texel into registers.
  float4 c0, c1, c2, gq;
  auto fetch = [&](int item) {
    long long row; int x0, npx;
    item_span(item, row, x0, npx);
    if (tid * 4 < npx) {
      const int quad = static_cast<int>((row * g.W + x0) >> 2) + tid;   // pixel quad index
      c0 = tex1Dfetch<float4>(args.in_tex, 3 * quad);
      c1 = tex1Dfetch<float4>(args.in_tex, 3 * quad + 1);
      c2 = tex1Dfetch<float4>(args.in_tex, 3 * quad + 2);
      gq = tex1Dfetch<float4>(args.guide_tex, quad);
    }
  };

  if (tid == 0) {
    const uint32_t bytes = static_cast<uint32_t>(pl.row_floats) * 4u;
    mbar_expect_tx(&gridbar[0], bytes);
    tma_load_1d(raw0, args.yslab + static_cast<size_t>(r_begin) * pl.row_floats, bytes, &gridbar[0]);
    prefetch_l2(0);
    if (nitems > 1) prefetch_l2(1);
    if (nitems > 2) prefetch_l2(2);
  }
  fetch(0);

  const float gd_f = static_cast<float>(g.gd);
  const int x_stride = g.gd * kGc;
  const float* slab = raw0;
  int tex_row = 0;

  for (int item = 0; item < nitems; ++item) {
    long long row; int x0, npx;
    item_span(item, row, x0, npx);
    if (x0 == 0) {  // new image row: its slab row (double-buffered one row ahead)
      const int rowk = item / pl.nseg;
      const int cur = rowk & 1;
      slab = raw0 + cur * pl.row_floats;
      mbar_wait(&gridbar[cur], static_cast<uint32_t>(rowk >> 1) & 1u);
      if (tid == 0 && row + 1 < r_end) {
        const uint32_t bytes = static_cast<uint32_t>(pl.row_floats) * 4u;
        mbar_expect_tx(&gridbar[cur ^ 1], bytes);
        tma_load_1d(raw0 + (cur ^ 1) * pl.row_floats,
                    args.yslab + static_cast<size_t>(row + 1) * pl.row_floats, bytes, &gridbar[cur ^ 1]);
      }
      tex_row = static_cast<int>(row) * (pl.row_floats / 4);
    }
    unsigned char* otile = stage_base + static_cast<size_t>(item % kTexInStages) * pl.stage_bytes;

    if (tid * 4 < npx) {
      const float pr[4] = {c0.x, c0.w, c1.z, c2.y};
      const float pg[4] = {c0.y, c1.x, c1.w, c2.z};
      const float pb[4] = {c0.z, c1.y, c2.x, c2.w};
      const float gv[4] = {gq.x, gq.y, gq.z, gq.w};
      float o_r[4], o_g[4], o_b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const Axis ax = spatial_axis(x0 + 4 * tid + i, g.scale_x);
        const Axis az = range_axis(gv[i], gd_f);
        const int xo0 = clampi(ax.i0, 0, g.gw - 1) * x_stride;
        const int xo1 = clampi(ax.i0 + 1, 0, g.gw - 1) * x_stride;
        const int zo0 = clampi(az.i0, 0, g.gd - 1) * kGc;
        const int zo1 = clampi(az.i0 + 1, 0, g.gd - 1) * kGc;
        float wz0, wz1;
        smoothed_weights(az.f, wz0, wz1);
        const float wx1 = ax.f, wx0 = 1.0f - ax.f;
        blend_apply<kTexChunks>(slab, args.slab_tex, tex_row, xo0 + zo0, xo0 + zo1, xo1 + zo0,
                                xo1 + zo1, wx0 * wz0, wx0 * wz1, wx1 * wz0, wx1 * wz1, pr[i], pg[i],
                                pb[i], o_r[i], o_g[i], o_b[i]);
      }
      store_quad<kPxF32>(otile, tid, o_r, o_g, o_b);
      fence_proxy_async_smem();
    }
    // Next item's pixels: issued here, when nothing else is live; they land during the barrier.
    if (item + 1 < nitems) fetch(item + 1);
    // Thread 0: the output stage the NEXT item writes must have been drained by its last store.
    if (tid == 0) tma_store_wait_read<kTexInStages - 2>();
    __syncthreads();

    if (tid == 0) {
      const size_t pix = static_cast<size_t>(row) * g.W + x0;
      tma_store_1d(args.out + pix * 12, otile, static_cast<uint32_t>(npx) * 12u);
      tma_store_commit();
      if (item + 3 < nitems) prefetch_l2(item + 3);
    }
  }
  if (tid == 0) tma_store_wait_all<0>();
}

// =========================================================================================
// Warp-specialised form of the texture-assisted row kernel (HDRNET_VARIANT_TEX_WS).
// =========================================================================================
// ncu stall sampling of the block-synchronous kernel: ~20 % of samples sit in synchronisation
// (the per-item __syncthreads before the bulk store, and all 256 threads spinning on the TMA
// barrier).  Here nothing is block-synchronous after start-up:
//   * warp 8 (one lane) is the PRODUCER: it issues every TMA load -- per item the RGB + guide
//     segment into the stage ring, per image row the y-pre-blended slab row (from the pre-pass
//     workspace) into one of two slab buffers -- gated by stage_free[] / slab_free[] mbarriers;
//   * warps 0..7 are CONSUMERS and never wait for each other: a warp waits for its stage
//     (full[]) and slab (slab_full[]), processes its own 128 pixels in place, issues ITS OWN
//     bulk store (lane 0), and one item later -- once cp.async.bulk.wait_group.read says the
//     store has drained the tile -- arrives on stage_free[]; after a row's last item it arrives
//     on slab_free[].  stage_free / slab_free count kWsConsumerWarps arrivals per phase.
constexpr int kTexChunksWs = 4;

template <class GuideFn, int kTexChunks, int kWsConsumerWarps>
__global__ void __launch_bounds__((kWsConsumerWarps + 1) * 32, 2)
slice_apply_rows_ws_kernel(const TmaArgs args, const __grid_constant__ GuideFn guide_fn) {
  static_assert(kTexChunks > 0, "the warp-specialised kernel reads slab rows from the workspace");
  constexpr bool kGuideIn = GuideFn::kFromInput;
  extern __shared__ __align__(128) unsigned char smem[];
  const SliceGeom& g = args.g;
  const TmaPlan& pl = args.p;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  uint64_t* full = reinterpret_cast<uint64_t*>(smem);   // [kMaxStages]  TMA landed
  uint64_t* stage_free = full + kMaxStages;              // [kMaxStages]  all consumers done + stored
  uint64_t* slab_full = stage_free + kMaxStages;         // [2]
  uint64_t* slab_free = slab_full + 2;                   // [2]
  float* raw0 = reinterpret_cast<float*>(smem + pl.off_raw);
  unsigned char* stage_base = smem + pl.off_stage;

  const long long total_rows = static_cast<long long>(g.B) * g.rows;
  const long long r_begin = total_rows * blockIdx.x / gridDim.x;
  const long long r_end = total_rows * (blockIdx.x + 1) / gridDim.x;
  const int nitems = static_cast<int>(r_end - r_begin) * pl.nseg;
  if (nitems <= 0) return;

  if (tid == 0) {
    for (int s = 0; s < pl.stages; ++s) { mbar_init(&full[s], 1); mbar_init(&stage_free[s], kWsConsumerWarps); }
    for (int b = 0; b < 2; ++b) { mbar_init(&slab_full[b], 1); mbar_init(&slab_free[b], kWsConsumerWarps); }
    fence_mbar_init();
  }
  __syncthreads();  // the only block-wide barrier

  const int NS = pl.stages;
  const uint32_t slab_bytes = static_cast<uint32_t>(pl.row_floats) * 4u;
  auto stage_rgb = [&](int s) { return stage_base + static_cast<size_t>(s) * pl.stage_bytes; };
  auto stage_guide = [&](int s) { return stage_rgb(s) + pl.off_guide; };
  auto item_span = [&](int item, long long& row, int& x0, int& npx) {
    const int rr = item / pl.nseg;
    const int seg = item - rr * pl.nseg;
    row = r_begin + rr;
    x0 = seg * pl.seg_px;
    npx = min(pl.seg_px, g.W - x0);
  };
  auto arrive = [&](uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
  };

  if (warp == kWsConsumerWarps) {
    // ------------------------------- producer ---------------------------------------------
    if (lane != 0) return;
    for (int item = 0; item < nitems; ++item) {
      long long row; int x0, npx;
      item_span(item, row, x0, npx);
      const int s = item % NS;
      const int use = item / NS;
      if (x0 == 0) {  // first item of an image row: its slab row, two buffers deep
        const int rowk = item / pl.nseg, rb = rowk & 1, v = rowk >> 1;
        if (v >= 1) mbar_wait(&slab_free[rb], static_cast<uint32_t>(v - 1) & 1u);
        mbar_expect_tx(&slab_full[rb], slab_bytes);
        tma_load_1d(raw0 + rb * pl.row_floats, args.yslab + static_cast<size_t>(row) * pl.row_floats,
                    slab_bytes, &slab_full[rb]);
      }
      if (use >= 1) mbar_wait(&stage_free[s], static_cast<uint32_t>(use - 1) & 1u);
      const size_t pix = static_cast<size_t>(row) * g.W + x0;
      mbar_expect_tx(&full[s], static_cast<uint32_t>(npx) * (kGuideIn ? 16u : 12u));
      tma_load_1d(stage_rgb(s), args.input + pix * 12, static_cast<uint32_t>(npx) * 12u, &full[s]);
      if (kGuideIn)
        tma_load_1d(stage_guide(s), args.guide + pix, static_cast<uint32_t>(npx) * 4u, &full[s]);
    }
    return;
  }

  // --------------------------------- consumers ----------------------------------------------
  const int px0w = warp * 128;   // this warp's pixels inside a segment
  const float* slab = raw0;
  int tex_row = 0;
  for (int item = 0; item < nitems; ++item) {
    long long row; int x0, npx;
    item_span(item, row, x0, npx);
    const int s = item % NS;
    const int rowk = item / pl.nseg, rb = rowk & 1;
    mbar_wait(&full[s], static_cast<uint32_t>(item / NS) & 1u);
    if (x0 == 0) {
      mbar_wait(&slab_full[rb], static_cast<uint32_t>(rowk >> 1) & 1u);
      slab = raw0 + rb * pl.row_floats;
      tex_row = static_cast<int>(row) * (pl.row_floats / 4);
    }
    const int q = (px0w >> 2) + lane;
    if (q * 4 < npx)
      process_quad<GuideFn, kTexChunks>(args, guide_fn, stage_rgb(s), stage_rgb(s), stage_guide(s),
                                        slab, tex_row, row, x0, q);
    __syncwarp();
    if (lane == 0) {
      const int nw = min(128, npx - px0w);
      if (nw > 0) {
        const size_t pix = static_cast<size_t>(row) * g.W + x0 + px0w;
        tma_store_1d(args.out + pix * 12, stage_rgb(s) + static_cast<size_t>(px0w) * 12,
                     static_cast<uint32_t>(nw) * 12u);
      }
      tma_store_commit();            // one (possibly empty) group per item keeps the counting simple
      if (item >= 1) {
        tma_store_wait_read<1>();    // this warp's store of item-1 has drained its tile
        arrive(&stage_free[(item - 1) % NS]);
      }
      if (x0 + pl.seg_px >= g.W) arrive(&slab_free[rb]);  // row finished: slab no longer read here
    }
  }
  if (lane == 0) tma_store_wait_all<0>();
}

template <int kTexChunks, int kThreads, int kMinBlocks = 2>
static int launch_texin(const TmaArgs& a, cudaStream_t stream) {
  auto kern = slice_apply_rows_texin_kernel<kTexChunks, kThreads, kMinBlocks>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       a.p.smem_bytes);
  if (e != cudaSuccess) return static_cast<int>(e);
  kern<<<a.p.ctas, kThreads, a.p.smem_bytes, stream>>>(a);
  return static_cast<int>(cudaGetLastError());
}

template <class GuideFn, int kConsumerWarps>
static int launch_ws_n(const TmaArgs& a, const GuideFn& fn, cudaStream_t stream) {
  auto kern = slice_apply_rows_ws_kernel<GuideFn, kTexChunksWs, kConsumerWarps>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       a.p.smem_bytes);
  if (e != cudaSuccess) return static_cast<int>(e);
  kern<<<a.p.ctas, (kConsumerWarps + 1) * 32, a.p.smem_bytes, stream>>>(a, fn);
  return static_cast<int>(cudaGetLastError());
}

template <class GuideFn>
static int launch_ws(const TmaArgs& a, const GuideFn& fn, cudaStream_t stream) {
  // consumer warps = planned threads / 32 (8 for the 256-thread plan, 15 for the 480-quad one)
  // 15 consumer warps + the producer warp = 512 threads: 64 registers at two CTAs per SM
  if (a.p.threads == 480) return launch_ws_n<GuideFn, 15>(a, fn, stream);
  return launch_ws_n<GuideFn, 8>(a, fn, stream);
}

int launch_texin_form(const TmaArgs& a, int chunks, cudaStream_t stream) {
      if (a.p.threads == 512) {
        switch (chunks) {
          case 3: return launch_texin<3, 512>(a, stream);
          case 5: return launch_texin<5, 512>(a, stream);
          default: return launch_texin<kTexChunksDefault, 512>(a, stream);
        }
      }
      if (a.p.resident == 4) return launch_texin<kTexChunksDefault, kTmaThreads, 4>(a, stream);
      if (a.p.resident == 3) return launch_texin<kTexChunksDefault, kTmaThreads, 3>(a, stream);
      return launch_texin<kTexChunksDefault, kTmaThreads>(a, stream);
}

int launch_ws_form(const TmaArgs& a, cudaStream_t stream) { return launch_ws(a, GuideFromInput{}, stream); }

}  // namespace hdrnet_b200
