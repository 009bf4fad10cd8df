// conv_tcgen05.cu -- 3x3 / 1x1 convolution of the coefficient network as an implicit GEMM on
// the 5th-generation tensor cores (tcgen05.mma, accumulator in TMEM), with fp32 parity.
//
// The north star asks for the low-res branch's dense contractions on tcgen05 "only because
// they are genuine dense contractions".  A plain TF32 (10-bit mantissa) product would lose the
// float32 parity the rest of the path keeps, so every operand is split
//     a = a_hi + a_lo,  a_hi = a with the low 13 mantissa bits cleared (exactly a TF32 value),
//                       a_lo = a - a_hi (exact in fp32; its own TF32 truncation is 2^-21 of a)
// and three MMAs accumulate a_hi*b_hi + a_hi*b_lo + a_lo*b_hi into the same fp32 TMEM
// accumulator ("3xTF32"): relative error ~1e-6 per product, i.e. float32-grade results from
// the tensor pipe at 3x the MMA count -- irrelevant here, the network is latency-bound.
//
// Shape: out[m][n] = sum_k A[m][k] * W[k][n] with m = output pixel (b, oy, ox), n = output
// channel, k = (ky, kx, ci).  One CTA (128 threads) owns a 128-pixel x N tile, N = Cout <= 256:
//   * A and W chunks (32 k-values) are gathered by the threads (im2col with TF SAME padding),
//     split hi/lo and written to shared memory in the canonical K-major no-swizzle UMMA
//     layout (8-row x 16-byte core matrices; SBO = 128 B between 8-row groups, LBO = rows*16 B
//     between 16-byte k-chunks);
//   * one elected thread issues 4 k-steps x 3 tcgen05.mma.kind::tf32 (M=128, N, K=8) per chunk
//     and a tcgen05.commit onto an mbarrier that gates the next chunk's overwrite;
//   * epilogue: tcgen05.ld 32x32b (each warp its 32 TMEM lanes = 32 pixels), bias + ReLU,
//     float4 stores.
// Two kernels: conv2d_tcgen05_kernel gathers the weights itself, one chunk in flight (any HWIO
// weight buffer); conv2d_tcgen05_packed_kernel takes weights PRE-PACKED per 32-k chunk into the
// canonical layout, already split hi/lo (hdrnet_conv2d_tc_pack_f32, once per model), so that a
// chunk's B operand is ONE TMA bulk copy, and runs a 3-stage ring: the im2col gather of chunk
// c+1 is prefetched into registers while chunk c's MMAs run asynchronously, each stage being
// recycled on the mbarrier its tcgen05.commit arrives on.
#include <cuda_runtime.h>

#include <cstdint>

#include "common.cuh"

namespace hdrnet_b200 {

constexpr int kTcThreads = 128;
constexpr int kTcM = 128;        // pixels per CTA (= TMEM lanes)
constexpr int kTcKc = 32;        // k-values per staged chunk (8 x 16-byte k-chunks, 4 MMA k-steps)
constexpr int kTcMaxN = 256;

struct TcConvArgs {
  const float* in;
  const float* w;     // [k][k][Cin][Cout]
  const float* bias;  // [Cout] or nullptr
  float* out;
  int B, H, W, Cin, OH, OW, Cout, k, stride, pad_t, pad_l, relu;
  int ncols;          // TMEM columns allocated (power of two >= max(32, Cout))
};

__device__ __forceinline__ uint64_t make_kmajor_desc(uint32_t smem_addr, uint32_t lbo_bytes,
                                                     uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3fffu);          // start address   [0,14)
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3fffu) << 16;    // leading offset  [16,30)
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3fffu) << 32;    // stride offset   [32,46)
  d |= static_cast<uint64_t>(1) << 46;                             // descriptor version (sm_100)
  return d;                                                        // layout_type = SWIZZLE_NONE
}

__device__ __forceinline__ void split_tf32(float v, float& hi, float& lo) {
  hi = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
  lo = v - hi;
}

__global__ void __launch_bounds__(kTcThreads, 1)
conv2d_tcgen05_kernel(const TcConvArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t mma_bar;
  __shared__ uint32_t tmem_base_smem;

  const int tid = threadIdx.x, warp = tid >> 5;
  const int N = a.Cout;
  float* a_hi = reinterpret_cast<float*>(smem);                       // [8 chunks][128 rows][4]
  float* a_lo = a_hi + kTcM * kTcKc;
  float* b_hi = a_lo + kTcM * kTcKc;                                  // [8 chunks][N rows][4]
  float* b_lo = b_hi + N * kTcKc;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(&tmem_base_smem)),
                 "r"(static_cast<uint32_t>(a.ncols)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    mbar_init(&mma_bar, 1);
    fence_mbar_init();
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem_base = tmem_base_smem;

  // This thread's A row = output pixel.
  const long long total_px = static_cast<long long>(a.B) * a.OH * a.OW;
  const long long q = static_cast<long long>(blockIdx.x) * kTcM + tid;
  const bool pv = q < total_px;
  const long long qq = pv ? q : 0;
  const int ox = static_cast<int>(qq % a.OW);
  const int oy = static_cast<int>((qq / a.OW) % a.OH);
  const int ob = static_cast<int>(qq / (static_cast<long long>(a.OW) * a.OH));

  const int K = a.k * a.k * a.Cin;
  const int nchunks = (K + kTcKc - 1) / kTcKc;
  // instruction descriptor: D = F32, A = B = TF32, both K-major, N >> 3, M >> 4
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(N >> 3) << 17) |
                         (static_cast<uint32_t>(kTcM >> 4) << 24);
  const uint32_t a_lbo = kTcM * 16, b_lbo = static_cast<uint32_t>(N) * 16, sbo = 128;
  uint32_t phase = 0;

  for (int ch = 0; ch < nchunks; ++ch) {
    const int k0 = ch * kTcKc;
    // ---- gather A: row `tid`, 8 chunks of 4 consecutive k (same (ky,kx), consecutive ci) ----
#pragma unroll
    for (int c = 0; c < kTcKc / 4; ++c) {
      const int kk = k0 + 4 * c;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (pv && kk < K) {
        const int t = kk / a.Cin, ci = kk - t * a.Cin;
        const int ky = t / a.k, kx = t - ky * a.k;
        const int iy = oy * a.stride - a.pad_t + ky, ix = ox * a.stride - a.pad_l + kx;
        if (iy >= 0 && iy < a.H && ix >= 0 && ix < a.W)
          v = __ldg(reinterpret_cast<const float4*>(
              a.in + ((static_cast<size_t>(ob) * a.H + iy) * a.W + ix) * a.Cin + ci));
      }
      float4 h, l;
      split_tf32(v.x, h.x, l.x); split_tf32(v.y, h.y, l.y);
      split_tf32(v.z, h.z, l.z); split_tf32(v.w, h.w, l.w);
      const int off = c * (kTcM * 4) + (tid >> 3) * 32 + (tid & 7) * 4;  // floats
      *reinterpret_cast<float4*>(a_hi + off) = h;
      *reinterpret_cast<float4*>(a_lo + off) = l;
    }
    // ---- gather B: element (n, k) = W[k][n]; thread -> (n = e % N, chunk = e / N) --------------
    for (int e = tid; e < N * (kTcKc / 4); e += kTcThreads) {
      const int n = e % N, c = e / N;
      float4 v;
      float* vp = reinterpret_cast<float*>(&v);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int kk = k0 + 4 * c + j;
        vp[j] = (kk < K) ? __ldg(a.w + static_cast<size_t>(kk) * a.Cout + n) : 0.0f;
      }
      float4 h, l;
      split_tf32(v.x, h.x, l.x); split_tf32(v.y, h.y, l.y);
      split_tf32(v.z, h.z, l.z); split_tf32(v.w, h.w, l.w);
      const int off = c * (N * 4) + (n >> 3) * 32 + (n & 7) * 4;
      *reinterpret_cast<float4*>(b_hi + off) = h;
      *reinterpret_cast<float4*>(b_lo + off) = l;
    }
    fence_proxy_async_smem();        // generic-proxy smem writes -> visible to the tensor core
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");

    if (tid == 0) {
#pragma unroll
      for (int ks = 0; ks < kTcKc / 8; ++ks) {
        const uint64_t dah = make_kmajor_desc(smem_u32(a_hi) + ks * 2 * a_lbo, a_lbo, sbo);
        const uint64_t dal = make_kmajor_desc(smem_u32(a_lo) + ks * 2 * a_lbo, a_lbo, sbo);
        const uint64_t dbh = make_kmajor_desc(smem_u32(b_hi) + ks * 2 * b_lbo, b_lbo, sbo);
        const uint64_t dbl = make_kmajor_desc(smem_u32(b_lo) + ks * 2 * b_lbo, b_lbo, sbo);
        const uint32_t acc0 = (ch > 0 || ks > 0) ? 1u : 0u;
        const uint64_t da[3] = {dah, dah, dal};
        const uint64_t db[3] = {dbh, dbl, dbh};
#pragma unroll
        for (int t = 0; t < 3; ++t) {
          const uint32_t acc = (t > 0) ? 1u : acc0;
          asm volatile(
              "{\n\t"
              ".reg .pred p;\n\t"
              "setp.ne.b32 p, %4, 0;\n\t"
              "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
              "}\n" ::"r"(tmem_base),
              "l"(da[t]), "l"(db[t]), "r"(idesc), "r"(acc)
              : "memory");
        }
      }
      // arrives on the mbarrier once every MMA issued so far has completed
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                       smem_u32(&mma_bar))
                   : "memory");
    }
    mbar_wait(&mma_bar, phase);  // operands consumed: the staging buffers may be rewritten
    phase ^= 1u;
    asm volatile("tcgen05.fence::after_thread_sync;");
  }

  // ---- epilogue: TMEM -> registers -> bias / ReLU -> global ------------------------------------
  float* dst = a.out + static_cast<size_t>(qq) * a.Cout;
  for (int n0 = 0; n0 < N; n0 += 16) {
    uint32_t r[16];
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + n0;
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
          "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
          "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (pv) {
#pragma unroll
      for (int j = 0; j < 16; j += 4) {
        float4 o;
        float* op = reinterpret_cast<float*>(&o);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float v = __uint_as_float(r[j + e]) + (a.bias ? __ldg(a.bias + n0 + j + e) : 0.0f);
          op[e] = a.relu ? fmaxf(v, 0.0f) : v;
        }
        *reinterpret_cast<float4*>(dst + n0 + j) = o;
      }
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"(static_cast<uint32_t>(a.ncols)));
  }
}


// =========================================================================================
// Pipelined form with pre-packed weights
// =========================================================================================
constexpr int kTcStages = 3;

// packed[chunk][half: hi, lo][kchunk 0..7][n 0..N-1][4]: chunk c's B operand (both halves) is one
// contiguous block of 2 * N * 32 floats in exactly the shared-memory layout the MMA reads.
__global__ void __launch_bounds__(256)
conv_tc_pack_kernel(const float* __restrict__ w, float* __restrict__ packed, int K, int N,
                    int nchunks) {
  const long long total = static_cast<long long>(nchunks) * 8 * N;
  for (long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; e < total;
       e += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int n = static_cast<int>(e % N);
    const int c = static_cast<int>((e / N) % 8);
    const int ch = static_cast<int>(e / (static_cast<long long>(N) * 8));
    float4 h, l;
    float* hp = reinterpret_cast<float*>(&h);
    float* lp = reinterpret_cast<float*>(&l);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int kk = ch * kTcKc + 4 * c + j;
      const float v = (kk < K) ? __ldg(w + static_cast<size_t>(kk) * N + n) : 0.0f;
      split_tf32(v, hp[j], lp[j]);
    }
    float* base = packed + static_cast<size_t>(ch) * 2 * N * kTcKc;
    const int off = c * (N * 4) + (n >> 3) * 32 + (n & 7) * 4;
    *reinterpret_cast<float4*>(base + off) = h;
    *reinterpret_cast<float4*>(base + N * kTcKc + off) = l;
  }
}

struct TcPackedArgs {
  const float* in;
  const float* packed;  // conv_tc_pack_kernel output
  const float* bias;
  float* out;
  int B, H, W, Cin, OH, OW, Cout, k, stride, pad_t, pad_l, relu, ncols;
};

__global__ void __launch_bounds__(kTcThreads, 1)
conv2d_tcgen05_packed_kernel(const TcPackedArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t b_full[kTcStages];    // TMA: weights of the stage have landed
  __shared__ uint64_t mma_done[kTcStages];  // tcgen05.commit: the stage's operands were consumed
  __shared__ uint32_t tmem_base_smem;

  const int tid = threadIdx.x, warp = tid >> 5;
  const int N = a.Cout;
  const int a_floats = kTcM * kTcKc;         // one half (hi or lo) of the A stage
  const int b_floats = N * kTcKc;
  const int stage_floats = 2 * a_floats + 2 * b_floats;
  float* ring = reinterpret_cast<float*>(smem);

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(&tmem_base_smem)),
                 "r"(static_cast<uint32_t>(a.ncols)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    for (int s = 0; s < kTcStages; ++s) { mbar_init(&b_full[s], 1); mbar_init(&mma_done[s], 1); }
    fence_mbar_init();
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem_base = tmem_base_smem;

  const long long total_px = static_cast<long long>(a.B) * a.OH * a.OW;
  const long long q = static_cast<long long>(blockIdx.x) * kTcM + tid;
  const bool pv = q < total_px;
  const long long qq = pv ? q : 0;
  const int ox = static_cast<int>(qq % a.OW);
  const int oy = static_cast<int>((qq / a.OW) % a.OH);
  const int ob = static_cast<int>(qq / (static_cast<long long>(a.OW) * a.OH));
  const int K = a.k * a.k * a.Cin;
  const int nchunks = (K + kTcKc - 1) / kTcKc;
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(N >> 3) << 17) |
                         (static_cast<uint32_t>(kTcM >> 4) << 24);
  const uint32_t a_lbo = kTcM * 16, b_lbo = static_cast<uint32_t>(N) * 16, sbo = 128;
  const uint32_t b_bytes = static_cast<uint32_t>(2 * b_floats) * 4u;
  const uint32_t acc_stride = static_cast<uint32_t>(a.ncols) / 4u;  // columns between accumulators

  // im2col gather of this thread's row for the NEXT chunk in sequence, into registers.  The
  // (ky, kx, ci) position advances incrementally -- 4 channels per 16-byte k-chunk -- because
  // runtime integer divisions here cost ~900 instructions per warp per chunk (ncu).
  int g_ky = 0, g_kx = 0, g_ci = 0;
  const int iy0 = oy * a.stride - a.pad_t, ix0 = ox * a.stride - a.pad_l;
  const float* in_b = a.in + static_cast<size_t>(ob) * a.H * a.W * a.Cin;
  auto gather = [&](float4 (&v)[kTcKc / 4]) {
#pragma unroll
    for (int c = 0; c < kTcKc / 4; ++c) {
      v[c] = make_float4(0.f, 0.f, 0.f, 0.f);
      const int iy = iy0 + g_ky, ix = ix0 + g_kx;
      if (pv && g_ky < a.k && iy >= 0 && iy < a.H && ix >= 0 && ix < a.W)
        v[c] = __ldg(reinterpret_cast<const float4*>(
            in_b + (static_cast<size_t>(iy) * a.W + ix) * a.Cin + g_ci));
      g_ci += 4;
      if (g_ci >= a.Cin) {
        g_ci = 0;
        if (++g_kx == a.k) { g_kx = 0; ++g_ky; }  // g_ky == k: past K, zero-fill
      }
    }
  };

  auto stage_b = [&](int st) { return ring + static_cast<size_t>(st) * stage_floats + 2 * a_floats; };
  auto issue_b = [&](int ch) {  // thread 0: weights of chunk `ch` -> its stage, one TMA bulk copy
    const int st = ch % kTcStages;
    mbar_expect_tx(&b_full[st], b_bytes);
    tma_load_1d(stage_b(st), a.packed + static_cast<size_t>(ch) * 2 * b_floats, b_bytes, &b_full[st]);
  };
  // Weights run two chunks ahead of the MMAs, the im2col gather one chunk ahead (in registers).
  if (tid == 0) {
    issue_b(0);
    if (nchunks > 1) issue_b(1);
  }
  float4 cur[kTcKc / 4], nxt[kTcKc / 4];
  gather(cur);
  for (int ch = 0; ch < nchunks; ++ch) {
    const int s = ch % kTcStages;
    const uint32_t use = static_cast<uint32_t>(ch / kTcStages);
    float* a_hi = ring + static_cast<size_t>(s) * stage_floats;
    float* a_lo = a_hi + a_floats;
    // the MMAs that read this stage kTcStages chunks ago must have completed
    if (use > 0) mbar_wait(&mma_done[s], (use - 1) & 1u);
    if (ch + 1 < nchunks) gather(nxt);  // loads in flight across the stores + barrier
#pragma unroll
    for (int c = 0; c < kTcKc / 4; ++c) {
      float4 h, l;
      split_tf32(cur[c].x, h.x, l.x); split_tf32(cur[c].y, h.y, l.y);
      split_tf32(cur[c].z, h.z, l.z); split_tf32(cur[c].w, h.w, l.w);
      const int off = c * (kTcM * 4) + (tid >> 3) * 32 + (tid & 7) * 4;
      *reinterpret_cast<float4*>(a_hi + off) = h;
      *reinterpret_cast<float4*>(a_lo + off) = l;
    }
    fence_proxy_async_smem();
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    if (tid == 0) {
      mbar_wait(&b_full[s], use & 1u);
      const uint32_t ah = smem_u32(a_hi), al = smem_u32(a_lo), bh = smem_u32(stage_b(s));
      const uint32_t bl = bh + static_cast<uint32_t>(b_floats) * 4u;
#pragma unroll
      for (int ks = 0; ks < kTcKc / 8; ++ks) {
        const uint64_t dah = make_kmajor_desc(ah + ks * 2 * a_lbo, a_lbo, sbo);
        const uint64_t dal = make_kmajor_desc(al + ks * 2 * a_lbo, a_lbo, sbo);
        const uint64_t dbh = make_kmajor_desc(bh + ks * 2 * b_lbo, b_lbo, sbo);
        const uint64_t dbl = make_kmajor_desc(bl + ks * 2 * b_lbo, b_lbo, sbo);
        const uint64_t da[3] = {dah, dah, dal};
        const uint64_t db[3] = {dbh, dbl, dbh};
#pragma unroll
        for (int t = 0; t < 3; ++t) {
          // hi*hi, hi*lo and lo*hi go to THREE accumulators (TMEM columns t*acc_stride..):
          // into one accumulator every MMA would wait for the previous one (~128 clk each).
          const uint32_t acc = (ch > 0 || ks > 0) ? 1u : 0u;
          asm volatile(
              "{\n\t"
              ".reg .pred p;\n\t"
              "setp.ne.b32 p, %4, 0;\n\t"
              "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
              "}\n" ::"r"(tmem_base + static_cast<uint32_t>(t) * acc_stride),
              "l"(da[t]), "l"(db[t]), "r"(idesc), "r"(acc)
              : "memory");
        }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                       smem_u32(&mma_done[s]))
                   : "memory");
      // refill the stage chunk ch+2 will use: its last reader was chunk ch-1
      if (ch + 2 < nchunks) {
        if (ch >= 1) {
          const int pc = ch - 1;
          mbar_wait(&mma_done[pc % kTcStages], static_cast<uint32_t>(pc / kTcStages) & 1u);
        }
        issue_b(ch + 2);
      }
    }
#pragma unroll
    for (int c = 0; c < kTcKc / 4; ++c) cur[c] = nxt[c];
  }
  // all MMAs complete when the last chunk's commit has arrived (commits are ordered)
  {
    const int last = nchunks - 1;
    mbar_wait(&mma_done[last % kTcStages], static_cast<uint32_t>(last / kTcStages) & 1u);
    asm volatile("tcgen05.fence::after_thread_sync;");
  }

  float* dst = a.out + static_cast<size_t>(qq) * a.Cout;
  for (int n0 = 0; n0 < N; n0 += 16) {
    float sum[16];
#pragma unroll
    for (int t = 0; t < 3; ++t) {
      uint32_t r[16];
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16) +
                             static_cast<uint32_t>(t) * acc_stride + n0;
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
            "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
            "=r"(r[14]), "=r"(r[15])
          : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int e = 0; e < 16; ++e) sum[e] = (t == 0) ? __uint_as_float(r[e]) : sum[e] + __uint_as_float(r[e]);
    }
    if (pv) {
#pragma unroll
      for (int j = 0; j < 16; j += 4) {
        float4 o;
        float* op = reinterpret_cast<float*>(&o);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float v = sum[j + e] + (a.bias ? __ldg(a.bias + n0 + j + e) : 0.0f);
          op[e] = a.relu ? fmaxf(v, 0.0f) : v;
        }
        *reinterpret_cast<float4*>(dst + n0 + j) = o;
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"(static_cast<uint32_t>(a.ncols)));
  }
}

static bool tc_shape_ok(int Cin, int Cout) {
  return Cin % 4 == 0 && Cout % 16 == 0 && Cout <= kTcMaxN && Cout >= 16;
}

// Returns HDRNET_E_UNSUPPORTED when the shapes do not suit this path (the caller then runs the
// CUDA-core kernel in cnn.cu).
int launch_conv_tcgen05(const float* in, const float* w, const float* bias, float* out, int B,
                        int H, int W, int Cin, int Cout, int k, int stride, int relu, int OH,
                        int OW, int pad_t, int pad_l, cudaStream_t stream) {
  if (Cin % 4 != 0 || Cout % 16 != 0 || Cout > kTcMaxN || Cout < 16) return HDRNET_E_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(in) & 15u) || (reinterpret_cast<uintptr_t>(out) & 15u))
    return HDRNET_E_UNSUPPORTED;
  TcConvArgs a;
  a.in = in; a.w = w; a.bias = bias; a.out = out;
  a.B = B; a.H = H; a.W = W; a.Cin = Cin; a.OH = OH; a.OW = OW; a.Cout = Cout; a.k = k;
  a.stride = stride; a.pad_t = pad_t; a.pad_l = pad_l; a.relu = relu;
  int ncols = 32;
  while (ncols < Cout) ncols <<= 1;
  a.ncols = ncols;
  const size_t smem = static_cast<size_t>(2) * (kTcM + Cout) * kTcKc * sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(conv2d_tcgen05_kernel,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(smem));
  if (e != cudaSuccess) return static_cast<int>(e);
  const long long total_px = static_cast<long long>(B) * OH * OW;
  conv2d_tcgen05_kernel<<<static_cast<unsigned>((total_px + kTcM - 1) / kTcM), kTcThreads, smem,
                          stream>>>(a);
  return static_cast<int>(cudaGetLastError());
}

}  // namespace hdrnet_b200

using namespace hdrnet_b200;

extern "C" {

size_t hdrnet_conv2d_tc_packed_bytes(int k, int Cin, int Cout) {
  if ((k != 1 && k != 3) || !tc_shape_ok(Cin, Cout)) return 0;
  const int K = k * k * Cin;
  const size_t nchunks = (K + kTcKc - 1) / kTcKc;
  return nchunks * 2 * static_cast<size_t>(Cout) * kTcKc * sizeof(float);
}

int hdrnet_conv2d_tc_pack_f32(const float* w, float* packed, int k, int Cin, int Cout,
                              void* stream) {
  if (hdrnet_conv2d_tc_packed_bytes(k, Cin, Cout) == 0) return HDRNET_E_UNSUPPORTED;
  if (!w || !packed) return HDRNET_E_NULL_POINTER;
  const int K = k * k * Cin, nchunks = (K + kTcKc - 1) / kTcKc;
  const long long total = static_cast<long long>(nchunks) * 8 * Cout;
  conv_tc_pack_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0,
                        static_cast<cudaStream_t>(stream)>>>(w, packed, K, Cout, nchunks);
  return static_cast<int>(cudaGetLastError());
}

int hdrnet_conv2d_nhwc_tc_f32(const float* in, const float* packed_w, const float* bias,
                              float* out, int B, int H, int W, int Cin, int Cout, int k,
                              int stride, int relu, void* stream) {
  if (B < 0 || H < 1 || W < 1 || Cin < 1 || Cout < 1) return HDRNET_E_BAD_SHAPE;
  if ((k != 1 && k != 3) || (stride != 1 && stride != 2) || !tc_shape_ok(Cin, Cout))
    return HDRNET_E_UNSUPPORTED;
  if (B == 0) return HDRNET_OK;
  if (!in || !packed_w || !out) return HDRNET_E_NULL_POINTER;
  if ((reinterpret_cast<uintptr_t>(in) & 15u) || (reinterpret_cast<uintptr_t>(out) & 15u) ||
      (reinterpret_cast<uintptr_t>(packed_w) & 15u))
    return HDRNET_E_UNSUPPORTED;
  TcPackedArgs a;
  a.in = in; a.packed = packed_w; a.bias = bias; a.out = out;
  a.B = B; a.H = H; a.W = W; a.Cin = Cin; a.Cout = Cout; a.k = k; a.stride = stride; a.relu = relu;
  a.OH = (H + stride - 1) / stride;
  a.OW = (W + stride - 1) / stride;
  int tot = (a.OH - 1) * stride + k - H; if (tot < 0) tot = 0; a.pad_t = tot / 2;
  tot = (a.OW - 1) * stride + k - W; if (tot < 0) tot = 0; a.pad_l = tot / 2;
  int ncols = 32;
  while (ncols < Cout) ncols <<= 1;
  a.ncols = 4 * ncols;  // three accumulators at a power-of-two stride (<= 512 columns: Cout <= 128)
  const size_t smem = static_cast<size_t>(kTcStages) * 2 * (kTcM + Cout) * kTcKc * sizeof(float);
  if (smem > 200 * 1024) return HDRNET_E_UNSUPPORTED;  // Cout > 128: use the unpacked kernel
  cudaError_t e = cudaFuncSetAttribute(conv2d_tcgen05_packed_kernel,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(smem));
  if (e != cudaSuccess) return static_cast<int>(e);
  const long long total_px = static_cast<long long>(B) * a.OH * a.OW;
  conv2d_tcgen05_packed_kernel<<<static_cast<unsigned>((total_px + kTcM - 1) / kTcM), kTcThreads,
                                 smem, static_cast<cudaStream_t>(stream)>>>(a);
  return static_cast<int>(cudaGetLastError());
}

}  // extern "C"
