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
// This first version is deliberately unpipelined (one chunk in flight): the tensor pipe idles
// while threads gather.  It is opt-in (HDRNET_CONV_TCGEN05=1) until it is double-buffered.
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
