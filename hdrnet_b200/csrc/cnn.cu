// cnn.cu -- the low-resolution coefficient network: replaces the TF
// conv / fully_connected layers of HDRNetCurves._coefficients (hdrnet/models.py:62-142,
// hdrnet/layers.py:25-93) with three hand-written fp32 kernels:
//
//   conv2d_nhwc_kernel   k x k (1 or 3), stride 1/2, TF 'SAME' padding (asymmetric for
//                        stride 2 on even extents), HWIO weights, bias + ReLU epilogue.
//                        Batch norm is folded into weights/bias on the host (inference form).
//   fc_kernel            x[B,I] @ W[I,O] + b (+ReLU), weights streamed once for 8 images.
//   fuse_predict_kernel  fusion relu(local + global) (models.py:122-125), the 1x1 prediction
//                        conv (:129-132) and the unroll_grid permutation (:134-139) in one
//                        pass, writing the [B,gh,gw,gd,n_out*n_in] grid slice-apply reads.
//
// The whole network is ~83 MFLOP per image: launch-latency bound, not tensor bound (DESIGN.md
// section 5).  fp32 CUDA-core math keeps the coefficients within float32 round-off of the
// float64-accumulated oracle, which bf16 / tf32 tensor-core math could not.
#include <cooperative_groups.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdlib>

#include "hdrnet_b200.h"

namespace hdrnet_b200 {

constexpr int kConvThreads = 128;  // 32 pixel groups x 4 channel groups

struct ConvArgs {
  const float* in;
  const float* w;     // [k][k][Cin][Cout]
  const float* bias;  // [Cout] or nullptr
  float* out;
  int B, H, W, Cin, OH, OW, Cout, k, stride, pad_t, pad_l, relu, ci_chunk;
  int w_vec;  // Cout % 4 == 0 and 16-byte aligned weights: cp.async staging
};

// Register tile per thread: kConvPx output pixels x kConvCo output channels.  <2, 8> is the
// throughput shape (64 px x 32 channels per CTA); <1, 4> quadruples the CTA count for the
// tiny late layers at batch 1, which are otherwise a handful of CTAs on 148 SMs.
template <int kConvPx, int kConvCo>
__global__ void __launch_bounds__(kConvThreads)
conv2d_nhwc_kernel(const ConvArgs a) {
  constexpr int kConvTilePx = 32 * kConvPx;
  constexpr int kConvTileCo = 4 * kConvCo;
  static_assert(kConvCo == 4 || kConvCo == 8, "channel tile is one or two float4");
  extern __shared__ __align__(16) float wsm[];  // [k*k][ci_chunk][kConvTileCo]
  const int tid = threadIdx.x;
  const int pg = tid & 31, cg = tid >> 5;
  const int co0 = blockIdx.y * kConvTileCo;
  const long long total_px = static_cast<long long>(a.B) * a.OH * a.OW;
  const long long tile_px0 = static_cast<long long>(blockIdx.x) * kConvTilePx;

  // This thread's output pixels.
  int pb[kConvPx], py[kConvPx], px[kConvPx];
  bool pv[kConvPx];
#pragma unroll
  for (int p = 0; p < kConvPx; ++p) {
    const long long q = tile_px0 + pg * kConvPx + p;
    pv[p] = q < total_px;
    const long long qq = pv[p] ? q : 0;
    px[p] = static_cast<int>(qq % a.OW);
    py[p] = static_cast<int>((qq / a.OW) % a.OH);
    pb[p] = static_cast<int>(qq / (static_cast<long long>(a.OW) * a.OH));
  }
  float acc[kConvPx][kConvCo];
#pragma unroll
  for (int p = 0; p < kConvPx; ++p)
#pragma unroll
    for (int c = 0; c < kConvCo; ++c) acc[p][c] = 0.0f;

  const int kk = a.k * a.k;
  const bool vec_in = (a.Cin % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.in) & 15u) == 0);
  for (int ci0 = 0; ci0 < a.Cin; ci0 += a.ci_chunk) {
    const int cn = min(a.ci_chunk, a.Cin - ci0);
    __syncthreads();
    // Stage weights [kk][cn][32 co] (zero-fill channels beyond Cout).  16-byte cp.async
    // (LDGSTS) keeps every copy of the tile in flight at once: the tile is up to 72 KB and a
    // register-staged loop would serialise on global-load latency.
    if (a.w_vec) {
      for (int e4 = tid; e4 < kk * cn * (kConvTileCo / 4); e4 += kConvThreads) {
        const int co = (e4 % (kConvTileCo / 4)) * 4;
        const int ci = (e4 / (kConvTileCo / 4)) % cn;
        const int t = e4 / ((kConvTileCo / 4) * cn);
        float* dst = wsm + static_cast<size_t>(e4) * 4;
        if (co0 + co < a.Cout) {
          const float* src = a.w + (static_cast<size_t>(t) * a.Cin + ci0 + ci) * a.Cout + co0 + co;
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(
                           static_cast<uint32_t>(__cvta_generic_to_shared(dst))),
                       "l"(src)
                       : "memory");
        } else {
          *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
      asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
    } else {
      for (int e = tid; e < kk * cn * kConvTileCo; e += kConvThreads) {
        const int co = e % kConvTileCo;
        const int ci = (e / kConvTileCo) % cn;
        const int t = e / (kConvTileCo * cn);
        const int gco = co0 + co;
        wsm[e] = (gco < a.Cout)
                     ? __ldg(a.w + (static_cast<size_t>(t) * a.Cin + ci0 + ci) * a.Cout + gco)
                     : 0.0f;
      }
    }
    __syncthreads();
    for (int t = 0; t < kk; ++t) {
      const int ky = t / a.k, kx = t - ky * a.k;
      const float* src[kConvPx];
      bool ok[kConvPx];
#pragma unroll
      for (int p = 0; p < kConvPx; ++p) {
        const int iy = py[p] * a.stride - a.pad_t + ky;
        const int ix = px[p] * a.stride - a.pad_l + kx;
        ok[p] = pv[p] && iy >= 0 && iy < a.H && ix >= 0 && ix < a.W;
        src[p] = a.in + ((static_cast<size_t>(pb[p]) * a.H + (ok[p] ? iy : 0)) * a.W +
                         (ok[p] ? ix : 0)) * a.Cin + ci0;
      }
      const float* wt = wsm + static_cast<size_t>(t) * cn * kConvTileCo + cg * kConvCo;
      if (vec_in && (cn % 4 == 0)) {
        for (int ci = 0; ci < cn; ci += 4) {
          float xin[kConvPx][4];
#pragma unroll
          for (int p = 0; p < kConvPx; ++p) {
            const float4 v = ok[p] ? __ldg(reinterpret_cast<const float4*>(src[p] + ci))
                                   : make_float4(0.f, 0.f, 0.f, 0.f);
            xin[p][0] = v.x; xin[p][1] = v.y; xin[p][2] = v.z; xin[p][3] = v.w;
          }
#pragma unroll
          for (int d = 0; d < 4; ++d) {
            const float4 w0 = *reinterpret_cast<const float4*>(wt + (ci + d) * kConvTileCo);
            const float4 w1 = (kConvCo == 8)
                                  ? *reinterpret_cast<const float4*>(wt + (ci + d) * kConvTileCo + 4)
                                  : w0;
            const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
            for (int p = 0; p < kConvPx; ++p)
#pragma unroll
              for (int c = 0; c < kConvCo; ++c) acc[p][c] = fmaf(xin[p][d], wv[c], acc[p][c]);
          }
        }
      } else {
        for (int ci = 0; ci < cn; ++ci) {
          const float4 w0 = *reinterpret_cast<const float4*>(wt + ci * kConvTileCo);
          const float4 w1 = (kConvCo == 8)
                                ? *reinterpret_cast<const float4*>(wt + ci * kConvTileCo + 4)
                                : w0;
          const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
          for (int p = 0; p < kConvPx; ++p) {
            const float x = ok[p] ? __ldg(src[p] + ci) : 0.0f;
#pragma unroll
            for (int c = 0; c < kConvCo; ++c) acc[p][c] = fmaf(x, wv[c], acc[p][c]);
          }
        }
      }
    }
  }

  // Epilogue: bias, ReLU, store.
#pragma unroll
  for (int p = 0; p < kConvPx; ++p) {
    if (!pv[p]) continue;
    float* dst = a.out + ((static_cast<size_t>(pb[p]) * a.OH + py[p]) * a.OW + px[p]) * a.Cout;
#pragma unroll
    for (int c = 0; c < kConvCo; ++c) {
      const int gco = co0 + cg * kConvCo + c;
      if (gco < a.Cout) {
        float v = acc[p][c] + (a.bias ? __ldg(a.bias + gco) : 0.0f);
        if (a.relu) v = fmaxf(v, 0.0f);
        dst[gco] = v;
      }
    }
  }
}

// ---- fully connected --------------------------------------------------------------------
constexpr int kFcThreads = 256;   // 64 outputs x 4 k-slices
constexpr int kFcOut = 64;
constexpr int kFcSlices = 4;
constexpr int kFcBatch = 8;       // images per pass (weights streamed once for all of them)
constexpr int kFcChunk = 512;     // inputs staged per step: 8 x 512 x 4 B = 16 KB

__global__ void __launch_bounds__(kFcThreads)
fc_kernel(const float* __restrict__ in, const float* __restrict__ w,
          const float* __restrict__ bias, float* __restrict__ out, int B, int I, int O,
          int relu) {
  __shared__ float xs[kFcBatch][kFcChunk];
  __shared__ float red[kFcSlices][kFcBatch][kFcOut];
  const int tid = threadIdx.x;
  const int o = tid % kFcOut, ks = tid / kFcOut;
  const int go = blockIdx.x * kFcOut + o;
  const int b0 = blockIdx.y * kFcBatch;
  const int nb = min(kFcBatch, B - b0);
  float acc[kFcBatch];
#pragma unroll
  for (int b = 0; b < kFcBatch; ++b) acc[b] = 0.0f;

  for (int i0 = 0; i0 < I; i0 += kFcChunk) {
    const int n = min(kFcChunk, I - i0);
    __syncthreads();
    for (int e = tid; e < kFcBatch * kFcChunk; e += kFcThreads) {
      const int b = e / kFcChunk, i = e % kFcChunk;
      xs[b][i] = (b < nb && i < n) ? __ldg(in + static_cast<size_t>(b0 + b) * I + i0 + i) : 0.0f;
    }
    __syncthreads();
    if (go < O) {
      // k-slices interleave so consecutive rows of W stream from consecutive threads' loops
#pragma unroll 8
      for (int i = ks; i < n; i += kFcSlices) {
        const float wv = __ldg(w + static_cast<size_t>(i0 + i) * O + go);
#pragma unroll
        for (int b = 0; b < kFcBatch; ++b) acc[b] = fmaf(xs[b][i], wv, acc[b]);
      }
    }
  }
#pragma unroll
  for (int b = 0; b < kFcBatch; ++b) red[ks][b][o] = acc[b];
  __syncthreads();
  for (int e = tid; e < kFcBatch * kFcOut; e += kFcThreads) {
    const int b = e / kFcOut, oo = e % kFcOut;
    const int goo = blockIdx.x * kFcOut + oo;
    if (b < nb && goo < O) {
      float v = ((red[0][b][oo] + red[1][b][oo]) + (red[2][b][oo] + red[3][b][oo])) +
                (bias ? __ldg(bias + goo) : 0.0f);
      if (relu) v = fmaxf(v, 0.0f);
      out[static_cast<size_t>(b0 + b) * O + goo] = v;
    }
  }
}


// ---- fully connected, split-K over a thread-block cluster ----------------------------------------
// At batch <= 8 a fully connected layer is a weight stream (fc1: 1 MB) that a handful of CTAs
// cannot pull fast enough (52 us measured with 4 CTAs).  Here the K dimension is split over the
// CTAs of a cluster (up to 8): every CTA streams its slice of W with 128-bit loads (16 rows x 16
// float4 columns in flight per pass), reduces its 16 row-lanes in shared memory, and the
// cluster's rank 0 sums the per-CTA partials through DISTRIBUTED SHARED MEMORY in a fixed order
// (deterministic; no atomics), then applies bias / ReLU.
constexpr int kFcCThreads = 256;
constexpr int kFcCOut = 64;       // outputs per cluster (16 float4 columns)
constexpr int kFcCRows = 16;      // k rows streamed in parallel
constexpr int kFcCMaxSlice = 256; // inputs per CTA staged in shared memory (static smem <= 48 KB)

__global__ void __launch_bounds__(kFcCThreads)
fc_cluster_kernel(const float* __restrict__ in, const float* __restrict__ w,
                  const float* __restrict__ bias, float* __restrict__ out, int B, int I, int O,
                  int relu, int slice) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  __shared__ float xs[kFcBatch][kFcCMaxSlice];
  __shared__ float red[kFcCRows][kFcBatch][kFcCOut];   // 32 KB
  __shared__ float partial[kFcBatch][kFcCOut];
  const int tid = threadIdx.x;
  const int col = tid % 16, krow = tid / 16;
  const int o0 = blockIdx.x * kFcCOut + col * 4;
  const int b0 = blockIdx.z * kFcBatch;
  const int nb = min(kFcBatch, B - b0);
  const unsigned rank = cluster.block_rank();
  const int k0 = static_cast<int>(rank) * slice;
  const int kn = max(0, min(slice, I - k0));

  for (int e = tid; e < kFcBatch * slice; e += kFcCThreads) {
    const int b = e / slice, i = e % slice;
    xs[b][i] = (b < nb && i < kn) ? __ldg(in + static_cast<size_t>(b0 + b) * I + k0 + i) : 0.0f;
  }
  __syncthreads();

  float acc[kFcBatch][4];
#pragma unroll
  for (int b = 0; b < kFcBatch; ++b) acc[b][0] = acc[b][1] = acc[b][2] = acc[b][3] = 0.0f;
  if (o0 < O) {
#pragma unroll 4
    for (int i = krow; i < kn; i += kFcCRows) {
      const float4 wv = __ldg(reinterpret_cast<const float4*>(w + static_cast<size_t>(k0 + i) * O + o0));
#pragma unroll
      for (int b = 0; b < kFcBatch; ++b) {
        const float x = xs[b][i];
        acc[b][0] = fmaf(x, wv.x, acc[b][0]);
        acc[b][1] = fmaf(x, wv.y, acc[b][1]);
        acc[b][2] = fmaf(x, wv.z, acc[b][2]);
        acc[b][3] = fmaf(x, wv.w, acc[b][3]);
      }
    }
  }
#pragma unroll
  for (int b = 0; b < kFcBatch; ++b)
    *reinterpret_cast<float4*>(&red[krow][b][col * 4]) = make_float4(acc[b][0], acc[b][1], acc[b][2], acc[b][3]);
  __syncthreads();
  for (int e = tid; e < kFcBatch * kFcCOut; e += kFcCThreads) {
    const int b = e / kFcCOut, o = e % kFcCOut;
    float s = 0.0f;
#pragma unroll
    for (int r = 0; r < kFcCRows; ++r) s += red[r][b][o];
    partial[b][o] = s;
  }
  cluster.sync();  // every CTA's partial is complete and visible cluster-wide
  if (rank == 0) {
    const unsigned nranks = cluster.num_blocks();
    for (int e = tid; e < kFcBatch * kFcCOut; e += kFcCThreads) {
      const int b = e / kFcCOut, o = e % kFcCOut;
      const int go = blockIdx.x * kFcCOut + o;
      if (b < nb && go < O) {
        float s = 0.0f;
        for (unsigned r = 0; r < nranks; ++r)
          s += cluster.map_shared_rank(&partial[0][0], r)[b * kFcCOut + o];
        s += bias ? __ldg(bias + go) : 0.0f;
        out[static_cast<size_t>(b0 + b) * O + go] = relu ? fmaxf(s, 0.0f) : s;
      }
    }
  }
  cluster.sync();  // keep the remote shared memory alive until rank 0 has read it
}

// ---- fusion + prediction + unroll_grid ----------------------------------------------------
constexpr int kFpThreads = 256;
constexpr int kFpCells = 8;  // grid cells per CTA (one warp each)

__global__ void __launch_bounds__(kFpThreads)
fuse_predict_kernel(const float* __restrict__ local, const float* __restrict__ global_feat,
                    const float* __restrict__ w, const float* __restrict__ bias,
                    float* __restrict__ grid, int B, int cells_per_image, int C, int gd,
                    int n_out, int n_in, int stage_w) {
  extern __shared__ __align__(16) float sm[];  // [w[C][O] when stage_w] then f[kFpCells][C]
  const int O = gd * n_out * n_in;
  // stage_w == 0: the prediction weights do not fit shared memory next to the features (e.g.
  // HDRNetGaussianPyrNN with channel_multiplier 4: C = 256, O = 288, 295 KB): read them through L1
  float* fsm = sm + (stage_w ? static_cast<size_t>(C) * O : 0);
  const float* wsm = stage_w ? sm : w;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (stage_w)
    for (int e = tid; e < C * O; e += kFpThreads) sm[e] = __ldg(w + e);
  const long long total = static_cast<long long>(B) * cells_per_image;
  const long long cell = static_cast<long long>(blockIdx.x) * kFpCells + warp;
  const bool valid = cell < total;
  if (valid) {
    const int b = static_cast<int>(cell / cells_per_image);
    for (int c = lane; c < C; c += 32)
      fsm[warp * C + c] = fmaxf(__ldg(local + cell * C + c) + __ldg(global_feat + static_cast<size_t>(b) * C + c), 0.0f);
  }
  __syncthreads();
  if (!valid) return;
  const float* f = fsm + warp * C;
  for (int o = lane; o < O; o += 32) {
    float acc = 0.0f;
    for (int c = 0; c < C; ++c) acc = fmaf(f[c], wsm[c * O + o], acc);
    acc += bias ? __ldg(bias + o) : 0.0f;
    // unroll_grid (models.py:134-139): prediction channel o = (j*n_out + i)*gd + z
    const int z = o % gd;
    const int i = (o / gd) % n_out;
    const int j = o / (gd * n_out);
    grid[((cell * gd + z) * n_out + i) * n_in + j] = acc;
  }
}

// conv_tcgen05.cu: tensor-core (tcgen05 / TMEM, 3xTF32) implicit-GEMM form of the same layer.
int launch_conv_tcgen05(const float* in, const float* w, const float* bias, float* out, int B,
                        int H, int W, int Cin, int Cout, int k, int stride, int relu, int OH,
                        int OW, int pad_t, int pad_l, cudaStream_t stream);

static void same_pad(int size, int k, int s, int* out, int* before) {
  *out = (size + s - 1) / s;
  int total = (*out - 1) * s + k - size;
  if (total < 0) total = 0;
  *before = total / 2;
}

template <int kPx, int kCo>
static int launch_conv(ConvArgs a, cudaStream_t stream) {
  constexpr int kTilePx = 32 * kPx, kTileCo = 4 * kCo;
  // weights staged per input-channel chunk: k*k*chunk*kTileCo floats <= 72 KB
  int chunk = a.Cin;
  const int max_chunk = (72 * 1024 / 4) / (a.k * a.k * kTileCo);
  if (chunk > max_chunk) chunk = max_chunk / 4 * 4;
  a.ci_chunk = chunk;
  const size_t smem = static_cast<size_t>(a.k) * a.k * chunk * kTileCo * sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(conv2d_nhwc_kernel<kPx, kCo>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(smem));
  if (e != cudaSuccess) return static_cast<int>(e);
  const long long total_px = static_cast<long long>(a.B) * a.OH * a.OW;
  dim3 grid(static_cast<unsigned>((total_px + kTilePx - 1) / kTilePx),
            static_cast<unsigned>((a.Cout + kTileCo - 1) / kTileCo));
  conv2d_nhwc_kernel<kPx, kCo><<<grid, kConvThreads, smem, stream>>>(a);
  return static_cast<int>(cudaGetLastError());
}

}  // namespace hdrnet_b200

using namespace hdrnet_b200;

extern "C" {

int hdrnet_conv2d_nhwc_f32(const float* in, const float* w, const float* bias, float* out, int B,
                           int H, int W, int Cin, int Cout, int k, int stride, int relu,
                           void* stream) {
  if (B < 0 || H < 1 || W < 1 || Cin < 1 || Cout < 1) return HDRNET_E_BAD_SHAPE;
  if ((k != 1 && k != 3) || (stride != 1 && stride != 2)) return HDRNET_E_UNSUPPORTED;
  if (B == 0) return HDRNET_OK;
  if (!in || !w || !out) return HDRNET_E_NULL_POINTER;
  ConvArgs a;
  a.in = in; a.w = w; a.bias = bias; a.out = out;
  a.B = B; a.H = H; a.W = W; a.Cin = Cin; a.Cout = Cout; a.k = k; a.stride = stride; a.relu = relu;
  same_pad(H, k, stride, &a.OH, &a.pad_t);
  same_pad(W, k, stride, &a.OW, &a.pad_l);
  a.w_vec = (Cout % 4 == 0) && ((reinterpret_cast<uintptr_t>(w) & 15u) == 0);
  {  // Tensor-core path (conv_tcgen05.cu).  Each 128-pixel tile runs a fixed-latency chunk loop,
     // so it wins once there are about as many tiles as SMs (measured on B200: 2x faster at 128
     // tiles, 3x slower at 2); HDRNET_CONV_TCGEN05=1 / =0 forces it on / off.
    const char* e = std::getenv("HDRNET_CONV_TCGEN05");
    const long long tiles = (static_cast<long long>(B) * a.OH * a.OW + 127) / 128;
    const bool want = e ? (e[0] == '1') : (tiles >= 96);
    if (want) {
      const int rc = launch_conv_tcgen05(in, w, bias, out, B, H, W, Cin, Cout, k, stride, relu,
                                         a.OH, a.OW, a.pad_t, a.pad_l,
                                         static_cast<cudaStream_t>(stream));
      if (rc != HDRNET_E_UNSUPPORTED) return rc;
    }
  }
  const long long total_px = static_cast<long long>(B) * a.OH * a.OW;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long long big_ctas = ((total_px + 63) / 64) * ((Cout + 31) / 32);
  return (big_ctas >= 2LL * sms) ? launch_conv<2, 8>(a, static_cast<cudaStream_t>(stream))
                                 : launch_conv<1, 4>(a, static_cast<cudaStream_t>(stream));
}

int hdrnet_fc_f32(const float* in, const float* w, const float* bias, float* out, int B, int I,
                  int O, int relu, void* stream) {
  if (B < 0 || I < 1 || O < 1) return HDRNET_E_BAD_SHAPE;
  if (B == 0) return HDRNET_OK;
  if (!in || !w || !out) return HDRNET_E_NULL_POINTER;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // Cluster split-K form: needs float4 rows of W (O % 4 == 0, aligned) and a K worth splitting.
  int ksplit = 1;
  while (ksplit < 8 && I / (ksplit * 2) >= 64) ksplit *= 2;
  const int slice = (I + ksplit - 1) / ksplit;
  const char* env = std::getenv("HDRNET_FC_CLUSTER");
  const bool allow = !(env && env[0] == '0');
  if (allow && O % 4 == 0 && (reinterpret_cast<uintptr_t>(w) & 15u) == 0 && ksplit >= 2 &&
      slice <= kFcCMaxSlice) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((O + kFcCOut - 1) / kFcCOut, ksplit, (B + kFcBatch - 1) / kFcBatch);
    cfg.blockDim = dim3(kFcCThreads);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 1;
    attr[0].val.clusterDim.y = ksplit;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, fc_cluster_kernel, in, w, bias, out, B, I, O, relu, slice);
    return static_cast<int>(e != cudaSuccess ? e : cudaGetLastError());
  }
  dim3 grid((O + kFcOut - 1) / kFcOut, (B + kFcBatch - 1) / kFcBatch);
  fc_kernel<<<grid, kFcThreads, 0, st>>>(in, w, bias, out, B, I, O, relu);
  return static_cast<int>(cudaGetLastError());
}

int hdrnet_fuse_predict_f32(const float* local, const float* global_feat, const float* w,
                            const float* bias, float* grid, int B, int gh, int gw, int C, int gd,
                            int n_out, int n_in, void* stream) {
  if (B < 0 || gh < 1 || gw < 1 || C < 1 || gd < 1 || n_out < 1 || n_in < 1) return HDRNET_E_BAD_SHAPE;
  if (B == 0) return HDRNET_OK;
  if (!local || !global_feat || !w || !grid) return HDRNET_E_NULL_POINTER;
  const int O = gd * n_out * n_in;
  size_t smem = (static_cast<size_t>(C) * O + static_cast<size_t>(kFpCells) * C) * sizeof(float);
  const int stage_w = smem <= 200 * 1024;
  if (!stage_w) smem = static_cast<size_t>(kFpCells) * C * sizeof(float);
  if (smem > 200 * 1024) return HDRNET_E_UNSUPPORTED;
  cudaError_t e = cudaFuncSetAttribute(fuse_predict_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(smem));
  if (e != cudaSuccess) return static_cast<int>(e);
  const long long cells = static_cast<long long>(B) * gh * gw;
  fuse_predict_kernel<<<static_cast<unsigned>((cells + kFpCells - 1) / kFpCells), kFpThreads, smem,
                        static_cast<cudaStream_t>(stream)>>>(local, global_feat, w, bias, grid, B,
                                                             gh * gw, C, gd, n_out, n_in, stage_w);
  return static_cast<int>(cudaGetLastError());
}

}  // extern "C"
