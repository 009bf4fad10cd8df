// cnn_persistent.cu -- the whole coefficient network of HDRNetCurves._coefficients
// (hdrnet/models.py:62-142) as ONE persistent cooperative kernel, for small batches.
//
// Why: the network is 83 MFLOP per image.  As twelve launches (cnn.cu / conv_tcgen05.cu) it takes
// 177 us at batch 1 -- each late layer is a few dozen CTAs walking a 576-deep K loop against L2
// latency (profiles/r02_all_kernels_ncu.md: conv2d_nhwc_kernel<1,4> 31 us, fuse_predict 20 us,
// fc_cluster 11 us) -- and that is 85 % of a 1080p frame's model time (BASELINE.json config 2).
// Here one grid of one CTA per SM runs all layers, separated by grid barriers:
//
//   stage 1..n_ds   splat conv i            3x3 stride 2, TF SAME (asymmetric pads), ReLU
//   stage n_ds+1    global conv1 || local conv1      (both read the splat features)
//   stage n_ds+2    global conv2 || local conv2
//   stage n_ds+3    fc1 -> fc2 -> fc3 inside ONE CTA (block barriers only; 0.6 MFLOP)
//   stage n_ds+4    fusion relu(local + global) + 1x1 prediction + unroll_grid  -> the grid
//
// Work unit = one WARP task: 32 / min(Cout, 32) output pixels x min(Cout, 32) output channels
// (lane = channel: weight rows [tap][ci][co] are read coalesced, the pixel's input values are
// warp-broadcast 128-bit loads), accumulated sequentially over (tap, ci) in fp32 -- the same
// summation order as cnn.cu's kernels.  Every layer writes a buffer of its own in the caller's
// scratch: no address is written twice inside a launch, so L1-cached loads of activations
// written by other CTAs are safe behind the grid barrier.  Batch norm is folded on the host.
#include <cooperative_groups.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>

#include "hdrnet_b200.h"

namespace cg = cooperative_groups;

namespace hdrnet_b200 {

constexpr int kPcThreads = 256;
constexpr int kPcMaxSplat = 6;

struct PcConv {
  const float* in;    // [B, H, W, Cin]
  const float* w;     // [3, 3, Cin, Cout]
  const float* bias;  // [Cout] or nullptr
  float* out;         // [B, OH, OW, Cout]
  int H, W, Cin, OH, OW, Cout, stride, pad_t, pad_l, relu;
};

struct PcArgs {
  PcConv splat[kPcMaxSplat];
  PcConv gconv[2], lconv[2];
  const float* fc_w[3];
  const float* fc_b[3];
  float* fc_out[3];
  int fc_in[3], fc_outn[3];
  const float* pred_w;   // [C][O]
  const float* pred_b;   // [O]
  float* grid;           // [B, gh, gw, gd, n_out, n_in]
  int B, n_ds, C8, gd, n_out, n_in;
};

__device__ __forceinline__ int pc_tasks(const PcConv& L, int B) {
  const int cos = min(L.Cout, 32);
  const int ppw = 32 / cos;
  const long long npx = static_cast<long long>(B) * L.OH * L.OW;
  return static_cast<int>((npx + ppw - 1) / ppw) * ((L.Cout + 31) / 32);
}

// One warp task of a 3x3 convolution (TF SAME, stride 1 / 2).
__device__ __forceinline__ void pc_conv_task(const PcConv& L, int B, int task, int lane) {
  const int cos = min(L.Cout, 32), ppw = 32 / cos, cgroups = (L.Cout + 31) / 32;
  const int cgi = task % cgroups;
  const long long px = static_cast<long long>(task / cgroups) * ppw + lane / cos;
  const int co = cgi * 32 + lane % cos;
  const long long npx = static_cast<long long>(B) * L.OH * L.OW;
  const bool valid = px < npx;
  const long long q = valid ? px : 0;
  const int ox = static_cast<int>(q % L.OW), oy = static_cast<int>((q / L.OW) % L.OH);
  const int b = static_cast<int>(q / (static_cast<long long>(L.OW) * L.OH));
  float acc = 0.0f;
  const float* wco = L.w + co;
  for (int ky = 0; ky < 3; ++ky) {
    const int iy = oy * L.stride - L.pad_t + ky;
    for (int kx = 0; kx < 3; ++kx) {
      const int ix = ox * L.stride - L.pad_l + kx;
      if (iy < 0 || iy >= L.H || ix < 0 || ix >= L.W) continue;   // zero padding (per lane: pixels differ)
      const float* ip = L.in + ((static_cast<size_t>(b) * L.H + iy) * L.W + ix) * L.Cin;
      const float* wp = wco + static_cast<size_t>(ky * 3 + kx) * L.Cin * L.Cout;
      if ((L.Cin & 3) == 0) {
#pragma unroll 4
        for (int ci = 0; ci < L.Cin; ci += 4) {
          const float4 v = *reinterpret_cast<const float4*>(ip + ci);
          acc = fmaf(v.x, __ldg(wp + static_cast<size_t>(ci) * L.Cout), acc);
          acc = fmaf(v.y, __ldg(wp + static_cast<size_t>(ci + 1) * L.Cout), acc);
          acc = fmaf(v.z, __ldg(wp + static_cast<size_t>(ci + 2) * L.Cout), acc);
          acc = fmaf(v.w, __ldg(wp + static_cast<size_t>(ci + 3) * L.Cout), acc);
        }
      } else {
        for (int ci = 0; ci < L.Cin; ++ci) acc = fmaf(ip[ci], __ldg(wp + static_cast<size_t>(ci) * L.Cout), acc);
      }
    }
  }
  if (L.bias) acc += __ldg(L.bias + co);
  if (L.relu) acc = fmaxf(acc, 0.0f);
  if (valid) L.out[static_cast<size_t>(px) * L.Cout + co] = acc;
}

// out[b][o] = in[b][:] . w[:, o] + bias (+ReLU); a warp owns 32 consecutive outputs of one image.
__device__ __forceinline__ void pc_fc_task(const float* in, const float* w, const float* bias, float* out,
                                           int I, int O, int relu, int task, int lane) {
  const int og = (O + 31) / 32;
  const int b = task / og, o = (task % og) * 32 + lane;
  const bool valid = o < O;
  const int oc = valid ? o : 0;
  const float* x = in + static_cast<size_t>(b) * I;
  float acc = 0.0f;
#pragma unroll 8
  for (int k = 0; k < I; ++k) acc = fmaf(x[k], __ldg(w + static_cast<size_t>(k) * O + oc), acc);
  if (bias) acc += __ldg(bias + oc);
  if (relu) acc = fmaxf(acc, 0.0f);
  if (valid) out[static_cast<size_t>(b) * O + o] = acc;
}

__global__ void __launch_bounds__(kPcThreads, 1)
coefficients_persistent_kernel(const PcArgs a) {
  cg::grid_group grid = cg::this_grid();
  const int lane = threadIdx.x & 31;
  const int warps_per_cta = kPcThreads / 32;
  const int gwarp = blockIdx.x * warps_per_cta + (threadIdx.x >> 5);
  const int nwarps = gridDim.x * warps_per_cta;

  // ---- splat: 3x3 stride-2 convolutions down to spatial_bin x spatial_bin (models.py:69-82) ----
  for (int i = 0; i < a.n_ds; ++i) {
    const int nt = pc_tasks(a.splat[i], a.B);
    for (int t = gwarp; t < nt; t += nwarps) pc_conv_task(a.splat[i], a.B, t, lane);
    grid.sync();
  }
  // ---- global conv k || local conv k (models.py:86-93, :109-118): one task list ----
  for (int k = 0; k < 2; ++k) {
    const int ng = pc_tasks(a.gconv[k], a.B), nl = pc_tasks(a.lconv[k], a.B);
    for (int t = gwarp; t < ng + nl; t += nwarps) {
      if (t < nl) pc_conv_task(a.lconv[k], a.B, t, lane);
      else pc_conv_task(a.gconv[k], a.B, t - nl, lane);
    }
    grid.sync();
  }
  // ---- fc1 -> fc2 -> fc3 (models.py:94-104) inside CTA 0; flatten = the NHWC buffer as it lies ----
  if (blockIdx.x == 0) {
    const int w = threadIdx.x >> 5;
    const float* in = a.gconv[1].out;
    for (int f = 0; f < 3; ++f) {
      const int nt = a.B * ((a.fc_outn[f] + 31) / 32);
      for (int t = w; t < nt; t += warps_per_cta)
        pc_fc_task(in, a.fc_w[f], a.fc_b[f], a.fc_out[f], a.fc_in[f], a.fc_outn[f], f < 2, t, lane);
      in = a.fc_out[f];
      __syncthreads();   // block-scope visibility of the activations just written
    }
  }
  grid.sync();
  // ---- fusion + 1x1 prediction + unroll_grid (models.py:122-139) ----
  {
    const PcConv& L = a.lconv[1];
    const int O = a.gd * a.n_out * a.n_in, og = (O + 31) / 32;
    const int cells = L.OH * L.OW;
    const long long ncell = static_cast<long long>(a.B) * cells;
    const float* glob = a.fc_out[2];
    for (long long t = gwarp; t < ncell * og; t += nwarps) {
      const long long cell = t / og;
      const int o = static_cast<int>(t % og) * 32 + lane;
      const bool valid = o < O;
      const int oc = valid ? o : 0;
      const int b = static_cast<int>(cell / cells);
      const float* lp = L.out + static_cast<size_t>(cell) * a.C8;
      const float* gp = glob + static_cast<size_t>(b) * a.C8;
      float acc = 0.0f;
#pragma unroll 4
      for (int c = 0; c < a.C8; ++c)
        acc = fmaf(fmaxf(lp[c] + gp[c], 0.0f), __ldg(a.pred_w + static_cast<size_t>(c) * O + oc), acc);
      acc += a.pred_b ? __ldg(a.pred_b + oc) : 0.0f;
      // prediction channel o = (j * n_out + i) * gd + z  ->  grid[b, y, x, z, i, j]
      const int z = oc % a.gd, i = (oc / a.gd) % a.n_out, j = oc / (a.gd * a.n_out);
      if (valid) a.grid[((cell * a.gd + z) * a.n_out + i) * a.n_in + j] = acc;
    }
  }
}

static void pc_same_pad(int size, int s, int* out, int* before) {
  *out = (size + s - 1) / s;
  int total = (*out - 1) * s + 3 - size;
  if (total < 0) total = 0;
  *before = total / 2;
}

static bool pc_pow2(int v) { return v >= 1 && (v & (v - 1)) == 0; }

struct PcDims {
  int n_ds, sb, c_splat[kPcMaxSplat], c8, f1, f2, O;
  size_t act_floats;   // scratch floats
};

// Channel plan of models.py:62-142: splat i has cm * 2^i * gd channels; everything after 8 * cm * gd.
static bool pc_dims(int S, int spatial_bin, int gd, int cm, int n_out, int n_in, int B, PcDims* d) {
  if (S < 1 || spatial_bin < 4 || gd < 1 || cm < 1 || S % spatial_bin) return false;
  int n_ds = 0;
  for (int s = S; s > spatial_bin; s >>= 1) { if (s & 1) return false; ++n_ds; }
  if (n_ds < 1 || n_ds > kPcMaxSplat || (spatial_bin << n_ds) != S) return false;
  d->n_ds = n_ds;
  d->sb = spatial_bin;
  size_t fl = 0;
  for (int i = 0; i < n_ds; ++i) {
    d->c_splat[i] = cm * (1 << i) * gd;
    if (!pc_pow2(d->c_splat[i]) || (d->c_splat[i] > 32 && d->c_splat[i] % 32)) return false;
    const int sp = S >> (i + 1);
    fl += static_cast<size_t>(B) * sp * sp * d->c_splat[i];
  }
  d->c8 = 8 * cm * gd;
  d->f1 = 32 * cm * gd;
  d->f2 = 16 * cm * gd;
  d->O = gd * n_out * n_in;
  if (!pc_pow2(d->c8) || d->c8 % 4) return false;
  const int sb = spatial_bin;
  fl += static_cast<size_t>(B) * (sb / 2) * (sb / 2) * d->c8;        // global conv1
  fl += static_cast<size_t>(B) * (sb / 4) * (sb / 4) * d->c8;        // global conv2 (= the flattened fc input)
  fl += 2 * static_cast<size_t>(B) * sb * sb * d->c8;                // local conv1, conv2
  fl += static_cast<size_t>(B) * (d->f1 + d->f2 + d->c8);            // fc1, fc2, fc3
  d->act_floats = fl + 64;
  return true;
}

}  // namespace hdrnet_b200

using namespace hdrnet_b200;

extern "C" {

size_t hdrnet_coefficients_scratch_bytes(int B, int net_input_size, int spatial_bin, int luma_bins,
                                         int channel_multiplier, int n_out, int n_in) {
  PcDims d;
  if (B < 1 || !pc_dims(net_input_size, spatial_bin, luma_bins, channel_multiplier, n_out, n_in, B, &d)) return 0;
  return d.act_floats * sizeof(float);
}

int hdrnet_coefficients_f32(const float* lowres, float* grid, const float* const* weights,
                            const float* const* biases, int n_layers, void* scratch, size_t scratch_bytes,
                            int B, int net_input_size, int spatial_bin, int luma_bins,
                            int channel_multiplier, int n_out, int n_in, void* stream) {
  if (B < 0 || n_out < 1 || n_in < 1) return HDRNET_E_BAD_SHAPE;
  if (B == 0) return HDRNET_OK;
  if (!lowres || !grid || !weights || !biases || !scratch) return HDRNET_E_NULL_POINTER;
  PcDims d;
  if (!pc_dims(net_input_size, spatial_bin, luma_bins, channel_multiplier, n_out, n_in, B, &d))
    return HDRNET_E_UNSUPPORTED;
  if (n_layers != d.n_ds + 8) return HDRNET_E_BAD_SHAPE;
  if (scratch_bytes < d.act_floats * sizeof(float)) return HDRNET_E_BAD_SHAPE;
  for (int i = 0; i < n_layers; ++i)
    if (!weights[i]) return HDRNET_E_NULL_POINTER;
  if ((reinterpret_cast<uintptr_t>(scratch) & 15u) || (reinterpret_cast<uintptr_t>(lowres) & 15u))
    return HDRNET_E_UNSUPPORTED;

  PcArgs a = {};
  float* cur = static_cast<float*>(scratch);
  auto take = [&](size_t n) { float* p = cur; cur += (n + 3) & ~static_cast<size_t>(3); return p; };
  auto conv = [&](PcConv& L, const float* in, int H, int W, int Cin, int Cout, int stride, int relu,
                  const float* w, const float* b) {
    L.in = in; L.w = w; L.bias = b; L.H = H; L.W = W; L.Cin = Cin; L.Cout = Cout; L.stride = stride; L.relu = relu;
    pc_same_pad(H, stride, &L.OH, &L.pad_t);
    pc_same_pad(W, stride, &L.OW, &L.pad_l);
    L.out = take(static_cast<size_t>(B) * L.OH * L.OW * Cout);
  };
  // layer order of `weights` / `biases`: splat conv1..n_ds, global conv1, conv2, fc1, fc2, fc3,
  // local conv1, conv2, prediction conv1
  const float* in = lowres;
  int H = net_input_size, C = 3, li = 0;
  for (int i = 0; i < d.n_ds; ++i, ++li) {
    conv(a.splat[i], in, H, H, C, d.c_splat[i], 2, 1, weights[li], biases[li]);
    in = a.splat[i].out; H = a.splat[i].OH; C = d.c_splat[i];
  }
  const float* splat = in;
  conv(a.gconv[0], splat, H, H, C, d.c8, 2, 1, weights[li], biases[li]); ++li;
  conv(a.gconv[1], a.gconv[0].out, a.gconv[0].OH, a.gconv[0].OW, d.c8, d.c8, 2, 1, weights[li], biases[li]); ++li;
  const int fc_in0 = a.gconv[1].OH * a.gconv[1].OW * d.c8;
  const int fin[3] = {fc_in0, d.f1, d.f2}, fout[3] = {d.f1, d.f2, d.c8};
  for (int f = 0; f < 3; ++f, ++li) {
    a.fc_w[f] = weights[li]; a.fc_b[f] = biases[li]; a.fc_in[f] = fin[f]; a.fc_outn[f] = fout[f];
    a.fc_out[f] = take(static_cast<size_t>(B) * fout[f]);
  }
  conv(a.lconv[0], splat, H, H, C, d.c8, 1, 1, weights[li], biases[li]); ++li;
  conv(a.lconv[1], a.lconv[0].out, H, H, d.c8, d.c8, 1, 0, weights[li], biases[li]); ++li;
  a.pred_w = weights[li]; a.pred_b = biases[li];
  a.grid = grid;
  a.B = B; a.n_ds = d.n_ds; a.C8 = d.c8; a.gd = luma_bins; a.n_out = n_out; a.n_in = n_in;
  if (static_cast<size_t>(cur - static_cast<float*>(scratch)) > d.act_floats) return HDRNET_E_BAD_SHAPE;

  int dev = 0, sms = 0, per_sm = 0, coop = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
  if (!coop) return HDRNET_E_UNSUPPORTED;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, coefficients_persistent_kernel,
                                                                kPcThreads, 0);
  if (e != cudaSuccess) return static_cast<int>(e);
  if (per_sm < 1) return HDRNET_E_UNSUPPORTED;
  void* kargs[] = {&a};
  e = cudaLaunchCooperativeKernel(reinterpret_cast<void*>(coefficients_persistent_kernel), dim3(sms), dim3(kPcThreads),
                                  kargs, 0, static_cast<cudaStream_t>(stream));
  return static_cast<int>(e);
}

}  // extern "C"
