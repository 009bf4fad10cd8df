// tmem_paths.cu -- microbenchmark / semantics probe for a tensor-core "gather":
//   D[128 px][32] = A[128 px][8 z-weights] x B[8 z][32 coefficients]   (tcgen05.mma kind::tf32)
// with A written to TENSOR MEMORY by the threads (tcgen05.st), B in shared memory, D read back with
// tcgen05.ld.  Questions: (1) is the A-in-TMEM operand layout lane=row / column=k?  (2) what do
// tcgen05.ld, tcgen05.st and the tiny MMA cost per SM when 4..16 warps drive them?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I hdrnet_b200/csrc -I include -o tools/ubench/bin/tmem_paths tools/ubench/tmem_paths.cu
// CAUTION (found in round 1): the first version of the rate loops indexed the register array with a
// runtime value (r[it & 31]); nvcc then keeps the array in LOCAL memory and every iteration
// spills / reloads it through the LSU -- its "tcgen05.ld = 51 B/clk/SM" was spill traffic, not
// tensor memory (profiles/r01_tmem_paths.txt).  With static indices LDTM.x16 sustains ~900 B/clk/SM
// (profiles/r01_tmem_paths_v2.txt, section 3).  Check `cuobjdump -sass | grep -c STL` = 0 in a loop
// before believing its number.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "common.cuh"

using namespace hdrnet_b200;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint64_t kmajor_desc(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3fffu);
  d |= static_cast<uint64_t>((lbo >> 4) & 0x3fffu) << 16;
  d |= static_cast<uint64_t>((sbo >> 4) & 0x3fffu) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  return d;
}

__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(cols));
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
}
__device__ __forceinline__ void tmem_free(uint32_t base, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(cols));
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr),
               "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
      "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

constexpr int kN = 32;  // two x cells x 16 (12 coefficients + 4 pad)
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(kN >> 3) << 17) |
                            (static_cast<uint32_t>(128 >> 4) << 24);

// B element (n, k) lives at float offset: per-cell blocks of [n-group 2][k-chunk 2][8 n][4 k].
__host__ __device__ inline int b_off(int n, int k) { return (n >> 3) * 64 + (k >> 2) * 32 + (n & 7) * 4 + (k & 3); }

// ---- 1. semantics: D = A x B^T with A in TMEM ---------------------------------------------------
__global__ void __launch_bounds__(128) semantics_kernel(const float* A, const float* Bm, float* D) {
  __shared__ __align__(128) float b_s[kN * 8];
  __shared__ uint64_t bar;
  __shared__ uint32_t tbase_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) tmem_alloc(&tbase_s, 64);
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  for (int e = tid; e < kN * 8; e += 128) b_s[b_off(e / 8, e % 8)] = Bm[e];  // Bm[n][k]
  fence_proxy_async_smem();
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tb = tbase_s;
  const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;
  uint32_t a[8];
  for (int k = 0; k < 8; ++k) a[k] = __float_as_uint(A[tid * 8 + k]);
  tmem_st8(tb + lane_base + 32, a);   // A at columns [32, 40)
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  if (tid == 0) {
    mma_ts(tb, tb + 32, kmajor_desc(smem_u32(b_s), 128, 256), kIdesc, 0);
    mma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  asm volatile("tcgen05.fence::after_thread_sync;");
  uint32_t r[32];
  tmem_ld32(tb + lane_base, r);
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int n = 0; n < kN; ++n) D[tid * kN + n] = __uint_as_float(r[n]);
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) tmem_free(tb, 64);
}

// ---- 2. throughput loops ------------------------------------------------------------------------
// mode 0: tcgen05.ld x32 ; 1: tcgen05.st x16 ; 2: MMA only (thread 0) ; 3: ld + st + MMA together
// (the shape of the real kernel: per 128-px tile 16 columns stored, 3 MMAs, 32 columns loaded).
template <int MODE>
__global__ void __launch_bounds__(1024) rate_kernel(int iters, long long* cycles, float* sink) {
  __shared__ __align__(128) float b_s[kN * 8];
  __shared__ uint64_t bar;
  __shared__ uint32_t tbase_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) tmem_alloc(&tbase_s, 512);
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  for (int e = tid; e < kN * 8; e += blockDim.x) b_s[e] = 0.0f;
  fence_proxy_async_smem();
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tb = tbase_s;
  const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
  const uint32_t col0 = static_cast<uint32_t>((warp >> 2) * 64) & 511u;  // each warpgroup its own 64 columns
  uint32_t r[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) r[i] = tid + i;
  float acc = 0.f;
  const uint64_t bdesc = kmajor_desc(smem_u32(b_s), 128, 256);
  __syncthreads();
  const long long t0 = clock64();
  if (MODE == 0) {
    for (int it = 0; it < iters; ++it) {
      tmem_ld32(tb + lane_base + col0 + (it & 1) * 32, r);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      acc += __uint_as_float(r[0]) + __uint_as_float(r[17]) + __uint_as_float(r[31]);
    }
  } else if (MODE == 1) {
    for (int it = 0; it < iters; ++it) {
      r[0] += it; r[9] ^= it;
      tmem_st16(tb + lane_base + col0 + (it & 3) * 16, r);
      if ((it & 3) == 3) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  } else if (MODE == 2) {
    if (tid == 0) {
      for (int it = 0; it < iters; ++it) mma_ts(tb + (it & 7) * 32, tb + 256 + (it & 7) * 8, bdesc, kIdesc, 0);
      mma_commit(&bar);
      mbar_wait(&bar, 0);
    }
  } else {
    // every warpgroup: st 16 columns of A, (thread 0 of the warpgroup) 3 MMAs, everyone ld 32 columns.
    // No cross-thread ordering is enforced here: this measures pipe throughput, not a correct pipeline.
    for (int it = 0; it < iters; ++it) {
      r[0] += it; r[9] ^= it;
      tmem_st16(tb + lane_base + col0 + 32, r);
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      if ((tid & 127) == 0) {
        mma_ts(tb + col0, tb + col0 + 32, bdesc, kIdesc, 0);
        mma_ts(tb + col0, tb + col0 + 40, bdesc, kIdesc, 1);
        mma_ts(tb + col0, tb + col0 + 32, bdesc, kIdesc, 1);
      }
      tmem_ld32(tb + lane_base + col0, r);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      acc += __uint_as_float(r[0]) + __uint_as_float(r[17]) + __uint_as_float(r[31]);
    }
    if (tid == 0) { mma_commit(&bar); mbar_wait(&bar, 0); }
  }
  const long long t1 = clock64();
  if (tid == 0) cycles[blockIdx.x] = t1 - t0;
  sink[blockIdx.x * blockDim.x + tid] = acc + __uint_as_float(r[5]);
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) tmem_free(tb, 512);
}

// ---- 3. is tensor memory a THIRD on-chip path next to the LSU? ----------------------------------
// The fused slice-apply kernel is bound by the shared-memory data pipe (LSU, 128 B/clk/SM) with
// the texture pipe (64 B/clk/SM) as the only helper so far.  Candidate: read the staged INPUT
// tile (16 B/px) through tensor memory -- tcgen05.cp copies it shared -> TMEM, tcgen05.ld brings it
// to registers.  Questions: does tcgen05.ld overlap with LDS.128 (mode 1 vs 0 + 2), what does
// tcgen05.cp cost (mode 3), and does it take shared-memory bandwidth away from LDS (mode 4)?
//   mode 0: 4 x LDS.128 per thread and iteration (64 B)         mode 2: both, same iteration
//   mode 1: tcgen05.ld 32x32b.x16 per thread and iteration (64 B)
//   mode 3: thread 0 issues tcgen05.cp 128x128b (2 KB each), 8 per iteration, nobody else works
//   mode 4: mode 0 in warps 1.., mode 3 in warp 0
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_cp_128x128b(uint32_t taddr, uint64_t sdesc) {
  asm volatile("tcgen05.cp.cta_group::1.128x128b [%0], %1;" ::"r"(taddr), "l"(sdesc) : "memory");
}

template <int MODE>
__global__ void __launch_bounds__(1024) path_kernel(int iters, long long* cycles, float* sink) {
  extern __shared__ __align__(128) unsigned char dsm[];   // 32 KB of "tile" data
  __shared__ uint64_t bar;
  __shared__ uint32_t tbase_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) tmem_alloc(&tbase_s, 512);
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  for (int e = tid; e < 8192; e += blockDim.x) reinterpret_cast<float*>(dsm)[e] = static_cast<float>(e & 255);
  fence_proxy_async_smem();
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tb = tbase_s;
  const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
  const uint32_t col0 = static_cast<uint32_t>((warp >> 2) * 64) & 511u;
  uint32_t r[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) r[i] = 0;
  float acc = 0.f;
  const bool cp_warp = (MODE == 3) || (MODE == 4 && warp == 0);
  const bool lds_warp = (MODE == 0) || (MODE == 2) || (MODE == 4 && warp != 0);
  const bool ld_warp = (MODE == 1) || (MODE == 2);
  const uint32_t smem_base = smem_u32(dsm);
  __syncthreads();
  const long long t0 = clock64();
  if (cp_warp) {
    if (tid == 0) {
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u)   // 128 rows x 16 B, core matrices 128 B apart, into 4 columns each
          tmem_cp_128x128b(tb + 256 + ((it & 1) * 8 + u) * 4, kmajor_desc(smem_base + u * 2048, 128, 128));
      }
      mma_commit(&bar);
      mbar_wait(&bar, 0);
    }
  }
  if (lds_warp || ld_warp) {
    for (int it = 0; it < iters; ++it) {
      if (lds_warp) {
        // 48-byte lane stride, as the kernel's RGB reads: conflict-free quarter-warps
        const uint32_t a = smem_base + ((static_cast<uint32_t>(tid) * 48u + static_cast<uint32_t>(it) * 16u) & 32767u & ~15u);
        float4 v0 = lds128(reinterpret_cast<const void*>(__cvta_shared_to_generic(a)));
        float4 v1 = lds128(reinterpret_cast<const void*>(__cvta_shared_to_generic((a + 4096u) & 32767u)));
        float4 v2 = lds128(reinterpret_cast<const void*>(__cvta_shared_to_generic((a + 8192u) & 32767u)));
        float4 v3 = lds128(reinterpret_cast<const void*>(__cvta_shared_to_generic((a + 12288u) & 32767u)));
        acc += v0.x + v1.y + v2.z + v3.w;
      }
      if (ld_warp) {
        tmem_ld16(tb + lane_base + col0 + (it & 3) * 16, r);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        acc += __uint_as_float(r[0]) + __uint_as_float(r[7]) + __uint_as_float(r[15]);
      }
    }
  }
  const long long t1 = clock64();
  if (tid == 0) cycles[blockIdx.x] = t1 - t0;                       // cp rate (modes 3, 4) / loop time
  if (tid == blockDim.x - 1) cycles[148 + blockIdx.x] = t1 - t0;    // a worker warp's loop time
  sink[blockIdx.x * blockDim.x + tid] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) tmem_free(tb, 512);
}

template <int MODE>
static void run_path(const char* name, int threads, long long* d_cyc, float* d_sink) {
  const int iters = 1024;
  CK(cudaFuncSetAttribute(path_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768));
  for (int rep = 0; rep < 2; ++rep) {
    path_kernel<MODE><<<148, threads, 32768>>>(iters, d_cyc, d_sink);
    CK(cudaDeviceSynchronize());
  }
  long long c[296];
  CK(cudaMemcpy(c, d_cyc, sizeof(c), cudaMemcpyDeviceToHost));
  double a0 = 0, a1 = 0;
  for (int i = 0; i < 148; ++i) { a0 += c[i]; a1 += c[148 + i]; }
  a0 /= 148.0 * iters; a1 /= 148.0 * iters;
  printf("%-44s threads %4d: thread 0 %8.1f clk/iter, last thread %8.1f clk/iter", name, threads, a0, a1);
  if (MODE <= 2) printf("  -> %6.1f B/clk/SM per 64 B stream", 64.0 * threads / a1);
  if (MODE == 3) printf("  -> %6.1f B/clk/SM (8 x 2 KB per iteration)", 16384.0 / a0);
  if (MODE == 4) printf("  -> cp %6.1f B/clk/SM, LDS %6.1f B/clk/SM", 16384.0 / a0, 64.0 * (threads - 32) / a1);
  printf("\n");
}

template <int MODE>
static void run_rate(const char* name, int threads, double bytes_per_iter_per_warp, long long* d_cyc, float* d_sink) {
  const int iters = 2048;
  rate_kernel<MODE><<<148, threads>>>(iters, d_cyc, d_sink);
  CK(cudaDeviceSynchronize());
  rate_kernel<MODE><<<148, threads>>>(iters, d_cyc, d_sink);
  CK(cudaDeviceSynchronize());
  long long c[148];
  CK(cudaMemcpy(c, d_cyc, sizeof(c), cudaMemcpyDeviceToHost));
  double avg = 0;
  for (int i = 0; i < 148; ++i) avg += c[i];
  avg /= 148;
  const double per_iter = avg / iters;
  printf("%-34s threads %4d: %8.1f clk/iter", name, threads, per_iter);
  if (bytes_per_iter_per_warp > 0) printf("  -> %7.1f B/clk/SM", bytes_per_iter_per_warp * (threads / 32) / per_iter);
  printf("\n");
}

int main() {
  // ---- semantics ----
  std::vector<float> A(128 * 8), B(kN * 8), D(128 * kN), ref(128 * kN);
  for (int i = 0; i < 128; ++i) for (int k = 0; k < 8; ++k) A[i * 8 + k] = static_cast<float>(((i * 3 + k * 5) % 7) - 3);
  for (int n = 0; n < kN; ++n) for (int k = 0; k < 8; ++k) B[n * 8 + k] = static_cast<float>(((n * 2 + k * 3) % 5) - 2);
  for (int i = 0; i < 128; ++i) for (int n = 0; n < kN; ++n) {
    float s = 0; for (int k = 0; k < 8; ++k) s += A[i * 8 + k] * B[n * 8 + k];
    ref[i * kN + n] = s;
  }
  float *dA, *dB, *dD; long long* d_cyc; float* d_sink;
  CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dD, D.size() * 4));
  CK(cudaMalloc(&d_cyc, 296 * 8)); CK(cudaMalloc(&d_sink, 148 * 1024 * 4));
  CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
  semantics_kernel<<<1, 128>>>(dA, dB, dD);
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
  int bad = 0;
  for (size_t i = 0; i < D.size(); ++i) if (D[i] != ref[i]) ++bad;
  printf("semantics: A-in-TMEM (lane=row, column=k) x B(smem, per-cell K-major blocks): %d / %zu mismatches\n", bad, D.size());
  if (bad) {
    for (int i = 0; i < 4; ++i) { printf(" row %d got:", i); for (int n = 0; n < 8; ++n) printf(" %g", D[i * kN + n]); printf("  want:"); for (int n = 0; n < 8; ++n) printf(" %g", ref[i * kN + n]); printf("\n"); }
  }
  // ---- rates ----
  for (int threads : {128, 256, 512, 1024}) run_rate<0>("tcgen05.ld 32x32b.x32 (4 KB/warp)", threads, 4096.0, d_cyc, d_sink);
  for (int threads : {128, 256, 512, 1024}) run_rate<1>("tcgen05.st 32x32b.x16 (2 KB/warp)", threads, 2048.0, d_cyc, d_sink);
  run_rate<2>("tcgen05.mma M128 N32 K8 tf32, A TMEM", 128, 0, d_cyc, d_sink);
  for (int threads : {128, 256, 512, 1024}) run_rate<3>("tile loop: st16 + 3 MMA + ld32", threads, 0, d_cyc, d_sink);
  // ---- third path? ----
  for (int threads : {512, 1024}) {
    run_path<0>("4 x LDS.128 (64 B/thread)", threads, d_cyc, d_sink);
    run_path<1>("tcgen05.ld x16 (64 B/thread)", threads, d_cyc, d_sink);
    run_path<2>("4 x LDS.128 + tcgen05.ld x16 together", threads, d_cyc, d_sink);
  }
  run_path<3>("tcgen05.cp 128x128b smem -> TMEM alone", 128, d_cyc, d_sink);
  run_path<4>("tcgen05.cp (warp 0) + 4 x LDS.128 (others)", 512, d_cyc, d_sink);
  return 0;
}
