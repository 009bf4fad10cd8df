// host_path.cu -- slice-apply on HOST buffers: the end-to-end path a CPU-tensor caller of
// the reference op takes (TF feeds the full-res float image H2D every frame and fetches the
// result back: hdrnet/bin/run.py:185).
//
// A context owns kSlots pipeline slots (device staging for one row band of guide / input /
// output, one stream each) plus a device copy of the grid.  An image is cut into row bands
// (rows of one image are contiguous in NHWC); band k runs on slot k % kSlots:
//     H2D(guide, input) -> slice-apply kernel (global y via y_off) -> D2H(out)
// so that, with page-locked host buffers, the upload of band k+1, the kernel of band k and
// the download of band k-1 overlap on the two copy engines and the SMs.
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>
#include <new>

#include "common.cuh"

namespace hdrnet_b200 {
int launch_slice_apply(const float* grid, const float* guide, const float* input, float* out,
                       int B, int H, int W, int rows, int y_off, int gh, int gw, int gd,
                       int n_in, int n_out, int has_offset, int variant, cudaStream_t stream);
}  // namespace hdrnet_b200

namespace {
constexpr int kSlots = 3;
constexpr uint32_t kCtxMagic = 0x48445242u;  // "HDRB"
}  // namespace

struct hdrnet_host_ctx {
  uint32_t magic;
  int device;
  size_t band_pixels;       // requested staging capacity per slot, pixels
  size_t slot_bytes;        // allocated bytes per slot
  size_t grid_bytes;        // allocated bytes of grid_dev
  unsigned char* slot_mem[kSlots];
  cudaStream_t stream[kSlots];
  cudaEvent_t grid_ready;
  float* grid_dev;
};

namespace {

void free_buffers(hdrnet_host_ctx* c) {
  for (int s = 0; s < kSlots; ++s) {
    if (c->slot_mem[s]) cudaFree(c->slot_mem[s]);
    c->slot_mem[s] = nullptr;
  }
  c->slot_bytes = 0;
  if (c->grid_dev) cudaFree(c->grid_dev);
  c->grid_dev = nullptr;
  c->grid_bytes = 0;
}

inline size_t align256(size_t v) { return (v + 255) & ~static_cast<size_t>(255); }

}  // namespace

extern "C" {

int hdrnet_host_ctx_create(hdrnet_host_ctx** out, size_t max_band_pixels) {
  if (!out) return HDRNET_E_NULL_POINTER;
  *out = nullptr;
  hdrnet_host_ctx* c = new (std::nothrow) hdrnet_host_ctx();
  if (!c) return static_cast<int>(cudaErrorMemoryAllocation);
  c->magic = kCtxMagic;
  c->band_pixels = max_band_pixels ? max_band_pixels : (static_cast<size_t>(4) << 20);
  c->slot_bytes = 0;
  c->grid_bytes = 0;
  c->grid_dev = nullptr;
  for (int s = 0; s < kSlots; ++s) { c->slot_mem[s] = nullptr; c->stream[s] = nullptr; }
  cudaError_t e = cudaGetDevice(&c->device);
  for (int s = 0; s < kSlots && e == cudaSuccess; ++s)
    e = cudaStreamCreateWithFlags(&c->stream[s], cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->grid_ready, cudaEventDisableTiming);
  if (e != cudaSuccess) {
    for (int s = 0; s < kSlots; ++s)
      if (c->stream[s]) cudaStreamDestroy(c->stream[s]);
    delete c;
    return static_cast<int>(e);
  }
  *out = c;
  return HDRNET_OK;
}

int hdrnet_host_ctx_destroy(hdrnet_host_ctx* c) {
  if (!c || c->magic != kCtxMagic) return HDRNET_E_BAD_CONTEXT;
  for (int s = 0; s < kSlots; ++s) cudaStreamSynchronize(c->stream[s]);
  free_buffers(c);
  for (int s = 0; s < kSlots; ++s) cudaStreamDestroy(c->stream[s]);
  cudaEventDestroy(c->grid_ready);
  c->magic = 0;
  delete c;
  return HDRNET_OK;
}

int hdrnet_slice_apply_host_f32(hdrnet_host_ctx* c, const float* grid, const float* guide,
                                const float* input, float* out, int B, int H, int W, int gh,
                                int gw, int gd, int n_in, int n_out, int has_offset) {
  if (!c || c->magic != kCtxMagic) return HDRNET_E_BAD_CONTEXT;
  if (B < 0 || H < 0 || W < 0 || gh < 1 || gw < 1 || gd < 1 || n_in < 1 || n_out < 1)
    return HDRNET_E_BAD_SHAPE;
  if (static_cast<long long>(B) * H * W == 0) return HDRNET_OK;
  if (!grid || !guide || !input || !out) return HDRNET_E_NULL_POINTER;

  const int J = n_in + (has_offset ? 1 : 0);
  const size_t grid_image = static_cast<size_t>(gh) * gw * gd * n_out * J;
  const size_t grid_bytes = grid_image * B * sizeof(float);

  // Band geometry: whole rows, at most band_pixels pixels, at least one row.
  size_t band_rows = c->band_pixels / static_cast<size_t>(W);
  if (band_rows < 1) band_rows = 1;
  if (band_rows > static_cast<size_t>(H)) band_rows = H;
  const size_t band_px = band_rows * W;
  const size_t guide_b = align256(band_px * 4);
  const size_t in_b = align256(band_px * 4 * n_in);
  const size_t out_b = align256(band_px * 4 * n_out);
  const size_t need = guide_b + in_b + out_b;

  cudaError_t e = cudaSuccess;
  if (need > c->slot_bytes || grid_bytes > c->grid_bytes) {
    for (int s = 0; s < kSlots; ++s) cudaStreamSynchronize(c->stream[s]);
    free_buffers(c);
    for (int s = 0; s < kSlots && e == cudaSuccess; ++s)
      e = cudaMalloc(reinterpret_cast<void**>(&c->slot_mem[s]), need);
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&c->grid_dev), grid_bytes);
    if (e != cudaSuccess) { free_buffers(c); return static_cast<int>(e); }
    c->slot_bytes = need;
    c->grid_bytes = grid_bytes;
  }

  e = cudaMemcpyAsync(c->grid_dev, grid, grid_bytes, cudaMemcpyHostToDevice, c->stream[0]);
  if (e == cudaSuccess) e = cudaEventRecord(c->grid_ready, c->stream[0]);
  for (int s = 1; s < kSlots && e == cudaSuccess; ++s)
    e = cudaStreamWaitEvent(c->stream[s], c->grid_ready, 0);
  if (e != cudaSuccess) return static_cast<int>(e);

  int rc = HDRNET_OK;
  long long k = 0;
  for (int b = 0; b < B && rc == HDRNET_OK; ++b) {
    for (int y0 = 0; y0 < H && rc == HDRNET_OK; y0 += static_cast<int>(band_rows), ++k) {
      const int rows = (H - y0 < static_cast<int>(band_rows)) ? (H - y0) : static_cast<int>(band_rows);
      const int s = static_cast<int>(k % kSlots);
      cudaStream_t st = c->stream[s];
      float* d_guide = reinterpret_cast<float*>(c->slot_mem[s]);
      float* d_in = reinterpret_cast<float*>(c->slot_mem[s] + guide_b);
      float* d_out = reinterpret_cast<float*>(c->slot_mem[s] + guide_b + in_b);
      const size_t pix0 = (static_cast<size_t>(b) * H + y0) * W;
      const size_t npx = static_cast<size_t>(rows) * W;
      e = cudaMemcpyAsync(d_guide, guide + pix0, npx * 4, cudaMemcpyHostToDevice, st);
      if (e == cudaSuccess)
        e = cudaMemcpyAsync(d_in, input + pix0 * n_in, npx * 4 * n_in, cudaMemcpyHostToDevice, st);
      if (e != cudaSuccess) { rc = static_cast<int>(e); break; }
      rc = hdrnet_b200::launch_slice_apply(c->grid_dev + static_cast<size_t>(b) * grid_image,
                                           d_guide, d_in, d_out, 1, H, W, rows, y0, gh, gw, gd,
                                           n_in, n_out, has_offset, HDRNET_VARIANT_AUTO, st);
      if (rc != HDRNET_OK) break;
      e = cudaMemcpyAsync(out + pix0 * n_out, d_out, npx * 4 * n_out, cudaMemcpyDeviceToHost, st);
      if (e != cudaSuccess) rc = static_cast<int>(e);
    }
  }
  for (int s = 0; s < kSlots; ++s) {
    e = cudaStreamSynchronize(c->stream[s]);
    if (e != cudaSuccess && rc == HDRNET_OK) rc = static_cast<int>(e);
  }
  return rc;
}

}  // extern "C"
