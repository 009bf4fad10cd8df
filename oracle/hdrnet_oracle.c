/*
 * hdrnet_oracle.c -- CPU restatement of google/hdrnet's bilateral slice hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity oracle for hdrnet_b200's CUDA
 * kernels.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may build, load or call it.  The product (hdrnet_b200/) never
 * does: it fails loudly when its CUDA library is missing.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this port against
 *   (1) the reference's own jax/bilateral_slice.py executed under a numpy stand-in for
 *       jax (oracle/jax_shim.py; fixtures committed in tests/golden/ with the generating
 *       script tests/golden/make_golden.py),
 *   (2) the reference's own C++ loops (hdrnet/ops/bilateral_slice{,_apply}.cc) compiled
 *       unmodified into oracle/_ref/ against a stand-in for the un-vendored nda header
 *       (oracle/Makefile) -- bit-exact agreement is asserted,
 *   (3) the known-answer test hdrnet/test/ops_test.py:61-86 (test_interpolate).
 *
 * Every function follows the reference loops op for op in float32, in the same
 * accumulation order, so that agreement with (2) is bit-exact.  Build with
 * -ffp-contract=off (see oracle/Makefile) so the host compiler cannot fuse a*b+c.
 *
 * Layouts are the TF op's: row-major, last index fastest
 *   grid  [B, gh, gw, gd, gc]      (hdrnet/ops/bilateral_slice_apply_op.cc:201-207)
 *   guide [B, H, W]
 *   input [B, H, W, n_in]
 *   out   [B, H, W, n_out]   with grid channel c = i * (n_in + has_offset) + j.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define HDRNET_ORACLE_API __attribute__((visibility("default")))

/* hdrnet/ops/numerics.h:53-57 (LerpWeight) == jax/numerics.py:20-40 */
static inline float lerp_weight(float x, float xs) {
  const float dx = x - xs;
  const float abs_dx = fabsf(dx);
  return fmaxf(1.0f - abs_dx, 0.0f);
}

/* hdrnet/ops/numerics.h:83-85 (SmoothedAbs) == jax/numerics.py:43-45 */
static inline float smoothed_abs(float x) { return sqrtf(x * x + 1.0e-8f); }

/* hdrnet/ops/numerics.h:89-91 (SmoothedAbsGrad) */
static inline float smoothed_abs_grad(float x) { return x / sqrtf(x * x + 1.0e-8f); }

/* hdrnet/ops/numerics.h:108-113 (SmoothedLerpWeight) == jax/numerics.py:63-89 */
static inline float smoothed_lerp_weight(float x, float xs) {
  const float dx = x - xs;
  const float abs_dx = smoothed_abs(dx);
  return fmaxf(1.0f - abs_dx, 0.0f);
}

/* hdrnet/ops/numerics.h:116-126 (SmoothedLerpWeightGrad) */
static inline float smoothed_lerp_weight_grad(float x, float xs) {
  const float dx = x - xs;
  const float abs_dx = smoothed_abs(dx);
  if (abs_dx > 1.0f) return 0.0f;
  return smoothed_abs_grad(dx);
}

/* hdrnet/ops/numerics.h:72-80 (MirrorBoundary) */
static inline int mirror_boundary(int x, int extent) {
  if (x < 0) return -x - 1;
  if (x >= extent) return 2 * extent - 1 - x;
  return x;
}

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

HDRNET_ORACLE_API int hdrnet_oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/*
 * Cell indices (gx0, gy0, gz0) exactly as the reference computes them:
 * hdrnet/ops/bilateral_slice_apply.cc:38-49 (scale = float(gw)/W, (x+0.5f)*scale,
 * floor(g-0.5f)).  idx is [B, H, W, 3] int32 = (gx0, gy0, gz0), unclamped.
 */
HDRNET_ORACLE_API void hdrnet_oracle_slice_indices(const float* guide, int32_t* idx, int B, int H,
                                                   int W, int gh, int gw, int gd) {
  const float scale_x = (float)gw / W;
  const float scale_y = (float)gh / H;
#pragma omp parallel for collapse(2) schedule(static)
  for (int b = 0; b < B; ++b) {
    for (int y = 0; y < H; ++y) {
      for (int x = 0; x < W; ++x) {
        const size_t p = ((size_t)b * H + y) * W + x;
        const float gxf = (x + 0.5f) * scale_x;
        const float gyf = (y + 0.5f) * scale_y;
        const float gzf = guide[p] * gd;
        idx[3 * p + 0] = (int)floorf(gxf - 0.5f);
        idx[3 * p + 1] = (int)floorf(gyf - 0.5f);
        idx[3 * p + 2] = (int)floorf(gzf - 0.5f);
      }
    }
  }
}

/* hdrnet/ops/bilateral_slice.cc:25-70 (BilateralSlice) */
HDRNET_ORACLE_API void hdrnet_oracle_slice(const float* grid, const float* guide, float* out,
                                           int B, int H, int W, int gh, int gw, int gd, int gc) {
  const float scale_x = (float)gw / W;
  const float scale_y = (float)gh / H;
#pragma omp parallel for collapse(2) schedule(static)
  for (int b = 0; b < B; ++b) {
    for (int y = 0; y < H; ++y) {
      for (int x = 0; x < W; ++x) {
        const size_t p = ((size_t)b * H + y) * W + x;
        const float gxf = (x + 0.5f) * scale_x;
        const float gyf = (y + 0.5f) * scale_y;
        const float gzf = guide[p] * gd;
        const int gx0 = (int)floorf(gxf - 0.5f);
        const int gy0 = (int)floorf(gyf - 0.5f);
        const int gz0 = (int)floorf(gzf - 0.5f);
        for (int c = 0; c < gc; ++c) {
          float value = 0.0f;
          for (int gy = gy0; gy < gy0 + 2; ++gy) {
            const int gyc = clampi(gy, 0, gh - 1);
            const float wy = lerp_weight(gy + 0.5f, gyf);
            for (int gx = gx0; gx < gx0 + 2; ++gx) {
              const int gxc = clampi(gx, 0, gw - 1);
              const float wx = lerp_weight(gx + 0.5f, gxf);
              for (int gz = gz0; gz < gz0 + 2; ++gz) {
                const int gzc = clampi(gz, 0, gd - 1);
                const float wz = smoothed_lerp_weight(gz + 0.5f, gzf);
                const size_t g = ((((size_t)b * gh + gyc) * gw + gxc) * gd + gzc) * gc + c;
                value += wx * wy * wz * grid[g];
              }
            }
          }
          out[p * gc + c] = value;
        }
      }
    }
  }
}

/* hdrnet/ops/bilateral_slice_apply.cc:24-82 (BilateralSliceApply) */
HDRNET_ORACLE_API void hdrnet_oracle_slice_apply(const float* grid, const float* guide,
                                                 const float* input, float* out, int B, int H,
                                                 int W, int gh, int gw, int gd, int n_in,
                                                 int n_out, int has_offset) {
  const int J = n_in + (has_offset ? 1 : 0); /* grid_input_channels */
  const int gc = n_out * J;
  const float scale_x = (float)gw / W;
  const float scale_y = (float)gh / H;
#pragma omp parallel for collapse(2) schedule(static)
  for (int b = 0; b < B; ++b) {
    for (int y = 0; y < H; ++y) {
      for (int x = 0; x < W; ++x) {
        const size_t p = ((size_t)b * H + y) * W + x;
        const float gxf = (x + 0.5f) * scale_x;
        const float gyf = (y + 0.5f) * scale_y;
        const float gzf = guide[p] * gd;
        const int gx0 = (int)floorf(gxf - 0.5f);
        const int gy0 = (int)floorf(gyf - 0.5f);
        const int gz0 = (int)floorf(gzf - 0.5f);
        for (int i = 0; i < n_out; ++i) {
          float value = 0.0f;
          for (int j = 0; j < J; ++j) {
            float grid_sample = 0.0f;
            for (int gy = gy0; gy < gy0 + 2; ++gy) {
              const int gyc = clampi(gy, 0, gh - 1);
              const float wy = lerp_weight(gy + 0.5f, gyf);
              for (int gx = gx0; gx < gx0 + 2; ++gx) {
                const int gxc = clampi(gx, 0, gw - 1);
                const float wx = lerp_weight(gx + 0.5f, gxf);
                for (int gz = gz0; gz < gz0 + 2; ++gz) {
                  const int gzc = clampi(gz, 0, gd - 1);
                  const float wz = smoothed_lerp_weight(gz + 0.5f, gzf);
                  const size_t g =
                      ((((size_t)b * gh + gyc) * gw + gxc) * gd + gzc) * gc + (size_t)i * J + j;
                  grid_sample += wx * wy * wz * grid[g];
                }
              }
            }
            if (j < n_in) {
              value += grid_sample * input[p * n_in + j];
            } else {
              value += grid_sample;
            }
          }
          out[p * n_out + i] = value;
        }
      }
    }
  }
}

/*
 * Backward loops (row f-1 "next"; restated now so the forward oracle and the future
 * VJP kernels share one checker).
 */

/* hdrnet/ops/bilateral_slice_apply.cc:84-138 (BilateralSliceApplyGridGrad) */
HDRNET_ORACLE_API void hdrnet_oracle_slice_apply_grid_grad(const float* guide, const float* input,
                                                           const float* ct, float* vjp, int B,
                                                           int H, int W, int gh, int gw, int gd,
                                                           int n_in, int n_out, int has_offset) {
  const int J = n_in + (has_offset ? 1 : 0);
  const int gc = n_out * J;
  const float scale_x = (float)W / gw;
  const float scale_y = (float)H / gh;
#pragma omp parallel for collapse(3) schedule(dynamic)
  for (int b = 0; b < B; ++b) {
    for (int gy = 0; gy < gh; ++gy) {
      for (int gx = 0; gx < gw; ++gx) {
        const int x0 = (int)floorf(scale_x * (gx + 0.5f - 1.0f));
        const int x1e = (int)ceilf(scale_x * (gx + 0.5f + 1.0f));
        const int y0 = (int)floorf(scale_y * (gy + 0.5f - 1.0f));
        const int y1e = (int)ceilf(scale_y * (gy + 0.5f + 1.0f));
        for (int gz = 0; gz < gd; ++gz) {
          for (int i = 0; i < n_out; ++i) {
            for (int j = 0; j < J; ++j) {
              float v = 0.0f;
              for (int y = y0; y < y1e; ++y) {
                const int ym = mirror_boundary(y, H);
                const float gyf = (y + 0.5f) / scale_y;
                const float wy = lerp_weight(gy + 0.5f, gyf);
                for (int x = x0; x < x1e; ++x) {
                  const int xm = mirror_boundary(x, W);
                  const float gxf = (x + 0.5f) / scale_x;
                  const float wx = lerp_weight(gx + 0.5f, gxf);
                  const size_t p = ((size_t)b * H + ym) * W + xm;
                  const float gzf = guide[p] * gd;
                  float wz = smoothed_lerp_weight(gz + 0.5f, gzf);
                  if ((gz == 0 && gzf < 0.5f) || (gz == gd - 1 && gzf > gd - 0.5f)) wz = 1.0f;
                  const float iv = (j < n_in) ? input[p * n_in + j] : 1.0f;
                  const float gv = wx * wy * wz * iv;
                  v += gv * ct[p * n_out + i];
                }
              }
              vjp[((((size_t)b * gh + gy) * gw + gx) * gd + gz) * gc + (size_t)i * J + j] = v;
            }
          }
        }
      }
    }
  }
}

/* hdrnet/ops/bilateral_slice_apply.cc:140-206 (BilateralSliceApplyGuideGrad) */
HDRNET_ORACLE_API void hdrnet_oracle_slice_apply_guide_grad(const float* grid, const float* guide,
                                                            const float* input, const float* ct,
                                                            float* vjp, int B, int H, int W,
                                                            int gh, int gw, int gd, int n_in,
                                                            int n_out, int has_offset) {
  const int J = n_in + (has_offset ? 1 : 0);
  const int gc = n_out * J;
  const float scale_x = (float)gw / W;
  const float scale_y = (float)gh / H;
#pragma omp parallel for collapse(2) schedule(static)
  for (int b = 0; b < B; ++b) {
    for (int y = 0; y < H; ++y) {
      for (int x = 0; x < W; ++x) {
        const size_t p = ((size_t)b * H + y) * W + x;
        const float gxf = (x + 0.5f) * scale_x;
        const float gyf = (y + 0.5f) * scale_y;
        const float gzf = guide[p] * gd;
        const int gx0 = (int)floorf(gxf - 0.5f);
        const int gy0 = (int)floorf(gyf - 0.5f);
        const int gz0 = (int)floorf(gzf - 0.5f);
        float v = 0.0f;
        for (int i = 0; i < n_out; ++i) {
          float grad_value = 0.0f;
          for (int j = 0; j < J; ++j) {
            float grid_sample = 0.0f;
            for (int gy = gy0; gy < gy0 + 2; ++gy) {
              const int gyc = clampi(gy, 0, gh - 1);
              const float wy = lerp_weight(gy + 0.5f, gyf);
              for (int gx = gx0; gx < gx0 + 2; ++gx) {
                const int gxc = clampi(gx, 0, gw - 1);
                const float wx = lerp_weight(gx + 0.5f, gxf);
                for (int gz = gz0; gz < gz0 + 2; ++gz) {
                  const int gzc = clampi(gz, 0, gd - 1);
                  const float dwz = gd * smoothed_lerp_weight_grad(gz + 0.5f, gzf);
                  const size_t g =
                      ((((size_t)b * gh + gyc) * gw + gxc) * gd + gzc) * gc + (size_t)i * J + j;
                  grid_sample += wx * wy * dwz * grid[g];
                }
              }
            }
            const float iv = (j < n_in) ? input[p * n_in + j] : 1.0f;
            grad_value += grid_sample * iv;
          }
          v += grad_value * ct[p * n_out + i];
        }
        vjp[p] = v;
      }
    }
  }
}

/* hdrnet/ops/bilateral_slice_apply.cc:208-259 (BilateralSliceApplyInputGrad) */
HDRNET_ORACLE_API void hdrnet_oracle_slice_apply_input_grad(const float* grid, const float* guide,
                                                            const float* ct, float* vjp, int B,
                                                            int H, int W, int gh, int gw, int gd,
                                                            int n_in, int n_out, int has_offset) {
  const int J = n_in + (has_offset ? 1 : 0);
  const int gc = n_out * J;
  const float scale_x = (float)gw / W;
  const float scale_y = (float)gh / H;
#pragma omp parallel for collapse(2) schedule(static)
  for (int b = 0; b < B; ++b) {
    for (int y = 0; y < H; ++y) {
      for (int x = 0; x < W; ++x) {
        const size_t p = ((size_t)b * H + y) * W + x;
        const float gxf = (x + 0.5f) * scale_x;
        const float gyf = (y + 0.5f) * scale_y;
        const float gzf = guide[p] * gd;
        const int gx0 = (int)floorf(gxf - 0.5f);
        const int gy0 = (int)floorf(gyf - 0.5f);
        const int gz0 = (int)floorf(gzf - 0.5f);
        for (int j = 0; j < n_in; ++j) {
          float v = 0.0f;
          for (int i = 0; i < n_out; ++i) {
            float grad_value = 0.0f;
            for (int gy = gy0; gy < gy0 + 2; ++gy) {
              const int gyc = clampi(gy, 0, gh - 1);
              const float wy = lerp_weight(gy + 0.5f, gyf);
              for (int gx = gx0; gx < gx0 + 2; ++gx) {
                const int gxc = clampi(gx, 0, gw - 1);
                const float wx = lerp_weight(gx + 0.5f, gxf);
                for (int gz = gz0; gz < gz0 + 2; ++gz) {
                  const int gzc = clampi(gz, 0, gd - 1);
                  const float wz = smoothed_lerp_weight(gz + 0.5f, gzf);
                  const size_t g =
                      ((((size_t)b * gh + gyc) * gw + gxc) * gd + gzc) * gc + (size_t)i * J + j;
                  grad_value += wx * wy * wz * grid[g];
                }
              }
            }
            v += grad_value * ct[p * n_out + i];
          }
          vjp[p * n_in + j] = v;
        }
      }
    }
  }
}

/* hdrnet/ops/bilateral_slice.cc:72-118 (BilateralSliceGridGrad) */
HDRNET_ORACLE_API void hdrnet_oracle_slice_grid_grad(const float* guide, const float* ct,
                                                     float* vjp, int B, int H, int W, int gh,
                                                     int gw, int gd, int gc) {
  const float scale_x = (float)W / gw;
  const float scale_y = (float)H / gh;
#pragma omp parallel for collapse(3) schedule(dynamic)
  for (int b = 0; b < B; ++b) {
    for (int gy = 0; gy < gh; ++gy) {
      for (int gx = 0; gx < gw; ++gx) {
        const int x0 = (int)floorf(scale_x * (gx + 0.5f - 1.0f));
        const int x1e = (int)ceilf(scale_x * (gx + 0.5f + 1.0f));
        const int y0 = (int)floorf(scale_y * (gy + 0.5f - 1.0f));
        const int y1e = (int)ceilf(scale_y * (gy + 0.5f + 1.0f));
        for (int gz = 0; gz < gd; ++gz) {
          for (int c = 0; c < gc; ++c) {
            float v = 0.0f;
            for (int y = y0; y < y1e; ++y) {
              const int ym = mirror_boundary(y, H);
              const float gyf = (y + 0.5f) / scale_y;
              const float wy = lerp_weight(gy + 0.5f, gyf);
              for (int x = x0; x < x1e; ++x) {
                const int xm = mirror_boundary(x, W);
                const float gxf = (x + 0.5f) / scale_x;
                const float wx = lerp_weight(gx + 0.5f, gxf);
                const size_t p = ((size_t)b * H + ym) * W + xm;
                const float gzf = guide[p] * gd;
                float wz = smoothed_lerp_weight(gz + 0.5f, gzf);
                if ((gz == 0 && gzf < 0.5f) || (gz == gd - 1 && gzf > gd - 0.5f)) wz = 1.0f;
                v += wz * wx * wy * ct[p * gc + c]; /* reference's order, :111 */
              }
            }
            vjp[((((size_t)b * gh + gy) * gw + gx) * gd + gz) * gc + c] = v;
          }
        }
      }
    }
  }
}

/* hdrnet/ops/bilateral_slice.cc:120-168 (BilateralSliceGuideGrad) */
HDRNET_ORACLE_API void hdrnet_oracle_slice_guide_grad(const float* grid, const float* guide,
                                                      const float* ct, float* vjp, int B, int H,
                                                      int W, int gh, int gw, int gd, int gc) {
  const float scale_x = (float)gw / W;
  const float scale_y = (float)gh / H;
#pragma omp parallel for collapse(2) schedule(static)
  for (int b = 0; b < B; ++b) {
    for (int y = 0; y < H; ++y) {
      for (int x = 0; x < W; ++x) {
        const size_t p = ((size_t)b * H + y) * W + x;
        const float gxf = (x + 0.5f) * scale_x;
        const float gyf = (y + 0.5f) * scale_y;
        const float gzf = guide[p] * gd;
        const int gx0 = (int)floorf(gxf - 0.5f);
        const int gy0 = (int)floorf(gyf - 0.5f);
        const int gz0 = (int)floorf(gzf - 0.5f);
        float v = 0.0f;
        for (int c = 0; c < gc; ++c) {
          float grid_sample = 0.0f;
          for (int gy = gy0; gy < gy0 + 2; ++gy) {
            const int gyc = clampi(gy, 0, gh - 1);
            const float wy = lerp_weight(gy + 0.5f, gyf);
            for (int gx = gx0; gx < gx0 + 2; ++gx) {
              const int gxc = clampi(gx, 0, gw - 1);
              const float wx = lerp_weight(gx + 0.5f, gxf);
              for (int gz = gz0; gz < gz0 + 2; ++gz) {
                const int gzc = clampi(gz, 0, gd - 1);
                const float dwz = gd * smoothed_lerp_weight_grad(gz + 0.5f, gzf);
                const size_t g = ((((size_t)b * gh + gyc) * gw + gxc) * gd + gzc) * gc + c;
                grid_sample += wx * wy * dwz * grid[g];
              }
            }
          }
          v += grid_sample * ct[p * gc + c];
        }
        vjp[p] = v;
      }
    }
  }
}
