/*
 * hdrnet_b200.h -- C-ABI of libhdrnet_b200.so: the B200 (sm_100a) implementation of
 * google/hdrnet's bilateral-slice hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch / TF types.  Each entry
 * point cites the reference interface it replaces.  All tensors use the reference op's
 * layout (TF row-major NHWC, last index fastest; hdrnet/ops/bilateral_slice_apply_op.cc:
 * 201-227):
 *
 *     grid   [B, gh, gw, gd, gc]   float32   gc = n_out * (n_in + has_offset), c = i*J + j
 *     guide  [B, H, W]             float32   expected in [0, 1], not enforced
 *     input  [B, H, W, n_in]       float32
 *     out    [B, H, W, n_out]      float32   (slice-apply)   /  [B, H, W, gc]  (slice)
 *
 * Conventions (replacing TF's OpKernel contract):
 *   - the caller owns and allocates every buffer, including outputs;
 *   - `*_f32` device entry points take DEVICE pointers valid on the current CUDA device and
 *     launch asynchronously on `stream` (a cudaStream_t passed as void*; NULL = default
 *     stream); they never allocate, never synchronise, and are re-entrant;
 *   - every function returns 0 on success, a negative HDRNET_E_* code for a contract
 *     violation (the conditions the reference raises InvalidArgument for), or a positive
 *     cudaError_t if a launch failed (the reference's Internal("... kernel failed."));
 *   - empty outputs (B*H*W == 0) succeed without launching (bilateral_slice_apply.cu.cc:
 *     373-379).
 */
#ifndef HDRNET_B200_H_
#define HDRNET_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HDRNET_B200_ABI_VERSION 1

/* Exported from libhdrnet_b200.so (the library is built with -fvisibility=hidden). */
#if defined(__GNUC__)
#define HDRNET_API __attribute__((visibility("default")))
#else
#define HDRNET_API
#endif

/* Contract violations (negative; positive return values are cudaError_t). */
#define HDRNET_OK 0
#define HDRNET_E_NULL_POINTER (-1)   /* a required pointer is NULL                         */
#define HDRNET_E_BAD_SHAPE (-2)      /* a dimension is negative, or gh/gw/gd/gc/n_* is < 1 */
#define HDRNET_E_BAD_CHANNELS (-3)   /* gc != n_out * (n_in + has_offset)                  */
#define HDRNET_E_TOO_LARGE (-4)      /* an extent does not fit the kernels' 32-bit indices */
#define HDRNET_E_UNSUPPORTED (-5)    /* the requested variant cannot run these shapes      */
#define HDRNET_E_BAD_CONTEXT (-6)    /* invalid / destroyed host-path context              */

/* Kernel selection for the *_variant debug entry points. */
#define HDRNET_VARIANT_AUTO 0    /* what the plain entry points use                        */
#define HDRNET_VARIANT_GENERIC 1 /* one thread per pixel, any shape / alignment            */
#define HDRNET_VARIANT_TMA 2     /* persistent TMA-staged row kernel (needs W % 4 == 0,    */
                                 /* 16-byte aligned buffers; n_in == 3, n_out == 3,        */
                                 /* has_offset for slice-apply)                            */
#define HDRNET_VARIANT_TEX 4     /* texture-assisted TMA kernel: a pre-pass writes the      */
                                 /* y-pre-blended slab rows to a caller workspace; the row  */
                                 /* kernel then fetches part of each pixel's corner data    */
                                 /* through the texture pipe, which does not share the      */
                                 /* shared-memory crossbar.  Needs hdrnet_slice_apply_f32_ws*/
#define HDRNET_VARIANT_TEX_ASYNC 7 /* the texture-assisted kernel with an ISSUER warp: math      */
                                 /* warps that never wait for each other (mbarrier arrive    */
                                 /* instead of block barriers) + one warp that issues every  */
                                 /* bulk copy; per-quad index arithmetic.  Same requirements */
                                 /* as HDRNET_VARIANT_TEX.  (Values 3, 5, 6, 8-11 named forms */
                                 /* that measured slower; they live in tools/experiments/.)   */
HDRNET_API int hdrnet_b200_abi_version(void);

/* Human-readable text for a return code of this library (static storage). */
HDRNET_API const char* hdrnet_b200_error_string(int code);

/*
 * Fused slice + affine apply.  Replaces the BilateralSliceApply op:
 *   Python   hdrnet/hdrnet_ops.py:31          hdrnet_ops.bilateral_slice_apply
 *   OpKernel hdrnet/ops/bilateral_slice_apply_op.cc:140-235  (Compute, GpuDevice)
 *   kernel   hdrnet/ops/bilateral_slice_apply.cu.cc:36-126, launcher :368-382
 * out[b,y,x,i] = sum_j trilerp(grid[..., i*J + j]) * (j < n_in ? input[b,y,x,j] : 1).
 */
HDRNET_API int hdrnet_slice_apply_f32(const float* grid, const float* guide, const float* input,
                           float* out, int B, int H, int W, int gh, int gw, int gd, int n_in,
                           int n_out, int has_offset, void* stream);

/* Same, with the kernel variant forced (tests exercise every variant on the same inputs). */
HDRNET_API int hdrnet_slice_apply_f32_variant(const float* grid, const float* guide, const float* input,
                                   float* out, int B, int H, int W, int gh, int gw, int gd,
                                   int n_in, int n_out, int has_offset, int variant,
                                   void* stream);

/*
 * Same op with a caller-provided device workspace (the library itself never allocates):
 * required by HDRNET_VARIANT_TEX, ignored by the other variants.
 * hdrnet_slice_apply_workspace_bytes() = B * H * gw * gd * 48.
 */
HDRNET_API size_t hdrnet_slice_apply_workspace_bytes(int B, int H, int gw, int gd);
HDRNET_API int hdrnet_slice_apply_f32_ws(const float* grid, const float* guide,
                                         const float* input, float* out, int B, int H, int W,
                                         int gh, int gw, int gd, int n_in, int n_out,
                                         int has_offset, int variant, void* workspace,
                                         size_t workspace_bytes, void* stream);

/*
 * Row band of the same op: guide / input / out hold `rows` image rows per image, starting at
 * image row `y_off` of images that are H rows tall ([B, rows, W(, c)] dense); the grid is whole.
 * The op is pointwise in (x, y) -- the kernel's only use of y is the grid coordinate
 * (y + 0.5) * gh / H, hdrnet/ops/bilateral_slice_apply.cu.cc:52, :75, :80, :88-90 -- so bands need no halo
 * and the bands of an image, computed anywhere, are bit for bit the rows of the whole-image call
 * by the same kernel.  This is the multi-GPU fallback when there are fewer images than GPUs
 * (each rank takes a row band and a copy of the 98 KB grid) and what the
 * host path streams.  workspace (optional, may be NULL / 0): hdrnet_slice_apply_workspace_bytes(B,
 * rows, gw, gd) bytes.  rows == H, y_off == 0 is hdrnet_slice_apply_f32_ws.
 */
HDRNET_API int hdrnet_slice_apply_rows_f32_ws(const float* grid, const float* guide,
                                              const float* input, float* out, int B, int H, int W,
                                              int rows, int y_off, int gh, int gw, int gd, int n_in,
                                              int n_out, int has_offset, int variant,
                                              void* workspace, size_t workspace_bytes, void* stream);

/*
 * Un-fused slice.  Replaces the BilateralSlice op:
 *   Python   hdrnet/hdrnet_ops.py:30          hdrnet_ops.bilateral_slice
 *   OpKernel hdrnet/ops/bilateral_slice_op.cc:120-174
 *   kernel   hdrnet/ops/bilateral_slice.cu.cc:34-91, launcher :230-244
 * out[b,y,x,c] = trilerp(grid[..., c]) at ((x+.5)*gw/W, (y+.5)*gh/H, guide*gd).
 */
HDRNET_API int hdrnet_slice_f32(const float* grid, const float* guide, float* out, int B, int H, int W,
                     int gh, int gw, int gd, int gc, void* stream);

HDRNET_API int hdrnet_slice_f32_variant(const float* grid, const float* guide, float* out, int B, int H,
                             int W, int gh, int gw, int gd, int gc, int variant, void* stream);

/*
 * Vector-Jacobian products (training).  Replace the BilateralSliceApplyGrad / BilateralSliceGrad
 * ops: Python registration hdrnet/hdrnet_ops.py:34-48; OpKernels
 * hdrnet/ops/bilateral_slice_apply_op.cc:249-362 and bilateral_slice_op.cc:183-256; kernels
 * hdrnet/ops/bilateral_slice_apply.cu.cc:128-364 (+ launcher :384-417) and
 * bilateral_slice.cu.cc:93-227 (+ :246-272).  codomain_tangent has the forward output's shape;
 * grid_vjp / guide_vjp / input_vjp have the shapes of grid / guide / input.  Semantics follow
 * the reference exactly (mirror boundary of the grid-VJP footprint, wz = 1 at the depth
 * borders, dwz = gd * SmoothedLerpWeightGrad); results are deterministic (no atomics).
 */
HDRNET_API int hdrnet_slice_apply_grad_f32(const float* grid, const float* guide,
                                           const float* input, const float* codomain_tangent,
                                           float* grid_vjp, float* guide_vjp, float* input_vjp,
                                           int B, int H, int W, int gh, int gw, int gd, int n_in,
                                           int n_out, int has_offset, void* stream);

HDRNET_API int hdrnet_slice_grad_f32(const float* grid, const float* guide,
                                     const float* codomain_tangent, float* grid_vjp,
                                     float* guide_vjp, int B, int H, int W, int gh, int gw,
                                     int gd, int gc, void* stream);

/*
 * Debug: the unclamped lower cell indices (gx0, gy0, gz0) the kernels use, written as
 * idx[b,y,x,0..2] int32.  It runs the same device functions as the slice kernels, so the
 * bit-exactness of the index arithmetic (bilateral_slice_apply.cu.cc:73-80;
 * jax/bilateral_slice.py:317-327) can be asserted against the oracle.
 */
HDRNET_API int hdrnet_slice_indices_i32(const float* guide, int32_t* idx, int B, int H, int W, int gh,
                             int gw, int gd, void* stream);

/*
 * Kernel-introspection for benchmarks: which variant AUTO would pick for these shapes, and
 * the launch geometry of the TMA kernel (CTAs, threads, dynamic shared memory bytes).
 */
HDRNET_API int hdrnet_slice_apply_plan(int B, int H, int W, int gh, int gw, int gd, int n_in, int n_out,
                            int has_offset, int* variant, int* ctas, int* threads,
                            int* smem_bytes);
/* Same, for a call that lends a workspace (hdrnet_slice_apply_f32_ws). */
HDRNET_API int hdrnet_slice_apply_plan_ws(int B, int H, int W, int gh, int gw, int gd, int n_in,
                                          int n_out, int has_offset, int with_workspace,
                                          int* variant, int* ctas, int* threads, int* smem_bytes);

/*
 * Full-resolution guidance maps (input [npix, 3] float32 RGB -> guide [npix] float32).
 * Coefficient arrays are HOST pointers, read before the call returns (they travel in the
 * kernel argument block).
 *
 * Curves guide.  Replaces HDRNetCurves._guide, hdrnet/models.py:145-190:
 *   t = rgb . ccm + ccm_bias;  u_c = sum_k slopes[c][k] * relu(t_c - shifts[c][k]);
 *   guide = clip(sum_c mix[c] * u_c + mix_bias, 0, 1)
 * ccm[3][3] is indexed [in][out] (tf.matmul(x, ccm), :156); shifts/slopes are [3][16]
 * (the reference's variables `shifts` [1,1,3,16] and `slopes` [1,1,1,3,16] flattened).
 */
HDRNET_API int hdrnet_guide_curves_f32(const float* input, float* guide, long long npix,
                                       const float* ccm, const float* ccm_bias,
                                       const float* shifts, const float* slopes,
                                       const float* mix, float mix_bias, void* stream);

/*
 * Pointwise-NN guide.  Replaces HDRNetPointwiseNNGuide._guide, hdrnet/models.py:199-210:
 *   h_f = relu(sum_c x_c * w1[c][f] + b1[f]);  guide = sigmoid(sum_f h_f * w2[f] + b2)
 * with conv1's batch norm already folded into w1[3][feats] / b1[feats] by the caller
 * (inference form, hdrnet/bin/freeze_graph.py:141-142).  feats <= 32.
 */
HDRNET_API int hdrnet_guide_nn_f32(const float* input, float* guide, long long npix,
                                   const float* w1, const float* b1, const float* w2, float b2,
                                   int feats, void* stream);

/*
 * Model-path forms of slice-apply: the guide is computed per pixel INSIDE the kernel from the
 * full-res RGB (the guide map never touches HBM: 24 B/px instead of 28 B/px + a guide pass).
 * Replaces HDRNetCurves.inference / HDRNetPointwiseNNGuide.inference's `_guide` + `_output`
 * (hdrnet/models.py:43-59, :145-196, :199-210).  n_in = 3, n_out = 3, has_offset.
 * guide_out: optional [B,H,W] dump of the guide (run.py --debug); may be NULL when the
 * shapes suit the fused kernel (W % 4 == 0, W >= 128, 16-byte aligned buffers); other shapes
 * run guide kernel + generic slice-apply and then REQUIRE guide_out as the intermediate.
 * Coefficient arrays are host pointers, as for hdrnet_guide_*_f32.
 */
HDRNET_API int hdrnet_slice_apply_curves_f32(const float* grid, const float* input, float* out,
                                             float* guide_out, int B, int H, int W, int gh,
                                             int gw, int gd, const float* ccm,
                                             const float* ccm_bias, const float* shifts,
                                             const float* slopes, const float* mix,
                                             float mix_bias, void* stream);

HDRNET_API int hdrnet_slice_apply_nn_f32(const float* grid, const float* input, float* out,
                                         float* guide_out, int B, int H, int W, int gh, int gw,
                                         int gd, const float* w1, const float* b1,
                                         const float* w2, float b2, int feats, void* stream);

/*
 * Coefficient network layers (device pointers; activations NHWC, float32).  Replace the TF
 * layers of HDRNetCurves._coefficients, hdrnet/models.py:62-142 / hdrnet/layers.py:25-93.
 * Batch norm (inference) is folded into w / bias by the caller.
 *
 * conv2d: k in {1, 3}, stride in {1, 2}, TF 'SAME' padding (total = max((ceil(in/s)-1)*s +
 * k - in, 0), floor(total/2) before, rest after), weights HWIO [k][k][Cin][Cout], optional
 * bias (NULL = none), optional ReLU.  out is [B, ceil(H/s), ceil(W/s), Cout].
 */
HDRNET_API int hdrnet_conv2d_nhwc_f32(const float* in, const float* w, const float* bias,
                                      float* out, int B, int H, int W, int Cin, int Cout, int k,
                                      int stride, int relu, void* stream);

/*
 * Tensor-core form of conv2d (tcgen05.mma kind::tf32 with 3xTF32 operand splitting, accumulator
 * in TMEM; float32-grade results).  Weights are packed ONCE per model into per-chunk hi/lo tiles
 * in the MMA's shared-memory layout (hdrnet_conv2d_tc_packed_bytes() bytes, device memory owned
 * by the caller); the layer call then needs Cin % 4 == 0, Cout % 16 == 0, 16 <= Cout <= 128.
 * hdrnet_conv2d_nhwc_f32 also reaches an unpacked tensor-core kernel on its own when the layer
 * has >= 96 tiles of 128 pixels (HDRNET_CONV_TCGEN05=0/1 overrides).
 */
HDRNET_API size_t hdrnet_conv2d_tc_packed_bytes(int k, int Cin, int Cout);
HDRNET_API int hdrnet_conv2d_tc_pack_f32(const float* w, float* packed, int k, int Cin, int Cout,
                                         void* stream);
HDRNET_API int hdrnet_conv2d_nhwc_tc_f32(const float* in, const float* packed_w, const float* bias,
                                         float* out, int B, int H, int W, int Cin, int Cout,
                                         int k, int stride, int relu, void* stream);

/* fully_connected: out[B,O] = in[B,I] @ w[I,O] + bias (+ReLU). */
HDRNET_API int hdrnet_fc_f32(const float* in, const float* w, const float* bias, float* out,
                             int B, int I, int O, int relu, void* stream);

/*
 * Fusion + prediction + unroll_grid (models.py:122-139) in one pass:
 *   f = relu(local[b,y,x,:] + global[b,:]);  p[o] = sum_c f[c] * w[c][o] + bias[o]
 *   grid[b,y,x,z,i,j] = p[(j*n_out + i)*gd + z]        (n_in counts the offset column)
 * local [B,gh,gw,C], global [B,C], w [C][gd*n_out*n_in], grid [B,gh,gw,gd,n_out*n_in].
 */
HDRNET_API int hdrnet_fuse_predict_f32(const float* local, const float* global_feat,
                                       const float* w, const float* bias, float* grid, int B,
                                       int gh, int gw, int C, int gd, int n_out, int n_in,
                                       void* stream);

/*
 * The WHOLE coefficient network (splat convs, global convs + fcs, local convs, fusion, prediction,
 * unroll_grid) behind one call: replaces HDRNetCurves._coefficients, hdrnet/models.py:62-142, with
 * batch norm folded by the caller.  At small batch the twelve layers are 8 launches -- the global
 * and the local branch (both read the splat features) share a launch per depth, fc1-fc3 run in one
 * 8-CTA cluster with activations in distributed shared memory -- chained with programmatic
 * dependent launch so that each layer's launch and weight fetch overlap the previous layer's tail
 * (batch 1: 65 us against 150 us for twelve per-layer calls, tools/time_cnn.py).
 *   lowres [B, S, S, 3] float32 -> grid [B, sb, sb, gd, n_out, n_in] float32
 *   weights / biases: HOST arrays of n_layers = n_ds + 8 DEVICE pointers (n_ds = log2(S / sb)), in
 *     the order splat conv1..n_ds, global conv1, conv2, fc1, fc2, fc3, local conv1, conv2,
 *     prediction conv1; conv weights HWIO, fc / prediction weights [in][out]; a bias may be NULL;
 *   scratch: hdrnet_coefficients_scratch_bytes(...) bytes of device memory, 16-byte aligned (every
 *     layer's activations, each in a buffer of its own; the library never allocates).  0 bytes =
 *     S / sb is not a power of two >= 2 (HDRNET_E_UNSUPPORTED from the call).
 * Layers whose shape a fast form does not take (channels % 4, fc widths not powers of two, large
 * batches) run the general kernels inside the same call.
 */
HDRNET_API size_t hdrnet_coefficients_scratch_bytes(int B, int net_input_size, int spatial_bin,
                                                    int luma_bins, int channel_multiplier, int n_out,
                                                    int n_in);
HDRNET_API int hdrnet_coefficients_f32(const float* lowres, float* grid, const float* const* weights,
                                       const float* const* biases, int n_layers, void* scratch,
                                       size_t scratch_bytes, int B, int net_input_size,
                                       int spatial_bin, int luma_bins, int channel_multiplier,
                                       int n_out, int n_in, void* stream);

/*
 * Bilinear resize, align_corners=True, NHWC, with an optional fused add (`add` has the output's
 * shape, or NULL).  Replaces tf.image.resize_images(BILINEAR, align_corners=True) in
 * HDRNetGaussianPyrNN._multiscale_input / ._output (hdrnet/models.py:249-289).
 */
HDRNET_API int hdrnet_resize_bilinear_f32(const float* in, const float* add, float* out, int B,
                                          int H, int W, int C, int OH, int OW, void* stream);

/* The model-path forms with a lent workspace (B*H*gw*gd*48 bytes, as for
 * hdrnet_slice_apply_f32_ws): large images then run the texture-assisted kernel. */
HDRNET_API int hdrnet_slice_apply_curves_f32_ws(const float* grid, const float* input, float* out,
                                                float* guide_out, int B, int H, int W, int gh,
                                                int gw, int gd, const float* ccm,
                                                const float* ccm_bias, const float* shifts,
                                                const float* slopes, const float* mix,
                                                float mix_bias, void* workspace,
                                                size_t workspace_bytes, void* stream);
HDRNET_API int hdrnet_slice_apply_nn_f32_ws(const float* grid, const float* input, float* out,
                                            float* guide_out, int B, int H, int W, int gh, int gw,
                                            int gd, const float* w1, const float* b1,
                                            const float* w2, float b2, int feats, void* workspace,
                                            size_t workspace_bytes, void* stream);

/*
 * Pixel storage formats of the model-path forms below.  The
 * reference's CLI decodes an image to uint8 / uint16, converts it on the host with
 * skimage.img_as_float (v / 255, v / 65535; hdrnet/bin/run.py:156-164), feeds float32 to the
 * graph and casts the prediction with tf.cast(255 * clip(x, 0, 1), tf.uint8) (run.py:95).  These
 * entry points keep the full-resolution image in its integer format on both sides: 3 + 3 bytes
 * per pixel cross PCIe and HBM instead of 12 + 12.
 */
#define HDRNET_PX_F32 0 /* float32 [B,H,W,3]                                              */
#define HDRNET_PX_U8 1  /* uint8   [B,H,W,3]; as input: img_as_float; as output: the cast  */
#define HDRNET_PX_U16 2 /* uint16  [B,H,W,3]; input only                                  */

/*
 * hdrnet_slice_apply_{curves,nn}_f32_ws with `input` in in_fmt and `out` in out_fmt
 * (HDRNET_PX_F32 or HDRNET_PX_U8).  The conversion of a code value is bit-exact with
 * float32(float64(v) / 255) (resp. 65535); the uint8 cast truncates like tf.cast.  (u8 | u16) ->
 * u8 with W % 16 == 0 (W % 8 for u16) and 16-byte aligned buffers runs the persistent row kernel;
 * every other combination runs a one-thread-per-pixel fused kernel.  guide_out (optional)
 * receives the float32 guide map.  f32 -> f32 is hdrnet_slice_apply_{curves,nn}_f32_ws itself.
 */
HDRNET_API int hdrnet_slice_apply_curves_px_ws(const float* grid, const void* input, int in_fmt,
                                               void* out, int out_fmt, float* guide_out, int B,
                                               int H, int W, int gh, int gw, int gd,
                                               const float* ccm, const float* ccm_bias,
                                               const float* shifts, const float* slopes,
                                               const float* mix, float mix_bias, void* workspace,
                                               size_t workspace_bytes, void* stream);
HDRNET_API int hdrnet_slice_apply_nn_px_ws(const float* grid, const void* input, int in_fmt,
                                           void* out, int out_fmt, float* guide_out, int B, int H,
                                           int W, int gh, int gw, int gd, const float* w1,
                                           const float* b1, const float* w2, float b2, int feats,
                                           void* workspace, size_t workspace_bytes, void* stream);

/*
 * Low-resolution network input straight from the decoded image: nearest-neighbour resize of
 * image [B,H,W,3] (fmt) to lowres [B,SH,SW,3] float32, img_as_float applied on the fly.
 * Replaces skimage.transform.resize(im, [S, S], order=0) on the float image (run.py:168-169):
 * output sample i reads input floor((i + 0.5) * H / SH).
 */
HDRNET_API int hdrnet_lowres_nearest_f32(const void* image, int fmt, float* lowres, int B, int H,
                                         int W, int SH, int SW, void* stream);

/*
 * Host-buffer path (what a CPU-tensor caller of the reference op gets: TF copies feeds to
 * the GPU and fetches back, hdrnet/bin/run.py:185).  A context owns device staging buffers
 * and streams; the call splits the batch into row bands, and pipelines H2D copy -> kernel
 * -> D2H copy across streams.  Host buffers should be page-locked for the copies to
 * overlap (pageable memory works but serialises).  Blocks until `out` is complete.
 */
typedef struct hdrnet_host_ctx hdrnet_host_ctx;

/* max_band_pixels: staging capacity per pipeline slot, in pixels (0 = default 4 Mi px). */
HDRNET_API int hdrnet_host_ctx_create(hdrnet_host_ctx** ctx, size_t max_band_pixels);
HDRNET_API int hdrnet_host_ctx_destroy(hdrnet_host_ctx* ctx);

HDRNET_API int hdrnet_slice_apply_host_f32(hdrnet_host_ctx* ctx, const float* grid, const float* guide,
                                const float* input, float* out, int B, int H, int W, int gh,
                                int gw, int gd, int n_in, int n_out, int has_offset);

#ifdef __cplusplus
} /* extern "C" */
#endif

#endif /* HDRNET_B200_H_ */
