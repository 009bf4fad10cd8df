"""numpy restatement of the reference's model graphs -- TEST INFRASTRUCTURE ONLY.

Follows hdrnet/models.py (HDRNetCurves, HDRNetPointwiseNNGuide) and hdrnet/layers.py
(conv, fc) with TensorFlow's semantics restated by hand:

* ``conv``  -> tf.contrib.layers.convolution2d, padding='SAME' (layers.py:25-59): TF SAME
  pads ``max((ceil(in/s)-1)*s + k - in, 0)`` in total, floor(half) BEFORE and the rest
  AFTER (asymmetric for stride 2 on even inputs); weights are HWIO; NHWC activations.
* ``fc``    -> tf.contrib.layers.fully_connected (layers.py:62-93): x @ W[in,out] + b.
* batch norm (inference): tf.contrib.layers.batch_norm with center=True, scale=False
  (layers.py:47-54): y = (x - moving_mean) / sqrt(moving_variance + eps) + beta, eps=1e-3
  (the contrib default; freeze_graph.py:134 reads it from the graph, :141-142 folds it
  without a gamma).
* flatten  -> tf.reshape(NHWC -> [bs, h*w*c]) (models.py:94-95): channel fastest.

Parity status: the reference has no test that exercises models.py and TensorFlow cannot run
here, so there is no reference-side golden vector for the CNN / guide numerics.  This restatement
is pinned instead by (i) oracle/model_torch.py, an independently written torch.nn.functional
restatement that must agree with it to 1e-6 of range on every graph, and (ii) hand-computed known
answers for the TensorFlow conventions a restatement can get wrong (tests/test_model_kats.py).
All sums accumulate in float64 and the result is rounded to float32 once, so this file is
the "exact" answer a float32 kernel is compared against with a stated tolerance.

Weights are a flat dict keyed by the reference's TF variable names under ``inference/``
(hdrnet/bin/run.py:92, freeze_graph.py:108-165), e.g.
``inference/coefficients/splat/conv1/weights`` [3,3,3,8] (HWIO).
"""
from __future__ import annotations

import math

import numpy as np

BN_EPS = 1e-3

DEFAULT_PARAMS = dict(  # hdrnet/bin/train.py:224-236
    model_name="HDRNetCurves", net_input_size=256, output_resolution=[512, 512],
    batch_norm=False, channel_multiplier=1, guide_complexity=16, luma_bins=8, spatial_bin=16)

N_OUT = 3       # HDRNetCurves.n_out(), models.py:34-36
N_IN = 4        # HDRNetCurves.n_in() = 3 + 1 (offset), models.py:38-40


# ---- layer primitives ------------------------------------------------------------------------
def same_pads(size: int, k: int, s: int):
    out = -(-size // s)
    total = max((out - 1) * s + k - size, 0)
    return out, total // 2, total - total // 2


def conv2d_same(x, w, stride=1):
    """x [B,H,W,Cin] float, w [k,k,Cin,Cout] HWIO -> [B,ceil(H/s),ceil(W/s),Cout] float64."""
    x = np.asarray(x, np.float64)
    w = np.asarray(w, np.float64)
    B, H, W, C = x.shape
    k = w.shape[0]
    oh, pt, pb = same_pads(H, k, stride)
    ow, pl, pr = same_pads(W, k, stride)
    xp = np.pad(x, ((0, 0), (pt, pb), (pl, pr), (0, 0)))
    out = np.zeros((B, oh, ow, w.shape[3]), np.float64)
    for dy in range(k):
        for dx in range(k):
            patch = xp[:, dy:dy + (oh - 1) * stride + 1:stride, dx:dx + (ow - 1) * stride + 1:stride, :]
            out += patch @ w[dy, dx]
    return out


def batch_norm_inference(x, beta, mean, var, eps=BN_EPS):
    return (x - mean) / np.sqrt(np.asarray(var, np.float64) + eps) + beta


def _layer_post(x, wts, scope, use_bn, use_bias, relu):
    if use_bn:
        x = batch_norm_inference(x, wts[scope + "/BatchNorm/beta"], wts[scope + "/BatchNorm/moving_mean"],
                                 wts[scope + "/BatchNorm/moving_variance"])
    elif use_bias:
        x = x + np.asarray(wts[scope + "/biases"], np.float64)
    if relu:
        x = np.maximum(x, 0.0)
    return x.astype(np.float32)  # TF materialises every activation in float32


def conv(x, wts, scope, stride=1, use_bias=True, batch_norm=False, relu=True):
    """hdrnet/layers.py:25-59."""
    y = conv2d_same(x, wts[scope + "/weights"], stride)
    return _layer_post(y, wts, scope, batch_norm, use_bias, relu)


def fc(x, wts, scope, use_bias=True, batch_norm=False, relu=True):
    """hdrnet/layers.py:62-93."""
    y = np.asarray(x, np.float64) @ np.asarray(wts[scope + "/weights"], np.float64)
    return _layer_post(y, wts, scope, batch_norm, use_bias, relu)


# ---- model graphs ----------------------------------------------------------------------------
def coefficients(lowres, wts, params, prefix="inference/coefficients", n_out=N_OUT):
    """HDRNetCurves._coefficients, hdrnet/models.py:62-142.  lowres [B,S,S,3] ->
    [B, sb, sb, gd, n_out, n_in] (sb = spatial_bin)."""
    gd = params["luma_bins"]
    bn = bool(params["batch_norm"])
    n_ds = int(np.log2(params["net_input_size"] / params["spatial_bin"]))
    x = np.asarray(lowres, np.float32)
    bs = x.shape[0]
    for i in range(n_ds):                                     # splat, :69-82
        x = conv(x, wts, f"{prefix}/splat/conv{i + 1}", stride=2, batch_norm=bn and i > 0)
    splat = x
    g = splat
    for i in range(2):                                        # global, :86-105
        g = conv(g, wts, f"{prefix}/global/conv{i + 1}", stride=2, batch_norm=bn)
    g = g.reshape(bs, -1)                                     # NHWC flatten, :94-95
    g = fc(g, wts, f"{prefix}/global/fc1", batch_norm=bn)
    g = fc(g, wts, f"{prefix}/global/fc2", batch_norm=bn)
    g = fc(g, wts, f"{prefix}/global/fc3", relu=False)
    loc = conv(splat, wts, f"{prefix}/local/conv1", batch_norm=bn)            # local, :109-118
    loc = conv(loc, wts, f"{prefix}/local/conv2", use_bias=False, relu=False)
    fusion = np.maximum(loc + g[:, None, None, :], 0.0).astype(np.float32)    # :122-125
    pred = conv(fusion, wts, f"{prefix}/prediction/conv1", relu=False)        # 1x1, :129-132
    # unroll_grid, :134-139: channel (j*n_out + i)*gd + z -> [b, y, x, z, i, j]
    B, sh, sw, _ = pred.shape
    return np.ascontiguousarray(pred.reshape(B, sh, sw, N_IN, n_out, gd).transpose(0, 1, 2, 5, 4, 3))


def guide_curves(fullres, wts, prefix="inference/guide"):
    """HDRNetCurves._guide, hdrnet/models.py:145-190.  [B,H,W,3] -> [B,H,W]."""
    x = np.asarray(fullres, np.float64)
    ccm = np.asarray(wts[prefix + "/ccm"], np.float64)                         # [3,3]
    t = x @ ccm + np.asarray(wts[prefix + "/ccm_bias"], np.float64)           # :156-160
    t = t.astype(np.float32).astype(np.float64)
    shifts = np.asarray(wts[prefix + "/shifts"], np.float64).reshape(3, -1)    # [1,1,3,16]
    slopes = np.asarray(wts[prefix + "/slopes"], np.float64).reshape(3, -1)    # [1,1,1,3,16]
    u = (slopes * np.maximum(t[..., None] - shifts, 0.0)).sum(-1)              # :175
    u = u.astype(np.float32).astype(np.float64)
    w = np.asarray(wts[prefix + "/channel_mixing/weights"], np.float64).reshape(3)
    b = float(np.asarray(wts[prefix + "/channel_mixing/biases"]).reshape(-1)[0])
    return np.clip(u @ w + b, 0.0, 1.0).astype(np.float32)                     # :177-188


def guide_nn(fullres, wts, prefix="inference/guide"):
    """HDRNetPointwiseNNGuide._guide, hdrnet/models.py:199-210 (conv1 always batch-normed)."""
    h = conv(fullres, wts, prefix + "/conv1", batch_norm=True)
    y = conv2d_same(h, wts[prefix + "/conv2/weights"]) + np.asarray(wts[prefix + "/conv2/biases"], np.float64)
    return (1.0 / (1.0 + np.exp(-y)))[..., 0].astype(np.float32)


# ---- synthetic weights (local_laplacian_sample is not in the tree) ---------------------------
def resize_bilinear_ac(x, oh, ow):
    """tf.image.resize_images(..., BILINEAR, align_corners=True), TF1 legacy kernel semantics."""
    x = np.asarray(x, np.float64)
    B, H, W, C = x.shape
    sy = (H - 1) / (oh - 1) if oh > 1 else 0.0
    sx = (W - 1) / (ow - 1) if ow > 1 else 0.0
    ys = np.arange(oh) * np.float32(sy)
    xs = np.arange(ow) * np.float32(sx)
    ys = ys.astype(np.float32).astype(np.float64)
    xs = xs.astype(np.float32).astype(np.float64)
    y0 = np.floor(ys).astype(int); x0 = np.floor(xs).astype(int)
    y1 = np.minimum(y0 + 1, H - 1); x1 = np.minimum(x0 + 1, W - 1)
    fy = (ys - y0)[None, :, None, None]; fx = (xs - x0)[None, None, :, None]
    top = x[:, y0][:, :, x0] + (x[:, y0][:, :, x1] - x[:, y0][:, :, x0]) * fx
    bot = x[:, y1][:, :, x0] + (x[:, y1][:, :, x1] - x[:, y1][:, :, x0]) * fx
    return (top + (bot - top) * fy).astype(np.float32)


def gaussian_pyr_inference(lowres, fullres, wts, params, slice_apply):
    """HDRNetGaussianPyrNN.inference, hdrnet/models.py:213-289 (Python-2 `reversed(zip())`
    restated: coarsest level first, using output rows 0..2)."""
    coeffs = coefficients(lowres, wts, params, n_out=9)           # [B,gh,gw,gd,9,4]
    lvls = [np.asarray(fullres, np.float32)]
    h, w = lvls[0].shape[1:3]
    for _ in range(2):
        h, w = h // 2, w // 2
        lvls.append(resize_bilinear_ac(lvls[-1], h, w))
    guides = [guide_nn(lvl, wts, f"inference/guide/level_{i}") for i, lvl in enumerate(lvls)]
    B, gh, gw, gd = coeffs.shape[:4]
    current = None
    for il in range(3):
        src = 2 - il
        c = np.ascontiguousarray(coeffs[:, :, :, :, il * 3:(il + 1) * 3, :]).reshape(B, gh, gw, gd, 12)
        out_lvl = slice_apply(c, guides[src], lvls[src], True)
        if il == 0:
            current = out_lvl
        else:
            current = resize_bilinear_ac(current, out_lvl.shape[1], out_lvl.shape[2]) + out_lvl
    return current.astype(np.float32), coeffs, guides


def make_weights(params, seed=0, model_name=None, trained_like=True):
    """Seeded random weights with the reference's variable names and shapes.
    ``trained_like`` perturbs the guide parameters away from their identity init so the test
    exercises every term (init: models.py:150-186)."""
    rng = np.random.RandomState(seed)
    gd, cm = params["luma_bins"], params["channel_multiplier"]
    bn = bool(params["batch_norm"])
    model_name = model_name or params.get("model_name", "HDRNetCurves")
    w = {}

    def add_conv(scope, k, cin, cout, use_bias=True, use_bn=False):
        fan_in = k * k * cin
        w[scope + "/weights"] = (rng.randn(k, k, cin, cout) * math.sqrt(2.0 / fan_in)).astype(np.float32)
        add_post(scope, cout, use_bias, use_bn)

    def add_fc(scope, cin, cout, use_bias=True, use_bn=False):
        w[scope + "/weights"] = (rng.randn(cin, cout) * math.sqrt(2.0 / cin)).astype(np.float32)
        add_post(scope, cout, use_bias, use_bn)

    def add_post(scope, cout, use_bias, use_bn):
        if use_bn:
            w[scope + "/BatchNorm/beta"] = (0.1 * rng.randn(cout)).astype(np.float32)
            w[scope + "/BatchNorm/moving_mean"] = (0.1 * rng.randn(cout)).astype(np.float32)
            w[scope + "/BatchNorm/moving_variance"] = (0.5 + rng.rand(cout)).astype(np.float32)
        elif use_bias:
            w[scope + "/biases"] = (0.05 * rng.randn(cout)).astype(np.float32)

    p = "inference/coefficients"
    n_ds = int(np.log2(params["net_input_size"] / params["spatial_bin"]))
    cin = 3
    for i in range(n_ds):
        cout = cm * (2 ** i) * gd
        add_conv(f"{p}/splat/conv{i + 1}", 3, cin, cout, use_bn=bn and i > 0)
        cin = cout
    splat_c = cin
    c8 = 8 * cm * gd
    gin = splat_c
    for i in range(2):
        add_conv(f"{p}/global/conv{i + 1}", 3, gin, c8, use_bn=bn)
        gin = c8
    sb = params["spatial_bin"]
    flat = (sb // 4) * (sb // 4) * c8
    add_fc(f"{p}/global/fc1", flat, 32 * cm * gd, use_bn=bn)
    add_fc(f"{p}/global/fc2", 32 * cm * gd, 16 * cm * gd, use_bn=bn)
    add_fc(f"{p}/global/fc3", 16 * cm * gd, c8)
    add_conv(f"{p}/local/conv1", 3, splat_c, c8, use_bn=bn)
    add_conv(f"{p}/local/conv2", 3, c8, c8, use_bias=False)
    n_out = 9 if model_name == "HDRNetGaussianPyrNN" else N_OUT
    add_conv(f"{p}/prediction/conv1", 1, c8, gd * n_out * N_IN)

    g = "inference/guide"
    if model_name == "HDRNetGaussianPyrNN":
        nf = params["guide_complexity"]
        for lvl in range(3):
            add_conv(f"{g}/level_{lvl}/conv1", 1, 3, nf, use_bn=True)
            add_conv(f"{g}/level_{lvl}/conv2", 1, nf, 1)
    elif model_name == "HDRNetCurves":
        npts = 16
        ccm = np.identity(3, dtype=np.float32)
        shifts = np.tile(np.linspace(0, 1, npts, endpoint=False, dtype=np.float32)[None, None, None, :],
                         (1, 1, 3, 1))
        slopes = np.zeros((1, 1, 1, 3, npts), np.float32)
        slopes[..., 0] = 1.0
        mix = np.full((1, 1, 3, 1), 1.0 / 3.0, np.float32)
        bias = np.zeros((1,), np.float32)
        ccm_bias = np.zeros((3,), np.float32)
        if trained_like:
            ccm = ccm + (0.1 * rng.randn(3, 3)).astype(np.float32)
            ccm_bias = (0.02 * rng.randn(3)).astype(np.float32)
            shifts = (shifts + 0.02 * rng.randn(*shifts.shape)).astype(np.float32)
            slopes = (slopes + 0.15 * rng.randn(*slopes.shape)).astype(np.float32)
            mix = (mix + 0.05 * rng.randn(*mix.shape)).astype(np.float32)
            bias = (0.02 * rng.randn(1)).astype(np.float32)
        w.update({g + "/ccm": ccm, g + "/ccm_bias": ccm_bias, g + "/shifts": shifts,
                  g + "/slopes": slopes, g + "/channel_mixing/weights": mix,
                  g + "/channel_mixing/biases": bias})
    else:
        nf = params["guide_complexity"]
        add_conv(g + "/conv1", 1, 3, nf, use_bn=True)
        add_conv(g + "/conv2", 1, nf, 1)
    return w


def inference(lowres, fullres, wts, params, slice_apply):
    """HDRNetCurves.inference / HDRNetPointwiseNNGuide.inference, models.py:43-59."""
    coeffs = coefficients(lowres, wts, params)
    if params.get("model_name", "HDRNetCurves") == "HDRNetCurves":
        guide = guide_curves(fullres, wts)
    else:
        guide = guide_nn(fullres, wts)
    B, gh, gw, gd = coeffs.shape[:4]
    grid = coeffs.reshape(B, gh, gw, gd, N_OUT * N_IN)   # layers.py:141-144
    return slice_apply(grid, guide, np.asarray(fullres, np.float32), True), coeffs, guide
