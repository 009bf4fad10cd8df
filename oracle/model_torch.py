"""torch-CPU restatement of the reference's model graphs -- TEST INFRASTRUCTURE ONLY.

A SECOND, independently written restatement of hdrnet/models.py + hdrnet/layers.py (SURVEY.md
section 8c asked for it), used to cross-check ``oracle/model_np.py``: the two share no layer code
-- this one goes through ``torch.nn.functional`` on NCHW tensors (``F.pad`` with explicit
asymmetric SAME pads + ``F.conv2d`` on OIHW weights, ``F.interpolate(align_corners=True)``,
``torch.split`` / ``torch.stack`` for ``unroll_grid`` exactly as the reference's ``tf.split`` /
``tf.stack``), the numpy one through strided slicing and matmuls on NHWC arrays.  Agreement to
1e-6 of range on every graph (tests/test_oracle.py), plus the hand-computed known answers in
tests/test_model_kats.py, is what pins the model oracle; TensorFlow itself cannot run here.

TensorFlow semantics restated (citations: hdrnet/layers.py:25-93, hdrnet/models.py:62-289):
  * conv 'SAME': out = ceil(in / s); total pad = max((out - 1) s + k - in, 0); floor(total / 2)
    BEFORE, the rest AFTER (asymmetric for stride 2 on even extents) -- torch's own padding=1
    would be symmetric and is NOT used;
  * weights HWIO -> OIHW by permute(3, 2, 0, 1); activations NHWC at every interface;
  * batch norm (inference, center=True, scale=False, eps = 1e-3): (x - mean) / sqrt(var + eps) + beta;
  * flatten: NHWC -> [B, h * w * c] (channel fastest), models.py:94-95;
  * every layer's activation is materialised in float32 (as TF does); sums run in float64.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-3
N_OUT = 3
N_IN = 4


def _t(a) -> torch.Tensor:
    return torch.from_numpy(np.ascontiguousarray(np.asarray(a, np.float64)))


def same_pad_amounts(size: int, k: int, s: int):
    out = (size + s - 1) // s
    total = max((out - 1) * s + k - size, 0)
    return total // 2, total - total // 2


def conv2d_same(x, w, stride=1) -> np.ndarray:
    """x [B,H,W,Cin], w [k,k,Cin,Cout] (HWIO) -> [B,ceil(H/s),ceil(W/s),Cout] float64."""
    xt = _t(x).permute(0, 3, 1, 2)                       # NHWC -> NCHW
    wt = _t(w).permute(3, 2, 0, 1)                       # HWIO -> OIHW
    k = wt.shape[-1]
    pt, pb = same_pad_amounts(xt.shape[2], k, stride)
    pl, pr = same_pad_amounts(xt.shape[3], k, stride)
    xt = F.pad(xt, (pl, pr, pt, pb))                     # (left, right, top, bottom)
    y = F.conv2d(xt, wt, bias=None, stride=stride, padding=0)
    return y.permute(0, 2, 3, 1).contiguous().numpy()


def _post(y, wts, scope, use_bn, use_bias, relu) -> np.ndarray:
    y = torch.from_numpy(np.asarray(y, np.float64))
    if use_bn:
        y = F.batch_norm(y.reshape(-1, y.shape[-1]), _t(wts[scope + "/BatchNorm/moving_mean"]),
                         _t(wts[scope + "/BatchNorm/moving_variance"]), weight=None,
                         bias=_t(wts[scope + "/BatchNorm/beta"]), training=False, eps=BN_EPS).reshape(y.shape)
    elif use_bias:
        y = y + _t(wts[scope + "/biases"])
    if relu:
        y = F.relu(y)
    return y.to(torch.float32).numpy()


def conv(x, wts, scope, stride=1, use_bias=True, batch_norm=False, relu=True):
    return _post(conv2d_same(x, wts[scope + "/weights"], stride), wts, scope, batch_norm, use_bias, relu)


def fc(x, wts, scope, use_bias=True, batch_norm=False, relu=True):
    y = F.linear(_t(x), _t(wts[scope + "/weights"]).t())     # x @ W[in, out]
    return _post(y.numpy(), wts, scope, batch_norm, use_bias, relu)


def coefficients(lowres, wts, params, prefix="inference/coefficients", n_out=N_OUT):
    """HDRNetCurves._coefficients, hdrnet/models.py:62-142."""
    gd = params["luma_bins"]
    bn = bool(params["batch_norm"])
    n_ds = int(np.log2(params["net_input_size"] / params["spatial_bin"]))
    x = np.asarray(lowres, np.float32)
    bs = x.shape[0]
    for i in range(n_ds):
        x = conv(x, wts, f"{prefix}/splat/conv{i + 1}", stride=2, batch_norm=bn and i > 0)
    splat = x
    g = splat
    for i in range(2):
        g = conv(g, wts, f"{prefix}/global/conv{i + 1}", stride=2, batch_norm=bn)
    g = torch.from_numpy(g).reshape(bs, -1).numpy()             # tf.reshape of the NHWC tensor
    g = fc(g, wts, f"{prefix}/global/fc1", batch_norm=bn)
    g = fc(g, wts, f"{prefix}/global/fc2", batch_norm=bn)
    g = fc(g, wts, f"{prefix}/global/fc3", relu=False)
    loc = conv(splat, wts, f"{prefix}/local/conv1", batch_norm=bn)
    loc = conv(loc, wts, f"{prefix}/local/conv2", use_bias=False, relu=False)
    fusion = F.relu(torch.from_numpy(loc) + torch.from_numpy(g).reshape(bs, 1, 1, -1)).numpy()
    pred = torch.from_numpy(conv(fusion, wts, f"{prefix}/prediction/conv1", relu=False))
    # unroll_grid (models.py:134-139), op for op:
    #   tf.stack(tf.split(x, n_out * n_in, axis=3), axis=4); tf.stack(tf.split(., n_in, axis=4), axis=5)
    pieces = torch.split(pred, gd, dim=3)
    assert len(pieces) == n_out * N_IN
    cur = torch.stack(pieces, dim=4)
    cur = torch.stack(torch.split(cur, n_out, dim=4), dim=5)
    return cur.contiguous().numpy()


def guide_curves(fullres, wts, prefix="inference/guide"):
    """HDRNetCurves._guide, hdrnet/models.py:145-190."""
    x = _t(fullres)
    shape = x.shape
    t = (x.reshape(-1, 3) @ _t(wts[prefix + "/ccm"]) + _t(wts[prefix + "/ccm_bias"])).reshape(shape)
    t = t.to(torch.float32).to(torch.float64).unsqueeze(4)                       # tf.expand_dims(., 4)
    shifts = _t(wts[prefix + "/shifts"]).reshape(1, 1, 3, -1)
    slopes = _t(wts[prefix + "/slopes"]).reshape(1, 1, 1, 3, -1)
    u = (slopes * F.relu(t - shifts)).sum(dim=4)
    u = u.to(torch.float32).to(torch.float64)
    w = _t(wts[prefix + "/channel_mixing/weights"]).reshape(1, 3, 1, 1)          # 1x1 conv, HWIO [1,1,3,1]
    y = F.conv2d(u.permute(0, 3, 1, 2), w, bias=_t(np.asarray(wts[prefix + "/channel_mixing/biases"]).reshape(-1)))
    return torch.clamp(y[:, 0], 0.0, 1.0).to(torch.float32).numpy()


def guide_nn(fullres, wts, prefix="inference/guide"):
    """HDRNetPointwiseNNGuide._guide, hdrnet/models.py:199-210."""
    h = conv(fullres, wts, prefix + "/conv1", batch_norm=True)
    y = torch.from_numpy(conv2d_same(h, wts[prefix + "/conv2/weights"])) + _t(wts[prefix + "/conv2/biases"])
    return torch.sigmoid(y)[..., 0].to(torch.float32).numpy()


def resize_bilinear_ac(x, oh, ow):
    """tf.image.resize_images(BILINEAR, align_corners=True): source = dst * (in - 1) / (out - 1)."""
    xt = _t(x).permute(0, 3, 1, 2)
    y = F.interpolate(xt, size=(oh, ow), mode="bilinear", align_corners=True)
    return y.permute(0, 2, 3, 1).to(torch.float32).contiguous().numpy()


def gaussian_pyr_inference(lowres, fullres, wts, params, slice_apply):
    """HDRNetGaussianPyrNN.inference, hdrnet/models.py:213-289."""
    coeffs = coefficients(lowres, wts, params, n_out=9)
    lvls = [np.asarray(fullres, np.float32)]
    h, w = lvls[0].shape[1:3]
    for _ in range(2):
        h, w = h // 2, w // 2
        lvls.append(resize_bilinear_ac(lvls[-1], h, w))
    guides = [guide_nn(lvl, wts, f"inference/guide/level_{i}") for i, lvl in enumerate(lvls)]
    B, gh, gw, gd = coeffs.shape[:4]
    current = None
    for il, (lvl, guide_lvl) in enumerate(reversed(list(zip(lvls, guides)))):    # Python-2 reversed(zip())
        c = np.ascontiguousarray(coeffs[:, :, :, :, il * 3:(il + 1) * 3, :]).reshape(B, gh, gw, gd, 12)
        out_lvl = slice_apply(c, guide_lvl, lvl, True)
        current = out_lvl if il == 0 else resize_bilinear_ac(current, out_lvl.shape[1], out_lvl.shape[2]) + out_lvl
    return current.astype(np.float32), coeffs, guides


def inference(lowres, fullres, wts, params, slice_apply):
    """HDRNetCurves.inference / HDRNetPointwiseNNGuide.inference, models.py:43-59."""
    coeffs = coefficients(lowres, wts, params)
    if params.get("model_name", "HDRNetCurves") == "HDRNetCurves":
        guide = guide_curves(fullres, wts)
    else:
        guide = guide_nn(fullres, wts)
    B, gh, gw, gd = coeffs.shape[:4]
    grid = coeffs.reshape(B, gh, gw, gd, N_OUT * N_IN)
    return slice_apply(grid, guide, np.asarray(fullres, np.float32), True), coeffs, guide
