"""Known-answer tests that pin the MODEL oracle.

The reference has no test for hdrnet/models.py and TensorFlow cannot run here, so the model graphs
are pinned three ways:
  1. two independently written restatements -- oracle/model_np.py (numpy, NHWC slicing) and
     oracle/model_torch.py (torch.nn.functional, NCHW) -- must agree to 1e-6 of range on every
     graph and parameter set;
  2. HAND-COMPUTED vectors for exactly the TensorFlow conventions a restatement can get wrong:
     the asymmetric stride-2 'SAME' padding on even and odd extents, the NHWC flatten -> fc order,
     the `unroll_grid` channel map ch = (j * n_out + i) * gd + z, align_corners bilinear resize;
  3. (-m gpu) the CUDA layers against the same hand-computed vectors.
Every expected number below was worked out by hand from the TF op definitions cited next to it.
"""
import numpy as np
import pytest
import torch

import oracle
from oracle import model_np as M
from oracle import model_torch as T

RESTATEMENTS = [M, T]
IDS = ["model_np", "model_torch"]


# ---- hand-computed vectors ---------------------------------------------------------------------
def _row(vals):
    return np.asarray(vals, np.float32).reshape(1, 1, -1, 1)


def _k3(vals):
    """3 x 3 kernel whose middle row is `vals` (the restatements take square kernels, as the
    reference's layers do).  On a one-row image (H = 1: one zero row above and below, stride 1 or
    2) only the x direction matters."""
    k = np.zeros((3, 3, 1, 1), np.float32)
    k[1, :, 0, 0] = vals
    return k


@pytest.mark.parametrize("m", RESTATEMENTS, ids=IDS)
def test_same_padding_stride2_even_extent_pads_after_only(m):
    """tf 'SAME', W = 4, k = 3, s = 2: out = 2, total pad = (2-1)*2 + 3 - 4 = 1 -> 0 before, 1 after.
    Windows [1,2,3] and [3,4,0]: 6 and 7.  (Symmetric padding would give 3 and 9.)"""
    got = m.conv2d_same(_row([1, 2, 3, 4]), _k3([1, 1, 1]), stride=2)
    assert got.shape == (1, 1, 2, 1)
    assert got.reshape(-1).tolist() == [6.0, 7.0]


@pytest.mark.parametrize("m", RESTATEMENTS, ids=IDS)
def test_same_padding_stride2_odd_extent_pads_one_each_side(m):
    """W = 5, k = 3, s = 2: out = 3, total = (3-1)*2 + 3 - 5 = 2 -> 1 before, 1 after.
    Windows [0,1,2], [2,3,4], [4,5,0]: 3, 9, 9."""
    got = m.conv2d_same(_row([1, 2, 3, 4, 5]), _k3([1, 1, 1]), stride=2)
    assert got.reshape(-1).tolist() == [3.0, 9.0, 9.0]


@pytest.mark.parametrize("m", RESTATEMENTS, ids=IDS)
def test_same_padding_stride1_and_kernel_orientation(m):
    """Stride 1, k = 3: one zero each side.  An asymmetric kernel (1, 10, 100) fixes the
    orientation: TF correlates (no flip): out[x] = 1*in[x-1] + 10*in[x] + 100*in[x+1].
    in = [1,2,3]: [0+10+200, 1+20+300, 2+30+0] = [210, 321, 32]."""
    got = m.conv2d_same(_row([1, 2, 3]), _k3([1, 10, 100]), stride=1)
    assert got.reshape(-1).tolist() == [210.0, 321.0, 32.0]


@pytest.mark.parametrize("m", RESTATEMENTS, ids=IDS)
def test_same_padding_2d_even_extent_rows_and_columns(m):
    """4 x 4 image v[y][x] = 10 y + x, 3 x 3 ones, stride 2 -> 2 x 2, pad bottom / right only:
       out[0][0] = sum y,x in 0..2 = 3 * (0 + 10 + 20) + 3 * (0 + 1 + 2) = 99
       out[0][1] = x in {2, 3}   : 2 * 30 + 3 * (2 + 3)                  = 75
       out[1][0] = y in {2, 3}   : 3 * (20 + 30) + 2 * 3                 = 156
       out[1][1] = y, x in {2, 3}: 2 * 50 + 2 * 5                        = 110"""
    v = (10 * np.arange(4)[:, None] + np.arange(4)[None, :]).astype(np.float32).reshape(1, 4, 4, 1)
    got = m.conv2d_same(v, np.ones((3, 3, 1, 1), np.float32), stride=2)
    assert got.reshape(2, 2).tolist() == [[99.0, 75.0], [156.0, 110.0]]


@pytest.mark.parametrize("m", RESTATEMENTS, ids=IDS)
def test_hwio_weight_layout(m):
    """Weights are [kh, kw, Cin, Cout]: a 1 x 1 conv with w[0,0,ci,co] = 10 ci + co on the pixel
    (1, 2) gives out[co] = 1 * (0 + co) + 2 * (10 + co) = 20 + 3 co."""
    w = (10 * np.arange(2)[:, None] + np.arange(3)[None, :]).astype(np.float32).reshape(1, 1, 2, 3)
    x = np.asarray([1, 2], np.float32).reshape(1, 1, 1, 2)
    assert m.conv2d_same(x, w).reshape(-1).tolist() == [20.0, 23.0, 26.0]


@pytest.mark.parametrize("m", RESTATEMENTS, ids=IDS)
def test_flatten_is_nhwc_channel_fastest_and_unroll_grid_channel_map(m):
    """A whole coefficient network (models.py:62-142) whose weights make two conventions visible:
      * global fc1 reads ONE element of the flattened [B, 2, 2, C] tensor, flat index 5 * ... --
        see below -- so the fc input order must be (y * W + x) * C + c (tf.reshape of NHWC);
      * the 1x1 prediction conv has zero weights and bias[ch] = ch, so after `unroll_grid`
        coeffs[b, y, x, z, i, j] must equal (j * n_out + i) * gd + z."""
    p = dict(M.DEFAULT_PARAMS, net_input_size=16, spatial_bin=8, luma_bins=2)
    wts = M.make_weights(p, seed=0)
    gd, n_out, n_in = 2, 3, 4
    pre = "inference/coefficients"
    wts[pre + "/prediction/conv1/weights"] = np.zeros_like(wts[pre + "/prediction/conv1/weights"])
    wts[pre + "/prediction/conv1/biases"] = np.arange(gd * n_out * n_in, dtype=np.float32)
    low = np.random.RandomState(0).rand(1, 16, 16, 3).astype(np.float32)
    c = m.coefficients(low, wts, p)
    assert c.shape == (1, 8, 8, gd, n_out, n_in)
    z, i, j = np.meshgrid(np.arange(gd), np.arange(n_out), np.arange(n_in), indexing="ij")
    assert np.array_equal(c[0, 3, 5], ((j * n_out + i) * gd + z).astype(np.float32))

    # flatten order through the real graph: global conv2 output is [1, 2, 2, 16]; make fc1 the
    # selector of flat element 37 = (y=1, x=0, c=5) and the following layers pass it on
    sel = np.zeros_like(wts[pre + "/global/fc1/weights"])
    sel[37, 0] = 1.0
    wts[pre + "/global/fc1/weights"] = sel
    wts[pre + "/global/fc1/biases"] = np.zeros_like(wts[pre + "/global/fc1/biases"])
    wts2 = dict(wts)
    # probe: the activation the graph feeds to fc1
    x = low
    for k in range(1):
        x = m.conv(x, wts, f"{pre}/splat/conv{k + 1}", stride=2)
    g = x
    for k in range(2):
        g = m.conv(g, wts, f"{pre}/global/conv{k + 1}", stride=2)
    assert g.shape == (1, 2, 2, 16)
    want = max(float(g[0, 1, 0, 5]), 0.0)
    got = m.fc(g.reshape(1, -1), wts2, pre + "/global/fc1")
    assert got[0, 0] == np.float32(want) and not got[0, 1:].any()


@pytest.mark.parametrize("m", RESTATEMENTS, ids=IDS)
def test_resize_bilinear_align_corners_known_answers(m):
    """align_corners=True: src = dst * (in - 1) / (out - 1).  [0, 10, 20, 30] -> 7 samples at
    0, .5, 1, ... : 0, 5, 10, 15, 20, 25, 30;  -> 2 samples: the two corners; 5 -> 2 likewise
    (the pyramid's `sz / 2` levels, models.py:249-258)."""
    x = np.asarray([0, 10, 20, 30], np.float32).reshape(1, 1, 4, 1)
    assert m.resize_bilinear_ac(x, 1, 7).reshape(-1).tolist() == [0, 5, 10, 15, 20, 25, 30]
    assert m.resize_bilinear_ac(x, 1, 2).reshape(-1).tolist() == [0, 30]
    y = np.asarray([1, 2, 4, 8, 16], np.float32).reshape(1, 5, 1, 1)
    assert m.resize_bilinear_ac(y, 2, 1).reshape(-1).tolist() == [1, 16]
    # 2-D: the centre of a 2 x 2 image is the mean of its corners
    q = np.asarray([[0, 2], [4, 10]], np.float32).reshape(1, 2, 2, 1)
    assert m.resize_bilinear_ac(q, 3, 3)[0, 1, 1, 0] == 4.0


@pytest.mark.parametrize("m", RESTATEMENTS, ids=IDS)
def test_curves_guide_known_answer(m):
    """models.py:145-190 on one pixel, by hand.  rgb = (0.2, 0.5, 0.9), ccm = I, bias 0, every
    channel: shifts (0, 0.25, 0.5, 0.75), slopes (1, 2, -1, 0.5) [padded with zero slopes]:
       u(0.2) = 0.2;  u(0.5) = 0.5 + 2 * 0.25 = 1.0;  u(0.9) = 0.9 + 2 * 0.65 - 0.4 + 0.5 * 0.15 = 1.875
    mix = (0.5, 0.25, 0.125), bias 0.01: 0.1 + 0.25 + 0.234375 + 0.01 = 0.594375."""
    shifts = np.zeros((1, 1, 3, 16), np.float32)
    slopes = np.zeros((1, 1, 1, 3, 16), np.float32)
    shifts[..., :4] = [0, 0.25, 0.5, 0.75]
    shifts[..., 4:] = 2.0
    slopes[..., :4] = [1, 2, -1, 0.5]
    wts = {"inference/guide/ccm": np.identity(3, dtype=np.float32),
           "inference/guide/ccm_bias": np.zeros(3, np.float32),
           "inference/guide/shifts": shifts, "inference/guide/slopes": slopes,
           "inference/guide/channel_mixing/weights": np.asarray([0.5, 0.25, 0.125], np.float32).reshape(1, 1, 3, 1),
           "inference/guide/channel_mixing/biases": np.asarray([0.01], np.float32)}
    px = np.asarray([0.2, 0.5, 0.9], np.float32).reshape(1, 1, 1, 3)
    assert abs(float(m.guide_curves(px, wts)[0, 0, 0]) - 0.594375) < 2e-7
    # clip_by_value: a large mix pushes it to 1, a negative bias to 0
    wts["inference/guide/channel_mixing/biases"] = np.asarray([5.0], np.float32)
    assert m.guide_curves(px, wts)[0, 0, 0] == 1.0
    wts["inference/guide/channel_mixing/biases"] = np.asarray([-5.0], np.float32)
    assert m.guide_curves(px, wts)[0, 0, 0] == 0.0


@pytest.mark.parametrize("m", RESTATEMENTS, ids=IDS)
def test_batch_norm_inference_has_no_gamma(m):
    """layers.py:47-54 (center=True, scale=False), eps = 1e-3: x = 3, mean 1, var 0.999, beta 0.5
    -> (3 - 1) / sqrt(1.0) + 0.5 = 2.5; then ReLU."""
    wts = {"s/weights": np.ones((1, 1, 1, 1), np.float32),
           "s/BatchNorm/beta": np.asarray([0.5], np.float32),
           "s/BatchNorm/moving_mean": np.asarray([1.0], np.float32),
           "s/BatchNorm/moving_variance": np.asarray([0.999], np.float32)}
    x = np.asarray([3.0], np.float32).reshape(1, 1, 1, 1)
    assert abs(float(m.conv(x, wts, "s", batch_norm=True)[0, 0, 0, 0]) - 2.5) < 1e-6
    x = np.asarray([-3.0], np.float32).reshape(1, 1, 1, 1)
    assert m.conv(x, wts, "s", batch_norm=True)[0, 0, 0, 0] == 0.0          # relu((-4) + 0.5)


# ---- the two restatements agree on every graph ----------------------------------------------------
PARAM_SETS = {
    "default_small": dict(M.DEFAULT_PARAMS, net_input_size=64, spatial_bin=16),
    "bn_small": dict(M.DEFAULT_PARAMS, batch_norm=True, net_input_size=64, spatial_bin=8, luma_bins=4),
    "cm2": dict(M.DEFAULT_PARAMS, channel_multiplier=2, net_input_size=64, spatial_bin=16),
    "nn_guide": dict(M.DEFAULT_PARAMS, model_name="HDRNetPointwiseNNGuide", batch_norm=True,
                     net_input_size=64, spatial_bin=16),
    "pyramid": dict(M.DEFAULT_PARAMS, model_name="HDRNetGaussianPyrNN", net_input_size=64, spatial_bin=16),
}


def _rel(a, b):
    return float(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)).max() / max(np.abs(b).max(), 1e-30))


@pytest.mark.parametrize("name", list(PARAM_SETS))
def test_numpy_and_torch_restatements_agree(name):
    p = PARAM_SETS[name]
    wts = M.make_weights(p, seed=5)
    rng = np.random.RandomState(6)
    S = p["net_input_size"]
    low = rng.rand(2, S, S, 3).astype(np.float32)
    full = rng.rand(2, 37, 52, 3).astype(np.float32)
    sa = oracle.port().bilateral_slice_apply
    if name == "pyramid":
        a, ca, ga = M.gaussian_pyr_inference(low, full, wts, p, sa)
        b, cb, gb = T.gaussian_pyr_inference(low, full, wts, p, sa)
        for x, y in zip(ga, gb):
            assert np.abs(x - y).max() <= 1e-6      # levels 1, 2 sit behind a bilinear resize
    else:
        a, ca, ga = M.inference(low, full, wts, p, sa)
        b, cb, gb = T.inference(low, full, wts, p, sa)
        assert np.abs(ga - gb).max() <= 2e-7
    assert ca.shape == cb.shape and _rel(ca, cb) <= 1e-6, _rel(ca, cb)
    assert _rel(a, b) <= 1e-5      # through the slice: a guide difference of 1e-7 moves the depth coordinate
    x = rng.rand(1, 9, 7, 3).astype(np.float32)
    for oh, ow in ((4, 3), (18, 14), (1, 1), (9, 7)):
        assert np.abs(M.resize_bilinear_ac(x, oh, ow) - T.resize_bilinear_ac(x, oh, ow)).max() <= 2e-7


# ---- the CUDA layers against the same hand-computed vectors ------------------------------------------
def _cuda(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.gpu
def test_cuda_layers_reproduce_the_hand_computed_vectors():
    from hdrnet_b200 import models
    for vals, want in (([1, 2, 3, 4], [6, 7]), ([1, 2, 3, 4, 5], [3, 9, 9])):
        got = models._conv(_cuda(_row(vals)), (_cuda(_k3([1, 1, 1])), None), stride=2, relu=False)
        assert got.cpu().numpy().reshape(-1).tolist() == want
    got = models._conv(_cuda(_row([1, 2, 3])), (_cuda(_k3([1, 10, 100])), None), stride=1, relu=False)
    assert got.cpu().numpy().reshape(-1).tolist() == [210.0, 321.0, 32.0]
    v = (10 * np.arange(4)[:, None] + np.arange(4)[None, :]).astype(np.float32).reshape(1, 4, 4, 1)
    got = models._conv(_cuda(v), (_cuda(np.ones((3, 3, 1, 1), np.float32)), None), stride=2, relu=False)
    assert got.cpu().numpy().reshape(2, 2).tolist() == [[99.0, 75.0], [156.0, 110.0]]
    w = (10 * np.arange(2)[:, None] + np.arange(3)[None, :]).astype(np.float32).reshape(1, 1, 2, 3)
    x = np.asarray([1, 2], np.float32).reshape(1, 1, 1, 2)
    assert models._conv(_cuda(x), (_cuda(w), None), relu=False).cpu().numpy().reshape(-1).tolist() == [20.0, 23.0, 26.0]
    x = np.asarray([0, 10, 20, 30], np.float32).reshape(1, 1, 4, 1)
    assert models._resize(_cuda(np.tile(x, (1, 1, 1, 3))), 1, 7).cpu().numpy()[0, 0, :, 0].tolist() == [0, 5, 10, 15, 20, 25, 30]
    # unroll_grid through the fused fusion + prediction kernel
    p = dict(M.DEFAULT_PARAMS, net_input_size=16, spatial_bin=8, luma_bins=2)
    wts = M.make_weights(p, seed=0)
    pre = "inference/coefficients"
    wts[pre + "/prediction/conv1/weights"] = np.zeros_like(wts[pre + "/prediction/conv1/weights"])
    wts[pre + "/prediction/conv1/biases"] = np.arange(2 * 3 * 4, dtype=np.float32)
    low = np.random.RandomState(0).rand(1, 16, 16, 3).astype(np.float32)
    c = models.HDRNetCurves._coefficients(_cuda(low), dict(p, weights=wts)).cpu().numpy()
    z, i, j = np.meshgrid(np.arange(2), np.arange(3), np.arange(4), indexing="ij")
    assert np.array_equal(c[0, 3, 5], ((j * 3 + i) * 2 + z).astype(np.float32))
