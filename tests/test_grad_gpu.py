"""GPU tests of the VJP kernels against the compiled reference loops
(oracle/_ref: hdrnet/ops/bilateral_slice_apply.cc:84-259, bilateral_slice.cc:72-168) or their
bit-exact C restatement, plus the reference's own numeric-vs-analytic criterion through
torch.autograd (hdrnet/hdrnet_ops_test.py:173-180, :361-408)."""
import numpy as np
import pytest
import torch

import oracle
from hdrnet_b200 import hdrnet_ops
from util import assert_parity, rand_case

pytestmark = pytest.mark.gpu


def cuda(a, grad=False):
    t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
    return t.requires_grad_(grad)


# (B, H, W, gh, gw, gd, n_in, n_out, has_offset)
APPLY_CASES = [
    (3, 8, 5, 6, 3, 7, 3, 4, True),      # hdrnet_ops_test.py:185-195 (generic grid-VJP path: gc=16)
    (3, 30, 25, 16, 12, 8, 3, 3, True),  # default test extents: fast column kernel (gd<=8, gc<=12)
    (2, 40, 64, 4, 4, 8, 3, 3, True),    # several pixels per cell, mirror boundary on all sides
    (2, 33, 47, 5, 3, 4, 3, 4, False),   # no offset
    (1, 16, 16, 2, 2, 1, 1, 1, True),    # gd = 1: both depth borders at once
]


@pytest.mark.parametrize("case", APPLY_CASES, ids=str)
def test_slice_apply_vjps_match_reference(case):
    B, H, W, gh, gw, gd, n_in, n_out, ho = case
    grid, guide, inp = rand_case(5, B, H, W, gh, gw, gd, n_in, n_out, ho, signed=True)
    rng = np.random.RandomState(6)
    ct = rng.randn(B, H, W, n_out).astype(np.float32)
    want = oracle.port().bilateral_slice_apply_grad(grid, guide, inp, ct, ho)
    g, u, i = cuda(grid, True), cuda(guide, True), cuda(inp, True)
    out = hdrnet_ops.bilateral_slice_apply(g, u, i, ho)
    out.backward(cuda(ct))
    for got, ref, name in zip((g.grad, u.grad, i.grad), want, ("grid", "guide", "input")):
        if name == "guide" and gd == 1:
            # Degenerate: both depth corners clamp to cell 0, the two derivative terms cancel
            # and the reference's own result is float32 rounding noise (|vjp| ~ 1e-7 of the
            # terms).  Bound it against the scale of the TERMS instead of the cancelled sum.
            scale = float(np.abs(grid).max() * np.abs(ct).max() * np.abs(inp).max() * gd * 4)
            assert np.abs(got.cpu().numpy() - ref).max() <= 1e-6 * scale
            continue
        assert_parity(got.cpu().numpy(), ref, rtol=2e-5, what=f"{case} {name} VJP")


@pytest.mark.parametrize("case", [(3, 30, 25, 16, 12, 8, 12), (2, 21, 36, 5, 4, 6, 2), (1, 9, 7, 3, 3, 9, 5)],
                         ids=str)
def test_slice_vjps_match_reference(case):
    B, H, W, gh, gw, gd, gc = case
    rng = np.random.RandomState(7)
    grid = rng.randn(B, gh, gw, gd, gc).astype(np.float32)
    guide = rng.rand(B, H, W).astype(np.float32)
    ct = rng.randn(B, H, W, gc).astype(np.float32)
    want = oracle.port().bilateral_slice_grad(grid, guide, ct)
    g, u = cuda(grid, True), cuda(guide, True)
    hdrnet_ops.bilateral_slice(g, u).backward(cuda(ct))
    assert_parity(g.grad.cpu().numpy(), want[0], rtol=2e-5, what="grid VJP")
    assert_parity(u.grad.cpu().numpy(), want[1], rtol=2e-5, what="guide VJP")


@pytest.mark.parametrize("name", ["vjp_0", "vjp_2"])
def test_slice_vjps_match_reference_jax_golden(name):
    """tests/golden/vjp_*.npz: the reference's own JAX VJP functions (jax/bilateral_slice.py:26-108,
    :257-295; tests/golden/make_golden.py) at the reference's default test extents and on a coarse
    grid -- the same bar as against the C++ loops above."""
    from util import load_golden
    z = load_golden(name)
    g, u = cuda(z["grid"], True), cuda(z["guide"], True)
    hdrnet_ops.bilateral_slice(g, u).backward(cuda(z["codomain_tangent"]))
    # global bar only: the per-element statistic is asserted against the C++ loops above
    assert_parity(g.grad.cpu().numpy(), z["grid_vjp"], rtol=2e-5, what=f"{name} grid VJP", elem_rtol=None)
    assert_parity(u.grad.cpu().numpy(), z["guide_vjp"], rtol=2e-5, what=f"{name} guide VJP", elem_rtol=None)


def test_grad_shapes_follow_reference_contract():
    """hdrnet_ops_test.py:125-135, :304-315 (test_grad_shape)."""
    grid, guide, inp = rand_case(1, 3, 30, 25, 16, 12, 8)
    g, u, i = cuda(grid, True), cuda(guide, True), cuda(inp, True)
    hdrnet_ops.bilateral_slice_apply(g, u, i, True).sum().backward()
    assert g.grad.shape == g.shape and u.grad.shape == u.shape and i.grad.shape == i.shape


def test_analytic_vs_numeric_gradient_error():
    """The reference's criterion (compute_gradient_error <= 1e-2, hdrnet_ops_test.py:361-363):
    central differences of the CUDA forward vs the CUDA VJPs, for grid and input (the forward is
    piecewise linear in both, so float32 differences are accurate)."""
    grid, guide, inp = rand_case(3, 1, 12, 10, 3, 3, 4, 3, 3, True)
    guide = (0.1 + 0.8 * guide).astype(np.float32)
    rng = np.random.RandomState(4)
    ct = rng.rand(1, 12, 10, 3).astype(np.float32)
    g, u, i = cuda(grid, True), cuda(guide, True), cuda(inp, True)
    hdrnet_ops.bilateral_slice_apply(g, u, i, True).backward(cuda(ct))

    def loss(gg, ii):
        with torch.no_grad():
            o = hdrnet_ops.bilateral_slice_apply(cuda(gg), cuda(guide), cuda(ii), True)
        return float((o.double() * cuda(ct).double()).sum())

    eps = 1e-2
    for arr, grad, which in ((grid, g.grad, 0), (inp, i.grad, 1)):
        gflat = grad.cpu().numpy().reshape(-1)
        for k in rng.choice(arr.size, 10, replace=False):
            hi, lo = arr.copy(), arr.copy()
            hi.reshape(-1)[k] += eps
            lo.reshape(-1)[k] -= eps
            num = (loss(hi, inp) - loss(lo, inp)) / (2 * eps) if which == 0 else \
                  (loss(grid, hi) - loss(grid, lo)) / (2 * eps)
            assert abs(num - gflat[k]) <= 1e-2 * max(1.0, abs(num))


def test_backward_is_deterministic():
    grid, guide, inp = rand_case(9, 2, 64, 96, 8, 8, 8)
    ct = np.random.RandomState(1).rand(2, 64, 96, 3).astype(np.float32)
    grads = []
    for _ in range(2):
        g, u, i = cuda(grid, True), cuda(guide, True), cuda(inp, True)
        hdrnet_ops.bilateral_slice_apply(g, u, i, True).backward(cuda(ct))
        grads.append((g.grad.clone(), u.grad.clone(), i.grad.clone()))
    for a, b in zip(*grads):
        assert torch.equal(a, b)


def test_grid_vjp_is_zero_when_there_are_no_pixels():
    """B > 0 with H == 0 (or W == 0): no pixel contributes, the grid VJP is a tensor of
    zeros (the reference's kernels write 0 for every cell) -- not uninitialised memory."""
    for H, W in ((0, 8), (8, 0)):
        grid = torch.randn(2, 4, 4, 8, 12, device="cuda", requires_grad=True)
        guide = torch.rand(2, H, W, device="cuda", requires_grad=True)
        inp = torch.randn(2, H, W, 3, device="cuda", requires_grad=True)
        out = hdrnet_ops.bilateral_slice_apply(grid, guide, inp, True)
        assert out.shape == (2, H, W, 3)
        out.sum().backward()
        assert grid.grad is not None and not grid.grad.any()
        g2 = torch.randn(2, 4, 4, 8, 12, device="cuda", requires_grad=True)
        hdrnet_ops.bilateral_slice(g2, guide.detach()).sum().backward()
        assert not g2.grad.any()


@pytest.mark.parametrize("name", ["grid", "guide", "both"])
def test_sgd_convergence_bounds_of_the_reference_tests(name):
    """hdrnet/test/ops_test.py:189-322 (test_grid_optimize, test_guide_optimize, test_optimize_both):
    plain gradient descent on sum((target - slice(grid, guide))^2) through the CUDA forward and VJP
    kernels (torch.autograd) reaches the loss bounds the reference asserts.  tests/test_oracle.py runs
    the same cases through the reference's own loops."""
    from util import sgd_case
    c = sgd_case(name)
    grid = cuda(c["grid"], "grid" in c["trained"])
    v = cuda(c["guide"], "guide" in c["trained"])
    target = cuda(c["target"])
    params = [t for t in (grid, v) if t.requires_grad]

    def forward():
        return hdrnet_ops.bilateral_slice(grid, torch.sigmoid(v) if c["sigmoid"] else v)

    for _ in range(c["steps"]):
        loss = (target - forward()).square().sum()
        grads = torch.autograd.grad(loss, params)
        with torch.no_grad():
            for p, g in zip(params, grads):
                p -= c["lr"] * g
    with torch.no_grad():
        final = float((target - forward()).square().sum())
    assert final < c["bound"], f"{name}: final loss {final:.3e} >= {c['bound']:.1e}"
