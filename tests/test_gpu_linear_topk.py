"""GPU parity of the fused Linear + TopK kernel (csrc/linear_topk.cu).

Two claims, tested apart:
  * the Linear: the kernel's own fp32 result (`linear_out`) against an fp64 product of the same operands, at the
    accuracy of an fp32 GEMM (the kernel multiplies on TF32 tensor cores with a two-term operand split);
  * the selection: bit-exact against the oracle's TopK applied to the kernel's own `linear_out` -- ties, rows past
    the last full tile, all k -- and, on tie-free rows, the same SET as torch.topk on torch's own fp32 Linear up to
    elements within the GEMMs' rounding distance of the threshold.
"""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "spgemm-new_b200"))
sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from maxk_b200 import ops as o
    return o


def _oracle():
    from oracle import oracle as O
    return O


@pytest.mark.parametrize("n,in_dim", [(1, 256), (127, 256), (128, 64), (129, 256), (1000, 100), (5000, 256),
                                       (40000, 32), (3001, 516)])
def test_linear_matches_fp64_product(ops, n, in_dim):
    g = torch.Generator("cuda").manual_seed(n + in_dim)
    x = torch.randn(n, in_dim, device="cuda", generator=g)
    w = torch.randn(256, in_dim, device="cuda", generator=g) / in_dim ** 0.5
    b = torch.randn(256, device="cuda", generator=g)
    for bias in (b, None):
        lin = ops.linear_topk_cbsr(x, w, bias, 32, want_linear=True, want_cbsr=False)
        want = x.double() @ w.double().t() + (0 if bias is None else bias.double())
        scale = (x.double().abs() @ w.double().abs().t()) + 1.0          # sum of |products|: the error's yardstick
        err = ((lin.double() - want).abs() / scale).max().item()
        ref = ((torch.nn.functional.linear(x, w, bias).double() - want).abs() / scale).max().item()
        assert err < 4e-6, (n, in_dim, err, ref)                         # fp32 GEMM class (torch's own: ~1e-7)


@pytest.mark.parametrize("k", [1, 8, 16, 32, 64, 128, 200, 256])
def test_selection_is_the_oracles_on_the_kernels_own_product(ops, k):
    O = _oracle()
    g = torch.Generator("cuda").manual_seed(k)
    for n in (1, 33, 128, 300, 1025):
        x = torch.randn(n, 128, device="cuda", generator=g)
        w = torch.randn(256, 128, device="cuda", generator=g)
        b = torch.randn(256, device="cuda", generator=g)
        if n > 32:
            x[5] = 0.0                     # bias only: still tie-free
            x[7] = float("nan")            # a NaN row
            w[10] = w[11]                  # two equal columns everywhere -> ties whenever they straddle the threshold
            b[10] = b[11]
        v, i, lin = ops.linear_topk_cbsr(x, w, b, k, want_linear=True)
        want_v, want_i = O.topk_cbsr(lin.cpu().numpy(), k)
        np.testing.assert_array_equal(i.cpu().numpy(), want_i, err_msg=f"n={n} k={k}")
        np.testing.assert_array_equal(v.cpu().numpy().view(np.uint32), want_v.view(np.uint32))
        v2, i2 = ops.linear_topk_cbsr(x, w, b, k)                       # without the dense output
        assert torch.equal(i2, i) and torch.equal(v2.view(torch.int32), v.view(torch.int32))


def test_all_ties_and_zero_weight(ops):
    O = _oracle()
    x = torch.randn(300, 64, device="cuda")
    w = torch.zeros(256, 64, device="cuda")
    for k in (8, 32):
        v, i = ops.linear_topk_cbsr(x, w, None, k)                       # every row: 256 zeros -> first k columns
        assert torch.equal(i.long(), torch.arange(k, device="cuda").expand(300, k))
        assert torch.count_nonzero(v) == 0
    b = torch.arange(256, device="cuda", dtype=torch.float32)
    v, i = ops.linear_topk_cbsr(x, w, b, 8)
    assert torch.equal(i.long(), torch.arange(248, 256, device="cuda").expand(300, 8))


def test_against_torch_linear_then_topk(ops):
    """Same set as torch's Linear + topk except where the k-th / (k+1)-th values are closer than the two GEMMs'
    rounding; products-like width (in = 256) and 200 k rows, several tiles per SM."""
    g = torch.Generator("cuda").manual_seed(3)
    n = 200_000
    x = torch.randn(n, 256, device="cuda", generator=g)
    w = torch.randn(256, 256, device="cuda", generator=g) / 16
    b = torch.randn(256, device="cuda", generator=g) * 0.1
    y = torch.nn.functional.linear(x, w, b)
    for k in (16, 32):
        v, i = ops.linear_topk_cbsr(x, w, b, k)
        tv, ti = torch.topk(y, k, dim=1)
        order = ti.argsort(dim=1)
        same = (i.long() == ti.gather(1, order)).all(dim=1)
        # rows that differ: the gap at the threshold must be at rounding level
        gap = tv[:, -1] - torch.topk(y, k + 1, dim=1).values[:, -1]
        assert (gap[~same] < 1e-4).all(), float(gap[~same].max())
        assert same.float().mean() > 0.999
        assert torch.allclose(v[same], tv.gather(1, order)[same], rtol=1e-5, atol=1e-5)
        assert (i[:, 1:].int() > i[:, :-1].int()).all()                 # ascending, distinct


def test_autograd_node_uses_the_fused_forward(ops):
    from maxk_b200.autograd import LinearTopKCBSR
    g = torch.Generator("cuda").manual_seed(9)
    x = torch.randn(2000, 256, device="cuda", generator=g, requires_grad=True)
    lin = torch.nn.Linear(256, 256).cuda()
    sp_data, sp_index = LinearTopKCBSR.apply(x, lin.weight, lin.bias, 32)
    sp_data.sum().backward()
    # gradient of sum(top-k values): every selected column contributes its weight row
    dense = torch.zeros(2000, 256, device="cuda")
    dense.scatter_(1, sp_index.long(), 1.0)
    assert torch.allclose(x.grad, dense @ lin.weight.detach(), rtol=1e-4, atol=1e-4)
    assert torch.allclose(lin.bias.grad, dense.sum(0), rtol=1e-5, atol=1e-5)
    assert torch.allclose(lin.weight.grad, dense.t() @ x.detach(), rtol=1e-3, atol=1e-3)


@pytest.mark.parametrize("n,in_dim", [(1, 256), (130, 64), (1000, 100), (40001, 256), (300000, 256), (5000, 36)])
def test_linear_alone_and_weight_gradient(ops, n, in_dim):
    """maxk_linear_f32 (coalesced dense epilogue) and maxk_linear_wgrad against fp64 products of the same operands."""
    g = torch.Generator("cuda").manual_seed(7 * n + in_dim)
    x = torch.randn(n, in_dim, device="cuda", generator=g)
    w = torch.randn(256, in_dim, device="cuda", generator=g) / in_dim ** 0.5
    b = torch.randn(256, device="cuda", generator=g)
    y = ops.linear_f32(x, w, b)
    want = x.double() @ w.double().t() + b.double()
    scale = x.double().abs() @ w.double().abs().t() + 1.0
    assert ((y.double() - want).abs() / scale).max().item() < 4e-6
    res = torch.randn(n, 256, device="cuda", generator=g)
    y2 = ops.linear_f32(x, w, None, residual=res)                          # + residual in the epilogue
    want2 = x.double() @ w.double().t() + res.double()
    assert ((y2.double() - want2).abs() / scale).max().item() < 4e-6
    gy = torch.randn(n, 256, device="cuda", generator=g)
    dw = ops.linear_wgrad(gy, x)
    want = gy.double().t() @ x.double()
    scale = gy.double().abs().t() @ x.double().abs() + 1.0
    err = ((dw.double() - want).abs() / scale).max().item()
    ref = (((gy.t() @ x).double() - want).abs() / scale).max().item()
    assert err < 4e-6, (n, in_dim, err, ref)
    acc = dw.clone()
    ops.linear_wgrad(gy, x, out=acc, accumulate=True)                      # += semantics
    assert ((acc.double() - 2 * dw.double()).abs() / scale).max().item() < 1e-6   # the CTAs' partials add in any order


def test_tensor_core_linear_autograd(ops):
    from maxk_b200.autograd import TensorCoreLinear
    g = torch.Generator("cuda").manual_seed(11)
    x = torch.randn(3000, 256, device="cuda", generator=g, requires_grad=True)
    lin = torch.nn.Linear(256, 256).cuda()
    gy = torch.randn(3000, 256, device="cuda", generator=g)
    y = TensorCoreLinear.apply(x, lin.weight, lin.bias)
    y.backward(gy)
    got = (y.detach(), x.grad.clone(), lin.weight.grad.clone(), lin.bias.grad.clone())
    x.grad = None
    lin.zero_grad()
    y2 = lin(x)
    y2.backward(gy)
    for a, b_, tol in zip(got, (y2.detach(), x.grad, lin.weight.grad, lin.bias.grad), (1e-4, 1e-4, 2e-3, 1e-3)):
        assert torch.allclose(a, b_, rtol=1e-4, atol=tol), float((a - b_).abs().max())


def test_narrow_output_linear_backward(ops):
    """The classifier's backward on the tensor-core kernels (roles of the operands exchanged) against torch's."""
    from maxk_b200.autograd import output_linear
    g = torch.Generator("cuda").manual_seed(21)
    for classes in (47, 41, 112, 7):
        lin = torch.nn.Linear(256, classes).cuda()
        x = torch.randn(20_001, 256, device="cuda", generator=g, requires_grad=True)
        gy = torch.randn(20_001, classes, device="cuda", generator=g)
        y = output_linear(lin, x)
        y.backward(gy)
        got = (y.detach(), x.grad.clone(), lin.weight.grad.clone(), lin.bias.grad.clone())
        x.grad = None
        lin.zero_grad()
        y2 = lin(x)
        y2.backward(gy)
        for a, b_, tol in zip(got, (y2.detach(), x.grad, lin.weight.grad, lin.bias.grad), (1e-5, 1e-4, 5e-3, 1e-3)):
            assert a.shape == b_.shape and torch.allclose(a, b_, rtol=1e-4, atol=tol), float((a - b_).abs().max())
