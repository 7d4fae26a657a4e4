"""GPU parity of the LayerNorm kernels (csrc/layernorm.cu) against the fp64 oracle and torch.nn.LayerNorm."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "spgemm-new_b200"))
sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("d", [256, 128, 64, 36])
def test_layernorm_against_oracle(d):
    from maxk_b200 import ops
    from oracle import oracle as O
    rng = np.random.default_rng(d)
    for n in (1, 9, 1000, 70001):
        x = (rng.standard_normal((n, d)) * 3 + 1.5).astype(np.float32)
        gamma = rng.standard_normal(d).astype(np.float32)
        beta = rng.standard_normal(d).astype(np.float32)
        g = rng.standard_normal((n, d)).astype(np.float32)
        cx, cg, cb, cgr = (torch.from_numpy(a).cuda() for a in (x, gamma, beta, g))
        y, mean, rstd = ops.layernorm_forward(cx, cg, cb)
        wy, wm, wr = O.layernorm(x, gamma, beta)
        np.testing.assert_allclose(y.cpu().numpy(), wy, rtol=2e-5, atol=2e-5)
        np.testing.assert_allclose(mean.cpu().numpy(), wm, rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(rstd.cpu().numpy(), wr, rtol=1e-5, atol=1e-6)
        dx, dga, dbe = ops.layernorm_backward(cgr, cx, mean, rstd, cg)
        wdx, wdg, wdb = O.layernorm_grad(g, x, gamma)
        np.testing.assert_allclose(dx.cpu().numpy(), wdx, rtol=1e-4, atol=1e-4)
        np.testing.assert_allclose(dga.cpu().numpy(), wdg, rtol=1e-4, atol=2e-4 * max(1, n ** 0.5))
        np.testing.assert_allclose(dbe.cpu().numpy(), wdb, rtol=1e-4, atol=2e-4 * max(1, n ** 0.5))


def test_layer_norm_module_path_matches_torch():
    from maxk_b200.autograd import layer_norm
    g = torch.Generator("cuda").manual_seed(5)
    norm = torch.nn.LayerNorm(256).cuda()
    with torch.no_grad():
        norm.weight.copy_(torch.randn(256, device="cuda", generator=g))
        norm.bias.copy_(torch.randn(256, device="cuda", generator=g))
    x = torch.randn(50_000, 256, device="cuda", generator=g, requires_grad=True)
    gy = torch.randn(50_000, 256, device="cuda", generator=g)
    y = layer_norm(norm, x)
    y.backward(gy)
    got = (y.detach(), x.grad.clone(), norm.weight.grad.clone(), norm.bias.grad.clone())
    x.grad = None
    norm.zero_grad()
    y2 = norm(x)
    y2.backward(gy)
    for a, b, tol in zip(got, (y2.detach(), x.grad, norm.weight.grad, norm.bias.grad), (1e-5, 1e-5, 2e-2, 2e-2)):
        assert torch.allclose(a, b, rtol=1e-4, atol=tol), float((a - b).abs().max())
    assert layer_norm(torch.nn.Identity(), x) is x


@pytest.mark.parametrize("d,k", [(256, 32), (256, 8), (128, 16), (64, 64)])
def test_layernorm_topk_fused(d, k):
    """TopK_k(LayerNorm(x)) in one pass == the TopK kernel on the LayerNorm kernel's output, bit for bit; the backward on
    the k-sparse gradient == the dense LayerNorm backward on the scattered twin; the autograd node end to end."""
    from maxk_b200 import ops
    from maxk_b200.autograd import LayerNormTopKCBSR
    g = torch.Generator("cuda").manual_seed(d + k)
    for n in (1, 1000, 50_001):
        x = torch.randn(n, d, device="cuda", generator=g) * 2 - 0.5
        gamma = torch.randn(d, device="cuda", generator=g)
        beta = torch.randn(d, device="cuda", generator=g)
        y, mean, rstd = ops.layernorm_forward(x, gamma, beta)
        want_v, want_i = ops.topk_cbsr(y, k)
        v, i, m2, r2, y2 = ops.layernorm_topk_cbsr(x, gamma, beta, 1e-5, k, want_dense=True)
        assert torch.equal(y2.view(torch.int32), y.view(torch.int32))
        assert torch.equal(i, want_i) and torch.equal(v.view(torch.int32), want_v.view(torch.int32))
        assert torch.equal(m2, mean) and torch.equal(r2, rstd)
        g_data = torch.randn(n, k, device="cuda", generator=g)
        dx, dga, dbe = ops.layernorm_backward_cbsr(g_data, i, x, mean, rstd, gamma)
        dense_g = ops.cbsr_scatter(g_data, i, d)
        wdx, wdg, wdb = ops.layernorm_backward(dense_g, x, mean, rstd, gamma)
        assert torch.allclose(dx, wdx, rtol=1e-5, atol=1e-5)
        assert torch.allclose(dga, wdg, rtol=1e-4, atol=1e-3 * max(1, n ** 0.5) / 30)
        assert torch.allclose(dbe, wdb, rtol=1e-4, atol=1e-3 * max(1, n ** 0.5) / 30)
    norm = torch.nn.LayerNorm(d).cuda()
    x = torch.randn(5000, d, device="cuda", generator=g, requires_grad=True)
    sp_data, sp_index = LayerNormTopKCBSR.apply(x, norm.weight, norm.bias, norm.eps, k)
    w = torch.randn(5000, k, device="cuda", generator=g)
    (sp_data * w).sum().backward()
    got = (x.grad.clone(), norm.weight.grad.clone(), norm.bias.grad.clone())
    x.grad = None
    norm.zero_grad()
    yy = norm(x)
    (yy.gather(1, sp_index.long()) * w).sum().backward()
    for a, b, tol in zip(got, (x.grad, norm.weight.grad, norm.bias.grad), (1e-4, 5e-3, 5e-3)):
        assert torch.allclose(a, b, rtol=1e-4, atol=tol), float((a - b).abs().max())
