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
