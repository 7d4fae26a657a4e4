"""CBSR-native SAGE / GCN / GIN layers against a plain-torch dense restatement of
the reference models (utils/models.py:300-392, utils/models_old.py:26-67): MaxK mask via
torch.topk, aggregation via torch.sparse.mm, same weights, eval mode (dropout off)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from helpers import Problem

pytestmark = pytest.mark.gpu


def _graph(n=300, avg=12.0, seed=3):
    p = Problem(n=n, avg=avg, k=8, seed=seed, empty_frac=0.05)
    g = p.cuda()
    from maxk_b200.graph import Graph, edge_values
    G = Graph(g["indptr"], g["indices"])
    dense = {}
    for kind in ("sum", "mean", "gcn"):
        vals = edge_values(G.indptr, G.indices, kind, n)
        dense[kind] = torch.sparse_csr_tensor(G.indptr.long(), G.indices.long(), vals.double(),
                                              size=(n, n)).to_dense()
    return G, dense, n


def _maxk_dense(x, k):
    _, idx = x.topk(k, dim=1)
    return x * torch.zeros_like(x).scatter_(1, idx, 1.0)


def _ref_forward(model, kind, A, x, k):
    """fp64 dense restatement with the model's own parameters."""
    d = lambda t: t.double()  # noqa: E731
    lin = lambda l, v: F.linear(v, d(l.weight), None if l.bias is None else d(l.bias))  # noqa: E731
    name = type(model).__name__
    x = lin(model.lin_in, d(x))
    if name != "SAGE":
        x = x.relu()
    for i in range(model.num_layers):
        if name == "SAGE":
            h = _maxk_dense(x, k)
            x = lin(model.fc_self_layers[i], h) + lin(model.fc_neigh_layers[i], A["mean"] @ h)
        elif name == "GCN":
            h = _maxk_dense(lin(model.linlayers[i], x), k)
            x = A["gcn"] @ h + d(model.conv_bias[i])
        else:
            h = _maxk_dense(lin(model.linlayers[i], x), k)
            x = (1.0 + d(model.eps[i])) * h + A["sum"] @ h
        if model.norm:
            ln = model.normlayers[i]
            x = F.layer_norm(x, (x.shape[1],), d(ln.weight), d(ln.bias), ln.eps)
    return lin(model.lin_out, x)


@pytest.mark.parametrize("native", ["0", "1"])
@pytest.mark.parametrize("name", ["SAGE", "GIN", "GCN"])
def test_models_with_and_without_the_layers_on_the_cbsr_operand(name, native, monkeypatch):
    """MAXK_CBSR_NATIVE forces the dense-twin layers (0) or the CBSR-operand layers (1); the default (k <= 32: on)
    is what test_models_match_dense_restatement runs."""
    monkeypatch.setenv("MAXK_CBSR_NATIVE", native)
    test_models_match_dense_restatement(name, 2, True)


@pytest.mark.parametrize("name", ["SAGE", "GCN", "GIN"])
@pytest.mark.parametrize("layers,norm", [(1, False), (2, True)])
def test_models_match_dense_restatement(name, layers, norm):
    from utils import models as M
    torch.manual_seed(0)
    G, A, n = _graph()
    k, hid = 8, 64
    model = getattr(M, name)(20, hid, layers, 7, maxk=k, feat_drop=0.5, norm=norm).cuda().eval()
    feats = torch.randn(n, 20, device="cuda")
    target = torch.randn(n, 7, device="cuda")
    out = model(G, feats)
    loss = (out * target).sum()
    loss.backward()
    ref_model_params = [p.detach().clone().double().requires_grad_(True) for p in model.parameters()]
    # rebuild the reference through functional calls that read the same Parameter objects
    for p in model.parameters():
        p.grad_ours = p.grad.clone()
        p.grad = None
    ref = _ref_forward(model, name, A, feats, k)
    (ref * target.double()).sum().backward()
    np.testing.assert_allclose(out.detach().cpu().numpy(), ref.detach().cpu().numpy(), rtol=2e-4, atol=2e-4)
    for pname, p in model.named_parameters():
        np.testing.assert_allclose(p.grad_ours.cpu().numpy(), p.grad.cpu().numpy(), rtol=2e-3, atol=2e-3,
                                   err_msg=pname)
    del ref_model_params


def test_training_reduces_loss_and_dropout_is_on_cbsr_values():
    from utils import models as M
    torch.manual_seed(1)
    G, _, n = _graph(n=400, avg=15.0, seed=5)
    model = M.SAGE(16, 64, 2, 5, maxk=16, feat_drop=0.5, norm=True).cuda()
    feats = torch.randn(n, 16, device="cuda")
    labels = torch.randint(0, 5, (n,), device="cuda")
    opt = torch.optim.Adam(model.parameters(), lr=0.01)
    losses = []
    for _ in range(30):
        model.train()
        loss = F.cross_entropy(model(G, feats), labels)
        opt.zero_grad()
        loss.backward()
        opt.step()
        losses.append(float(loss))
    assert losses[-1] < 0.8 * losses[0]
    model.eval()
    a, b = model(G, feats), model(G, feats)
    # eval has no dropout; rows split over several warps' runs are combined with reductions
    # whose order is free (as with the reference's atomics), so equality is up to fp32 rounding
    assert torch.allclose(a, b, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("name", ["SAGE", "GCN", "GIN"])
def test_models_hidden256_on_the_own_dense_kernels(name, monkeypatch):
    """Width 256, k = 32: the layers as the epochs run them -- Linear + TopK in one tcgen05 kernel, tensor-core Linears
    and weight gradients, own LayerNorm (+ TopK), the classifier's backward on the tensor-core kernels -- against the
    fp64 dense restatement, and against the same model with every one of those switched back to the library.  A
    selection may differ on a row whose 32nd and 33rd values are closer than the two GEMMs' rounding, so the check
    is on the fraction of entries inside the tolerance."""
    from utils import models as M
    torch.manual_seed(2)
    G, A, n = _graph(n=700, avg=14.0, seed=9)
    k, hid = 32, 256
    model = getattr(M, name)(20, hid, 2, 7, maxk=k, feat_drop=0.5, norm=True).cuda().eval()
    feats = torch.randn(n, 20, device="cuda")
    target = torch.randn(n, 7, device="cuda")

    def run():
        for p in model.parameters():
            p.grad = None
        out = model(G, feats)
        (out * target).sum().backward()
        return out.detach().clone(), {pn: p.grad.clone() for pn, p in model.named_parameters()}

    out_own, grads_own = run()
    for var in ("MAXK_TC_LINEAR", "MAXK_LAYERNORM", "MAXK_LINEAR_TOPK_FUSED"):
        monkeypatch.setenv(var, "0")
    out_lib, grads_lib = run()
    for var in ("MAXK_TC_LINEAR", "MAXK_LAYERNORM", "MAXK_LINEAR_TOPK_FUSED"):
        monkeypatch.delenv(var)
    for p in model.parameters():
        p.grad = None
    ref = _ref_forward(model, name, A, feats, k)
    (ref * target.double()).sum().backward()

    def close_frac(a, b, tol):
        a, b = a.double(), b.double()
        return float(((a - b).abs() <= tol * (1.0 + b.abs())).double().mean())

    assert close_frac(out_own, ref.detach(), 5e-4) > 0.995
    assert close_frac(out_own, out_lib, 5e-4) > 0.995
    for pn, p in model.named_parameters():
        scale = float(p.grad.abs().max()) + 1e-6
        assert close_frac(grads_own[pn] / scale, p.grad / scale, 5e-3) > 0.99, pn
        assert close_frac(grads_own[pn] / scale, grads_lib[pn] / scale, 5e-3) > 0.99, pn
