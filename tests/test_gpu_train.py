"""The full-graph training loop of the package (maxk_b200/train.py; restates maxk_gnn_dgl.py:78-113 with the
hyper-parameters of scripts_train/*.sh) on scaled-down BASELINE shapes: one GPU, through the CUDA kernels."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape,model,k", [("products", "sage", 32), ("proteins", "gin", 64), ("flickr", "gcn", 16)])
def test_training_epochs_reduce_the_loss(shape, model, k):
    from maxk_b200 import _cabi
    from maxk_b200.train import Trainer
    scale = {"products": 0.004, "proteins": 0.05, "flickr": 0.1}[shape]
    tr = Trainer(shape, model=model, k=k, scale=scale, device="cuda")
    n0 = _cabi.launch_count()
    losses = [float(tr.epoch()) for _ in range(12)]
    assert _cabi.launch_count() - n0 >= 12 * 3 * 4, "every layer must go through TopK, SpGEMM and SSpMM kernels"
    assert all(math.isfinite(v) for v in losses), losses
    assert min(losses[-3:]) < losses[0], losses   # Adam on random labels still fits the training rows
    for p in tr.model.parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all()


def test_run_epochs_record():
    from maxk_b200.train import run_epochs
    rec = run_epochs("products", epochs=2, warmup=1, scale=0.002, device="cuda")
    assert rec["metric"] == "full_graph_epochs_per_s" and rec["n_gpus"] == 1
    assert rec["value"] > 0 and abs(rec["value"] * rec["ms_per_epoch"] - 1e3) < 1e-6
    assert rec["model"] == "sage" and rec["k"] == 32 and rec["layers"] == 3 and rec["dtype"] == "f32"


def test_epoch_matches_a_dense_restatement():
    """One training step (loss and every weight gradient) against plain torch on the dense masked twin: same
    weights, dropout off."""
    import torch.nn.functional as F

    from maxk_b200.graph import edge_values
    from maxk_b200.train import Trainer
    tr = Trainer("flickr", model="sage", k=8, scale=0.005, device="cuda", hidden=64)
    tr.model.p = 0.0
    G = tr.G
    n = G.n_rows
    A = torch.sparse_csr_tensor(G.indptr.long(), G.indices.long(), edge_values(G.indptr, G.indices, "mean", n).double(),
                                size=(n, n)).to_dense()
    m = tr.model
    loss = tr.loss_of(m(G, tr.feats))
    grads = torch.autograd.grad(loss, list(m.parameters()))
    d = lambda t: t.detach().double().requires_grad_(True)  # noqa: E731
    P = {name: d(p) for name, p in m.named_parameters()}
    lin = lambda pre, v: F.linear(v, P[pre + ".weight"], P.get(pre + ".bias"))  # noqa: E731
    x = lin("lin_in", tr.feats.double())
    for i in range(m.num_layers):
        _, idx = x.topk(8, dim=1)
        h = x * torch.zeros_like(x).scatter_(1, idx, 1.0)
        x = lin(f"fc_self_layers.{i}", h) + lin(f"fc_neigh_layers.{i}", A @ h)
        x = F.layer_norm(x, (x.shape[1],), P[f"normlayers.{i}.weight"], P[f"normlayers.{i}.bias"], 1e-5)
    logits = lin("lin_out", x)
    ref = F.cross_entropy(logits[tr.train_mask], tr.labels[tr.train_mask], reduction="sum") / tr.n_train.double()
    assert abs(float(loss) - float(ref)) <= 1e-4 * max(1.0, abs(float(ref)))
    ref_grads = torch.autograd.grad(ref, [P[name] for name, _ in m.named_parameters()])
    for (name, _), g, r in zip(m.named_parameters(), grads, ref_grads):
        scale = float(r.abs().max()) + 1e-12
        assert float((g.double() - r).abs().max()) <= 2e-3 * scale + 1e-7, name
