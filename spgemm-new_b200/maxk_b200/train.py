"""Full-graph MaxK-GNN training epochs on a synthetic BASELINE.json shape.

Restates the reference's training loop (maxk_gnn_dgl.py:78-113: model.train(), logits = model(g, x),
loss on the train mask, optimizer.zero_grad / backward / step, evaluation apart) without DGL / OGB, with
the hyper-parameters of scripts_train/*.sh (ogbn_products_maxk.sh:21-24: SAGE, 3 hidden layers, hidden 256,
k = 32, dropout 0.5, LayerNorm, Adam lr 0.003).  An EPOCH is forward + loss + backward + Adam step on the whole
graph; evaluation is excluded.

One GPU, or N GPUs with one process per GPU (torchrun): the adjacency is 1-D row-partitioned (nnz + row
cost), every layer exchanges the CBSR features and reduce-scatters the CBSR-shaped gradients
(DistGraph: our kernels over NVLink peer memory, or NCCL), and the weight gradients (~1 MB) are summed with
ONE flat all-reduce per step.

`run_epochs` is what `bench.py`'s `epoch` leg, `scripts/train_epoch.py` and tests/test_gpu_train.py call.
"""
from __future__ import annotations

import time
from typing import Dict, Optional

import torch
import torch.nn.functional as F

from . import graphs
from .graph import DistGraph, Graph

CONFIGS = {  # scripts_train/*.sh of the reference
    "products": dict(in_size=100, out_size=47, layers=3, lr=0.003, loss="ce", model="sage", k=32),
    "proteins": dict(in_size=8, out_size=112, layers=3, lr=0.01, loss="bce", model="gin", k=64),
    "reddit": dict(in_size=602, out_size=41, layers=4, lr=0.01, loss="ce", model="sage", k=32),
    "flickr": dict(in_size=500, out_size=7, layers=3, lr=0.001, loss="ce", model="sage", k=32),
}

# weight of one node row, in edges, when balancing the row partition: the dense Linear / LayerNorm work of
# an epoch is per row (measured on one B200, products shape: 78 ns per row vs 0.42 ns per edge, 3 layers)
DEFAULT_ROW_COST = 186.0


class Trainer:
    """Model + optimizer + synthetic features / labels of one rank's rows; `epoch()` is one training step
    on the full graph (maxk_gnn_dgl.py:78-96)."""

    def __init__(self, shape: str = "products", model: Optional[str] = None, k: Optional[int] = None,
                 hidden: int = 256, scale: float = 1.0, device=None, rank: int = 0, world: int = 1,
                 comm: str = "peer", group=None, row_cost: float = DEFAULT_ROW_COST, tf32: bool = False,
                 seed: int = 97, split: bool = True):
        from utils import models as M  # the reference-named module (utils/models.py) of this package
        import torch.distributed as dist
        self.dist = dist
        cfg = dict(CONFIGS[shape])
        self.cfg, self.shape, self.world, self.rank = cfg, shape, world, rank
        self.model_name = model or cfg["model"]
        self.k = int(k or cfg["k"])
        self.hidden = hidden
        self.tf32 = bool(tf32)
        dev = torch.device(device if device is not None else "cuda")
        assert dev.type == "cuda", "training runs on the GPU (no CPU fallback)"
        self.dev = dev
        g = graphs.make_graph(shape, seed=0, device=dev, values="ones", scale=scale)
        self.n, self.e = g["v_num"], g["e_num"]
        if world > 1:
            from .dist import partition_rows_by_nnz
            bounds = partition_rows_by_nnz(g["indptr"], world, align=16, row_cost=row_cost)
            if comm == "auto":
                # the fused kernel's TMA push hides a small CBSR block behind the own-column aggregation (Reddit shape:
                # 4.7 MB per rank at 8 GPUs); on large blocks NCCL's all-gather moves the bytes 2.2x faster (products
                # shape, 49 MB per rank: 547 vs ~250 GB/s; epoch 15.5 vs 16.1 ms at N = 8, 23.7 vs 24.7 ms at N = 4)
                block_bytes = max(bounds[r + 1] - bounds[r] for r in range(world)) * self.k * 5
                comm = "nccl" if block_bytes >= (16 << 20) else "peer"
            self.G = DistGraph(g["indptr"], g["indices"], bounds, rank, group=group, split=split, comm=comm)
            lo, hi = bounds[rank], bounds[rank + 1]
            self.comm = comm
        else:
            self.G = Graph(g["indptr"], g["indices"])
            lo, hi = 0, self.n
            self.comm = "none"
        del g
        torch.cuda.empty_cache()
        gen = torch.Generator(dev).manual_seed(seed + rank)  # --seed 97 (utils/config.py)
        self.feats = torch.randn(hi - lo, cfg["in_size"], device=dev, generator=gen)
        if cfg["loss"] == "ce":
            self.labels = torch.randint(0, cfg["out_size"], (hi - lo,), device=dev, generator=gen)
        else:
            self.labels = (torch.rand(hi - lo, cfg["out_size"], device=dev, generator=gen) < 0.1).float()
        self.train_mask = torch.rand(hi - lo, device=dev, generator=gen) < 0.1
        n_train = self.train_mask.sum().float()
        if world > 1:
            dist.all_reduce(n_train, group=group)
        self.n_train = n_train
        self.group = group
        torch.manual_seed(seed)  # identical initial weights on every rank
        cls = {"sage": M.SAGE, "gcn": M.GCN, "gin": M.GIN}[self.model_name]
        self.model = cls(cfg["in_size"], hidden, cfg["layers"], cfg["out_size"], maxk=self.k, feat_drop=0.5,
                         norm=True).to(dev)
        self.opt = torch.optim.Adam(self.model.parameters(), lr=cfg["lr"])
        self.params = list(self.model.parameters())
        self._flat = torch.empty(sum(p.numel() for p in self.params), device=dev)

    def loss_of(self, logits):
        cfg = self.cfg
        m = self.train_mask
        if cfg["loss"] == "ce":
            return F.cross_entropy(logits[m], self.labels[m], reduction="sum") / self.n_train
        return F.binary_cross_entropy_with_logits(logits[m], self.labels[m], reduction="sum") / (
            self.n_train * cfg["out_size"])

    def epoch(self):
        old = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = self.tf32
        try:
            self.model.train()
            logits = self.model(self.G, self.feats)
            loss = self.loss_of(logits)
            self.opt.zero_grad(set_to_none=False)
            loss.backward()
            if self.world > 1:  # one flat all-reduce of all weight gradients (the model is ~1 MB)
                torch.cat([p.grad.reshape(-1) for p in self.params], out=self._flat)
                self.dist.all_reduce(self._flat, group=self.group)
                off = 0
                for p in self.params:
                    p.grad.copy_(self._flat[off:off + p.numel()].view_as(p))
                    off += p.numel()
            self.opt.step()
        finally:
            torch.backends.cuda.matmul.allow_tf32 = old
        return loss

    def close(self):
        if hasattr(self.G, "close"):
            self.G.close()


def run_epochs(shape: str = "products", epochs: int = 5, warmup: int = 2, **kw) -> Dict:
    """Time `epochs` training epochs after `warmup` untimed ones (CUDA events on the launching stream, barrier
    + synchronize on both sides, max over ranks).  Returns the record every caller prints; on ranks != 0 too."""
    tr = Trainer(shape, **kw)
    dist, world, dev = tr.dist, tr.world, tr.dev
    loss = None
    for _ in range(warmup):
        loss = tr.epoch()
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier(group=tr.group)
        torch.cuda.synchronize(dev)
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    wall0 = time.perf_counter()
    t0.record()
    for _ in range(epochs):
        loss = tr.epoch()
    t1.record()
    torch.cuda.synchronize(dev)
    wall = (time.perf_counter() - wall0) * 1e3 / max(epochs, 1)
    ms = t0.elapsed_time(t1) / max(epochs, 1)
    t = torch.tensor([ms, wall], device=dev, dtype=torch.float64)
    lv = loss.detach().double().reshape(1)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=tr.group)
        dist.all_reduce(lv, group=tr.group)
    rec = {"metric": "full_graph_epochs_per_s", "value": 1e3 / float(t[0]), "unit": "epochs/s",
           "ms_per_epoch": float(t[0]), "ms_per_epoch_wall": float(t[1]), "n_gpus": world, "shape": shape,
           "nodes": tr.n, "edges": tr.e, "model": tr.model_name, "k": tr.k, "hidden": tr.hidden,
           "layers": tr.cfg["layers"], "epochs": epochs, "warmup": warmup,
           "dtype": "tf32-matmul" if tr.tf32 else "f32", "comm": tr.comm, "loss": float(lv),
           "peak_mem_GB": torch.cuda.max_memory_allocated(dev) / 1e9,
           "epoch": "forward + loss + backward + Adam step on the full graph, evaluation excluded "
                    "(maxk_gnn_dgl.py:78-96); hyper-parameters of scripts_train/*.sh"}
    tr.close()
    return rec
