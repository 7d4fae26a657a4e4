#!/usr/bin/env python3
"""Full-graph MaxK-GNN training epochs on a synthetic BASELINE.json shape (SURVEY 8f rank 2;
replaces maxk_gnn_dgl.py:78-113 without DGL/OGB): epoch = forward + loss + backward + Adam
step on the whole graph, evaluation excluded.  1 GPU, or N GPUs under torchrun (row
partition, CBSR all-gather / reduce-scatter inside every layer, one flat all-reduce of the
weight gradients per step).  Prints one JSON line with epochs/s.

    python scripts/train_epoch.py --shape products --model sage --k 32 --epochs 10
    torchrun --nproc-per-node 8 scripts/train_epoch.py --shape products ...
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "spgemm-new_b200"))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from maxk_b200 import graphs  # noqa: E402
from maxk_b200.dist import partition_rows_by_nnz  # noqa: E402
from maxk_b200.graph import DistGraph, Graph  # noqa: E402
from utils import models as M  # noqa: E402

CONFIGS = {  # scripts_train/*.sh
    "products": dict(in_size=100, out_size=47, layers=3, lr=0.003, loss="ce"),
    "proteins": dict(in_size=8, out_size=112, layers=3, lr=0.01, loss="bce"),
    "reddit": dict(in_size=602, out_size=41, layers=4, lr=0.01, loss="ce"),
    "flickr": dict(in_size=500, out_size=7, layers=3, lr=0.001, loss="ce"),
}

ap = argparse.ArgumentParser()
ap.add_argument("--shape", default="products")
ap.add_argument("--model", default="sage", choices=["sage", "gcn", "gin"])
ap.add_argument("--k", type=int, default=32)
ap.add_argument("--hidden", type=int, default=256)
ap.add_argument("--epochs", type=int, default=10)
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--scale", type=float, default=1.0)
ap.add_argument("--tf32", action="store_true", help="TF32 tensor cores for the dense Linear layers")
ap.add_argument("--no-split", action="store_true")
ap.add_argument("--comm", choices=("nccl", "peer"), default="nccl")
ap.add_argument("--row-cost", type=float, default=186.0,
                help="weight of one node row, in edges, when balancing the row partition: the dense "
                     "Linear/LayerNorm work of an epoch is per row (measured on one B200, products "
                     "shape: 78 ns per row vs 0.42 ns per edge over 3 layers); 0 = balance edges only")
a = ap.parse_args()

rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
torch.backends.cuda.matmul.allow_tf32 = a.tf32
torch.backends.cudnn.allow_tf32 = a.tf32

cfg = CONFIGS[a.shape]
g = graphs.make_graph(a.shape, seed=0, device=dev, values="ones", scale=a.scale)
n = g["v_num"]
if world > 1:
    bounds = partition_rows_by_nnz(g["indptr"], world, align=16, row_cost=a.row_cost)
    G = DistGraph(g["indptr"], g["indices"], bounds, rank, split=not a.no_split, comm=a.comm)
    lo, hi = bounds[rank], bounds[rank + 1]
else:
    G = Graph(g["indptr"], g["indices"])
    lo, hi = 0, n
e_total = g["e_num"]
del g
torch.cuda.empty_cache()

gen = torch.Generator(dev).manual_seed(97 + rank)  # --seed 97 (utils/config.py)
feats = torch.randn(hi - lo, cfg["in_size"], device=dev, generator=gen)
if cfg["loss"] == "ce":
    labels = torch.randint(0, cfg["out_size"], (hi - lo,), device=dev, generator=gen)
else:
    labels = (torch.rand(hi - lo, cfg["out_size"], device=dev, generator=gen) < 0.1).float()
train_mask = torch.rand(hi - lo, device=dev, generator=gen) < 0.1
n_train = train_mask.sum().float()
if world > 1:
    dist.all_reduce(n_train)

torch.manual_seed(97)  # identical initial weights on every rank
model = {"sage": M.SAGE, "gcn": M.GCN, "gin": M.GIN}[a.model](
    cfg["in_size"], a.hidden, cfg["layers"], cfg["out_size"], maxk=a.k, feat_drop=0.5, norm=True).to(dev)
opt = torch.optim.Adam(model.parameters(), lr=cfg["lr"])
params = [p for p in model.parameters()]


def epoch():
    model.train()
    logits = model(G, feats)
    if cfg["loss"] == "ce":
        loss = F.cross_entropy(logits[train_mask], labels[train_mask], reduction="sum") / n_train
    else:
        loss = F.binary_cross_entropy_with_logits(logits[train_mask], labels[train_mask],
                                                  reduction="sum") / (n_train * cfg["out_size"])
    opt.zero_grad(set_to_none=False)
    loss.backward()
    if world > 1:  # one flat all-reduce of all weight gradients (the model is ~2 MB)
        flat = torch.cat([p.grad.reshape(-1) for p in params])
        dist.all_reduce(flat)
        off = 0
        for p in params:
            p.grad.copy_(flat[off:off + p.numel()].view_as(p))
            off += p.numel()
    opt.step()
    return loss


for _ in range(a.warmup):
    loss = epoch()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0.record()
for _ in range(a.epochs):
    loss = epoch()
t1.record()
torch.cuda.synchronize()
ms = t0.elapsed_time(t1) / a.epochs
t = torch.tensor([ms], device=dev, dtype=torch.float64)
lv = loss.detach().double().reshape(1)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(lv)
if rank == 0:
    print(json.dumps({"metric": "full_graph_epochs_per_s", "value": 1e3 / float(t), "ms_per_epoch": float(t),
                      "n_gpus": world, "shape": a.shape, "nodes": n, "edges": e_total, "model": a.model,
                      "k": a.k, "hidden": a.hidden, "layers": cfg["layers"], "epochs": a.epochs,
                      "dtype": "tf32-matmul" if a.tf32 else "f32", "loss": float(lv),
                      "peak_mem_GB": torch.cuda.max_memory_allocated() / 1e9,
                      "epoch": "forward + loss + backward + Adam step, evaluation excluded"}), flush=True)
if world > 1:
    dist.destroy_process_group()
