#!/usr/bin/env python3
"""Full-graph MaxK-GNN training epochs on a synthetic BASELINE.json shape: command-line front of
maxk_b200.train.run_epochs (the loop itself lives in the package; bench.py's `epoch` leg calls the same function).

    python scripts/train_epoch.py --shape products --model sage --k 32 --epochs 10
    torchrun --nproc-per-node 8 scripts/train_epoch.py --shape products --comm peer ...
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "spgemm-new_b200"))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from maxk_b200 import train  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--shape", default="products", choices=sorted(train.CONFIGS))
ap.add_argument("--model", default=None, choices=["sage", "gcn", "gin"])
ap.add_argument("--k", type=int, default=None)
ap.add_argument("--hidden", type=int, default=256)
ap.add_argument("--epochs", type=int, default=10)
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--scale", type=float, default=1.0)
ap.add_argument("--tf32", action="store_true", help="TF32 tensor cores for the dense Linear layers")
ap.add_argument("--no-split", action="store_true")
ap.add_argument("--comm", choices=("nccl", "peer", "auto"), default="auto")
ap.add_argument("--row-cost", type=float, default=train.DEFAULT_ROW_COST)
a = ap.parse_args()

rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
rec = train.run_epochs(a.shape, epochs=a.epochs, warmup=a.warmup, model=a.model, k=a.k, hidden=a.hidden,
                       scale=a.scale, device=dev, rank=rank, world=world, comm=a.comm, row_cost=a.row_cost,
                       tf32=a.tf32, split=not a.no_split)
if rank == 0:
    print(json.dumps(rec), flush=True)
if world > 1:
    dist.destroy_process_group()
