#!/usr/bin/env python3
"""Which kernels a full-graph training epoch spends its time in (torch.profiler, one epoch after two warm-ups).

    python scripts/epoch_profile.py --shape products --model sage --k 32
"""
import argparse, os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "spgemm-new_b200"))
import torch
from torch.profiler import profile, ProfilerActivity
from maxk_b200 import train

ap = argparse.ArgumentParser()
ap.add_argument("--shape", default="products")
ap.add_argument("--model", default="sage")
ap.add_argument("--k", type=int, default=32)
ap.add_argument("--rows", type=int, default=22)
a = ap.parse_args()
rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(local)
if world > 1:   # torchrun: every rank trains, rank 0 prints its own kernels
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
tr = train.Trainer(a.shape, model=a.model, k=a.k, device=torch.device("cuda", local), rank=rank, world=world)
for _ in range(2):
    tr.epoch()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    tr.epoch()
    torch.cuda.synchronize()
rows = []
for e in prof.key_averages():
    t = getattr(e, "device_time_total", None)
    if t is None:
        t = getattr(e, "cuda_time_total", 0)
    rows.append((t / 1e3, e.count, e.key))
rows.sort(reverse=True)
total = sum(r[0] for r in rows)
if rank == 0:
  print(json.dumps({"shape": a.shape, "model": a.model, "k": a.k, "n_gpus": world, "kernel_ms_total": round(total, 2)}))
  for ms, cnt, key in rows[:a.rows]:
    print(f"{ms:9.3f} ms  {100 * ms / total:5.1f}%  x{cnt:<3d} {key[:110]}")
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
