#!/usr/bin/env python3
"""How long the HOST takes to issue one training epoch (no synchronisation inside), against the GPU time of the epoch:
at 8 GPUs the per-GPU work shrinks 8x and the host side does not."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "spgemm-new_b200"))
import torch
from maxk_b200 import train
for model in ("sage", "gcn"):
    tr = train.Trainer("products", model=model, k=32, device=torch.device("cuda", 0))
    for _ in range(3):
        tr.epoch()
    torch.cuda.synchronize()
    host, gpu = [], []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter(); a.record()
        tr.epoch()
        b.record(); t1 = time.perf_counter()
        torch.cuda.synchronize()
        host.append((t1 - t0) * 1e3); gpu.append(a.elapsed_time(b))
    print(json.dumps({"model": model, "host_issue_ms": round(min(host), 2), "gpu_ms": round(min(gpu), 2),
                      "launches": int(__import__("maxk_b200")._cabi.launch_count())}), flush=True)
    del tr
    torch.cuda.empty_cache()
