#!/usr/bin/env python3
"""Fused Linear + TopK on products- and Reddit-sized inputs (normal and uniform data), k = 16 / 32 / 64: one number each."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "spgemm-new_b200"))
import torch
from maxk_b200 import ops

def timeit(fn, iters=10, warm=3):
    for _ in range(warm): fn()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); s.record()
    for _ in range(iters): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / iters

g = torch.Generator("cuda").manual_seed(1)
out = {}
for name, n in (("products", 2449029), ("reddit", 232965)):
    for dist in ("normal", "uniform"):
        x = (torch.randn(n, 256, device="cuda", generator=g) if dist == "normal"
             else torch.rand(n, 256, device="cuda", generator=g))
        w = (torch.randn(256, 256, device="cuda", generator=g) / 16 if dist == "normal"
             else torch.eye(256, device="cuda"))          # identity weight: the output keeps the uniform distribution
        b = torch.randn(256, device="cuda", generator=g) * (0.1 if dist == "normal" else 0.0)
        for k in (16, 32, 64):
            out[f"{name}_{dist}_k{k}_ms"] = round(timeit(lambda: ops.linear_topk_cbsr(x, w, b, k)), 4)
        if dist == "normal":
            out[f"{name}_linear_only_ms"] = round(timeit(lambda: ops.linear_topk_cbsr(x, w, b, 32, want_linear=True, want_cbsr=False)), 4)
        del x
print(json.dumps(out), flush=True)
