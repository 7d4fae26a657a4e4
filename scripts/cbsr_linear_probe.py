#!/usr/bin/env python3
"""EXPERIMENTAL CBSR x dense products (csrc/cbsr_dense.cu) against torch's dense fp32 Linear on the scattered
twin: fc_self forward and the gradient of the kept values, on N rows of hidden 256.

    python scripts/cbsr_linear_probe.py [--n 2449029] [--ks 32,16,64] [--c 256]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "spgemm-new_b200"))
import torch  # noqa: E402

from maxk_b200 import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=2449029)   # ogbn-products-shaped
ap.add_argument("--ks", default="32,16,64")
ap.add_argument("--c", type=int, default=256)
ap.add_argument("--iters", type=int, default=10)
a = ap.parse_args()


def timeit(fn, iters=a.iters, warm=3):
    for _ in range(warm):
        fn()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


torch.backends.cuda.matmul.allow_tf32 = False
gen = torch.Generator("cuda").manual_seed(0)
n, d, c = a.n, 256, a.c
x = torch.randn(n, d, device="cuda", generator=gen)
w = torch.randn(c, d, device="cuda", generator=gen) * 0.06
g = torch.randn(n, c, device="cuda", generator=gen)
m = w.t().contiguous()
for k in [int(v) for v in a.ks.split(",")]:
    sd, si = ops.topk_cbsr(x, k)
    h = ops.cbsr_scatter(sd, si, d)
    y = torch.empty(n, c, device="cuda")
    dv = torch.empty(n, k, device="cuda")
    t_mm = timeit(lambda: ops.cbsr_matmul(sd, si, m, out=y))
    t_sd = timeit(lambda: ops.cbsr_sampled_dot(g, m, si, out=dv))
    t_scatter = timeit(lambda: ops.cbsr_scatter(sd, si, d))
    t_lin = timeit(lambda: torch.mm(h, m, out=y))
    gx = torch.empty(n, d, device="cuda")
    t_lin_bwd = timeit(lambda: torch.mm(g, w, out=gx))
    t_gather = timeit(lambda: ops.cbsr_gather(gx, si))
    err = float((ops.cbsr_matmul(sd, si, m) - torch.mm(h, m)).abs().max())
    err_b = float((ops.cbsr_sampled_dot(g, m, si) - ops.cbsr_gather(torch.mm(g, w), si)).abs().max())
    print(json.dumps({"n": n, "k": k, "c": c, "cbsr_matmul_ms": round(t_mm, 3), "scatter_plus_dense_mm_ms":
                      round(t_scatter + t_lin, 3), "dense_mm_ms": round(t_lin, 3),
                      "cbsr_sampled_dot_ms": round(t_sd, 3), "dense_mm_bwd_plus_gather_ms": round(t_lin_bwd + t_gather, 3),
                      "smem_bytes_model_ms": round(n * k * c * 4 / (148 * 128 * 1.9e9) * 1e3, 3),
                      "max_abs_diff_fwd": err, "max_abs_diff_bwd": err_b}), flush=True)
