#!/usr/bin/env python3
"""Short driver for ncu: build the Reddit-shaped problem, then run the hot-path kernels a
few times (work list, TopK, forward SpGEMM, backward SSpMM).  Not a benchmark -- numbers
printed under a profiler are never bench values."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "spgemm-new_b200"))
import torch  # noqa: E402

from maxk_b200 import graphs, ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--shape", default="reddit")
ap.add_argument("--k", type=int, default=32)
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--scale", type=float, default=1.0)
ap.add_argument("--depth", type=int, default=0, help="forward launch shape (maxk_set_forward_depth)")
a = ap.parse_args()
g = graphs.make_graph(a.shape, seed=0, device="cuda", values="uniform", scale=a.scale)
n = g["v_num"]
gen = torch.Generator("cuda").manual_seed(123)
x = torch.rand(n, 256, device="cuda", generator=gen)
dy = torch.rand(n, 256, device="cuda", generator=gen)
y = torch.empty(n, 256, device="cuda")
dxc = torch.empty(n, a.k, device="cuda")
ops.set_forward_depth(a.depth)
for _ in range(a.iters):
    w4 = ops.build_worklist(g["indptr"])
    sd, si = ops.topk_cbsr(x, a.k)
    ops.spgemm_forward_auto(w4, g["indices"], g["values"], sd, si, out=y, num_rows=n)   # bank plan where it pays
    ops.sspmm_backward(w4, g["indices"], g["values"], dy, si, out=dxc)
torch.cuda.synchronize()
print("ok", n, g["e_num"], w4.numel() // 4, float(y.sum()), float(dxc.sum()))
