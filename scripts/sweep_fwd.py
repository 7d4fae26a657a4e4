#!/usr/bin/env python3
"""Tuning sweep of the forward/backward kernels on one BASELINE.json shape (GPU only)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "spgemm-new_b200"))
import torch  # noqa: E402

from maxk_b200 import graphs, ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--shape", default="reddit")
ap.add_argument("--ks", default="32,16,8,64")
ap.add_argument("--copies", default="0,1,2,4,8,16")
ap.add_argument("--tpw", default="16")
ap.add_argument("--depth", default="0")
ap.add_argument("--bwd-tpw", default="")
ap.add_argument("--planned", action="store_true", help="also time the bank plan and the planned forward (k = 32, 64)")
ap.add_argument("--iters", type=int, default=10)
a = ap.parse_args()


def timeit(fn, iters, warm=3):
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


g = graphs.make_graph(a.shape, seed=0, device="cuda", values="uniform")
n, e = g["v_num"], g["e_num"]
w4 = ops.build_worklist(g["indptr"])
w = w4.numel() // 4
x = torch.rand(n, 256, device="cuda", generator=torch.Generator("cuda").manual_seed(123))
dy = torch.rand(n, 256, device="cuda", generator=torch.Generator("cuda").manual_seed(5))
y = torch.empty(n, 256, device="cuda")
peak = 6543.7
res = []
for k in [int(v) for v in a.ks.split(",")]:
    sd, si = ops.topk_cbsr(x, k)
    b = graphs.algorithmic_bytes(n, e, w, k)
    for tpw in [int(v) for v in a.tpw.split(",")]:
        for c in [int(v) for v in a.copies.split(",")]:
            for dep in [int(v) for v in a.depth.split(",")]:
                ops.set_forward_copies(c)
                ops.set_forward_depth(dep)
                ops.set_tuning(tpw, 0)
                t = timeit(lambda: ops.spgemm_forward(w4, g["indices"], g["values"], sd, si, out=y, num_rows=n), a.iters)
                r = {"kernel": "fwd", "shape": a.shape, "k": k, "copies": c, "tpw": tpw, "depth": dep,
                     "ms": round(t, 4), "frac": round(b["fwd"] / t / 1e6 / peak, 3)}
                res.append(r)
                print(json.dumps(r), flush=True)
    ops.set_forward_copies(0)
    ops.set_forward_depth(0)
    if a.planned and k in (32, 64):
        # the plan is made once per TopK, the planned forward runs once per layer: both are timed
        ref = ops.spgemm_forward(w4, g["indices"], g["values"], sd, si, num_rows=n)
        sa = ops.cbsr_bank_plan(si)
        t_plan = timeit(lambda: ops.cbsr_bank_plan(si, out=sa), a.iters)
        for dep in (1, 2):
            ops.set_forward_depth(dep)
            t = timeit(lambda: ops.spgemm_forward_planned(w4, g["indices"], g["values"], sd, sa, out=y, num_rows=n), a.iters)
            err = float((y - ref).abs().max() / ref.abs().max())
            r = {"kernel": "fwd_planned", "shape": a.shape, "k": k, "depth": dep,
                 "ms": round(t, 4), "plan_ms": round(t_plan, 4), "frac": round(b["fwd"] / t / 1e6 / peak, 3),
                 "frac_with_plan": round(b["fwd"] / (t + t_plan) / 1e6 / peak, 3), "rel_err_vs_default": err}
            res.append(r)
            print(json.dumps(r), flush=True)
        ops.set_forward_depth(0)
    if a.bwd_tpw:
        dxc = torch.empty(n, k, device="cuda")
        for tpw in [int(v) for v in a.bwd_tpw.split(",")]:
            ops.set_tuning(0, tpw)
            t = timeit(lambda: ops.sspmm_backward(w4, g["indices"], g["values"], dy, si, out=dxc), a.iters)
            r = {"kernel": "bwd", "shape": a.shape, "k": k, "tpw": tpw, "ms": round(t, 4),
                 "frac": round(b["bwd"] / t / 1e6 / peak, 3)}
            res.append(r)
            print(json.dumps(r), flush=True)
ops.set_tuning(16, 16)
