#!/usr/bin/env python3
"""Cost of decoding the backward's columns from the bank plan (what the multi-GPU peer path runs when the plan
travels in place of the selectors) against the plain selector bytes: Reddit and proteins shapes, k = 32 / 64."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "spgemm-new_b200"))
import torch  # noqa: E402

from maxk_b200 import graphs, ops  # noqa: E402


def timeit(fn, iters=10, warm=3):
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


for shape in ("reddit", "proteins"):
    g = graphs.make_graph(shape, seed=0, device="cuda", values="uniform")
    n = g["v_num"]
    w4 = ops.build_worklist(g["indptr"])
    x = torch.rand(n, 256, device="cuda", generator=torch.Generator("cuda").manual_seed(123))
    dy = torch.rand(n, 256, device="cuda", generator=torch.Generator("cuda").manual_seed(5))
    for k in (32, 64):
        sd, si = ops.topk_cbsr(x, k)
        sa = ops.cbsr_bank_plan(si)
        dxc = torch.empty(n, k, device="cuda")
        t0 = timeit(lambda: ops.sspmm_backward(w4, g["indices"], g["values"], dy, si, out=dxc))
        ref = dxc.clone()
        t1 = timeit(lambda: ops.sspmm_backward_planned(w4, g["indices"], g["values"], dy, sa, out=dxc))
        err = float((dxc - ref).abs().max() / ref.abs().max())
        print(json.dumps({"shape": shape, "k": k, "bwd_selectors_ms": round(t0, 4), "bwd_columns_from_plan_ms": round(t1, 4),
                          "rel_diff": err}), flush=True)
