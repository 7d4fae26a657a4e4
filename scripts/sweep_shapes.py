#!/usr/bin/env python3
"""k = 8 / 16 / 32 / 64 x the four BASELINE.json graph shapes on one B200: work list, TopK, bank plan, forward SpGEMM
(as the layers call it: ops.spgemm_forward_auto) and backward SSpMM, ms per launch (CUDA events, 10 iterations after 3
warm-ups, inputs resident) and the fraction of the measured HBM peak by SURVEY 8(d)'s algorithmic bytes.  Restates the
sweep of the reference's C++ bench (kernels/main.cu:111-116).  One JSON document on stdout.

    python scripts/sweep_shapes.py > gpurun_out/sweep.json      # committed as profiles/r02_sweep.json
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "spgemm-new_b200"))
import torch  # noqa: E402

from maxk_b200 import graphs, ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--shapes", default="reddit,products,proteins,flickr")
ap.add_argument("--ks", default="8,16,32,64")
ap.add_argument("--iters", type=int, default=10)
a = ap.parse_args()
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    peak_src = "measured"
except Exception:  # noqa: BLE001
    peak, peak_src = 6650.0, "fallback"


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


doc = {"peak_GBps": peak, "peak_source": peak_src, "protocol": "CUDA events, %d iterations after 3 warm-ups, inputs resident; "
       "X ~ U[0,1) seed 123, values ~ U[0,1), dY ~ U[0,1); frac = SURVEY 8(d) algorithmic bytes / time / peak" % a.iters,
       "rows": []}
for shape in a.shapes.split(","):
    g = graphs.make_graph(shape, seed=0, device="cuda", values="uniform")
    n, e = g["v_num"], g["e_num"]
    w4 = ops.build_worklist(g["indptr"])
    w = w4.numel() // 4
    t_wl = timeit(lambda: ops.build_worklist(g["indptr"]), a.iters)
    x = torch.rand(n, 256, device="cuda", generator=torch.Generator("cuda").manual_seed(123))
    dy = torch.rand(n, 256, device="cuda", generator=torch.Generator("cuda").manual_seed(5))
    y = torch.empty(n, 256, device="cuda")
    for k in [int(v) for v in a.ks.split(",")]:
        b = graphs.algorithmic_bytes(n, e, w, k)
        sd, si = ops.topk_cbsr(x, k)
        dxc = torch.empty(n, k, device="cuda")
        t_topk = timeit(lambda: ops.topk_cbsr(x, k, out=(sd, si)), a.iters)
        planned = ops.use_bank_plan(k, w, n, n)
        t_plan = None
        if planned:
            sa = torch.empty(n, k, dtype=torch.uint16, device="cuda")
            t_plan = timeit(lambda: ops.cbsr_bank_plan(si, out=sa), a.iters)
        t_fwd = timeit(lambda: ops.spgemm_forward_auto(w4, g["indices"], g["values"], sd, si, out=y, num_rows=n), a.iters)
        t_fwd_plain = timeit(lambda: ops.spgemm_forward(w4, g["indices"], g["values"], sd, si, out=y, num_rows=n), a.iters) \
            if planned else t_fwd
        t_bwd = timeit(lambda: ops.sspmm_backward(w4, g["indices"], g["values"], dy, si, out=dxc), a.iters)
        row = {"shape": shape, "nodes": n, "edges": e, "tasks": w, "k": k,
               "worklist_ms": round(t_wl, 4), "worklist_frac": round(b["meta"] / t_wl / 1e6 / peak, 3),
               "topk_ms": round(t_topk, 4), "topk_frac": round(b["topk"] / t_topk / 1e6 / peak, 3),
               "fwd_ms": round(t_fwd, 4), "fwd_frac": round(b["fwd"] / t_fwd / 1e6 / peak, 3),
               "fwd_path": ("bank plan + planned kernel (plan inside fwd_ms)" if planned else
                            "cp.async ring" if (w < 2 * n and n * 5 * k > (96 << 20)) else "default"),
               "bank_plan_ms": None if t_plan is None else round(t_plan, 4),
               "fwd_ms_without_plan": round(t_fwd_plain, 4),
               "bwd_ms": round(t_bwd, 4), "bwd_frac": round(b["bwd"] / t_bwd / 1e6 / peak, 3),
               "fwd_edges_per_s": e / (t_fwd * 1e-3), "bwd_edges_per_s": e / (t_bwd * 1e-3)}
        doc["rows"].append(row)
        print(json.dumps(row), file=sys.stderr, flush=True)
    del g, w4, x, dy, y
    torch.cuda.empty_cache()
print(json.dumps(doc, indent=1))
