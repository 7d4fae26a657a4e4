#!/usr/bin/env python3
"""The reference README's headline comparison, restated on a B200: MaxK-GNN SpGEMM / SSpMM
against cuSPARSE CSR SpMM on the dense (masked) features -- kernels/main.cu:163-185 and
main_runner_direct.py:152-203 (`analyze_speedups`).  cuSPARSE is reached through
maxk_cuda_kernels.cusparse_spmm (a library call, comparator only).  One JSON line per graph."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "spgemm-new_b200"))
import torch  # noqa: E402

import maxk_cuda_kernels as mk  # noqa: E402
from maxk_b200 import graphs, ops  # noqa: E402

shapes = sys.argv[1].split(",") if len(sys.argv) > 1 else ["reddit", "proteins", "products", "flickr"]
ks = [16, 32, 64]


def timeit(fn, iters=5, warm=2):
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


for shape in shapes:
    g = graphs.make_graph(shape, seed=0, device="cuda", values="uniform")
    n, e = g["v_num"], g["e_num"]
    w4 = ops.build_worklist(g["indptr"])
    x = torch.rand(n, 256, device="cuda", generator=torch.Generator("cuda").manual_seed(123))
    dy = torch.rand(n, 256, device="cuda", generator=torch.Generator("cuda").manual_seed(5))
    y = torch.empty(n, 256, device="cuda")
    # cuSPARSE forward does not depend on k (dense operand); the transpose graph serves the backward
    a_t = torch.sparse_csr_tensor(g["indptr"].long(), g["indices"].long(), g["values"], size=(n, n)) \
        .to_sparse_coo().t().coalesce().to_sparse_csr()
    t_cus_f = timeit(lambda: mk.cusparse_spmm(g["indptr"], g["indices"], g["values"], x))
    t_cus_b = timeit(lambda: torch.sparse.mm(a_t, dy))
    out = {"graph": shape, "nodes": n, "edges": e, "avg_deg": round(e / n, 1),
           "cusparse_fwd_ms": round(t_cus_f, 3), "cusparse_bwd_ms": round(t_cus_b, 3)}
    del a_t
    for k in ks:
        sd, si = ops.topk_cbsr(x, k)
        dxc = torch.empty(n, k, device="cuda")
        tf = timeit(lambda: ops.spgemm_forward(w4, g["indices"], g["values"], sd, si, out=y, num_rows=n))
        tb = timeit(lambda: ops.sspmm_backward(w4, g["indices"], g["values"], dy, si, out=dxc))
        out[f"k{k}"] = {"fwd_ms": round(tf, 3), "bwd_ms": round(tb, 3),
                        "fwd_speedup": round(t_cus_f / tf, 2), "bwd_speedup": round(t_cus_b / tb, 2)}
    print(json.dumps(out), flush=True)
    del g, x, dy, y
    torch.cuda.empty_cache()
