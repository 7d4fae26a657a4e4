#!/usr/bin/env python3
"""BASELINE.json configs[4]: R-MAT power-law sweep, GCN MaxK k=16 aggregation (sym-norm edge
values), forward SpGEMM + backward SSpMM.  One GPU by default; under torchrun the graph is row-
partitioned and the CBSR blocks are exchanged (--comm peer|nccl).  Prints one JSON line per graph."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "spgemm-new_b200"))
import torch  # noqa: E402

from maxk_b200 import graphs, ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--cases", default="20:5,20:20,20:100,20:500,22:5,22:20,22:100,24:5,24:20,25:5")
ap.add_argument("--k", type=int, default=16)
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--comm", choices=("peer", "nccl"), default="peer")
ap.add_argument("--partition", choices=("nnz", "cost"), default="cost",
                help="balance edges only, or edges + dense row traffic (dist.row_cost_in_edges)")
a = ap.parse_args()
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
    from maxk_b200.dist import DistAggregator, partition_rows_by_nnz, row_cost_in_edges
    from maxk_b200.peer import PeerAggregator
PEAK = 6543.7
k = a.k


def timeit(fn, iters, warm=2):
    for _ in range(warm):
        fn()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    t = torch.tensor([s.elapsed_time(e) / iters], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


def topk_chunked(n_rows, seed):
    """CBSR of U[0,1) features without materialising all of X (33 M x 256 fp32 = 34 GB)."""
    sd = torch.empty((n_rows, k), dtype=torch.float32, device=dev)
    si = torch.empty((n_rows, k), dtype=torch.uint8, device=dev)
    gen = torch.Generator(dev).manual_seed(seed)
    step = 1 << 20
    for r0 in range(0, n_rows, step):
        r1 = min(n_rows, r0 + step)
        x = torch.rand(r1 - r0, 256, device=dev, generator=gen)
        ops.topk_cbsr(x, k, out=(sd[r0:r1], si[r0:r1]))
    return sd, si


for case in a.cases.split(","):
    lg, deg = case.split(":")
    lg, deg = int(lg), float(deg)
    g = graphs.make_rmat(lg, deg, seed=0, device=dev, values="gcn")
    n, e = g["v_num"], g["e_num"]
    out = {"graph": f"rmat{lg}", "nodes": n, "edges": e, "avg_deg": deg, "k": k, "n_gpus": world}
    if world == 1:
        w4 = ops.build_worklist(g["indptr"])
        w = w4.numel() // 4
        sd, si = topk_chunked(n, 1)
        y = torch.empty(n, 256, device=dev)
        dy = torch.rand(n, 256, device=dev, generator=torch.Generator(dev).manual_seed(2))
        dxc = torch.empty(n, k, device=dev)
        tf = timeit(lambda: ops.spgemm_forward(w4, g["indices"], g["values"], sd, si, out=y, num_rows=n), a.iters)
        tb = timeit(lambda: ops.sspmm_backward(w4, g["indices"], g["values"], dy, si, out=dxc), a.iters)
        tm = timeit(lambda: ops.build_worklist(g["indptr"]), a.iters)
        out.update({"fwd_ms": round(tf, 4), "bwd_ms": round(tb, 4), "worklist_ms": round(tm, 4)})
    else:
        bounds = partition_rows_by_nnz(g["indptr"], world, align=16,
                                       row_cost=row_cost_in_edges(k) if a.partition == "cost" else 0.0)
        out["partition"] = a.partition
        lo, hi = bounds[rank], bounds[rank + 1]
        cls = PeerAggregator if a.comm == "peer" else DistAggregator
        agg = cls.from_global_csr(g["indptr"], g["indices"], g["values"], bounds, rank, k, 256)
        wt = torch.tensor([agg.num_tasks], device=dev, dtype=torch.int64)
        dist.all_reduce(wt)
        w = int(wt)
        sd, si = topk_chunked(hi - lo, 1 + rank)
        dy = torch.rand(hi - lo, 256, device=dev, generator=torch.Generator(dev).manual_seed(2 + rank))

        def fwd():
            gd, gi = agg.exchange(sd, si)
            agg.forward(gd, gi)
            return gi

        def both():
            agg.backward(dy, fwd())
        tf = timeit(fwd, a.iters)
        ts = timeit(both, a.iters)
        out.update({"fwd_ms": round(tf, 4), "bwd_ms": round(ts - tf, 4), "comm": a.comm})
        if a.comm == "peer":
            st = torch.tensor([agg.exchange_rows, sum(agg.send_dense), agg.n_local], device=dev, dtype=torch.int64)
            allst = [torch.zeros_like(st) for _ in range(world)]
            dist.all_gather(allst, st)
            out["rows_pushed_per_rank"] = [int(x[0]) for x in allst]
            out["dense_pairs_per_rank"] = [int(x[1]) for x in allst]
            out["rows_owned_per_rank"] = [int(x[2]) for x in allst]
        if hasattr(agg, "close"):
            agg.close()
        del agg
    b = graphs.algorithmic_bytes(n, e, w, k)
    out.update({"tasks": w,
                "fwd_GBps": round(b["fwd"] / out["fwd_ms"] / 1e6, 1), "fwd_frac": round(b["fwd"] / out["fwd_ms"] / 1e6 / PEAK / world, 3),
                "bwd_GBps": round(b["bwd"] / out["bwd_ms"] / 1e6, 1), "bwd_frac": round(b["bwd"] / out["bwd_ms"] / 1e6 / PEAK / world, 3),
                "edges_per_s": round(2 * e / ((out["fwd_ms"] + out["bwd_ms"]) * 1e-3))})
    if rank == 0:
        print(json.dumps(out), flush=True)
    del g
    torch.cuda.empty_cache()
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
