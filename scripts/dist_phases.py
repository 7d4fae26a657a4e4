#!/usr/bin/env python3
"""Per-phase timing of the multi-GPU aggregation step (torchrun, one rank per GPU): every phase
is bracketed by a synchronize + barrier, so the numbers are SERIAL costs (no overlap) -- they
say where the time is, the overlapped step is what bench.py measures."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "spgemm-new_b200"))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
from maxk_b200 import graphs, ops  # noqa: E402
from maxk_b200.dist import DistAggregator, partition_rows_by_nnz  # noqa: E402

shape = sys.argv[1] if len(sys.argv) > 1 else "reddit"
k = int(sys.argv[2]) if len(sys.argv) > 2 else 32
g = graphs.make_graph(shape, seed=0, device=dev, values="uniform")
bounds = partition_rows_by_nnz(g["indptr"], world, align=16)
agg = DistAggregator.from_global_csr(g["indptr"], g["indices"], g["values"], bounds, rank, k, 256)
gen = torch.Generator(dev).manual_seed(123 + rank)
x = torch.rand(agg.n_local, 256, device=dev, generator=gen)
dy = torch.rand(agg.n_local, 256, device=dev, generator=gen)
sd, si = ops.topk_cbsr(x, k)
lp, rp = agg.part


def timed(fn, iters=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record(); torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / iters], device=dev, dtype=torch.float64)
    mx, mn = t.clone(), t.clone()
    dist.all_reduce(mx, op=dist.ReduceOp.MAX); dist.all_reduce(mn, op=dist.ReduceOp.MIN)
    return round(float(mx), 4), round(float(mn), 4)


def copy_in():
    agg.loc_data[: agg.n_local].copy_(sd); agg.loc_index[: agg.n_local].copy_(si)


def gather():
    dist.all_gather_into_tensor(agg.g_data, agg.loc_data); dist.all_gather_into_tensor(agg.g_index, agg.loc_index)


own = agg.g_index[rank * agg.max_rows:(rank + 1) * agg.max_rows]
res = {"world": world, "shape": shape, "k": k,
       "tasks_local": lp["warp4"].numel() // 4, "tasks_remote": rp["warp4"].numel() // 4,
       "edges_local": lp["indices"].numel(), "edges_remote": rp["indices"].numel(), "rows": agg.n_local}
res["copy_in"] = timed(copy_in)
res["all_gather"] = timed(gather)
res["fwd_local"] = timed(lambda: ops.spgemm_forward(lp["warp4"], lp["indices"], lp["values"], agg.loc_data, agg.loc_index, feat_dim=256, num_rows=agg.n_local, out=agg.y))
res["fwd_remote_acc"] = timed(lambda: ops.spgemm_forward(rp["warp4"], rp["indices"], rp["values"], agg.g_data, agg.g_index, feat_dim=256, num_rows=agg.n_local, out=agg.y, accumulate=True))
res["bwd_remote"] = timed(lambda: ops.sspmm_backward(rp["warp4"], rp["indices"], rp["values"], dy, agg.g_index, out=agg.partial))
res["reduce_scatter"] = timed(lambda: dist.reduce_scatter_tensor(agg.rs_out, agg.partial))
res["bwd_local"] = timed(lambda: ops.sspmm_backward(lp["warp4"], lp["indices"], lp["values"], dy, own, out=None))


def step():
    gd, gi = agg.exchange(sd, si); agg.forward(gd, gi); agg.backward(dy, gi)


res["step_overlapped"] = timed(step, 20)
agg2 = DistAggregator.from_global_csr(g["indptr"], g["indices"], g["values"], bounds, rank, k, 256, split=False)


def step2():
    gd, gi = agg2.exchange(sd, si); agg2.forward(gd, gi); agg2.backward(dy, gi)


res["step_unsplit"] = timed(step2, 20)
res["fwd_unsplit"] = timed(lambda: ops.spgemm_forward(agg2.warp4, agg2.indices, agg2.values, agg2.g_data, agg2.g_index, feat_dim=256, num_rows=agg2.n_local, out=agg2.y))
res["bwd_unsplit"] = timed(lambda: ops.sspmm_backward(agg2.warp4, agg2.indices, agg2.values, dy, agg2.g_index, out=agg2.partial))
if rank == 0:
    print(json.dumps(res), flush=True)
dist.barrier()
dist.destroy_process_group()
