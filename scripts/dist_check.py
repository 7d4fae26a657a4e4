#!/usr/bin/env python3
"""Multi-GPU parity check (run under torchrun, one rank per GPU): the row-partitioned
DistAggregator (CBSR all-gather + reduce-scatter over NCCL) against the single-GPU kernels
on the same graph.  Exit code 0 = parity within fp32 rtol 1e-5 / atol 1e-6."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "spgemm-new_b200"))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from maxk_b200 import graphs, ops  # noqa: E402
from maxk_b200.dist import DistAggregator, partition_rows_by_nnz  # noqa: E402

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
ok = True
for shape, scale, k in (("flickr", 1.0, 32), ("reddit", 0.05, 16), ("products", 0.05, 64)):
    g = graphs.make_graph(shape, seed=1, device=dev, values="uniform", scale=scale)
    n = g["v_num"]
    x = torch.rand(n, 256, device=dev, generator=torch.Generator(dev).manual_seed(3))
    dy = torch.rand(n, 256, device=dev, generator=torch.Generator(dev).manual_seed(4))
    w4 = ops.build_worklist(g["indptr"])
    sd, si = ops.topk_cbsr(x, k)
    y_ref = ops.spgemm_forward(w4, g["indices"], g["values"], sd, si, num_rows=n)
    g_ref = ops.sspmm_backward(w4, g["indices"], g["values"], dy, si)
    bounds = partition_rows_by_nnz(g["indptr"], world)
    lo, hi = bounds[rank], bounds[rank + 1]
    for split in (False, True):
        agg = DistAggregator.from_global_csr(g["indptr"], g["indices"], g["values"], bounds, rank, k, 256,
                                             split=split)
        for _ in range(2):  # twice: buffers are reused across steps
            gd, gi = agg.topk_and_exchange(x[lo:hi].contiguous())
            y = agg.forward(gd, gi)
            dxc = agg.backward(dy[lo:hi].contiguous(), gi)
        torch.cuda.synchronize()
        e1 = torch.allclose(y, y_ref[lo:hi], rtol=1e-5, atol=1e-6)
        e2 = torch.allclose(dxc, g_ref[lo:hi], rtol=1e-5, atol=1e-6)
        flag = torch.tensor([int(e1 and e2)], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if rank == 0:
            print(f"{shape} k={k} split={split} world={world}: fwd {e1} bwd {e2} all-ranks {bool(flag.item())}",
                  flush=True)
        ok = ok and bool(flag.item())
dist.destroy_process_group()
sys.exit(0 if ok else 1)
