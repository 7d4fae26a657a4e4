#!/usr/bin/env python3
"""The two collectives of one layer of the NCCL transport at a given shape's sizes (torchrun): all-gather of the CBSR
block (k fp32 + k uint8 per node) and reduce-scatter of the CBSR-shaped partial gradient, under whatever NCCL_*
environment the caller set.  No graph, no kernels: buffers only."""
import json
import os
import sys

import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
nodes = int(sys.argv[1]) if len(sys.argv) > 1 else 2449029   # products
k = int(sys.argv[2]) if len(sys.argv) > 2 else 32
rows = -(-nodes // world)
loc_data = torch.rand(rows, k, device=dev)
loc_index = torch.zeros(rows, k, dtype=torch.uint8, device=dev)
g_data = torch.empty(rows * world, k, device=dev)
g_index = torch.empty(rows * world, k, dtype=torch.uint8, device=dev)
partial = torch.rand(rows * world, k, device=dev)
rs_out = torch.empty(rows, k, device=dev)


def timed(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record(); torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / iters], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return round(float(t), 4)


def gather():
    dist.all_gather_into_tensor(g_data, loc_data); dist.all_gather_into_tensor(g_index, loc_index)


res = {"world": world, "nodes": nodes, "k": k,
       "env": {n: v for n, v in os.environ.items() if n.startswith("NCCL_")},
       "all_gather_ms": timed(gather),
       "all_gather_data_only_ms": timed(lambda: dist.all_gather_into_tensor(g_data, loc_data)),
       "reduce_scatter_ms": timed(lambda: dist.reduce_scatter_tensor(rs_out, partial))}
recv = (world - 1) * rows * k * 5
res["all_gather_GBps_into_each_gpu"] = round(recv / res["all_gather_ms"] / 1e6, 1)
if rank == 0:
    print(json.dumps(res), flush=True)
dist.destroy_process_group()
