#!/usr/bin/env python3
"""Parity + timing of the NVLink peer-memory path (PeerAggregator: fused pull all-gather +
SpGEMM kernel, pull reduce-scatter) against the single-GPU kernels and against the NCCL path
(DistAggregator).  Run under torchrun, one rank per GPU.  Exit code 0 = parity."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "spgemm-new_b200"))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from maxk_b200 import graphs, ops  # noqa: E402
from maxk_b200.dist import DistAggregator, partition_rows_by_nnz  # noqa: E402
from maxk_b200.peer import PeerAggregator  # noqa: E402

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
quick = "--quick" in sys.argv
time_only = "--time-only" in sys.argv
ok = True
# (shape, scale, k, halo_threshold, homes): halo 0.5 = automatic choice per pair, 0 = whole blocks always,
# 2 = row lists always; homes 2 = the bank plan travels in place of the selectors, "auto" = ops.use_bank_plan
cases = (("flickr", 1.0, 32, 0.5, 0), ("reddit", 0.05, 16, 0.5, "auto"), ("products", 0.05, 64, 2.0, 0),
         ("reddit", 0.05, 32, 0.5, 2), ("reddit", 0.05, 64, 0.0, "auto"), ("products", 0.05, 32, 2.0, 2),
         ("reddit", 0.02, 8, 2.0, 0), ("products", 0.05, 32, 0.0, "auto"), ("flickr", 1.0, 16, 2.0, 0))
if time_only:
    cases = cases[:1]
for shape, scale, k, halo, homes in cases:
    g = graphs.make_graph(shape, seed=1, device=dev, values="uniform", scale=scale)
    n = g["v_num"]
    w4 = ops.build_worklist(g["indptr"])
    bounds = partition_rows_by_nnz(g["indptr"], world, align=16)
    lo, hi = bounds[rank], bounds[rank + 1]
    agg = PeerAggregator.from_global_csr(g["indptr"], g["indices"], g["values"], bounds, rank, k, 256,
                                         copy_ctas=(3 if k != 16 else -5), halo_threshold=halo, homes=homes)
    if rank == 0:
        print(f"peer {shape} k={k} halo_threshold={halo} homes={homes}->{agg.homes}: dense send pairs {sum(agg.send_dense)}, "
              f"rows pushed per step {agg.exchange_rows} of {agg.n_local} x {world - 1}", flush=True)
    for step in range(5):  # different inputs every step: a stale buffer would show
        x = torch.rand(n, 256, device=dev, generator=torch.Generator(dev).manual_seed(3 + step))
        dy = torch.rand(n, 256, device=dev, generator=torch.Generator(dev).manual_seed(40 + step))
        sd, si = ops.topk_cbsr(x, k)
        y_ref = ops.spgemm_forward(w4, g["indices"], g["values"], sd, si, num_rows=n)
        g_ref = ops.sspmm_backward(w4, g["indices"], g["values"], dy, si)
        if step % 2 == 0:
            gd, gi = agg.topk_and_exchange(x[lo:hi].contiguous())
        else:
            gd, gi = agg.exchange(sd[lo:hi].contiguous(), si[lo:hi].contiguous())
        y = agg.forward(gd, gi)
        dxc = agg.backward(dy[lo:hi].contiguous(), gi)
        torch.cuda.synchronize()
        e1 = torch.allclose(y, y_ref[lo:hi], rtol=1e-5, atol=1e-6)
        e2 = torch.allclose(dxc, g_ref[lo:hi], rtol=1e-5, atol=1e-6)
        flag = torch.tensor([int(e1 and e2)], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if rank == 0:
            print(f"peer {shape} k={k} world={world} step {step}: fwd {e1} bwd {e2} all-ranks {bool(flag.item())}",
                  flush=True)
        ok = ok and bool(flag.item())
    # the host-buffer entry point of a rank (bench.py's e2e at N > 1): one stream, then with the copy overlap
    from maxk_b200.host import ShardHostAggregator
    for overlap in (False, True):
        shard = ShardHostAggregator(agg, overlap=overlap)
        for step in range(3):
            x = torch.rand(n, 256, device=dev, generator=torch.Generator(dev).manual_seed(70 + step))
            dy = torch.rand(n, 256, device=dev, generator=torch.Generator(dev).manual_seed(80 + step))
            sd, si = ops.topk_cbsr(x, k)
            y_ref = ops.spgemm_forward(w4, g["indices"], g["values"], sd, si, num_rows=n)
            g_ref = ops.sspmm_backward(w4, g["indices"], g["values"], dy, si)
            y_h, dxc_h = shard.step(x[lo:hi].cpu().pin_memory(), dy[lo:hi].cpu().pin_memory())
            e1 = torch.allclose(y_h, y_ref[lo:hi].cpu(), rtol=1e-5, atol=1e-6)
            e2 = torch.allclose(dxc_h, g_ref[lo:hi].cpu(), rtol=1e-5, atol=1e-6)
            flag = torch.tensor([int(e1 and e2)], device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if rank == 0:
                print(f"shard-host {shape} k={k} world={world} overlap={overlap} step {step}: fwd {e1} bwd {e2} "
                      f"all-ranks {bool(flag.item())}", flush=True)
            ok = ok and bool(flag.item())
        del shard
    # ... and two steps in flight (submit / result), every result checked
    shard = ShardHostAggregator(agg)
    refs, prev = {}, None
    for step in range(5):
        x = torch.rand(n, 256, device=dev, generator=torch.Generator(dev).manual_seed(170 + step))
        dy = torch.rand(n, 256, device=dev, generator=torch.Generator(dev).manual_seed(180 + step))
        sd, si = ops.topk_cbsr(x, k)
        refs[step] = (ops.spgemm_forward(w4, g["indices"], g["values"], sd, si, num_rows=n)[lo:hi].cpu(),
                      ops.sspmm_backward(w4, g["indices"], g["values"], dy, si)[lo:hi].cpu())
        t = shard.submit(x[lo:hi].cpu().pin_memory(), dy[lo:hi].cpu().pin_memory())
        for done in ([prev] if prev is not None else []) + ([t] if step == 4 else []):
            y_h, dxc_h = shard.result(done)
            e1 = torch.allclose(y_h, refs[done][0], rtol=1e-5, atol=1e-6)
            e2 = torch.allclose(dxc_h, refs[done][1], rtol=1e-5, atol=1e-6)
            flag = torch.tensor([int(e1 and e2)], device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if rank == 0:
                print(f"shard-host pipelined {shape} k={k} world={world} step {done}: fwd {e1} bwd {e2} "
                      f"all-ranks {bool(flag.item())}", flush=True)
            ok = ok and bool(flag.item())
        prev = t
    del shard
    agg.close()
    del agg

if not quick:
    # timing on the Reddit-shaped graph: NCCL path vs peer-memory path, overlapped steps
    k = 32
    g = graphs.make_graph("reddit", seed=0, device=dev, values="uniform")
    bounds = partition_rows_by_nnz(g["indptr"], world, align=16)
    lo, hi = bounds[rank], bounds[rank + 1]
    gen = torch.Generator(dev).manual_seed(123 + rank)
    x = torch.rand(hi - lo, 256, device=dev, generator=gen)
    dy = torch.rand(hi - lo, 256, device=dev, generator=gen)
    sd, si = ops.topk_cbsr(x, k)

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

    res = {"world": world}
    nccl = DistAggregator.from_global_csr(g["indptr"], g["indices"], g["values"], bounds, rank, k, 256)

    def step_nccl():
        gd, gi = nccl.exchange(sd, si); nccl.forward(gd, gi); nccl.backward(dy, gi)
    res["nccl_step_ms"] = timed(step_nccl)
    del nccl
    for cc in (-64, 4, 8, 16, 32):
        peer = PeerAggregator.from_global_csr(g["indptr"], g["indices"], g["values"], bounds, rank, k, 256,
                                              copy_ctas=cc)

        def step_peer():
            gd, gi = peer.exchange(sd, si); peer.forward(gd, gi); peer.backward(dy, gi)

        def fwd_only():
            gd, gi = peer.exchange(sd, si); peer.forward(gd, gi)
        tag = f"tma{cc}" if cc > 0 else f"lsu{-cc}"
        res[f"peer_step_ms_{tag}"] = timed(step_peer)
        res[f"peer_fwd_ms_{tag}"] = timed(fwd_only)
        peer.close()
        del peer
    if rank == 0:
        print(json.dumps(res), flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
