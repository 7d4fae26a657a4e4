#!/usr/bin/env python3
"""`world` ranks of the peer-memory path on ONE GPU (PeerAggregator.local_group: plain device
tensors instead of CUDA IPC mappings, no torch.distributed): the multi-GPU kernels -- push of the
CBSR rows (TMA bulk copies or row lists), arrival flags, [own | other columns] task order, partial
zeroing, publish + pull reduce-scatter, epoch-parity buffers -- checked against the single-GPU
kernels with fresh inputs every step.

    python scripts/peer_local_check.py WORLD K HALO_THRESHOLD [STEPS [HOMES]]

HOMES = 2, 4, 8: the EXPERIMENTAL planned kernels (the bank plan travels in place of the selectors).

Run in its own process (tests/test_gpu_zz_peer_one_gpu.py does): the ranks' kernels wait for each other,
so every rank's step is enqueued on its own stream before anything synchronises, and a bounded
device-side wait that expires traps the whole CUDA context.  Exit code 0 = parity."""
import os
import sys

# One host thread enqueues the work of ALL ranks here.  With CUDA's default lazy module loading
# the first launch of a kernel can block the host until running kernels drain -- and a running
# kernel of rank 0 may be waiting for work of rank 1 that this very thread has not enqueued yet.
# (One process per GPU never meets this: a peer's flag is written by work that peer enqueued
# BEFORE its own waiting kernel.)  Load everything up front instead.
os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "spgemm-new_b200"), os.path.join(ROOT, "tests"), ROOT):
    sys.path.insert(0, p)
import torch  # noqa: E402

from helpers import ATOL, RTOL, Problem  # noqa: E402
from maxk_b200 import ops  # noqa: E402
from maxk_b200.dist import partition_rows_by_nnz  # noqa: E402
from maxk_b200.peer import PeerAggregator  # noqa: E402

world, k, halo = int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3])
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 4
homes = int(sys.argv[5]) if len(sys.argv) > 5 else 0
p = Problem(n=480, avg=35.0, k=k, seed=17 * world + k, empty_frac=0.1)
g = p.cuda()
w4 = g["warp4"]
bounds = partition_rows_by_nnz(g["indptr"], world, align=16)
aggs = PeerAggregator.local_group(g["indptr"], g["indices"], g["values"], bounds, k, 256,
                                  copy_ctas=2, halo_threshold=halo, homes=homes)
if halo == 0.0:
    assert all(sum(a.send_dense) == world - 1 for a in aggs)
if halo >= 2.0:
    assert all(sum(a.send_dense) == 0 for a in aggs)
streams = [torch.cuda.Stream() for _ in range(world)]
gen = torch.Generator("cuda").manual_seed(5)
ok = True
for step in range(steps):
    x = torch.randn(p.n, 256, device="cuda", generator=gen)
    dy = torch.rand(p.n, 256, device="cuda", generator=gen)
    sd, si = ops.topk_cbsr(x, k)
    y_ref = ops.spgemm_forward(w4, g["indices"], g["values"], sd, si, num_rows=p.n)
    g_ref = ops.sspmm_backward(w4, g["indices"], g["values"], dy, si)
    xs = [x[bounds[r]:bounds[r + 1]].contiguous() for r in range(world)]
    dys = [dy[bounds[r]:bounds[r + 1]].contiguous() for r in range(world)]
    sds = [sd[bounds[r]:bounds[r + 1]].contiguous() for r in range(world)]
    sis = [si[bounds[r]:bounds[r + 1]].contiguous() for r in range(world)]
    torch.cuda.synchronize()
    ys, gis, outs = [], [], []
    for r, a in enumerate(aggs):        # forward of every rank; nothing waits on the host
        with torch.cuda.stream(streams[r]):
            gd, gi = a.exchange(sds[r], sis[r]) if step % 2 else a.topk_and_exchange(xs[r])
            ys.append(a.forward(gd, gi))
            gis.append(gi)
    for r, a in enumerate(aggs):        # then the backward of every rank
        with torch.cuda.stream(streams[r]):
            outs.append(a.backward(dys[r], gis[r]))
    torch.cuda.synchronize()
    for r in range(world):
        lo, hi = bounds[r], bounds[r + 1]
        f = bool(torch.allclose(ys[r], y_ref[lo:hi], rtol=RTOL, atol=ATOL))
        b = bool(torch.allclose(outs[r], g_ref[lo:hi], rtol=RTOL, atol=ATOL))
        ok = ok and f and b
        if not (f and b):
            print(f"MISMATCH world={world} k={k} halo={halo} step={step} rank={r}: fwd {f} bwd {b}", flush=True)
print(f"peer_local_check world={world} k={k} halo={halo} steps={steps} homes={homes}: {'ok' if ok else 'FAILED'}", flush=True)
sys.exit(0 if ok else 1)
