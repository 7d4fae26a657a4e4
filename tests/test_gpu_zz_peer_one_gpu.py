"""The multi-GPU (peer-memory) kernels with 2-4 ranks living on ONE GPU, each case in its own process
(scripts/peer_local_check.py).  Last file of the `-m gpu` suite on purpose: the ranks' kernels wait for
each other through flag words, so a case needs the GPU to run kernels of different streams side by side
-- which CUDA makes likely, not certain.  A device-side wait that expires traps (it never hangs); such a
run says nothing about parity and is repeated, a MISMATCH is a failure at once."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_local_ranks(*argv, attempts=3):
    last = None
    for _ in range(attempts):
        last = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "peer_local_check.py"),
                               *[str(a) for a in argv]], capture_output=True, text=True, timeout=120)
        if last.returncode == 0:
            assert "ok" in last.stdout
            return
        assert "MISMATCH" not in last.stdout, (last.stdout[-600:], last.stderr[-1200:])
        if "timed out waiting for peer" not in last.stdout + last.stderr:
            break   # anything but an expired co-scheduling wait is not retried
    assert last.returncode == 0, (last.stdout[-600:], last.stderr[-1200:])


@pytest.mark.parametrize("world,k,halo", [(2, 32, 0.0), (2, 16, 2.0), (3, 32, 0.3), (3, 64, 2.0), (2, 8, 2.0),
                                          (4, 32, 0.5)])
def test_peer_kernels_multi_rank_on_one_gpu(world, k, halo):
    """TMA push / row-list push, arrival flags, [own | other columns] task order, partial zeroing,
    publish + pull reduce-scatter, epoch-parity buffers (PeerAggregator.local_group), against the
    single-GPU kernels with fresh inputs every step."""
    run_local_ranks(world, k, halo)


@pytest.mark.skipif(os.environ.get("MAXK_EXPERIMENTAL") != "1",
                    reason="planned kernels on the peer-memory path: written after round 1's GPU budget was spent, "
                           "checked with 2-4 ranks under the SIMT interpreter only; opt in with MAXK_EXPERIMENTAL=1")
@pytest.mark.parametrize("world,k,halo,homes", [(2, 32, 0.0, 8), (3, 32, 0.3, 2), (2, 64, 2.0, 4), (2, 16, 0.0, 4)])
def test_peer_planned_kernels_multi_rank_on_one_gpu(world, k, halo, homes):
    run_local_ranks(world, k, halo, 4, homes)


@pytest.mark.parametrize("overlap", [False, pytest.param(True, marks=pytest.mark.skipif(
    os.environ.get("MAXK_EXPERIMENTAL") != "1", reason="overlapped ShardHostAggregator: written after round 1's GPU "
    "budget was spent; opt in with MAXK_EXPERIMENTAL=1"))], ids=["one-stream", "overlap-experimental"])
def test_shard_host_aggregator_on_a_one_rank_job(overlap):
    """ShardHostAggregator.step (what bench.py times as e2e at N > 1) over the peer-memory path of a one-rank
    job: host rows in, host rows out, three steps with fresh inputs, against the oracle."""
    import socket

    import numpy as np
    import torch
    import torch.distributed as dist

    from helpers import Problem, assert_close
    from maxk_b200.host import ShardHostAggregator
    from maxk_b200.peer import PeerAggregator
    from oracle import oracle as O
    p = Problem(n=700, avg=50.0, k=32, seed=123, empty_frac=0.1)
    g = p.cuda()
    own_pg = not dist.is_initialized()
    if own_pg:
        with socket.socket() as sk_:
            sk_.bind(("127.0.0.1", 0))
            port = sk_.getsockname()[1]
        dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=0, world_size=1)
    try:
        agg = PeerAggregator(g["indptr"], g["indices"], g["values"], [0, p.n], 0, 32, 256, copy_ctas=2)
        shard = ShardHostAggregator(agg, overlap=overlap)
        assert shard.overlap == overlap
        rng = np.random.default_rng(4)
        for step in range(3):
            x = rng.random((p.n, 256), dtype=np.float32)
            dy = rng.random((p.n, 256), dtype=np.float32)
            sd, si = O.topk_cbsr(x, 32)
            y_h, dxc_h = shard.step(torch.from_numpy(x).pin_memory(), torch.from_numpy(dy).pin_memory())
            assert not y_h.is_cuda and not dxc_h.is_cuda
            assert_close(y_h, O.spgemm_forward(p.w4, p.idx, p.val, sd, si, 256, n_out=p.n), f"Y, step {step}")
            assert_close(dxc_h, O.sspmm_backward(p.w4, p.idx, p.val, dy, si), f"dXc, step {step}")
        agg.close()
    finally:
        if own_pg:
            dist.destroy_process_group()
