"""ShardHostAggregator over the peer-memory path of a ONE-rank job (no kernel waits for another rank's kernel).

The 2-4-ranks-on-one-GPU cases that used to live here enqueued mutually waiting kernels on one device, which
/opt/skills/guides/B200_PROFILING.md forbids (Xid 109 on this pool); they are gone.  The multi-rank logic is covered
by the SIMT multi-rank cases (tests/test_simt_python_layer.py), the protocol model (tests/test_peer_protocol_model.py)
and, on real multi-GPU hardware, by the `parity_ok` check bench.py runs at N > 1 outside its timed region."""
import os

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("overlap", [False, True], ids=["one-stream", "overlap"])
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


def test_shard_host_aggregator_two_steps_in_flight():
    """submit / result (what bench.py times as e2e at N > 1): five steps with fresh inputs, two in flight, every
    result against the oracle -- a slot reused too early or a download racing the next step's kernels shows here."""
    import socket

    import numpy as np
    import torch
    import torch.distributed as dist

    from helpers import Problem, assert_close
    from maxk_b200.host import ShardHostAggregator
    from maxk_b200.peer import PeerAggregator
    from oracle import oracle as O
    p = Problem(n=900, avg=60.0, k=32, seed=77, empty_frac=0.1)
    g = p.cuda()
    own_pg = not dist.is_initialized()
    if own_pg:
        with socket.socket() as sk_:
            sk_.bind(("127.0.0.1", 0))
            port = sk_.getsockname()[1]
        dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=0, world_size=1)
    try:
        agg = PeerAggregator(g["indptr"], g["indices"], g["values"], [0, p.n], 0, 32, 256, copy_ctas=2)
        shard = ShardHostAggregator(agg)
        assert shard.pipelined and len(shard.slots) == 2
        rng = np.random.default_rng(5)
        xs = [rng.random((p.n, 256), dtype=np.float32) for _ in range(5)]
        dys = [rng.random((p.n, 256), dtype=np.float32) for _ in range(5)]
        pinned = [(torch.from_numpy(x).pin_memory(), torch.from_numpy(dy).pin_memory()) for x, dy in zip(xs, dys)]

        def check(i, y_h, dxc_h):
            sd, si = O.topk_cbsr(xs[i], 32)
            assert_close(y_h, O.spgemm_forward(p.w4, p.idx, p.val, sd, si, 256, n_out=p.n), f"Y, step {i}")
            assert_close(dxc_h, O.sspmm_backward(p.w4, p.idx, p.val, dys[i], si), f"dXc, step {i}")

        prev = None
        for i in range(5):
            t = shard.submit(*pinned[i])
            if prev is not None:
                check(prev, *shard.result(prev))
            prev = t
        check(prev, *shard.result(prev))
        with pytest.raises(RuntimeError, match="no longer in flight"):
            shard.result(0)
        y_h, dxc_h = shard.step(*pinned[1])     # the synchronous form still works after the pipeline
        check(1, y_h, dxc_h)
        agg.close()
    finally:
        if own_pg:
            dist.destroy_process_group()
