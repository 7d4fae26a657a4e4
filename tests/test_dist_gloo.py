"""world_size-2 gloo tests (CPU) of the multi-GPU host logic in maxk_b200.dist: nnz-balanced
row partition, column remapping into the block-padded gathered CBSR buffer, local/remote
column split, CBSR all-gather and gradient reduce-scatter.  The operator set is a CPU
stand-in built on the oracle (the oracle as checker: tests only), so that everything except
the CUDA kernels themselves is exercised without a GPU."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "spgemm-new_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


class OracleOps:
    """CPU stand-in with the signatures of maxk_b200.ops that DistAggregator uses."""

    @staticmethod
    def build_worklist(indptr, warp_max_nz=64):
        from oracle import oracle as O
        return torch.from_numpy(O.worklist(indptr.numpy(), warp_max_nz))

    @staticmethod
    def topk_cbsr(x, k):
        from oracle import oracle as O
        v, i = O.topk_cbsr(x.numpy(), k)
        return torch.from_numpy(v), torch.from_numpy(i)

    @staticmethod
    def spgemm_forward(warp4, indices, values, sp_data, sp_index, feat_dim=256, num_rows=None,
                       out=None, accumulate=False):
        from oracle import oracle as O
        y = O.spgemm_forward(warp4.numpy(), indices.numpy(), values.numpy(), sp_data.numpy(),
                             sp_index.numpy(), feat_dim, n_out=num_rows).astype(np.float32)
        y = torch.from_numpy(y)
        if out is None:
            return y
        if accumulate:
            out.add_(y)
        else:
            out.copy_(y)
        return out

    @staticmethod
    def sspmm_backward(warp4, indices, values, dy, sp_index, out=None, accumulate=False):
        from oracle import oracle as O
        g = O.sspmm_backward(warp4.numpy(), indices.numpy(), values.numpy(), dy.numpy(),
                             sp_index.numpy()).astype(np.float32)
        g = torch.from_numpy(g)
        if out is None:
            return g
        if accumulate:
            out.add_(g)
        else:
            out.copy_(g)
        return out


def _problem():
    from oracle import oracle as O
    rng = np.random.default_rng(17)
    n, k, d = 700, 16, 64
    indptr, idx = O.synth_graph(n, 18.0, seed=17, empty_frac=0.1)
    val = rng.random(idx.shape[0], dtype=np.float32)
    x = rng.random((n, d), dtype=np.float32)
    dy = rng.random((n, d), dtype=np.float32)
    return n, k, d, indptr, idx, val, x, dy


def _worker(rank, world, port, split, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from maxk_b200.dist import DistAggregator, partition_rows_by_nnz
        n, k, d, indptr, idx, val, x, dy = _problem()
        tip, tix, tv = torch.from_numpy(indptr), torch.from_numpy(idx), torch.from_numpy(val)
        bounds = partition_rows_by_nnz(tip, world)
        agg = DistAggregator.from_global_csr(tip, tix, tv, bounds, rank, k, d, backend=OracleOps,
                                             split=split)
        lo, hi = bounds[rank], bounds[rank + 1]
        g_data, g_index = agg.topk_and_exchange(torch.from_numpy(x[lo:hi]))
        y = agg.forward(g_data, g_index).clone()
        dxc = agg.backward(torch.from_numpy(dy[lo:hi]), g_index).clone()
        # the per-rank host-buffer entry point (what bench.py times as e2e at N > 1): same results, twice
        from maxk_b200.host import ShardHostAggregator
        shard = ShardHostAggregator(agg)
        assert shard.h2d_bytes == 2 * (hi - lo) * d * 4 and shard.d2h_bytes == (hi - lo) * (d + k) * 4
        for _ in range(2):
            y_h, dxc_h = shard.step(torch.from_numpy(x[lo:hi]), torch.from_numpy(dy[lo:hi]))
            assert torch.equal(y_h, y) and torch.equal(dxc_h, dxc)
        q.put((rank, bounds, y.numpy(), dxc.numpy(), agg.num_tasks))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("split", [False, True])
def test_two_rank_partitioned_aggregation_matches_single_process(split):
    from oracle import oracle as O
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, split, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted([q.get(timeout=180) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    n, k, d, indptr, idx, val, x, dy = _problem()
    sp_data, sp_index = O.topk_cbsr(x, k)
    w4 = O.worklist(indptr)
    y_ref = O.spgemm_forward(w4, idx, val, sp_data, sp_index, d)
    g_ref = O.sspmm_backward(w4, idx, val, dy, sp_index)
    bounds = results[0][1]
    assert bounds[0] == 0 and bounds[-1] == n and bounds == results[1][1]
    nnz = [int(indptr[bounds[i + 1]] - indptr[bounds[i]]) for i in range(world)]
    assert abs(nnz[0] - nnz[1]) <= 0.1 * sum(nnz)          # balanced by nnz
    y = np.concatenate([r[2] for r in results])
    g = np.concatenate([r[3] for r in results])
    np.testing.assert_allclose(y, y_ref, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(g, g_ref, rtol=1e-5, atol=1e-6)


def _peer_plan_worker(rank, world, port, q):
    """Host logic of maxk_b200.peer on CPU: the plan ([own | other] task order over a concatenated
    CSR, rows referenced per owner) and the halo row-list exchange over gloo.  The kernels are
    replaced by the oracle: executing the plan's forward list on the FULL operand must give this
    rank's rows of A * X, and the lists must be exactly the rows the peers' edges reference."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle as O
        from maxk_b200.dist import partition_rows_by_nnz
        from maxk_b200.peer import PeerAggregator
        n, k, d, indptr, idx, val, x, dy = _problem()
        tip, tix, tv = torch.from_numpy(indptr), torch.from_numpy(idx), torch.from_numpy(val)
        bounds = partition_rows_by_nnz(tip, world, align=16)
        lo, hi = bounds[rank], bounds[rank + 1]
        e0, e1 = int(indptr[lo]), int(indptr[hi])
        a = PeerAggregator.__new__(PeerAggregator)
        a._plan((tip[lo:hi + 1] - tip[lo]).to(torch.int32), tix[e0:e1].contiguous(), tv[e0:e1].contiguous(),
                bounds, rank, k, d, 2, 0.3, backend=OracleOps)
        send = a._exchange_lists(None)
        sp_data, sp_index = O.topk_cbsr(x, k)
        y = O.spgemm_forward(a.f_warp4.numpy(), a.f_indices.numpy(), a.f_values.numpy(), sp_data, sp_index,
                             d, n_out=hi - lo)
        g = O.sspmm_backward(a.b_warp4.numpy(), a.b_indices.numpy(), a.b_values.numpy(), dy[lo:hi], sp_index)
        own_tasks = a.f_warp4.view(-1, 4)[: a.wait_from_task]
        own_cols_ok = True
        for row, loc, ln, _ in own_tasks.tolist():
            c = a.f_indices[loc:loc + ln]
            own_cols_ok &= bool(((c >= lo) & (c < hi)).all())
        q.put((rank, bounds, y, g, [s.numpy().copy() for s in send], [r.numpy().copy() for r in a._recv],
               own_cols_ok, a.wait_from_task, a.num_tasks))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_peer_plan_and_halo_lists_over_gloo(world):
    from oracle import oracle as O
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_peer_plan_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=180) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    n, k, d, indptr, idx, val, x, dy = _problem()
    sp_data, sp_index = O.topk_cbsr(x, k)
    w4 = O.worklist(indptr)
    y_ref = O.spgemm_forward(w4, idx, val, sp_data, sp_index, d)
    g_ref = O.sspmm_backward(w4, idx, val, dy, sp_index)
    bounds = res[0][1]
    assert all(b % 16 == 0 for b in bounds[:-1]) and bounds[-1] == n
    y = np.concatenate([r[2] for r in res])
    np.testing.assert_allclose(y, y_ref, rtol=1e-5, atol=1e-6)      # [own | other] order, global columns
    g = sum(r[3] for r in res)                                       # partial gradients add up
    np.testing.assert_allclose(g, g_ref, rtol=1e-5, atol=1e-6)
    for r in range(world):
        lo, hi = bounds[r], bounds[r + 1]
        assert res[r][6] and 0 < res[r][7] < res[r][8]               # tasks before wait_from touch own columns only
        for p in range(world):
            e0, e1 = int(indptr[bounds[p]]), int(indptr[bounds[p + 1]])
            cols_p = np.unique(idx[e0:e1])                           # what rank p's edges reference
            want = cols_p[(cols_p >= lo) & (cols_p < hi)] if p != r else cols_p[:0]
            np.testing.assert_array_equal(res[r][4][p], want.astype(np.int32), err_msg=f"send[{r}->{p}]")
            np.testing.assert_array_equal(res[p][5][r], want.astype(np.int32), err_msg=f"recv[{p}<-{r}]")


def test_partition_and_remap_helpers():
    from maxk_b200.dist import partition_rows_by_nnz, remap_columns
    indptr = torch.tensor([0, 10, 10, 11, 50, 90, 100], dtype=torch.int32)
    for world in (1, 2, 3, 4, 8):
        b = partition_rows_by_nnz(indptr, world)
        assert len(b) == world + 1 and b[0] == 0 and b[-1] == 6
        assert all(b[i] <= b[i + 1] for i in range(world))
    # alignment and the row-aware cost model (skewed graph: 4 hub rows, 4000 rows of degree 1)
    from maxk_b200.dist import row_cost_in_edges
    deg = torch.cat([torch.full((4,), 1000), torch.ones(4000)]).long()
    skew = torch.zeros(4005, dtype=torch.int64)
    torch.cumsum(deg, 0, out=skew[1:])
    by_nnz = partition_rows_by_nnz(skew, 2, align=16)
    by_cost = partition_rows_by_nnz(skew, 2, align=16, row_cost=row_cost_in_edges(16))
    assert by_nnz[1] % 16 == 0 and by_cost[1] % 16 == 0
    assert by_nnz[1] <= 16                      # edges only: one rank gets nearly every row
    assert 1000 < by_cost[1] < 3000             # rows weigh 11.6 edges each at k=16
    assert abs(row_cost_in_edges(32) - 2048 / 336) < 1e-9
    bounds = [0, 3, 4, 10]
    cols = torch.tensor([0, 2, 3, 4, 9], dtype=torch.int32)
    pos = remap_columns(cols, bounds, max_rows=6)
    assert pos.tolist() == [0, 2, 6, 12, 17]
