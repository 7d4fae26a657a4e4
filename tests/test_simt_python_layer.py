"""The Python host layer (maxk_b200.ops, the autograd Functions, the CBSR-native models) run on CPU
tensors against the kernel sources under the SIMT interpreter.

TEST DOUBLE, not a fallback: inside these tests only, `_cabi._lib` is pointed at
tests/simt/_build/libmaxk_simt.so (same C ABI, host pointers for device pointers) and the three
places where `ops` touches CUDA (`_cuda`, `_stream`, `torch.cuda.device`) are patched.  Outside the
fixture nothing changes: `_cabi.load()` binds lib/libmaxk_b200.so and raises when it is missing.
Purpose: the argument plumbing of the Python layer (the bank plan and its automatic use included),
checked against the oracle without a GPU.
"""
import contextlib
import ctypes

import numpy as np
import pytest
import torch

import simt_lib as S
from helpers import ATOL, RTOL, Problem
from oracle import oracle as O


@pytest.fixture()
def ops(monkeypatch):
    from maxk_b200 import _cabi
    from maxk_b200 import ops as _ops

    lib = S.load()
    for name, (res, args) in _cabi.PROTOTYPES.items():
        fn = getattr(lib, name)
        if args is not None:
            fn.restype, fn.argtypes = res, args
    monkeypatch.setattr(_cabi, "_lib", lib)

    def host_tensor(t, name, dtype=None, contiguous=True):
        _ops._req(isinstance(t, torch.Tensor), f"{name} must be a tensor")
        if dtype is not None:
            _ops._req(t.dtype == dtype, f"{name} must be {str(dtype).replace('torch.', '')}")
        if contiguous:
            _ops._req(t.is_contiguous(), f"{name} must be contiguous")

    monkeypatch.setattr(_ops, "_cuda", host_tensor)
    monkeypatch.setattr(_ops, "_stream", lambda t: None)
    monkeypatch.setattr(torch.cuda, "device", lambda *_a, **_k: contextlib.nullcontext())
    yield _ops
    lib.maxk_set_tuning(-1, -1)
    lib.maxk_set_forward_depth(0)
    lib.maxk_set_forward_copies(0)
    # restore what simt_lib expects of its own handle
    lib.maxk_last_error.restype = ctypes.c_char_p


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def close(got, want, what=""):
    np.testing.assert_allclose(got.detach().numpy().astype(np.float64), want, rtol=RTOL, atol=ATOL, err_msg=what)


def _problem(**kw):
    p = Problem(**kw)
    return p, dict(indptr=_t(p.indptr), indices=_t(p.idx), values=_t(p.val), x=_t(p.x), sp_data=_t(p.sp_data),
                   sp_index=_t(p.sp_index), warp4=_t(p.w4).reshape(-1), dy=_t(p.dy))


def test_ops_default_kernels_through_python(ops):
    p, g = _problem(n=150, avg=20.0, k=32, seed=4)
    w4 = ops.build_worklist(g["indptr"])
    assert torch.equal(w4, g["warp4"])
    v, i, dense = ops.topk_cbsr(g["x"], 32, want_dense=True)
    assert torch.equal(i, g["sp_index"]) and torch.equal(v, g["sp_data"])
    np.testing.assert_array_equal(dense.numpy(), O.maxk_dense(p.x, 32))
    y = ops.spgemm_forward(w4, g["indices"], g["values"], v, i, num_rows=p.n)
    close(y, p.fwd_oracle())
    dxc = ops.sspmm_backward(w4, g["indices"], g["values"], g["dy"], i)
    close(dxc, p.bwd_oracle())
    d = ops.cbsr_scatter(v, i, 256)
    np.testing.assert_array_equal(d.numpy(), O.cbsr_scatter(p.sp_data, p.sp_index, 256))
    assert torch.equal(ops.cbsr_gather(d, i), v)
    rec = ops.cbsr_pack(v, i)
    v2, i2 = ops.cbsr_unpack(rec, p.n, 32)
    assert torch.equal(v2, v) and torch.equal(i2, i)
    with pytest.raises(RuntimeError, match="accumulate=True needs"):
        ops.spgemm_forward(w4, g["indices"], g["values"], v, i, accumulate=True)


@pytest.mark.parametrize("k", [32, 64])
def test_ops_planned_kernels_through_python(ops, k):
    p, g = _problem(n=120, avg=30.0, k=k, seed=k)
    sp_addr = ops.cbsr_bank_plan(g["sp_index"])
    assert sp_addr.dtype == torch.uint16 and tuple(sp_addr.shape) == (p.n, k)
    assert int(sp_addr.to(torch.int64).max()) < 2 * 256
    y = ops.spgemm_forward_planned(g["warp4"], g["indices"], g["values"], g["sp_data"], sp_addr, num_rows=p.n)
    close(y, p.fwd_oracle(), "planned forward")
    ops.spgemm_forward_planned(g["warp4"], g["indices"], g["values"], g["sp_data"], sp_addr, num_rows=p.n,
                               out=y, accumulate=True)
    close(y, 2 * p.fwd_oracle(), "planned forward, accumulate")
    dxc = ops.sspmm_backward_planned(g["warp4"], g["indices"], g["values"], g["dy"], sp_addr)
    close(dxc, p.bwd_oracle(), "backward with the columns decoded from the plan")
    out = torch.empty_like(sp_addr)
    assert ops.cbsr_bank_plan(g["sp_index"], out=out) is out and torch.equal(out, sp_addr)
    with pytest.raises(RuntimeError, match="wrong shape"):
        ops.sspmm_backward_planned(g["warp4"], g["indices"], g["values"], g["dy"], sp_addr, out=torch.empty(3, 3))


def test_use_bank_plan_rule(ops):
    assert ops.use_bank_plan(32, 1_905_323, 232_965, 232_965)          # Reddit shape
    assert ops.use_bank_plan(64, 660_000, 132_534, 132_534)            # proteins shape
    assert not ops.use_bank_plan(32, 3_200_000, 2_449_029, 2_449_029)  # products: short rows, operand beyond the L2
    assert not ops.use_bank_plan(16, 1_905_323, 232_965, 232_965)      # k = 16 is not built
    assert not ops.use_bank_plan(32, 1_000_000, 100_000, 2_449_029)    # long rows but a 470 MB operand


@pytest.mark.parametrize("planned", [False, True])
def test_autograd_functions_with_and_without_the_plan(ops, monkeypatch, planned):
    """CBSRAggregate / TopKCBSR / SpGEMMFunction: same values and gradients whether the aggregation
    runs through the default forward kernel or (ops.use_bank_plan) through the bank plan."""
    from maxk_b200.autograd import CBSRAggregate, SpGEMMFunction, TopKCBSR
    calls = []
    real = ops.cbsr_bank_plan
    monkeypatch.setattr(ops, "use_bank_plan", lambda *a: planned)
    monkeypatch.setattr(ops, "cbsr_bank_plan", lambda *a, **kw: (calls.append(1), real(*a, **kw))[1])
    p, g = _problem(n=100, avg=15.0, k=32, seed=21)
    x = g["x"].clone().requires_grad_(True)
    sp_data, sp_index = TopKCBSR.apply(x, 32)
    y = CBSRAggregate.apply(sp_data, sp_index, g["warp4"], g["indices"], g["values"], 256, p.n)
    close(y, p.fwd_oracle())
    y.backward(g["dy"])
    want = O.cbsr_scatter(p.bwd_oracle().astype(np.float32), p.sp_index, 256)
    np.testing.assert_allclose(x.grad.numpy(), want, rtol=RTOL, atol=ATOL)
    x2 = g["x"].clone().requires_grad_(True)
    y2 = SpGEMMFunction.apply(x2, (g["indptr"], g["indices"], g["values"]), 32)
    close(y2, p.fwd_oracle())
    y2.backward(g["dy"])
    np.testing.assert_allclose(x2.grad.numpy(), want, rtol=RTOL, atol=ATOL)
    assert len(calls) == (2 if planned else 0)


class _HostGraph:
    """what maxk_b200.graph.Graph is to the models, on host tensors"""

    def __init__(self, p, g):
        from maxk_b200.graph import edge_values
        self.g, self.n = g, p.n
        self._vals = {kind: edge_values(g["indptr"], g["indices"], kind, p.n) for kind in ("sum", "mean", "gcn")}

    def aggregate(self, sp_data, sp_index, kind, feat_dim):
        from maxk_b200.autograd import CBSRAggregate
        return CBSRAggregate.apply(sp_data, sp_index, self.g["warp4"], self.g["indices"], self._vals[kind].contiguous(),
                                   feat_dim, self.n)


@pytest.mark.parametrize("name", ["SAGE", "GCN", "GIN"])
def test_models_train_step_same_with_and_without_the_plan(ops, monkeypatch, name):
    """One forward + backward of each CBSR-native model (hidden 64, k = 32): the planned kernels
    must give the loss and the parameter gradients of the default ones."""
    import utils.models as M
    p, g = _problem(n=80, avg=8.0, k=8, seed=5, empty_frac=0.0)
    G = _HostGraph(p, g)
    torch.manual_seed(0)
    model = getattr(M, name)(16, 64, 2, 5, maxk=32, feat_drop=0.0, norm=True)
    feats = torch.randn(p.n, 16)
    target = torch.randint(0, 5, (p.n,))
    results = []
    for planned in (False, True):
        monkeypatch.setattr(ops, "use_bank_plan", lambda *a, _p=planned: _p)
        model.zero_grad()
        loss = torch.nn.functional.cross_entropy(model(G, feats), target)
        loss.backward()
        results.append((float(loss.detach()), [q.grad.clone() for q in model.parameters()]))
    assert np.isfinite(results[0][0])
    assert abs(results[0][0] - results[1][0]) < 1e-5
    for a, b in zip(results[0][1], results[1][1]):
        torch.testing.assert_close(a, b, rtol=1e-4, atol=1e-5)


def test_cbsr_linear_function_matches_dense_linear(ops):
    """EXPERIMENTAL CBSRLinear (fc_self on the CBSR operand): value and all three gradients equal those of
    torch's Linear on the scattered twin."""
    from maxk_b200.autograd import CBSRLinear, TopKCBSR
    torch.manual_seed(3)
    x = torch.randn(90, 256)
    lin = torch.nn.Linear(256, 64)
    gy = torch.randn(90, 64)
    outs = []
    for sparse in (False, True):
        xs = x.clone().requires_grad_(True)
        lin.zero_grad()
        sp_data, sp_index = TopKCBSR.apply(xs, 32)
        if sparse:
            y = CBSRLinear.apply(sp_data, sp_index, lin.weight, lin.bias)
        else:
            y = lin(_scatter(ops, sp_data, sp_index))
        y.backward(gy)
        outs.append((y.detach(), xs.grad.clone(), lin.weight.grad.clone(), lin.bias.grad.clone()))
    for a, b, what in zip(outs[0], outs[1], ("y", "dx", "dW", "db")):
        torch.testing.assert_close(b, a, rtol=1e-4, atol=2e-5, msg=lambda m, w=what: f"{w}: {m}")
    y_nb = CBSRLinear.apply(sp_data.detach(), sp_index, lin.weight.detach(), None)
    torch.testing.assert_close(y_nb, outs[0][0] - lin.bias.detach(), rtol=1e-4, atol=2e-5)


def _scatter(ops, sp_data, sp_index):
    import utils.models as M
    return M._Scatter.apply(sp_data, sp_index, 256)


def test_cbsr_add_scaled_function_matches_the_dense_expression(ops):
    """EXPERIMENTAL CBSRAddScaled (GIN's (1 + eps) * h added in place): value and the gradients of y, sp_data, eps."""
    from maxk_b200.autograd import CBSRAddScaled, TopKCBSR
    torch.manual_seed(5)
    x, base, gy = torch.randn(60, 256), torch.randn(60, 256), torch.randn(60, 256)
    outs = []
    for fused in (False, True):
        xs, ys = x.clone().requires_grad_(True), base.clone().requires_grad_(True)
        eps = torch.nn.Parameter(torch.full((1,), 0.3))
        sp_data, sp_index = TopKCBSR.apply(xs, 32)
        if fused:
            out = CBSRAddScaled.apply(ys * 1.0, sp_data, sp_index, eps)
        else:
            out = (1.0 + eps) * _scatter(ops, sp_data, sp_index) + ys
        out.backward(gy)
        outs.append((out.detach(), xs.grad.clone(), ys.grad.clone(), eps.grad.clone()))
    for a, b, what in zip(outs[0], outs[1], ("value", "dx", "dy", "d eps")):
        torch.testing.assert_close(b, a, rtol=1e-5, atol=1e-5, msg=lambda m, w=what: f"{w}: {m}")


def test_linear_topk_function_matches_linear_then_topk(ops):
    """EXPERIMENTAL LinearTopKCBSR: same CBSR pair and same gradients of x, W, b as Linear followed by TopKCBSR."""
    from maxk_b200.autograd import LinearTopKCBSR, TopKCBSR
    torch.manual_seed(9)
    x0 = torch.randn(70, 48)
    lin = torch.nn.Linear(48, 256)
    g = torch.randn(70, 32)
    outs = []
    for fused in (False, True):
        x = x0.clone().requires_grad_(True)
        lin.zero_grad()
        if fused:
            sp_data, sp_index = LinearTopKCBSR.apply(x, lin.weight, lin.bias, 32)
        else:
            sp_data, sp_index = TopKCBSR.apply(lin(x), 32)
        sp_data.backward(g)
        outs.append((sp_data.detach(), sp_index, x.grad.clone(), lin.weight.grad.clone(), lin.bias.grad.clone()))
    assert torch.equal(outs[0][1], outs[1][1]) and torch.equal(outs[0][0], outs[1][0])
    for a, b, what in zip(outs[0][2:], outs[1][2:], ("dx", "dW", "db")):
        torch.testing.assert_close(b, a, rtol=1e-4, atol=1e-5, msg=lambda m, w=what: f"{w}: {m}")


@pytest.mark.parametrize("name", ["SAGE", "GIN", "GCN"])
def test_train_step_same_with_layers_on_the_cbsr_operand(ops, monkeypatch, name):
    import utils.models as M
    p, g = _problem(n=80, avg=8.0, k=8, seed=5, empty_frac=0.0)
    G = _HostGraph(p, g)
    torch.manual_seed(0)
    model = getattr(M, name)(16, 64, 2, 5, maxk=16, feat_drop=0.0, norm=True)
    feats = torch.randn(p.n, 16)
    target = torch.randint(0, 5, (p.n,))
    results = []
    for on in ("0", "1"):
        monkeypatch.setenv("MAXK_CBSR_NATIVE", on)
        model.zero_grad()
        loss = torch.nn.functional.cross_entropy(model(G, feats), target)
        loss.backward()
        results.append((float(loss.detach()), [q.grad.clone() for q in model.parameters()]))
    assert np.isfinite(results[0][0]) and abs(results[0][0] - results[1][0]) < 1e-5
    for a, b in zip(results[0][1], results[1][1]):
        torch.testing.assert_close(a, b, rtol=1e-4, atol=1e-5)


def test_worklist_cache_is_keyed_on_the_callers_tensor(ops):
    """An int64 indptr (what DGL / torch hand over) is converted once: the second call is a cache hit,
    an in-place change of the tensor is not."""
    from maxk_b200 import autograd as A
    p, g = _problem(n=50, avg=10.0, k=8, seed=2)
    A._WL_CACHE.clear()
    ip64 = g["indptr"].long()
    w1 = A.worklist_for(ip64)
    assert torch.equal(w1, g["warp4"]) and len(A._WL_CACHE) == 1
    assert A.worklist_for(ip64) is w1 and len(A._WL_CACHE) == 1
    assert A.worklist_for(g["indptr"]) is not w1 and len(A._WL_CACHE) == 2
    ip64[-1] -= 1   # bumps the version counter: a different graph
    w2 = A.worklist_for(ip64)
    assert w2 is not w1 and int(w2.reshape(-1, 4)[:, 2].sum()) == int(ip64[-1])
    A._WL_CACHE.clear()


def test_the_double_is_gone_after_the_fixture():
    from maxk_b200 import _cabi
    assert _cabi._lib is None or "libmaxk_b200.so" in str(_cabi._lib._name)


# ------------------------------------------------------------------- NVLink peer-memory path, on the CPU
@pytest.fixture()
def peer_group(ops, monkeypatch):
    """PeerAggregator.local_group on host tensors: every rank's fused forward / pull reduce-scatter
    kernel runs in its own host thread under the interpreter (the kernels of different ranks wait
    for each other through flag words, exactly as on the GPUs); dense pairs are pushed with loads and
    stores (copy_ctas < 0) or through the interpreter's stand-in for the TMA unit (mbarrier +
    cp.async.bulk, data landing as late as PTX allows)."""
    from concurrent.futures import ThreadPoolExecutor
    from maxk_b200 import peer as P

    registry = {}

    def wrap(ptr, shape, dtype, device):
        return registry[int(ptr)].view(*shape)

    monkeypatch.setattr(P, "_wrap", wrap)
    monkeypatch.setattr(P.PeerAggregator, "_stream", lambda self: None)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)

    def make(p, g, bounds, k, halo_threshold, copy_ctas, homes=0):
        world = len(bounds) - 1
        indptr, indices, values = g["indptr"].long(), g["indices"], g["values"]
        aggs = []
        for r in range(world):
            lo, hi = bounds[r], bounds[r + 1]
            e0, e1 = int(indptr[lo]), int(indptr[hi])
            a = P.PeerAggregator.__new__(P.PeerAggregator)
            a.group, a._shared, a.homes = None, [], homes
            a._plan((indptr[lo:hi + 1] - indptr[lo]).to(torch.int32), indices[e0:e1].contiguous(),
                    values[e0:e1].contiguous(), list(bounds), r, k, p.d, copy_ctas, halo_threshold)
            aggs.append(a)
        n_total = aggs[0].n_total

        def alloc(shape, dtype):
            ts = [torch.zeros(shape, dtype=dtype) for _ in range(world)]
            for t in ts:
                registry[t.data_ptr()] = t
            return [t.data_ptr() for t in ts]
        flag_ptrs = alloc((64,), torch.int32)
        tables = [(alloc((max(n_total, 1), k), torch.float32),
                   alloc((max(n_total, 1), k), torch.uint16 if homes else torch.uint8),
                   alloc((max(n_total, 1), k), torch.float32)) for _ in range(2)]
        for r, a in enumerate(aggs):
            a._wire([aggs[q]._recv[r] for q in range(world)], flag_ptrs, tables)
        return aggs

    pool = ThreadPoolExecutor(max_workers=4)
    yield make, pool
    pool.shutdown()


@pytest.mark.timeout(600)
@pytest.mark.parametrize("copy_ctas", [-2, 3])
@pytest.mark.parametrize("world,k,halo", [(2, 32, 0.0), (2, 16, 2.0), (3, 32, 0.3), (3, 64, 2.0), (2, 8, 2.0), (4, 32, 0.0)])
def test_peer_kernels_multi_rank_under_the_interpreter(peer_group, world, k, halo, copy_ctas):
    """Fused push all-gather + SpGEMM and pull reduce-scatter of `world` ranks (dense pairs, halo row
    lists, mixed) against the single-process oracle, three steps with fresh inputs (the buffer
    sets alternate by step parity)."""
    from maxk_b200.dist import partition_rows_by_nnz
    make, pool = peer_group
    p, g = _problem(n=96 * world, avg=14.0, k=k, seed=world * 100 + k, empty_frac=0.1)
    bounds = partition_rows_by_nnz(g["indptr"], world, align=16)
    aggs = make(p, g, bounds, k, halo, copy_ctas)
    rng = np.random.default_rng(k)
    for step in range(3):
        x = rng.random((p.n, p.d), dtype=np.float32)
        dy = rng.random((p.n, p.d), dtype=np.float32)
        sd, si = O.topk_cbsr(x, k)
        want_y = O.spgemm_forward(p.w4, p.idx, p.val, sd, si, p.d, n_out=p.n)
        want_g = O.sspmm_backward(p.w4, p.idx, p.val, dy, si)

        def fwd(r):
            lo, hi = bounds[r], bounds[r + 1]
            aggs[r].exchange(_t(sd[lo:hi]), _t(si[lo:hi]))
            return aggs[r].forward().clone()

        ys = list(pool.map(fwd, range(world)))
        for r in range(world):
            close(ys[r], want_y[bounds[r]:bounds[r + 1]], f"forward, step {step}, rank {r}")

        def bwd(r):
            lo, hi = bounds[r], bounds[r + 1]
            par = aggs[r]._fwd_epoch & 1
            return aggs[r].backward(_t(dy[lo:hi]), aggs[r].g_index[par]).clone()

        gs = list(pool.map(bwd, range(world)))
        for r in range(world):
            close(gs[r], want_g[bounds[r]:bounds[r + 1]], f"backward, step {step}, rank {r}")


@pytest.mark.timeout(600)
@pytest.mark.parametrize("world,k,halo", [(2, 32, 0.0), (3, 16, 2.0)])
def test_shard_host_aggregator_over_the_peer_path(peer_group, world, k, halo):
    """ShardHostAggregator.step (host rows in, host rows out; what bench.py times as e2e at N > 1): every rank
    runs its whole step -- TopK in place, push, fused forward, backward, pull reduce-scatter -- in its own
    thread, three steps in a row."""
    from maxk_b200.dist import partition_rows_by_nnz
    from maxk_b200.host import ShardHostAggregator
    make, pool = peer_group
    p, g = _problem(n=96 * world, avg=14.0, k=k, seed=world * 7 + k, empty_frac=0.1)
    bounds = partition_rows_by_nnz(g["indptr"], world, align=16)
    shards = [ShardHostAggregator(a) for a in make(p, g, bounds, k, halo, 2)]
    assert sum(s.h2d_bytes for s in shards) == 2 * p.n * p.d * 4
    assert sum(s.d2h_bytes for s in shards) == p.n * (p.d + k) * 4
    rng = np.random.default_rng(k)
    for step in range(3):
        x = rng.random((p.n, p.d), dtype=np.float32)
        dy = rng.random((p.n, p.d), dtype=np.float32)
        sd, si = O.topk_cbsr(x, k)
        want_y = O.spgemm_forward(p.w4, p.idx, p.val, sd, si, p.d, n_out=p.n)
        want_g = O.sspmm_backward(p.w4, p.idx, p.val, dy, si)

        def run(r):
            lo, hi = bounds[r], bounds[r + 1]
            y, gr = shards[r].step(_t(x[lo:hi]), _t(dy[lo:hi]))
            return y.clone(), gr.clone()

        outs = list(pool.map(run, range(world)))
        for r in range(world):
            close(outs[r][0], want_y[bounds[r]:bounds[r + 1]], f"Y, step {step}, rank {r}")
            close(outs[r][1], want_g[bounds[r]:bounds[r + 1]], f"dXc, step {step}, rank {r}")


@pytest.mark.timeout(600)
@pytest.mark.parametrize("world,k,halo,homes,copy_ctas", [(2, 32, 0.0, 2, 3), (3, 32, 0.3, 2, -2), (2, 64, 2.0, 2, 2),
                                                          (4, 32, 0.0, 2, 1)])
def test_peer_planned_kernels_multi_rank_under_the_interpreter(peer_group, world, k, halo, homes, copy_ctas):
    """The bank plan travels in place of the selectors (fused planned forward; the backward decodes its columns
    from the plan): same oracle, dense / halo / mixed pairs, TMA and load/store push."""
    from maxk_b200.dist import partition_rows_by_nnz
    make, pool = peer_group
    p, g = _problem(n=96 * world, avg=14.0, k=k, seed=world * 10 + k + homes, empty_frac=0.1)
    bounds = partition_rows_by_nnz(g["indptr"], world, align=16)
    aggs = make(p, g, bounds, k, halo, copy_ctas, homes)
    rng = np.random.default_rng(k + homes)
    for step in range(3):
        x = rng.random((p.n, p.d), dtype=np.float32)
        dy = rng.random((p.n, p.d), dtype=np.float32)
        sd, si = O.topk_cbsr(x, k)
        want_y = O.spgemm_forward(p.w4, p.idx, p.val, sd, si, p.d, n_out=p.n)
        want_g = O.sspmm_backward(p.w4, p.idx, p.val, dy, si)

        def fwd(r):
            lo, hi = bounds[r], bounds[r + 1]
            if step == 1:   # TopK straight into the peer-visible buffer, then the plan
                aggs[r].topk_and_exchange(_t(x[lo:hi]))
            else:
                aggs[r].exchange(_t(sd[lo:hi]), _t(si[lo:hi]))
            return aggs[r].forward().clone()

        ys = list(pool.map(fwd, range(world)))
        for r in range(world):
            close(ys[r], want_y[bounds[r]:bounds[r + 1]], f"forward, step {step}, rank {r}")

        def bwd(r):
            lo, hi = bounds[r], bounds[r + 1]
            par = aggs[r]._fwd_epoch & 1
            return aggs[r].backward(_t(dy[lo:hi]), aggs[r].g_index[par]).clone()

        gs = list(pool.map(bwd, range(world)))
        for r in range(world):
            close(gs[r], want_g[bounds[r]:bounds[r + 1]], f"backward, step {step}, rank {r}")


@pytest.mark.timeout(600)
def test_peer_tma_push_of_a_block_larger_than_the_staging_ring(peer_group):
    """One copy CTA pushes 100 KB (values) + 25 KB (selectors) through two 16 KB stages: the stage
    reuse (`cp.async.bulk.wait_group.read`) and the mbarrier phases of many chunks are exercised; the
    interpreter performs a bulk store only when its group is waited for, so a stage that is
    refilled too early pushes the wrong bytes."""
    from maxk_b200.dist import partition_rows_by_nnz
    make, pool = peer_group
    world, k = 2, 64
    p, g = _problem(n=800, avg=4.0, k=k, seed=77, empty_frac=0.0)
    bounds = partition_rows_by_nnz(g["indptr"], world, align=16)
    aggs = make(p, g, bounds, k, 0.0, 1)
    rng = np.random.default_rng(1)
    for step in range(2):
        x = rng.random((p.n, p.d), dtype=np.float32)
        sd, si = O.topk_cbsr(x, k)
        want_y = O.spgemm_forward(p.w4, p.idx, p.val, sd, si, p.d, n_out=p.n)

        def fwd(r):
            lo, hi = bounds[r], bounds[r + 1]
            aggs[r].exchange(_t(sd[lo:hi]), _t(si[lo:hi]))
            return aggs[r].forward().clone()

        ys = list(pool.map(fwd, range(world)))
        for r in range(world):
            close(ys[r], want_y[bounds[r]:bounds[r + 1]], f"forward, step {step}, rank {r}")
            par = aggs[r]._fwd_epoch & 1   # the whole gathered operand arrived intact
            np.testing.assert_array_equal(aggs[r].g_index[par].numpy(), si)
            np.testing.assert_array_equal(aggs[r].g_data[par].numpy(), sd)


