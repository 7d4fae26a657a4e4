"""GPU parity tests: the CUDA path, called through the C ABI (libmaxk_b200.so), against the
CPU oracle on the same seeded inputs, the committed golden vectors, and size-independent
properties at BASELINE.json's full sizes.  Bit-exact for TopK indices/values and the work
list; fp32 rtol 1e-5 / atol 1e-6 against the fp64 oracle for features and gradients."""
import os

import numpy as np
import pytest
import torch

from helpers import ATOL, RTOL, Problem, assert_close, topk_distribution_cases
from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from maxk_b200 import ops as _ops
    return _ops


def _cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


# ------------------------------------------------------------------------------ a4
def test_worklist_golden_and_oracle(ops, golden_dir):
    z = np.load(os.path.join(golden_dir, "worklist_cases.npz"))
    for key in z.files:
        if key.endswith("__indptr"):
            name = key[:-8]
            got = ops.build_worklist(_cu(z[key])).cpu().numpy()
            np.testing.assert_array_equal(got, z[name + "__warp4"], err_msg=name)
    for n, avg, seed, ef in ((1, 3.0, 0, 0.0), (1023, 9.0, 1, 0.2), (1024, 70.0, 2, 0.0),
                             (1025, 200.0, 3, 0.5), (5000, 33.0, 4, 0.1), (40000, 3.0, 5, 0.3)):
        indptr, _ = O.synth_graph(n, avg, seed=seed, empty_frac=ef)
        for nz in (64, 1, 7, 100):
            got = ops.build_worklist(_cu(indptr), nz).cpu().numpy()
            np.testing.assert_array_equal(got, O.worklist(indptr, nz), err_msg=f"n={n} nz={nz}")
    assert ops.build_worklist(_cu(np.zeros(1, np.int32))).numel() == 0
    assert ops.build_worklist(_cu(np.zeros(3000, np.int32))).numel() == 0
    hub = np.array([0, 0, 300000, 300001, 300001], np.int32)
    np.testing.assert_array_equal(ops.build_worklist(_cu(hub)).cpu().numpy(), O.worklist(hub))


# ------------------------------------------------------------------------------ a1
@pytest.mark.parametrize("d", [256, 128, 64, 32, 4])
def test_topk_bit_exact(ops, d):
    rng = np.random.default_rng(d)
    for n in (1, 7, 8, 333):
        x = rng.standard_normal((n, d)).astype(np.float32)
        for k in sorted({1, 2, 8, 16, 18, 32, 64, d} & set(range(1, d + 1))):
            want_v, want_i = O.topk_cbsr(x, k)
            v, i, dense = ops.topk_cbsr(_cu(x), k, want_dense=True)
            np.testing.assert_array_equal(i.cpu().numpy(), want_i, err_msg=f"n={n} d={d} k={k}")
            np.testing.assert_array_equal(v.cpu().numpy().view(np.uint32), want_v.view(np.uint32))
            np.testing.assert_array_equal(dense.cpu().numpy().view(np.uint32),
                                          O.maxk_dense(x, k).view(np.uint32))


def test_topk_ties_and_specials(ops):
    rng = np.random.default_rng(11)
    x = np.zeros((64, 256), np.float32)
    x[0, :] = 1.0
    x[1, 10:20] = 5.0
    x[2, :] = -1.0; x[2, 200] = np.nan; x[2, 7] = np.inf
    x[3, 0] = -0.0
    x[4, :] = -np.inf
    x[5, ::2] = np.nan
    x[6:] = rng.integers(0, 4, size=(58, 256)).astype(np.float32)        # heavy ties
    x[40:] -= 2.0                                                         # negative ties
    for k in (1, 8, 16, 31, 32, 33, 64, 200, 256):
        want_v, want_i = O.topk_cbsr(x, k)
        v, i = ops.topk_cbsr(_cu(x), k)
        np.testing.assert_array_equal(i.cpu().numpy(), want_i, err_msg=f"k={k}")
        np.testing.assert_array_equal(v.cpu().numpy().view(np.uint32), want_v.view(np.uint32))


def test_topk_goldens_from_torch(ops, golden_dir):
    z = np.load(os.path.join(golden_dir, "topk_cases.npz"))
    n = 0
    for key in z.files:
        if key.startswith("idx_d"):
            d, k = key[4:].split("_k")
            v, i = ops.topk_cbsr(_cu(z[f"x_{d}"]), int(k))
            np.testing.assert_array_equal(i.cpu().numpy(), z[key], err_msg=key)
            np.testing.assert_array_equal(v.cpu().numpy(), z[f"val_{d}_k{k}"])
            n += 1
    assert n >= 12
    # live cross-check against torch.topk on the device (tie-free rows)
    x = torch.rand(5000, 256, device="cuda")
    for k in (8, 32, 64):
        v, i = ops.topk_cbsr(x, k)
        tv, ti = torch.topk(x, k, dim=1)
        order = ti.argsort(dim=1)
        assert torch.equal(i.long(), ti.gather(1, order))
        assert torch.equal(v, tv.gather(1, order))


@pytest.mark.parametrize("lanes", [0, -1])
def test_topk_rows_per_warp(ops, lanes):
    """One row per warp and the grouped kernel (d = 128 / 64) on the same rows: oracle bits for ragged n, ties /
    NaN / infinities inside a warp of otherwise fast rows, unaligned output views; torch.topk on 100 k tie-free rows."""
    ops.set_topk_group(lanes)
    try:
        rng = np.random.default_rng(60 + lanes)
        for n in (1, 5, 8, 9, 33, 257, 1001):
            x = rng.standard_normal((n, 256)).astype(np.float32)
            if n > 8:
                x[::5] = rng.integers(0, 3, size=x[::5].shape).astype(np.float32)
                x[6, 17] = np.nan
                x[7, :] = -np.inf
                x[2, 5] = np.inf
                x[3, :] = np.nan
            for k in (1, 8, 18, 32, 64, 128, 129):
                want_v, want_i = O.topk_cbsr(x, k)
                v, i, dense = ops.topk_cbsr(_cu(x), k, want_dense=True)
                np.testing.assert_array_equal(i.cpu().numpy(), want_i, err_msg=f"n={n} k={k} lanes={lanes}")
                np.testing.assert_array_equal(v.cpu().numpy().view(np.uint32), want_v.view(np.uint32))
                np.testing.assert_array_equal(dense.cpu().numpy().view(np.uint32), O.maxk_dense(x, k).view(np.uint32))
                # caller-provided outputs that are not 16-byte aligned (the scalar copy-out)
                buf_v = torch.empty(n * k + 1, dtype=torch.float32, device="cuda")
                buf_i = torch.empty(n * k + 1, dtype=torch.uint8, device="cuda")
                ov, oi = buf_v[1:].view(n, k), buf_i[1:].view(n, k)
                ops.topk_cbsr(_cu(x), k, out=(ov, oi))
                np.testing.assert_array_equal(oi.cpu().numpy(), want_i)
                np.testing.assert_array_equal(ov.cpu().numpy().view(np.uint32), want_v.view(np.uint32))
        for gen in (torch.rand, torch.randn):
            x = gen(100_003, 256, device="cuda")
            for k in (8, 32, 64):
                v, i = ops.topk_cbsr(x, k)
                tv, ti = torch.topk(x, k, dim=1)
                order = ti.argsort(dim=1)
                assert torch.equal(i.long(), ti.gather(1, order))
                assert torch.equal(v, tv.gather(1, order))
        for d, k in ((128, 16), (128, 64), (64, 8), (64, 32)):
            x = torch.randn(50_001, d, device="cuda")
            v, i = ops.topk_cbsr(x, k)
            tv, ti = torch.topk(x, k, dim=1)
            order = ti.argsort(dim=1)
            assert torch.equal(i.long(), ti.gather(1, order))
            assert torch.equal(v, tv.gather(1, order))
    finally:
        ops.set_topk_group(-1)


def test_topk_u8(ops):
    rng = np.random.default_rng(3)
    for d in (256, 64):
        x = rng.integers(0, 256, size=(100, d), dtype=np.uint8)
        x[0] = 0; x[1] = 255
        for k in (1, 8, 32, 64):
            want_v, want_i = O.topk_cbsr(x, k)
            v, i = ops.topk_cbsr_u8(_cu(x), k)
            np.testing.assert_array_equal(i.cpu().numpy(), want_i)
            np.testing.assert_array_equal(v.cpu().numpy(), want_v)


# ------------------------------------------------------------------------------ a2 / a3
CASES = [
    dict(k=32), dict(k=64), dict(k=16), dict(k=8), dict(k=18), dict(k=96), dict(k=128),
    dict(k=1), dict(k=4), dict(k=256), dict(k=160),
    dict(k=32, d=128), dict(k=16, d=64), dict(k=32, selector="torch"), dict(k=64, selector="random"),
    dict(k=32, kind="uniform", avg=300.0, n=200, empty_frac=0.0),       # proteins-like dense rows
    dict(k=32, avg=3.0, n=3000, empty_frac=0.4),                          # products-like sparse rows
    dict(k=32, n=50, avg=2000.0, empty_frac=0.0),                         # hub rows spanning many runs
    dict(k=32, n=400, n_cols=1500, avg=25.0),                             # rectangular row partition
    dict(k=32, warp_max_nz=100), dict(k=16, warp_max_nz=200),             # foreign work lists, len > 64
    dict(k=32, warp_max_nz=5),
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: "-".join(f"{a}{b}" for a, b in c.items()))
def test_forward_backward_vs_oracle(ops, case):
    p = Problem(seed=len(str(case)), **case)
    g = p.cuda()
    y = ops.spgemm_forward(g["warp4"], g["indices"], g["values"], g["sp_data"], g["sp_index"],
                           feat_dim=p.d, num_rows=p.n)
    assert y.shape == (p.n, p.d)
    assert_close(y, p.fwd_oracle(), "forward")
    dxc = ops.sspmm_backward(g["warp4"], g["indices"], g["values"], g["dy"], g["sp_index"])
    assert dxc.shape == (p.n_cols, p.k)
    assert_close(dxc, p.bwd_oracle(), "backward")


@pytest.mark.parametrize("tpw", [1, 2, 3, 7, 64, 1000])
def test_run_length_invariance(ops, tpw):
    """The run length (tasks per warp) only changes which rows are shared between warps."""
    p = Problem(n=300, avg=150.0, k=32, seed=5, empty_frac=0.1)
    g = p.cuda()
    try:
        ops.set_tuning(tpw, tpw)
        y = ops.spgemm_forward(g["warp4"], g["indices"], g["values"], g["sp_data"], g["sp_index"],
                               num_rows=p.n)
        dxc = ops.sspmm_backward(g["warp4"], g["indices"], g["values"], g["dy"], g["sp_index"])
    finally:
        ops.set_tuning(-1, -1)
    assert_close(y, p.fwd_oracle())
    assert_close(dxc, p.bwd_oracle())


@pytest.mark.parametrize("copies", [1, 2, 4, 8, 16])
@pytest.mark.parametrize("k", [32, 64, 16, 8, 4, 2, 1, 18])
def test_forward_accumulator_copies(ops, k, copies):
    """Every (edge width, accumulator copies) variant of the forward kernel gives the same
    result; rows are flushed as plain stores, reductions (shared rows) and in accumulate mode."""
    p = Problem(n=150, avg=120.0, k=k, seed=7 * k + copies, empty_frac=0.1)
    g = p.cuda()
    want = p.fwd_oracle()
    try:
        ops.set_forward_copies(copies)
        for tpw in (16, 2):
            ops.set_tuning(tpw, 0)
            y = ops.spgemm_forward(g["warp4"], g["indices"], g["values"], g["sp_data"], g["sp_index"],
                                   num_rows=p.n)
            assert_close(y, want, f"k={k} copies={copies} tpw={tpw}")
        ops.spgemm_forward(g["warp4"], g["indices"], g["values"], g["sp_data"], g["sp_index"],
                           num_rows=p.n, out=y, accumulate=True)
        assert_close(y, 2 * want, "accumulate")
    finally:
        ops.set_forward_copies(0)
        ops.set_tuning(-1, -1)


@pytest.mark.parametrize("shape", [1, 2, 3])
@pytest.mark.parametrize("k", [32, 16, 64])
def test_forward_launch_shapes(ops, k, shape):
    """8-warp CTAs, 2-warp CTAs and the deeper gather pipeline give the same result, and rows
    without tasks are zero although the output is not memset (poisoned buffer)."""
    p = Problem(n=400, avg=20.0, k=k, seed=31 * k + shape, empty_frac=0.3)
    g = p.cuda()
    want = p.fwd_oracle()
    try:
        ops.set_forward_depth(shape)
        for tpw in (-1, 1, 5):
            ops.set_tuning(tpw, 0)
            y = torch.full((p.n, 256), float("nan"), device="cuda")
            ops.spgemm_forward(g["warp4"], g["indices"], g["values"], g["sp_data"], g["sp_index"],
                               num_rows=p.n, out=y)
            assert_close(y, want, f"k={k} shape={shape} tpw={tpw}")
    finally:
        ops.set_forward_depth(0)
        ops.set_tuning(-1, -1)


@pytest.mark.parametrize("shape", [8, 12])
@pytest.mark.parametrize("k", [32, 16, 64, 8])
def test_forward_ring_shapes(ops, k, shape):
    """cp.async gather ring (the automatic choice on operands larger than the L2), forced on a small problem."""
    p = Problem(n=400, avg=45.0, k=k, seed=13 * k + shape, empty_frac=0.2)
    g = p.cuda()
    want = p.fwd_oracle()
    try:
        ops.set_forward_depth(shape)
        for tpw in (-1, 1, 5):
            ops.set_tuning(tpw, 0)
            y = torch.full((p.n, 256), float("nan"), device="cuda")
            ops.spgemm_forward(g["warp4"], g["indices"], g["values"], g["sp_data"], g["sp_index"],
                               num_rows=p.n, out=y)
            assert_close(y, want, f"k={k} shape={shape} tpw={tpw}")
    finally:
        ops.set_forward_depth(0)
        ops.set_tuning(-1, -1)


@pytest.mark.parametrize("case", [dict(k=32), dict(k=64), dict(k=32, d=128), dict(k=32, selector="random"),
                                  dict(k=32, avg=3.0, n=3000, empty_frac=0.4), dict(k=64, n=50, avg=2000.0, empty_frac=0.0),
                                  dict(k=32, warp_max_nz=100)],
                         ids=lambda c: "-".join(f"{a}{b}" for a, b in c.items()))
def test_forward_through_the_bank_plan(ops, case):
    p = Problem(seed=len(str(case)) + 2, **case)
    g = p.cuda()
    want = p.fwd_oracle()
    try:
        sp_addr = ops.cbsr_bank_plan(g["sp_index"])
        a = sp_addr.long()
        assert int(a.max()) < 512
        col = torch.where(a < 256, a, ((a & 31) << 3) | ((a >> 5) & 7))   # common.cuh: plan_word_column
        assert torch.equal(col, g["sp_index"].long()), "every entry sits in one of the two homes of its column"
        for shape, tpw in ((0, -1), (1, 1), (2, 5)):
            ops.set_forward_depth(shape)
            ops.set_tuning(tpw, 0)
            y = torch.full((p.n, p.d), float("nan"), device="cuda")
            ops.spgemm_forward_planned(g["warp4"], g["indices"], g["values"], g["sp_data"], sp_addr,
                                       feat_dim=p.d, num_rows=p.n, out=y)
            assert_close(y, want, f"shape={shape} tpw={tpw}")
        ops.spgemm_forward_planned(g["warp4"], g["indices"], g["values"], g["sp_data"], sp_addr,
                                   feat_dim=p.d, num_rows=p.n, out=y, accumulate=True)
        assert_close(y, 2 * want, "accumulate")
    finally:
        ops.set_forward_depth(0)
        ops.set_tuning(-1, -1)


def test_bank_plan_spreads_the_banks(ops):
    """fullest bank of the 32 entries of one update: 3.3 on random columns, ~2.1 through the plan"""
    x = torch.rand(4000, 256, device="cuda", generator=torch.Generator("cuda").manual_seed(3))
    _, si = ops.topk_cbsr(x, 32)
    def fullest(words):
        b = (words.long() & 31)
        return torch.zeros(b.shape[0], 32, device="cuda").scatter_add_(1, b, torch.ones_like(b, dtype=torch.float32)).amax(1)
    plain, planned = fullest(si).mean().item(), fullest(ops.cbsr_bank_plan(si)).mean().item()
    assert plain > 3.0 and planned < 2.3, (plain, planned)


def test_forward_auto_takes_the_plan_where_it_pays(ops, monkeypatch):
    p = Problem(n=300, avg=200.0, k=32, seed=5, empty_frac=0.0)
    g = p.cuda()
    assert ops.use_bank_plan(32, g["warp4"].numel() // 4, p.n, p.n)
    calls = []
    real = ops.cbsr_bank_plan
    monkeypatch.setattr(ops, "cbsr_bank_plan", lambda *a, **kw: (calls.append(1), real(*a, **kw))[1])
    y = ops.spgemm_forward_auto(g["warp4"], g["indices"], g["values"], g["sp_data"], g["sp_index"], num_rows=p.n)
    assert calls == [1]
    assert_close(y, p.fwd_oracle(), "auto -> planned")
    q = Problem(n=3000, avg=5.0, k=32, seed=6)
    h = q.cuda()
    y = ops.spgemm_forward_auto(h["warp4"], h["indices"], h["values"], h["sp_data"], h["sp_index"], num_rows=q.n)
    assert calls == [1]
    assert_close(y, q.fwd_oracle(), "auto -> default")


@pytest.mark.parametrize("case", [dict(k=32), dict(k=64), dict(k=16), dict(k=8), dict(k=128), dict(k=18), dict(k=32, d=128),
                                  dict(k=32, selector="random"),
                                  dict(k=32, avg=3.0, n=3000, empty_frac=0.4), dict(k=32, n=50, avg=2000.0, empty_frac=0.0),
                                  dict(k=32, n=400, n_cols=1500, avg=25.0), dict(k=64, warp_max_nz=100)],
                         ids=lambda c: "-".join(f"{a}{b}" for a, b in c.items()))
def test_backward_with_columns_decoded_from_the_plan(ops, case):
    """what the multi-GPU peer path runs when the plan travels in place of the selectors"""
    p = Problem(seed=len(str(case)) + 4, **case)
    g = p.cuda()
    want = p.bwd_oracle()
    try:
        sp_addr = ops.cbsr_bank_plan(g["sp_index"])
        for tpw in (-1, 1, 5):
            ops.set_tuning(0, tpw)
            dxc = torch.full((p.n_cols, p.k), float("nan"), device="cuda")
            ops.sspmm_backward_planned(g["warp4"], g["indices"], g["values"], g["dy"], sp_addr, out=dxc)
            assert_close(dxc, want, f"tpw={tpw}")
        ops.sspmm_backward_planned(g["warp4"], g["indices"], g["values"], g["dy"], sp_addr, out=dxc, accumulate=True)
        assert_close(dxc, 2 * want, "accumulate")
    finally:
        ops.set_tuning(-1, -1)


@pytest.mark.parametrize("case", [dict(k=32), dict(k=64), dict(k=16), dict(k=18), dict(k=32, d=128),
                                  dict(k=32, avg=3.0, n=3000, empty_frac=0.4), dict(k=32, warp_max_nz=100)],
                         ids=lambda c: "-".join(f"{a}{b}" for a, b in c.items()))
def test_backward_run_lengths(ops, case):
    p = Problem(seed=len(str(case)) + 6, **case)
    g = p.cuda()
    try:
        for tpw in (-1, 1, 3):
            ops.set_tuning(0, tpw)
            dxc = torch.full((p.n_cols, p.k), float("nan"), device="cuda")
            ops.sspmm_backward(g["warp4"], g["indices"], g["values"], g["dy"], g["sp_index"], out=dxc)
            assert_close(dxc, p.bwd_oracle(), f"tpw={tpw}")
    finally:
        ops.set_tuning(-1, -1)


def test_topk_distributions_bit_exact(ops):
    """Full-width rows of ten value distributions x every k around the seeds' switch points (16 / 17, 32 / 33):
    indices and value bits equal to the oracle's, whichever way the search ended."""
    for name, x in topk_distribution_cases(3000, seed=23).items():
        xt = torch.from_numpy(x).cuda()
        for k in (1, 2, 3, 8, 16, 17, 24, 31, 32, 33, 64):
            want_v, want_i = O.topk_cbsr(x, k)
            v, i = ops.topk_cbsr(xt, k)
            np.testing.assert_array_equal(i.cpu().numpy(), want_i, err_msg=f"{name} k={k}")
            np.testing.assert_array_equal(v.cpu().numpy().view(np.uint32), want_v.view(np.uint32), err_msg=f"{name} k={k}")


def test_topk_full_width_rows(ops):
    """d = 256 (its own specialisation): bit-exact against the oracle, with tie / NaN / infinity rows (robust path)
    next to ordinary rows, odd row counts, dense output."""
    rng = np.random.default_rng(17)
    try:
        for n in (1, 2, 17, 4099):
            x = rng.standard_normal((n, 256)).astype(np.float32)
            x[::5] = rng.integers(0, 3, size=x[::5].shape).astype(np.float32)
            if n > 4:
                x[4, 17] = np.nan
                x[1, :] = -np.inf
                x[2, 5] = np.inf
            xt = torch.from_numpy(x).cuda()
            for k in (1, 8, 18, 32, 64, 256):
                want_v, want_i = O.topk_cbsr(x, k)
                v, i, dense = ops.topk_cbsr(xt, k, want_dense=True)
                np.testing.assert_array_equal(i.cpu().numpy(), want_i, err_msg=f"n={n} k={k}")
                np.testing.assert_array_equal(v.cpu().numpy().view(np.uint32), want_v.view(np.uint32))
                np.testing.assert_array_equal(dense.cpu().numpy().view(np.uint32), O.maxk_dense(x, k).view(np.uint32))
            for k in (3, 32, 40, 64, 100):   # MaxK.forward only, one k per seed configuration of the full-width kernel
                dense = ops.topk_cbsr(xt, k, want_dense=True, want_cbsr=False)
                np.testing.assert_array_equal(dense.cpu().numpy().view(np.uint32), O.maxk_dense(x, k).view(np.uint32),
                                              err_msg=f"dense only, n={n} k={k}")
    finally:
        pass


@pytest.mark.parametrize("case", [dict(n=5000, d=256, k=32, c=256), dict(n=3333, d=256, k=64, c=132),
                                  dict(n=40000, d=64, k=8, c=4), dict(n=700, d=36, k=18, c=260),
                                  dict(n=300, d=256, k=256, c=128), dict(n=1, d=128, k=1, c=8)],
                         ids=lambda c: "-".join(f"{a}{b}" for a, b in c.items()))
def test_cbsr_linear_products(ops, case):
    """fc_self on the CBSR operand and the gradient of its values against the dense Linear on the scattered
    twin (fp64 oracle); then the autograd Function against torch's own Linear."""
    from helpers import assert_close_sum
    n, d, k, c = case["n"], case["d"], case["k"], case["c"]
    rng = np.random.default_rng(n + d + k + c)
    sd, si = O.topk_cbsr(rng.standard_normal((n, d)).astype(np.float32), k)
    w = (rng.standard_normal((c, d)) * 0.2).astype(np.float32)
    b = rng.standard_normal(c).astype(np.float32)
    m = _cu(np.ascontiguousarray(w.T))
    y_terms = np.abs(O.cbsr_scatter(sd, si, d)).astype(np.float64) @ np.abs(w.T).astype(np.float64)
    y = ops.cbsr_matmul(_cu(sd), _cu(si), m, _cu(b))
    assert_close_sum(y.cpu().numpy(), O.cbsr_linear(sd, si, w, b), y_terms + np.abs(b), k + 1, "matmul + bias")
    ops.cbsr_matmul(_cu(sd), _cu(si), m, out=y, accumulate=True)
    assert_close_sum(y.cpu().numpy(), 2 * O.cbsr_linear(sd, si, w) + b, 2 * y_terms + np.abs(b), k + 2, "accumulate")
    g = rng.standard_normal((n, c)).astype(np.float32)
    g_terms = np.take_along_axis(np.abs(g).astype(np.float64) @ np.abs(w).astype(np.float64), si.astype(np.int64), axis=1)
    dv = ops.cbsr_sampled_dot(_cu(g), m, _cu(si))
    assert_close_sum(dv.cpu().numpy(), O.cbsr_linear_grad_values(g, si, w), g_terms, c, "sampled dot")
    # autograd Function against torch's Linear on the scattered twin
    from maxk_b200.autograd import CBSRLinear
    lin = torch.nn.Linear(d, c).cuda()
    outs = []
    for sparse in (False, True):
        v = _cu(sd).requires_grad_(True)
        lin.zero_grad()
        if sparse:
            out = CBSRLinear.apply(v, _cu(si), lin.weight, lin.bias)
        else:
            out = lin(torch.zeros(n, d, device="cuda").scatter(1, _cu(si).long(), v))
        out.backward(_cu(g))
        outs.append((out.detach(), v.grad.clone(), lin.weight.grad.clone(), lin.bias.grad.clone()))
    for a, b2, what in zip(outs[0], outs[1], ("y", "d sp_data", "dW", "db")):
        torch.testing.assert_close(b2, a, rtol=1e-4, atol=1e-4, msg=lambda s, w2=what: f"{w2}: {s}")
    # GIN's self term added in place: y += (1 + eps) * scatter(sp_data, sp_index)
    y0 = rng.standard_normal((n, d)).astype(np.float32)
    yt = _cu(y0)
    ops.cbsr_axpy_(yt, _cu(sd), _cu(si), 1.0, torch.full((1,), 0.25, device="cuda"))
    assert_close(yt, y0 + 1.25 * O.cbsr_scatter(sd, si, d).astype(np.float64), "axpy")


def test_forward_gap_rows_are_zeroed(ops):
    """Leading, interior and trailing rows without edges (and an output taller than the last
    task row) come back as zeros from a poisoned buffer."""
    n = 97
    deg = np.zeros(n, np.int64)
    deg[[5, 6, 40, 41, 42, 80]] = [3, 200, 1, 70, 64, 129]
    indptr = np.zeros(n + 1, np.int32)
    indptr[1:] = np.cumsum(deg)
    rng = np.random.default_rng(3)
    e = int(indptr[-1])
    indices = rng.integers(0, n, e).astype(np.int32)
    values = rng.random(e, dtype=np.float32)
    x = rng.standard_normal((n, 256)).astype(np.float32)
    sp_data, sp_index = O.topk_cbsr(x, 32)
    want = O.spgemm_forward(O.worklist(indptr), indices, values, sp_data, sp_index, d=256, n_out=n + 10)
    w4 = ops.build_worklist(_cu(indptr))
    for tpw in (-1, 1, 2, 16):
        ops.set_tuning(tpw, 0)
        y = torch.full((n + 10, 256), float("nan"), device="cuda")
        ops.spgemm_forward(w4, _cu(indices), _cu(values), _cu(sp_data), _cu(sp_index), num_rows=n + 10, out=y)
        assert_close(y, want, f"tpw={tpw}")
    ops.set_tuning(-1, -1)


def test_output_is_defined_not_accumulated(ops):
    """results do not depend on what the output buffers held before."""
    p = Problem(n=500, avg=20.0, k=32, seed=8, empty_frac=0.3)
    g = p.cuda()
    out = torch.full((p.n, 256), 7.0, device="cuda")
    ops.spgemm_forward(g["warp4"], g["indices"], g["values"], g["sp_data"], g["sp_index"],
                       num_rows=p.n, out=out)
    assert_close(out, p.fwd_oracle())
    dxc = torch.full((p.n, 32), -3.0, device="cuda")
    ops.sspmm_backward(g["warp4"], g["indices"], g["values"], g["dy"], g["sp_index"], out=dxc)
    assert_close(dxc, p.bwd_oracle())
    ops.sspmm_backward(g["warp4"], g["indices"], g["values"], g["dy"], g["sp_index"], out=dxc,
                       accumulate=True)
    assert_close(dxc, 2 * p.bwd_oracle())


def test_empty_inputs(ops):
    z32 = torch.zeros(0, dtype=torch.int32, device="cuda")
    sp_data = torch.rand(10, 32, device="cuda")
    sp_index = torch.arange(32, dtype=torch.uint8, device="cuda").repeat(10, 1).contiguous()
    y = ops.spgemm_forward(z32, z32, torch.zeros(0, device="cuda"), sp_data, sp_index)
    assert y.shape == (10, 256) and float(y.abs().sum()) == 0.0
    dxc = ops.sspmm_backward(z32, z32, torch.zeros(0, device="cuda"), torch.rand(10, 256, device="cuda"),
                             sp_index)
    assert dxc.shape == (10, 32) and float(dxc.abs().sum()) == 0.0
    v, i = ops.topk_cbsr(torch.zeros(0, 256, device="cuda"), 32)
    assert v.shape == (0, 32)


# ------------------------------------------------------------------------------ a5
def test_layout_kernels_exact(ops):
    for k, d in ((32, 256), (8, 64), (18, 128), (64, 256)):
        p = Problem(n=257, avg=5.0, k=k, d=d, seed=k)
        g = p.cuda()
        dense = ops.cbsr_scatter(g["sp_data"], g["sp_index"], d)
        np.testing.assert_array_equal(dense.cpu().numpy(), O.cbsr_scatter(p.sp_data, p.sp_index, d))
        back = ops.cbsr_gather(dense, g["sp_index"])
        np.testing.assert_array_equal(back.cpu().numpy(), p.sp_data)
        m = ops.cbsr_mask(g["x"], g["sp_index"])
        np.testing.assert_array_equal(m.cpu().numpy(), O.maxk_dense(p.x, k))
        rec = ops.cbsr_pack(g["sp_data"], g["sp_index"])
        assert rec.shape == (257, ops.cbsr_record_bytes(k))
        d2, i2 = ops.cbsr_unpack(rec, 257, k)
        assert torch.equal(d2, g["sp_data"]) and torch.equal(i2, g["sp_index"])


# ------------------------------------------------------------------ boundary / API level
def test_autograd_functions_vs_oracle():
    from maxk_b200 import MaxK, SpGEMMFunction
    p = Problem(n=400, avg=15.0, k=32, seed=21)
    g = p.cuda()
    x = g["x"].clone().requires_grad_(True)
    y = SpGEMMFunction.apply(x, (g["indptr"], g["indices"], g["values"]), 32)
    assert_close(y, p.fwd_oracle())
    y.backward(g["dy"])
    want = np.zeros((p.n, 256))
    np.put_along_axis(want, p.sp_index.astype(np.int64), p.bwd_oracle(), axis=1)
    assert_close(x.grad, want)
    x2 = g["x"].clone().requires_grad_(True)
    m = MaxK.apply(x2, 32)
    np.testing.assert_array_equal(m.detach().cpu().numpy(), O.maxk_dense(p.x, 32))
    m.backward(g["dy"])
    mask = O.cbsr_scatter(np.ones_like(p.sp_data), p.sp_index, 256)
    np.testing.assert_array_equal(x2.grad.cpu().numpy(), p.dy * mask)


def test_dropin_modules():
    import maxk_cuda_kernels as mk
    import spmm_kernels as sk
    from direct_kernel_interface import DirectMaxKKernels
    p = Problem(n=500, avg=40.0, k=32, seed=33)
    g = p.cuda()
    w4 = mk.generate_warp4_metadata(g["indptr"])
    np.testing.assert_array_equal(w4.cpu().numpy(), p.w4)
    nw = w4.numel() // 4
    v, i = mk.cuda_topk_maxk_float(g["x"], 32)
    assert i.dtype == torch.int32 and v.dtype == torch.float32
    np.testing.assert_array_equal(i.cpu().numpy().astype(np.uint8), p.sp_index)
    y = mk.spmm_maxk_forward(w4, g["indices"], g["values"], v, i.to(torch.uint8), nw, 32)
    assert y.shape == (500, 256)
    assert_close(y, p.fwd_oracle())
    gi = mk.spmm_maxk_backward(w4, g["indices"], g["values"], g["dy"], i.to(torch.uint8), nw, 32)
    assert_close(gi, p.bwd_oracle())
    assert mk.validate_spmm_maxk(w4, g["indices"], g["values"], v, i.to(torch.uint8),
                                 torch.from_numpy(p.fwd_oracle().astype(np.float32)).cuda(), nw, 32)
    times = mk.benchmark_spmm_maxk(w4, g["indices"], g["values"], v, i.to(torch.uint8), nw, 32, num_runs=2)
    assert len(times) == 2 and all(t > 0 for t in times)
    sel = mk.generate_sparse_selector(100, 256, 32)
    assert sel.shape == (100, 32) and sel.dtype == torch.uint8
    assert all(len(set(r.tolist())) == 32 for r in sel.cpu())
    # DirectMaxKKernels: validation against the dense-X CSR SpMM comparator, both TopK flavours
    dk = DirectMaxKKernels("synthetic")
    assert dk.load_warp4_metadata(indptr=g["indptr"]) and dk.num_warps == nw
    assert dk.validate_against_cusparse(g, g["x"], 32, use_cuda_topk=True)
    assert dk.validate_against_cusparse(g, g["x"], 32, use_cuda_topk=False)
    out, ms = dk.run_forward_kernel(g, g["x"], 32, timing=True)
    assert out.shape == (500, 256) and ms > 0
    assert_close(out, p.fwd_oracle())
    gin, ms = dk.run_backward_kernel(g, g["dy"], 32, timing=False, sparse_selector=g["sp_index"])
    assert_close(gin, p.bwd_oracle())
    # Path-A classes: caller-allocated outputs, selector passed as int32
    sd, ssel = sk.prepare_cbsr_format(g["x"], 32)
    assert ssel.dtype == torch.int32
    out = torch.zeros(500, 256, device="cuda")
    kern = sk.SpmmMaxK("graph", g["indptr"], g["indices"], g["values"], sd, out)
    kern.set_sparse_params(ssel, 32)
    assert kern.run_kernel(False, 256) == 0.0
    assert_close(out, p.fwd_oracle())
    gs = torch.zeros(500, 32, device="cuda")
    kb = sk.SpmmMaxKBackward("graph", g["indptr"], g["indices"], g["values"], g["dy"], gs)
    kb.set_sparse_params(ssel, 32)
    assert kb.run_kernel(True, 256) > 0.0
    assert_close(gs, p.bwd_oracle())
    np.testing.assert_array_equal(sk.topk_nonlinearity(g["x"], 32).cpu().numpy(), O.maxk_dense(p.x, 32))


@pytest.mark.parametrize("chunks", [1, 3, 4])
def test_host_aggregator_end_to_end(chunks):
    """Host buffers in, host buffers out (what bench.py's e2e times), chunked and overlapped."""
    from maxk_b200 import HostAggregator
    p = Problem(n=900, avg=60.0, k=32, seed=77, empty_frac=0.1)
    g = p.cuda()
    host = HostAggregator(g["indptr"], g["indices"], g["values"], 32, 256, "cuda", chunks=chunks)
    xh = torch.from_numpy(p.x).pin_memory()
    dyh = torch.from_numpy(p.dy).pin_memory()
    for _ in range(2):  # buffers are reused across steps
        y, dxc = host.step(xh, dyh)
        assert not y.is_cuda and not dxc.is_cuda
        np.testing.assert_allclose(y.numpy().astype(np.float64), p.fwd_oracle(), rtol=RTOL, atol=ATOL)
        np.testing.assert_allclose(dxc.numpy().astype(np.float64), p.bwd_oracle(), rtol=RTOL, atol=ATOL)
    assert host.h2d_bytes == 2 * 900 * 256 * 4 and host.d2h_bytes == 900 * 256 * 4 + 900 * 32 * 4


def test_host_aggregator_pipelined_submit_result():
    """submit/result keeps two steps in flight (uploads of step i+1 behind the kernels of step
    i); every step sees its own inputs and tickets expire when their buffers are reused."""
    from maxk_b200 import HostAggregator
    p = Problem(n=700, avg=40.0, k=32, seed=78, empty_frac=0.1)
    g = p.cuda()
    host = HostAggregator(g["indptr"], g["indices"], g["values"], 32, 256, "cuda", chunks=3)
    scales = [1.0, -2.0, 0.5, 3.0, -1.0]
    xs = [torch.from_numpy(p.x * s).pin_memory() for s in scales]
    dys = [torch.from_numpy(p.dy * (s + 1.5)).pin_memory() for s in scales]
    prev = None
    for i, s in enumerate(scales):
        t = host.submit(xs[i], dys[i])
        if prev is not None:
            j, tj = prev
            y, dxc = host.result(tj)
            pj = Problem(n=700, avg=40.0, k=32, seed=78, empty_frac=0.1)
            pj.x = p.x * scales[j]
            pj.dy = p.dy * (scales[j] + 1.5)
            pj.sp_data, pj.sp_index = O.topk_cbsr(pj.x, 32)
            np.testing.assert_allclose(y.numpy().astype(np.float64), pj.fwd_oracle(), rtol=RTOL, atol=ATOL)
            np.testing.assert_allclose(dxc.numpy().astype(np.float64), pj.bwd_oracle(), rtol=RTOL, atol=ATOL)
        prev = (i, t)
    host.result(prev[1])
    with pytest.raises(RuntimeError):
        host.result(0)


@pytest.mark.parametrize("k,copy_ctas", [(32, 3), (16, -2), (64, 1), (8, 2)])
def test_peer_memory_path_single_rank(ops, k, copy_ctas):
    """The fused (push all-gather + SpGEMM) kernel and the pull reduce-scatter of maxk_b200.peer
    on a one-rank job: IPC allocation, [own columns | other columns] task order, accumulate
    flushes, epoch-parity buffers.  The multi-rank exchange itself is checked by
    scripts/peer_check.py under torchrun (2 and 8 B200s, profiles/r01_peer_memory.md)."""
    import socket
    import torch.distributed as dist
    from maxk_b200.peer import PeerAggregator
    p = Problem(n=500, avg=70.0, k=k, seed=90 + k, empty_frac=0.1)
    g = p.cuda()
    own_pg = not dist.is_initialized()
    if own_pg:
        with socket.socket() as sk_:
            sk_.bind(("127.0.0.1", 0))
            port = sk_.getsockname()[1]
        dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=0, world_size=1)
    try:
        agg = PeerAggregator(g["indptr"], g["indices"], g["values"], [0, p.n], 0, k, 256,
                             copy_ctas=copy_ctas)
        for step in range(3):  # both buffer parities
            gd, gi = agg.exchange(g["sp_data"], g["sp_index"]) if step else agg.topk_and_exchange(g["x"])
            y = agg.forward(gd, gi)
            dxc = agg.backward(g["dy"], gi)
            assert_close(y, p.fwd_oracle(), f"fused forward step {step}")
            assert_close(dxc, p.bwd_oracle(), f"pull reduce-scatter step {step}")
        agg.close()
    finally:
        if own_pg:
            dist.destroy_process_group()


def test_reference_named_shims():
    """The reference's own extern "C" launch shims (cuda_kernel_wrappers.cu:38-91), same argument
    lists (dim3 by value), resolved from libmaxk_b200.so."""
    import ctypes
    from maxk_b200 import _cabi
    lib = _cabi.load()

    class Dim3(ctypes.Structure):
        _fields_ = [("x", ctypes.c_uint), ("y", ctypes.c_uint), ("z", ctypes.c_uint)]

    p = Problem(n=300, avg=30.0, k=32, seed=2)
    g = p.cuda()
    vp, ci = ctypes.c_void_p, ctypes.c_int
    f = lib.spmm_kernel_opt2_sparse_v3_wrapper
    f.restype = None
    f.argtypes = [vp, vp, vp, vp, vp, vp, ci, ci, ci, ci, ci, Dim3, Dim3, ci]
    out = torch.zeros(300, 256, device="cuda")
    torch.cuda.synchronize()
    nw = p.w4.shape[0] // 4
    f(g["warp4"].data_ptr(), g["indices"].data_ptr(), g["values"].data_ptr(), g["sp_data"].data_ptr(),
      g["sp_index"].data_ptr(), out.data_ptr(), 300, p.idx.shape[0], 256, 32, nw,
      Dim3((nw + 11) // 12, 1, 1), Dim3(384, 1, 1), 12 * 256 * 4)
    torch.cuda.synchronize()
    assert_close(out, p.fwd_oracle())
    b = lib.spmm_kernel_opt2_sparse_backward_v3_wrapper
    b.restype = None
    b.argtypes = f.argtypes
    gin = torch.zeros(300, 32, device="cuda")
    torch.cuda.synchronize()
    b(g["warp4"].data_ptr(), g["indices"].data_ptr(), g["values"].data_ptr(), g["dy"].data_ptr(),
      g["sp_index"].data_ptr(), gin.data_ptr(), 300, p.idx.shape[0], 256, 32, nw,
      Dim3((nw + 11) // 12, 1, 1), Dim3(384, 1, 1), 12 * 256 * 4)
    torch.cuda.synchronize()
    assert_close(gin, p.bwd_oracle())
    t = lib.topk_kernel_wrapper
    t.restype = None
    t.argtypes = [vp, vp, vp, ctypes.c_uint, ci, ci, Dim3, Dim3, ci]
    xu = torch.randint(0, 256, (64, 256), dtype=torch.uint8, device="cuda")
    val = torch.zeros(64, 32, dtype=torch.uint8, device="cuda")
    ind = torch.zeros(64, 32, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    t(xu.data_ptr(), val.data_ptr(), ind.data_ptr(), 32, 64, 256, Dim3(4, 1, 1), Dim3(512, 1, 1), 0)
    torch.cuda.synchronize()
    wv, wi = O.topk_cbsr(xu.cpu().numpy(), 32)
    np.testing.assert_array_equal(ind.cpu().numpy(), wi)
    np.testing.assert_array_equal(val.cpu().numpy(), wv)


# ------------------------------------------------ full-size, size-independent properties
@pytest.mark.parametrize("shape,k", [("reddit", 32), ("products", 32), ("proteins", 64)])
def test_full_size_properties(ops, shape, k):
    """BASELINE.json shapes at full size.  With values = 1, sp_data = 1 and selector columns
    0..k-1, Y[r, c<k] = deg(r) and dXc[j, :] (dY = 1) = in-degree(j): integers, exact in fp32
    below 2^24.  Then linearity + a float64 checksum on random data."""
    from maxk_b200 import graphs
    g = graphs.make_graph(shape, seed=0, device="cuda", values="ones")
    n, e = g["v_num"], g["e_num"]
    w4 = ops.build_worklist(g["indptr"])
    deg = (g["indptr"][1:] - g["indptr"][:-1]).float()
    assert int(w4.view(-1, 4)[:, 2].sum()) == e
    assert w4.numel() // 4 == int(((deg.long() + 63) // 64).sum())
    sp_index = torch.arange(k, dtype=torch.uint8, device="cuda").repeat(n, 1).contiguous()
    ones = torch.ones(n, k, device="cuda")
    y = ops.spgemm_forward(w4, g["indices"], g["values"], ones, sp_index)
    assert torch.equal(y[:, :k], deg[:, None].expand(n, k))
    assert float(y[:, k:].abs().max()) == 0.0
    del y
    dy = torch.ones(n, 256, device="cuda")
    dxc = ops.sspmm_backward(w4, g["indices"], g["values"], dy, sp_index)
    indeg = torch.bincount(g["indices"].long(), minlength=n).float()
    assert torch.equal(dxc, indeg[:, None].expand(n, k))
    del dxc, dy
    # random data: checksum of checksums in float64, and linearity f(2x) = 2 f(x) bit-exactly
    x = torch.rand(n, 256, device="cuda", generator=torch.Generator("cuda").manual_seed(123))
    sp_data, sp_index = ops.topk_cbsr(x, k)
    assert bool((sp_index[:, 1:] > sp_index[:, :-1]).all())          # ascending, distinct
    kth = torch.topk(x, k, dim=1).values[:, -1]
    assert torch.equal(sp_data.min(dim=1).values, kth)
    val = torch.rand(e, device="cuda", generator=torch.Generator("cuda").manual_seed(7))
    y = ops.spgemm_forward(w4, g["indices"], val, sp_data, sp_index)
    rowsum = sp_data.double().sum(dim=1)
    want = float((val.double() * rowsum[g["indices"].long()]).sum())
    got = float(y.double().sum())
    assert abs(got - want) <= 1e-6 * abs(want)
    y2 = ops.spgemm_forward(w4, g["indices"], val, sp_data * 2, sp_index)
    # same summation order is not guaranteed on shared rows (vector reductions), so compare
    # in tolerance there and bit-exactly in aggregate scale
    assert torch.allclose(y2, 2 * y, rtol=RTOL, atol=ATOL)
