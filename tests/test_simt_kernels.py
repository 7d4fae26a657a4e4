"""CPU-side logic and memory-safety tests of the CUDA kernel SOURCES (no GPU needed).

tests/simt/ compiles spgemm-new_b200/csrc/*.cu for the host and runs every CUDA thread as a
fibre (warp collectives and barriers are rendezvous points; lanes are NOT in lock step between
them, the weakest schedule CUDA allows).  Inputs and outputs sit between guard pages.  The cases
mirror tests/test_gpu_parity.py -- same oracle, same tolerances -- at sizes the interpreter
finishes in seconds.  This is a check of indexing, masks, barriers, tails and buffer bounds; the
parity tests proper are the `-m gpu` ones, which run the real library on a B200.
"""
import os

import numpy as np
import pytest

import simt_lib as S
from helpers import ATOL, RTOL, Problem, topk_distribution_cases
from oracle import oracle as O


def close(got, want, what=""):
    np.testing.assert_allclose(got.astype(np.float64), want, rtol=RTOL, atol=ATOL, err_msg=what)


@pytest.fixture(autouse=True)
def _reset_knobs():
    yield
    S.set_tuning(-1, -1)
    S.set_forward_depth(0)
    S.set_forward_copies(0)
    S.set_topk_group(-1)
    S.set_cp_async_eager(False)


# ------------------------------------------------------------------------------ a4
def test_simt_worklist(golden_dir):
    z = np.load(os.path.join(golden_dir, "worklist_cases.npz"))
    for key in z.files:
        if key.endswith("__indptr"):
            name = key[:-8]
            np.testing.assert_array_equal(S.worklist(z[key]), np.asarray(z[name + "__warp4"]).reshape(-1), err_msg=name)
    for n, avg, seed, ef in ((1, 3.0, 0, 0.0), (1023, 9.0, 1, 0.2), (1024, 70.0, 2, 0.0), (1025, 200.0, 3, 0.5),
                             (3000, 33.0, 4, 0.1)):
        indptr, _ = O.synth_graph(n, avg, seed=seed, empty_frac=ef)
        for nz in (64, 1, 7, 100):
            np.testing.assert_array_equal(S.worklist(indptr, nz), O.worklist(indptr, nz).reshape(-1),
                                          err_msg=f"n={n} nz={nz}")
    assert S.worklist(np.zeros(1, np.int32)).size == 0
    assert S.worklist(np.zeros(3000, np.int32)).size == 0
    hub = np.array([0, 0, 300000, 300001, 300001], np.int32)
    np.testing.assert_array_equal(S.worklist(hub), O.worklist(hub).reshape(-1))


# ------------------------------------------------------------------------------ a1
@pytest.mark.parametrize("d", [256, 128, 64, 32, 4])
def test_simt_topk_bit_exact(d):
    rng = np.random.default_rng(d)
    for n in (1, 7, 8, 61):
        x = rng.standard_normal((n, d)).astype(np.float32)
        for k in sorted({1, 2, 8, 16, 18, 32, 64, d} & set(range(1, d + 1))):
            want_v, want_i = O.topk_cbsr(x, k)
            v, i, dense = S.topk_f32(x, k, want_dense=True)
            np.testing.assert_array_equal(i, want_i, err_msg=f"n={n} d={d} k={k}")
            np.testing.assert_array_equal(v.view(np.uint32), want_v.view(np.uint32))
            np.testing.assert_array_equal(dense.view(np.uint32), O.maxk_dense(x, k).view(np.uint32))
    (dense,) = S.topk_f32(x, min(3, d), want_dense=True, want_cbsr=False)  # MaxK.forward only
    np.testing.assert_array_equal(dense.view(np.uint32), O.maxk_dense(x, min(3, d)).view(np.uint32))


def test_simt_topk_ties_and_specials():
    rng = np.random.default_rng(11)
    x = np.zeros((64, 256), np.float32)
    x[0, :] = 1.0
    x[1, 10:20] = 5.0
    x[2, :] = -1.0; x[2, 200] = np.nan; x[2, 7] = np.inf
    x[3, 0] = -0.0
    x[4, :] = -np.inf
    x[5, ::2] = np.nan
    x[6:] = rng.integers(0, 4, size=(58, 256)).astype(np.float32)
    x[40:] -= 2.0
    for k in (1, 8, 16, 31, 32, 33, 64, 200, 256):
        want_v, want_i = O.topk_cbsr(x, k)
        v, i = S.topk_f32(x, k)
        np.testing.assert_array_equal(i, want_i, err_msg=f"k={k}")
        np.testing.assert_array_equal(v.view(np.uint32), want_v.view(np.uint32))


def test_simt_topk_distributions_bit_exact():
    """Mirror of test_gpu_parity.test_topk_distributions_bit_exact (fewer rows): order-statistics seed, extraction
    finish and the hand-over to the robust search on ten value distributions."""
    for name, x in topk_distribution_cases(40, seed=23).items():
        for k in (1, 3, 16, 17, 32, 33, 64):
            want_v, want_i = O.topk_cbsr(x, k)
            v, i = S.topk_f32(x, k)
            np.testing.assert_array_equal(i, want_i, err_msg=f"{name} k={k}")
            np.testing.assert_array_equal(v.view(np.uint32), want_v.view(np.uint32), err_msg=f"{name} k={k}")


def test_simt_topk_full_width_rows():
    """d = 256 (its own specialisation): the oracle's bits for every n around the CTA boundaries, rows the fast path
    finishes next to rows that need the robust path (ties, NaN, infinities), the dense MaxK output included."""
    rng = np.random.default_rng(5)
    for n in (1, 2, 3, 15, 16, 17, 33):
        x = rng.standard_normal((n, 256)).astype(np.float32)
        x[::3] = rng.integers(0, 3, size=x[::3].shape).astype(np.float32)   # every third row: ties everywhere
        if n > 4:
            x[4, 17] = np.nan
            x[1, :] = -np.inf
            x[2, 5] = np.inf
        for k in (1, 8, 18, 32, 64, 255, 256):
            want_v, want_i = O.topk_cbsr(x, k)
            v, i, dense = S.topk_f32(x, k, want_dense=True)
            np.testing.assert_array_equal(i, want_i, err_msg=f"n={n} k={k}")
            np.testing.assert_array_equal(v.view(np.uint32), want_v.view(np.uint32))
            np.testing.assert_array_equal(dense.view(np.uint32), O.maxk_dense(x, k).view(np.uint32))
        for kk in (32, 40, 64, 100):   # MaxK.forward only, one k per seed configuration of the full-width kernel
            (dense,) = S.topk_f32(x, kk, want_dense=True, want_cbsr=False)
            np.testing.assert_array_equal(dense.view(np.uint32), O.maxk_dense(x, kk).view(np.uint32), err_msg=f"k={kk}")


@pytest.mark.parametrize("lanes", [0, -1])
def test_simt_topk_rows_per_warp(lanes):
    """The grouped kernel (d = 128 / 64: 4 / 2 lanes per row) and the one-row kernel on the same rows: every
    n around the warp and CTA boundaries, rows the fast path finishes next to rows that need the robust search
    (ties, NaN, infinities) inside one warp, both output copies (128-bit when sp_data / sp_index are aligned)."""
    S.set_topk_group(lanes)
    rng = np.random.default_rng(50 + lanes)
    for eager in (False, True):
        S.set_cp_async_eager(eager)
        for n in (1, 3, 4, 5, 8, 9, 31, 32, 33, 65):
            x = rng.standard_normal((n, 256)).astype(np.float32)
            if n > 8:
                x[::5] = rng.integers(0, 3, size=x[::5].shape).astype(np.float32)   # ties everywhere
                x[6, 17] = np.nan
                x[7, :] = -np.inf
                x[2, 5] = np.inf
                x[3, :] = np.nan
            for k in (1, 8, 18, 32, 64, 128, 129):
                want_v, want_i = O.topk_cbsr(x, k)
                for al in ((4, 1), (16, 4)):
                    v, i, dense = S.topk_f32(x, k, want_dense=True, out_align=al)
                    np.testing.assert_array_equal(i, want_i, err_msg=f"n={n} k={k} lanes={lanes}")
                    np.testing.assert_array_equal(v.view(np.uint32), want_v.view(np.uint32))
                    np.testing.assert_array_equal(dense.view(np.uint32), O.maxk_dense(x, k).view(np.uint32))
    for d, k in ((128, 16), (128, 64), (64, 8), (64, 32)):
        x = rng.uniform(-1, 1, size=(37, d)).astype(np.float32)
        x[5, :] = 0.25
        want_v, want_i = O.topk_cbsr(x, k)
        v, i = S.topk_f32(x, k)
        np.testing.assert_array_equal(i, want_i, err_msg=f"d={d} k={k} lanes={lanes}")
        np.testing.assert_array_equal(v.view(np.uint32), want_v.view(np.uint32))


def test_simt_topk_goldens_from_torch(golden_dir):
    z = np.load(os.path.join(golden_dir, "topk_cases.npz"))
    n = 0
    for key in z.files:
        if key.startswith("idx_d"):
            d, k = key[4:].split("_k")
            v, i = S.topk_f32(z[f"x_{d}"], int(k))
            np.testing.assert_array_equal(i, z[key], err_msg=key)
            np.testing.assert_array_equal(v, z[f"val_{d}_k{k}"])
            n += 1
    assert n >= 12


def test_simt_topk_u8():
    rng = np.random.default_rng(3)
    for d in (256, 64):
        x = rng.integers(0, 256, size=(40, d), dtype=np.uint8)
        x[0] = 0; x[1] = 255
        for k in (1, 8, 32, 64):
            want_v, want_i = O.topk_cbsr(x, k)
            v, i = S.topk_u8(x, k)
            np.testing.assert_array_equal(i, want_i)
            np.testing.assert_array_equal(v, want_v)


# ------------------------------------------------------------------------------ a2 / a3
CASES = [
    dict(k=32), dict(k=64), dict(k=16), dict(k=8), dict(k=18), dict(k=96), dict(k=128),
    dict(k=1), dict(k=4), dict(k=256), dict(k=160),
    dict(k=32, d=128), dict(k=16, d=64), dict(k=32, selector="torch"), dict(k=64, selector="random"),
    dict(k=32, kind="uniform", avg=300.0, n=40, empty_frac=0.0),
    dict(k=32, avg=3.0, n=1500, empty_frac=0.4),
    dict(k=32, n=20, avg=1500.0, empty_frac=0.0),
    dict(k=32, n=200, n_cols=700, avg=25.0),
    dict(k=32, warp_max_nz=100), dict(k=16, warp_max_nz=200),
    dict(k=32, warp_max_nz=5),
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: "-".join(f"{a}{b}" for a, b in c.items()))
def test_simt_forward_backward_vs_oracle(case):
    case = dict(n=250, **case) if "n" not in case else case
    p = Problem(seed=len(str(case)), **case)
    y = S.spgemm_forward(p.w4, p.idx, p.val, p.sp_data, p.sp_index, p.n, d=p.d)
    close(y, p.fwd_oracle(), "forward")
    dxc = S.sspmm_backward(p.w4, p.idx, p.val, p.dy, p.sp_index)
    assert dxc.shape == (p.n_cols, p.k)
    close(dxc, p.bwd_oracle(), "backward")


@pytest.mark.parametrize("case", [dict(k=32), dict(k=64), dict(k=16), dict(k=18), dict(k=8, d=64), dict(k=32, d=128),
                                  dict(k=32, d=36), dict(k=32, avg=3.0, n=1500, empty_frac=0.4),
                                  dict(k=32, n=20, avg=1500.0, empty_frac=0.0), dict(k=32, warp_max_nz=100),
                                  dict(k=32, warp_max_nz=5), dict(k=32, n=200, n_cols=700, avg=25.0)],
                         ids=lambda c: "-".join(f"{a}{b}" for a, b in c.items()))
def test_simt_backward_more_shapes(case):
    """The backward loads the next dY row into registers one task early: ragged widths, hub rows, foreign work lists."""
    case = dict(n=250, **case) if "n" not in case else case
    p = Problem(seed=len(str(case)) + 9, **case)
    for tpw in (-1, 1, 3):
        S.set_tuning(0, tpw)
        close(S.sspmm_backward(p.w4, p.idx, p.val, p.dy, p.sp_index), p.bwd_oracle(), f"tpw={tpw}")
    g2 = S.sspmm_backward(p.w4, p.idx, p.val, p.dy, p.sp_index, accumulate=True,
                          out_init=np.ones((p.n_cols, p.k), np.float32))
    close(g2, p.bwd_oracle() + 1.0, "accumulate")


@pytest.mark.parametrize("tpw", [1, 2, 3, 7, 64, 1000])
def test_simt_run_length_invariance(tpw):
    p = Problem(n=120, avg=150.0, k=32, seed=5, empty_frac=0.1)
    S.set_tuning(tpw, tpw)
    close(S.spgemm_forward(p.w4, p.idx, p.val, p.sp_data, p.sp_index, p.n), p.fwd_oracle())
    close(S.sspmm_backward(p.w4, p.idx, p.val, p.dy, p.sp_index), p.bwd_oracle())


@pytest.mark.parametrize("copies", [1, 2, 4, 8, 16])
@pytest.mark.parametrize("k", [32, 64, 16, 8, 4, 2, 1, 18])
def test_simt_forward_accumulator_copies(k, copies):
    p = Problem(n=60, avg=100.0, k=k, seed=7 * k + copies, empty_frac=0.1)
    want = p.fwd_oracle()
    S.set_forward_copies(copies)
    for tpw in (16, 2):
        S.set_tuning(tpw, 0)
        y = S.spgemm_forward(p.w4, p.idx, p.val, p.sp_data, p.sp_index, p.n)
        close(y, want, f"k={k} copies={copies} tpw={tpw}")
    y2 = S.spgemm_forward(p.w4, p.idx, p.val, p.sp_data, p.sp_index, p.n, accumulate=True, out_init=y)
    close(y2, 2 * want, "accumulate")


@pytest.mark.parametrize("shape", [1, 2, 3, 8, 12])
@pytest.mark.parametrize("k", [32, 16, 64, 8, 128])
def test_simt_forward_launch_shapes(k, shape):
    """All launch shapes (8, 12: cp.async gather ring -- the interpreter performs an asynchronous copy as
    late as PTX allows, at the issuing thread's wait), from a poisoned output buffer."""
    p = Problem(n=200, avg=35.0, k=k, seed=31 * k + shape, empty_frac=0.3)
    want = p.fwd_oracle()
    S.set_forward_depth(shape)
    for tpw in (-1, 1, 5):
        S.set_tuning(tpw, 0)
        close(S.spgemm_forward(p.w4, p.idx, p.val, p.sp_data, p.sp_index, p.n), want, f"k={k} shape={shape} tpw={tpw}")


@pytest.mark.parametrize("shape", [0, 1, 2, 8, 12])
def test_simt_forward_gap_rows_are_zeroed(shape):
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
    w4 = O.worklist(indptr)
    want = O.spgemm_forward(w4, indices, values, sp_data, sp_index, d=256, n_out=n + 10)
    S.set_forward_depth(shape)
    for tpw in (-1, 1, 2, 16):
        S.set_tuning(tpw, 0)
        close(S.spgemm_forward(w4, indices, values, sp_data, sp_index, n + 10), want, f"tpw={tpw}")


def test_simt_output_is_defined_not_accumulated():
    p = Problem(n=200, avg=20.0, k=32, seed=8, empty_frac=0.3)
    y = S.spgemm_forward(p.w4, p.idx, p.val, p.sp_data, p.sp_index, p.n, out_init=np.full((p.n, 256), 7.0, np.float32))
    close(y, p.fwd_oracle())
    dxc = S.sspmm_backward(p.w4, p.idx, p.val, p.dy, p.sp_index, out_init=np.full((p.n, 32), -3.0, np.float32))
    close(dxc, p.bwd_oracle())
    dxc2 = S.sspmm_backward(p.w4, p.idx, p.val, p.dy, p.sp_index, out_init=dxc, accumulate=True)
    close(dxc2, 2 * p.bwd_oracle())


def test_simt_unaligned_operands_take_the_scalar_paths():
    """sp_data not 16-byte / sp_index not 4-byte / dxc not 16-byte aligned: the library must fall
    back to its scalar variants instead of issuing misaligned vector accesses."""
    for k in (32, 64, 16, 20):
        p = Problem(n=90, avg=40.0, k=k, seed=k)
        close(S.spgemm_forward(p.w4, p.idx, p.val, p.sp_data, p.sp_index, p.n, data_align=4, index_align=1),
              p.fwd_oracle(), f"forward k={k}")
        close(S.sspmm_backward(p.w4, p.idx, p.val, p.dy, p.sp_index, dxc_align=4, index_align=1),
              p.bwd_oracle(), f"backward k={k}")


@pytest.mark.parametrize("eager", [False, True])
@pytest.mark.parametrize("shape", [8, 12])
def test_simt_ring_shapes_on_hard_cases(shape, eager):
    """cp.async ring variants: tasks longer than one ring (hub rows), foreign work lists with
    len > 64, single-edge rows, accumulate mode, and unaligned operands (falls back); with the
    asynchronous copies landing as late and as early as PTX allows."""
    S.set_forward_depth(shape)
    S.set_cp_async_eager(eager)
    for case in (dict(k=32, n=20, avg=1500.0, empty_frac=0.0), dict(k=16, warp_max_nz=200), dict(k=32, warp_max_nz=5),
                 dict(k=64, avg=3.0, n=700, empty_frac=0.4), dict(k=32, n=200, n_cols=700, avg=25.0),
                 dict(k=32, selector="random"), dict(k=64, d=128)):
        case = dict(n=150, **case) if "n" not in case else case
        p = Problem(seed=len(str(case)) + shape, **case)
        want = p.fwd_oracle()
        y = S.spgemm_forward(p.w4, p.idx, p.val, p.sp_data, p.sp_index, p.n, d=p.d)
        close(y, want, str(case))
        close(S.spgemm_forward(p.w4, p.idx, p.val, p.sp_data, p.sp_index, p.n, d=p.d, accumulate=True, out_init=y),
              2 * want, "accumulate " + str(case))
        close(S.spgemm_forward(p.w4, p.idx, p.val, p.sp_data, p.sp_index, p.n, d=p.d, data_align=4, index_align=1),
              want, "unaligned " + str(case))


# ------------------------------------------------------- forward through the bank plan
def _home_word(c, h):   # csrc/common.cuh: plan_home_word
    c = c.astype(np.int64)
    return c if h == 0 else 256 + (c % 8) * 32 + c // 8


def _word_column(w):    # csrc/common.cuh: plan_word_column
    w = w.astype(np.int64)
    return np.where(w < 256, w, ((w & 31) << 3) | ((w >> 5) & 7))


def _check_plan(sp_index, sp_addr):
    ok = (sp_addr == _home_word(sp_index, 0)) | (sp_addr == _home_word(sp_index, 1))
    assert ok.all(), "every entry sits in one of its two homes"
    for h in range(2):  # each home is a bijection of the 256 columns
        assert len(set(_home_word(np.arange(256), h).tolist())) == 256
    np.testing.assert_array_equal(_word_column(sp_addr), sp_index.astype(np.int64))


def _fullest_bank(addr, group, banks):
    """fullest bank of every group of `group` consecutive entries: [n, k/group]"""
    n, k = addr.shape
    a = (addr.astype(np.int64) % banks).reshape(-1, group)
    return np.stack([np.bincount(r, minlength=banks).max() for r in a]).reshape(n, k // group)


@pytest.mark.parametrize("k", [32, 64, 16, 8, 18, 96, 256])
def test_simt_bank_plan_is_valid_and_spreads_the_banks(k):
    rng = np.random.default_rng(k)
    n = 300
    sp_index = np.sort(np.argsort(rng.random((n, 256)), axis=1)[:, :k], axis=1).astype(np.uint8)
    sp_addr = S.cbsr_bank_plan(sp_index)
    _check_plan(sp_index, sp_addr)
    if k % 32 == 0 and k < 256:
        planned = _fullest_bank(sp_addr, 32, 32)
        plain = _fullest_bank(sp_index.astype(np.uint16), 32, 32)
        assert planned.mean() < plain.mean(), (planned.mean(), plain.mean())
        if k == 32:   # 32 random columns: 3.3 with one home; exact two-home matching 2.0; this greedy plan ~2.1
            assert plain.mean() > 3.0
            assert planned.mean() <= 2.25, planned.mean()
            assert planned.max() <= 4 and (planned >= 3).mean() < 0.25
    # a row that needs no choice stays in home 0
    ident = np.tile(np.arange(32, dtype=np.uint8), (3, 1))
    assert (S.cbsr_bank_plan(ident) == ident).all()
    # 32 columns of ONE bank in home 0 (c = 32 j + 5 has 8 members; pad with another class) spread over home 1
    cls = np.sort(np.concatenate([32 * np.arange(8) + 5, 32 * np.arange(8) + 6, 32 * np.arange(8) + 7,
                                  32 * np.arange(8) + 9])).astype(np.uint8)[None, :]
    assert _fullest_bank(S.cbsr_bank_plan(cls), 32, 32).max() <= 3 < _fullest_bank(cls.astype(np.uint16), 32, 32).max()


def test_simt_bank_plan_edge_cases():
    assert S.cbsr_bank_plan(np.zeros((0, 32), np.uint8)).shape == (0, 32)
    one = S.cbsr_bank_plan(np.array([[7]], np.uint8))
    assert one.shape == (1, 1) and int(one[0, 0]) in (7, 256 + 7 * 32)


PLANNED_CASES = [dict(k=32), dict(k=64), dict(k=32, d=128), dict(k=32, d=36), dict(k=32, selector="random"),
                 dict(k=64, selector="torch"), dict(k=32, kind="uniform", avg=300.0, n=40, empty_frac=0.0),
                 dict(k=32, avg=3.0, n=1500, empty_frac=0.4), dict(k=32, n=20, avg=1500.0, empty_frac=0.0),
                 dict(k=32, n=200, n_cols=700, avg=25.0), dict(k=32, warp_max_nz=100), dict(k=64, warp_max_nz=5)]


@pytest.mark.parametrize("case", PLANNED_CASES, ids=lambda c: "-".join(f"{a}{b}" for a, b in c.items()))
def test_simt_planned_forward_vs_oracle(case):
    case = dict(n=250, **case) if "n" not in case else case
    p = Problem(seed=len(str(case)) + 1, **case)
    want = p.fwd_oracle()
    sp_addr = S.cbsr_bank_plan(p.sp_index)
    for shape, tpw in ((0, -1), (1, 1), (2, 5), (1, 16)):
        S.set_forward_depth(shape)
        S.set_tuning(tpw, 0)
        y = S.spgemm_forward_planned(p.w4, p.idx, p.val, p.sp_data, sp_addr, p.n, d=p.d)
        close(y, want, f"shape={shape} tpw={tpw}")
    y2 = S.spgemm_forward_planned(p.w4, p.idx, p.val, p.sp_data, sp_addr, p.n, d=p.d, accumulate=True, out_init=y)
    close(y2, 2 * want, "accumulate")
    y3 = S.spgemm_forward_planned(p.w4, p.idx, p.val, p.sp_data, sp_addr, p.n, d=p.d, data_align=4, addr_align=2)
    close(y3, want, "operands aligned to their element size only")


def test_simt_planned_forward_gap_rows_and_empty():
    n = 97
    deg = np.zeros(n, np.int64)
    deg[[5, 6, 40, 41, 42, 80]] = [3, 200, 1, 70, 64, 129]
    indptr = np.zeros(n + 1, np.int32)
    indptr[1:] = np.cumsum(deg)
    rng = np.random.default_rng(3)
    e = int(indptr[-1])
    indices = rng.integers(0, n, e).astype(np.int32)
    values = rng.random(e, dtype=np.float32)
    sp_data, sp_index = O.topk_cbsr(rng.standard_normal((n, 256)).astype(np.float32), 32)
    sp_addr = S.cbsr_bank_plan(sp_index)
    w4 = O.worklist(indptr)
    want = O.spgemm_forward(w4, indices, values, sp_data, sp_index, d=256, n_out=n + 10)
    for tpw in (-1, 1, 2, 16):
        S.set_tuning(tpw, 0)
        close(S.spgemm_forward_planned(w4, indices, values, sp_data, sp_addr, n + 10), want, f"tpw={tpw}")
    z32 = np.zeros(0, np.int32)
    y = S.spgemm_forward_planned(z32, z32, np.zeros(0, np.float32), sp_data, sp_addr, 10)
    assert y.shape == (10, 256) and float(np.abs(y).sum()) == 0.0
    with pytest.raises(S.SimtError, match="is not built"):
        S.spgemm_forward_planned(w4, indices, values, sp_data[:, :20], sp_addr[:, :20], n)
    with pytest.raises(S.SimtError, match="is not built"):
        S.spgemm_forward_planned(w4, indices, values, np.tile(sp_data, (1, 4)), np.tile(sp_addr, (1, 4)), n)


BWD_PLANNED_CASES = [dict(k=16), dict(k=8), dict(k=16, d=64, avg=90.0), dict(k=8, warp_max_nz=7), dict(k=32), dict(k=64),
                     dict(k=96), dict(k=128), dict(k=18), dict(k=32, d=128), dict(k=32, d=36),
                     dict(k=32, selector="random"), dict(k=64, selector="torch"),
                     dict(k=32, kind="uniform", avg=300.0, n=40, empty_frac=0.0),
                     dict(k=32, avg=3.0, n=1500, empty_frac=0.4), dict(k=32, n=20, avg=1500.0, empty_frac=0.0),
                     dict(k=32, n=200, n_cols=700, avg=25.0), dict(k=32, warp_max_nz=100), dict(k=64, warp_max_nz=5)]


@pytest.mark.parametrize("case", BWD_PLANNED_CASES, ids=lambda c: "-".join(f"{a}{b}" for a, b in c.items()))
def test_simt_planned_backward_vs_oracle(case):
    """The backward reading its columns out of the bank plan (the peer path ships the plan in place of the
    selectors): vector form for k % 4 == 0 with aligned operands, scalar form otherwise."""
    case = dict(n=250, **case) if "n" not in case else case
    p = Problem(seed=len(str(case)) + 3, **case)
    want = p.bwd_oracle()
    sp_addr = S.cbsr_bank_plan(p.sp_index)
    for tpw in (-1, 1, 5):
        S.set_tuning(0, tpw)
        g = S.sspmm_backward_planned(p.w4, p.idx, p.val, p.dy, sp_addr)
        assert g.shape == (p.n_cols, p.k)
        close(g, want, f"tpw={tpw}")
    g2 = S.sspmm_backward_planned(p.w4, p.idx, p.val, p.dy, sp_addr, accumulate=True, out_init=g)
    close(g2, 2 * want, "accumulate")
    g3 = S.sspmm_backward_planned(p.w4, p.idx, p.val, p.dy, sp_addr, dxc_align=4, addr_align=2,
                                  out_init=np.full((p.n_cols, p.k), -3.0, np.float32))
    close(g3, want, "defined output, operands aligned to their element size only")


# ------------------------------------------------------------------------------ 8(f) rank 3
@pytest.mark.parametrize("case", [dict(n=70, d=256, k=32, c=256), dict(n=33, d=256, k=64, c=132), dict(n=17, d=64, k=8, c=4),
                                  dict(n=50, d=36, k=18, c=260), dict(n=5, d=256, k=256, c=128), dict(n=1, d=128, k=1, c=8),
                                  dict(n=40, d=256, k=16, c=48), dict(n=3000, d=32, k=4, c=8)],
                         ids=lambda c: "-".join(f"{a}{b}" for a, b in c.items()))
def test_simt_cbsr_linear_products(case):
    """EXPERIMENTAL csrc/cbsr_dense.cu: fc_self on the CBSR operand and the gradient of its values, against
    the dense Linear on the scattered twin (fp64 oracle)."""
    from helpers import assert_close_sum
    n, d, k, c = case["n"], case["d"], case["k"], case["c"]
    rng = np.random.default_rng(n + d + k + c)
    sd, si = O.topk_cbsr(rng.standard_normal((n, d)).astype(np.float32), k)
    w = (rng.standard_normal((c, d)) * 0.2).astype(np.float32)   # nn.Linear.weight layout [C, D]
    b = rng.standard_normal(c).astype(np.float32)
    m = np.ascontiguousarray(w.T)
    h_abs = np.abs(O.cbsr_scatter(sd, si, d)).astype(np.float64)
    y_terms = h_abs @ np.abs(w.T).astype(np.float64)
    y = S.cbsr_matmul(sd, si, m, bias=b)
    assert_close_sum(y, O.cbsr_linear(sd, si, w, b), y_terms + np.abs(b), k + 1, "matmul + bias")
    y0 = S.cbsr_matmul(sd, si, m)
    assert_close_sum(y0, O.cbsr_linear(sd, si, w), y_terms, k, "matmul")
    y2 = S.cbsr_matmul(sd, si, m, accumulate=True, out_init=y0)
    assert_close_sum(y2, 2 * O.cbsr_linear(sd, si, w), 2 * y_terms, k + 1, "matmul, accumulate")
    g = rng.standard_normal((n, c)).astype(np.float32)
    want = O.cbsr_linear_grad_values(g, si, w)
    g_terms = np.take_along_axis(np.abs(g).astype(np.float64) @ np.abs(w).astype(np.float64), si.astype(np.int64), axis=1)
    dv = S.cbsr_sampled_dot(g, m, si)
    assert dv.shape == (n, k)
    assert_close_sum(dv, want, g_terms, c, "sampled dot")
    dv2 = S.cbsr_sampled_dot(g, m, si, accumulate=True, out_init=np.ones((n, k), np.float32))
    assert_close_sum(dv2, want + 1.0, g_terms + 1.0, c + 1, "sampled dot, accumulate")


@pytest.mark.parametrize("n,d,k", [(70, 256, 32), (9, 64, 64), (33, 36, 18), (1, 4, 1), (20, 256, 256)])
def test_simt_cbsr_axpy_in_place(n, d, k):
    """y += (alpha + *alpha_dev) * scatter(sp_data, sp_index): GIN's self term without the dense twin."""
    rng = np.random.default_rng(n + d + k)
    sd, si = O.topk_cbsr(rng.standard_normal((n, d)).astype(np.float32), k)
    y = rng.standard_normal((n, d)).astype(np.float32)
    h = O.cbsr_scatter(sd, si, d).astype(np.float64)
    close(S.cbsr_axpy(y, sd, si, alpha=1.0, alpha_dev=0.25), y + 1.25 * h, "alpha on the device")
    close(S.cbsr_axpy(y, sd, si, alpha=-2.0), y - 2.0 * h, "host alpha only")
    if d < 256:   # a selector outside [0, d) contributes nothing
        si_bad, h_bad = si.copy(), h.copy()
        si_bad[0, 0] = 255
        h_bad[0, si[0, 0]] = 0.0
        close(S.cbsr_axpy(y, sd, si_bad, alpha=1.0), y + h_bad, "selector outside [0, d)")


def test_simt_cbsr_linear_products_edge_cases():
    rng = np.random.default_rng(3)
    m = rng.standard_normal((20, 8)).astype(np.float32)
    si = np.array([[0, 19, 20, 255]], np.uint8)               # two selectors outside [0, d): contribute nothing
    sd = np.array([[1.0, 2.0, 5.0, 7.0]], np.float32)
    np.testing.assert_allclose(S.cbsr_matmul(sd, si, m)[0], m[0] + 2 * m[19], rtol=1e-6)
    g = rng.standard_normal((1, 8)).astype(np.float32)
    dv = S.cbsr_sampled_dot(g, m, si)[0]
    np.testing.assert_allclose(dv[:2], [g[0] @ m[0], g[0] @ m[19]], rtol=1e-5)
    assert dv[2] == 0.0 and dv[3] == 0.0
    assert S.cbsr_matmul(np.zeros((0, 4), np.float32), np.zeros((0, 4), np.uint8), m).shape == (0, 8)
    with pytest.raises(S.SimtError, match="multiple of 4"):
        S.cbsr_matmul(sd, si, rng.standard_normal((20, 6)).astype(np.float32))
    with pytest.raises(S.SimtError, match="d=300"):
        S.cbsr_sampled_dot(np.zeros((1, 8), np.float32), np.zeros((300, 8), np.float32), si)


def test_simt_planned_backward_edge_cases():
    z32 = np.zeros(0, np.int32)
    sp_addr = np.tile(np.arange(32, dtype=np.uint16), (10, 1))
    g = S.sspmm_backward_planned(z32, z32, np.zeros(0, np.float32), np.ones((10, 256), np.float32), sp_addr)
    assert g.shape == (10, 32) and float(np.abs(g).sum()) == 0.0


def test_simt_empty_inputs():
    z32 = np.zeros(0, np.int32)
    sp_data = np.random.default_rng(0).random((10, 32), dtype=np.float32)
    sp_index = np.tile(np.arange(32, dtype=np.uint8), (10, 1))
    y = S.spgemm_forward(z32, z32, np.zeros(0, np.float32), sp_data, sp_index, 10)
    assert y.shape == (10, 256) and float(np.abs(y).sum()) == 0.0
    dxc = S.sspmm_backward(z32, z32, np.zeros(0, np.float32), np.ones((10, 256), np.float32), sp_index)
    assert dxc.shape == (10, 32) and float(np.abs(dxc).sum()) == 0.0
    v, i = S.topk_f32(np.zeros((0, 256), np.float32), 32)
    assert v.shape == (0, 32)


# ------------------------------------------------------------------------------ a5
def test_simt_layout_kernels_exact():
    for k, d in ((32, 256), (8, 64), (18, 128), (64, 256), (1, 4)):
        p = Problem(n=67, avg=5.0, k=k, d=d, seed=k)
        dense = S.cbsr_scatter(p.sp_data, p.sp_index, d)
        np.testing.assert_array_equal(dense, O.cbsr_scatter(p.sp_data, p.sp_index, d))
        np.testing.assert_array_equal(S.cbsr_gather(dense, p.sp_index), p.sp_data)
        np.testing.assert_array_equal(S.cbsr_mask(p.x, p.sp_index), O.maxk_dense(p.x, k))
        if k % 4 == 0:
            rec = S.cbsr_pack(p.sp_data, p.sp_index)
            d2, i2 = S.cbsr_unpack(rec, 67, k)
            np.testing.assert_array_equal(d2, p.sp_data)
            np.testing.assert_array_equal(i2, p.sp_index)


# ---------------------------------------------------------------- the interpreter itself
def test_simt_reports_out_of_bounds_accesses():
    """The guard pages work: an index array that is one element short is reported (with the
    buffer and the CUDA thread), and the next call runs normally."""
    p = Problem(n=80, avg=20.0, k=32, seed=3)
    with pytest.raises(S.SimtError, match="indices"):
        S.spgemm_forward(p.w4, p.idx[:-1], p.val, p.sp_data, p.sp_index, p.n)
    with pytest.raises(S.SimtError, match="dxc"):
        S.sspmm_backward(p.w4, p.idx, p.val, p.dy, p.sp_index, n_cbsr=p.n - 1)
    close(S.spgemm_forward(p.w4, p.idx, p.val, p.sp_data, p.sp_index, p.n), p.fwd_oracle())


def test_simt_library_is_not_the_product():
    """The interpreter build is test infrastructure: the product binds lib/libmaxk_b200.so only."""
    from maxk_b200 import _cabi
    src = open(_cabi.__file__).read()
    assert "simt" not in src.lower()
    assert os.path.basename(_cabi.LIB_PATH) == "libmaxk_b200.so"


# ------------------------------------------------------------------------------ LayerNorm of the layers
@pytest.mark.parametrize("d", [256, 128, 64, 36, 4])
def test_simt_layernorm_forward_backward(d):
    rng = np.random.default_rng(900 + d)
    for n in (1, 7, 8, 9, 300, 2500):        # 2500 rows: the persistent grid wraps around
        x = (rng.standard_normal((n, d)) * 3 + 1.5).astype(np.float32)
        gamma = rng.standard_normal(d).astype(np.float32)
        beta = rng.standard_normal(d).astype(np.float32)
        for ga, be in ((gamma, beta), (None, None)):
            y, mean, rstd = S.layernorm_forward(x, ga, be)
            wy, wm, wr = O.layernorm(x, ga, be)
            np.testing.assert_allclose(y, wy, rtol=2e-5, atol=2e-5)
            np.testing.assert_allclose(mean, wm, rtol=1e-5, atol=1e-6)
            np.testing.assert_allclose(rstd, wr, rtol=1e-5, atol=1e-6)
            g = rng.standard_normal((n, d)).astype(np.float32)
            dx, dga, dbe = S.layernorm_backward(g, x, mean, rstd, ga)
            wdx, wdg, wdb = O.layernorm_grad(g, x, ga)
            np.testing.assert_allclose(dx, wdx, rtol=1e-4, atol=1e-4)
            np.testing.assert_allclose(dga, wdg, rtol=1e-4, atol=1e-4 * max(1, n ** 0.5))
            np.testing.assert_allclose(dbe, wdb, rtol=1e-4, atol=1e-4 * max(1, n ** 0.5))
            (dx2, none_g, none_b) = S.layernorm_backward(g, x, mean, rstd, ga, want_affine=False)
            np.testing.assert_array_equal(dx2, dx)
            assert none_g is None and none_b is None


@pytest.mark.parametrize("d", [256, 128, 36])
def test_simt_layernorm_topk_and_its_sparse_backward(d):
    """LayerNorm + TopK in one pass: the selection is the oracle's TopK of the kernel's OWN normalised row (bit-exact),
    the normalised row is the LayerNorm oracle's; the backward on the k-sparse gradient equals the dense LayerNorm
    backward on its scattered twin."""
    rng = np.random.default_rng(40 + d)
    for n in (1, 9, 70):
        x = (rng.standard_normal((n, d)) * 2 - 0.5).astype(np.float32)
        gamma = rng.standard_normal(d).astype(np.float32)
        beta = rng.standard_normal(d).astype(np.float32)
        for k in sorted({1, 8, 32, d // 2, d} & set(range(1, d + 1))):
            r = S.layernorm_topk(x, gamma, beta, k, want_dense=True)
            wy, wm, wr = O.layernorm(x, gamma, beta)
            np.testing.assert_allclose(r["y"], wy, rtol=2e-5, atol=2e-5)
            want_v, want_i = O.topk_cbsr(r["y"], k)
            np.testing.assert_array_equal(r["sp_index"], want_i, err_msg=f"n={n} d={d} k={k}")
            np.testing.assert_array_equal(r["sp_data"].view(np.uint32), want_v.view(np.uint32))
            r2 = S.layernorm_topk(x, gamma, beta, k)          # without the dense output
            np.testing.assert_array_equal(r2["sp_index"], want_i)
            g_data = rng.standard_normal((n, k)).astype(np.float32)
            dx, dga, dbe = S.layernorm_backward_cbsr(g_data, r["sp_index"], x, r["mean"], r["rstd"], gamma)
            dense_g = O.cbsr_scatter(g_data, r["sp_index"], d)
            wdx, wdg, wdb = O.layernorm_grad(dense_g, x, gamma)
            np.testing.assert_allclose(dx, wdx, rtol=1e-4, atol=1e-4)
            np.testing.assert_allclose(dga, wdg, rtol=1e-4, atol=1e-4 * max(1, n ** 0.5))
            np.testing.assert_allclose(dbe, wdb, rtol=1e-4, atol=1e-4 * max(1, n ** 0.5))
