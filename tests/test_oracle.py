"""Pin the CPU oracle: golden vectors from the reference (generate_meta.py run
unmodified; torch.topk), internal consistency (C vs numpy vs dense twin), and the
Q1 known answer from the reference's debug.txt."""
import os

import numpy as np
import pytest

from oracle import oracle as O


def _cases(npz, suffix):
    return sorted(k[: -len(suffix)] for k in npz.files if k.endswith(suffix))


def test_worklist_matches_reference_generate_meta(golden_dir):
    z = np.load(os.path.join(golden_dir, "worklist_cases.npz"))
    names = _cases(z, "__indptr")
    assert len(names) >= 7
    for name in names:
        indptr, want = z[name + "__indptr"], z[name + "__warp4"]
        got = O.worklist(indptr)
        assert got.dtype == np.int32
        np.testing.assert_array_equal(got, want, err_msg=name)
        np.testing.assert_array_equal(O.worklist_numpy(indptr), want, err_msg=name)


def test_worklist_properties():
    indptr, idx = O.synth_graph(3000, 40.0, seed=3, empty_frac=0.1)
    w4 = O.worklist(indptr).reshape(-1, 4)
    deg = np.diff(indptr)
    assert w4[:, 2].sum() == indptr[-1]
    assert (w4[:, 2] >= 1).all() and (w4[:, 2] <= 64).all()
    assert (w4[:, 3] == 0).all()
    assert (np.diff(w4[:, 0]) >= 0).all()
    assert w4.shape[0] == int(((deg + 63) // 64).sum())
    # loc is the exclusive prefix of len (generate_meta.py's running counter)
    np.testing.assert_array_equal(w4[:, 1], np.cumsum(w4[:, 2]) - w4[:, 2])
    # other warp_max_nz values
    for nz in (1, 7, 32, 100):
        a = O.worklist(indptr, nz)
        np.testing.assert_array_equal(a, O.worklist_numpy(indptr, nz))


def test_worklist_empty_graph():
    assert O.worklist(np.zeros(1, np.int32)).shape == (0,)
    assert O.worklist(np.zeros(6, np.int32)).shape == (0,)


def test_topk_matches_torch_topk_goldens(golden_dir):
    z = np.load(os.path.join(golden_dir, "topk_cases.npz"))
    checked = 0
    for key in z.files:
        if key.startswith("idx_d"):
            d, k = key[4:].split("_k")
            x = z[f"x_{d}"]
            val, ind = O.topk_cbsr(x, int(k))
            np.testing.assert_array_equal(ind, z[key], err_msg=key)
            np.testing.assert_array_equal(val.view(np.uint32), z[f"val_{d}_k{k}"].view(np.uint32))
            checked += 1
        if key.startswith("negidx_"):
            k = int(key.split("_k")[1])
            val, ind = O.topk_cbsr(z["xneg_d256"], k)
            np.testing.assert_array_equal(ind, z[key], err_msg=key)
            np.testing.assert_array_equal(val, z[f"negval_d256_k{k}"])
            checked += 1
    assert checked >= 15


def test_topk_tie_rule_and_specials():
    x = np.zeros((4, 256), np.float32)
    x[0, :] = 1.0                      # all tied -> lowest k columns
    x[1, 10:20] = 5.0                  # 10 big, rest tied at 0 -> 10 + lowest others
    x[2, :] = -1.0
    x[2, 200] = np.nan                 # NaN is the largest (torch rule)
    x[2, 7] = np.inf
    x[3, 0] = -0.0                     # -0 == +0 tie
    x[3, 1:] = 0.0
    val, ind = O.topk_cbsr(x, 16)
    np.testing.assert_array_equal(ind[0], np.arange(16))
    np.testing.assert_array_equal(ind[1], np.r_[0:6, 10:20])
    assert 200 in ind[2] and 7 in ind[2]
    np.testing.assert_array_equal(ind[2], np.r_[0:7, 7, 8:15, 200][np.argsort(np.r_[0:7, 7, 8:15, 200])])
    np.testing.assert_array_equal(ind[3], np.arange(16))
    assert np.signbit(val[3, 0])       # bit copy keeps the sign of -0
    # ascending emission
    assert (np.diff(ind.astype(np.int32), axis=1) > 0).all()


def test_topk_k_equals_d_and_u8():
    rng = np.random.default_rng(0)
    x = rng.random((5, 64), dtype=np.float32)
    val, ind = O.topk_cbsr(x, 64)
    np.testing.assert_array_equal(ind, np.tile(np.arange(64, dtype=np.uint8), (5, 1)))
    np.testing.assert_array_equal(val, x)
    xu = rng.integers(0, 256, size=(33, 256), dtype=np.uint8)   # heavy ties
    v, i = O.topk_cbsr(xu, 32)
    for r in range(33):
        order = np.lexsort((np.arange(256), -xu[r].astype(np.int32)))[:32]
        np.testing.assert_array_equal(i[r], np.sort(order))
        np.testing.assert_array_equal(v[r], xu[r][np.sort(order)])


def _problem(n=700, avg=30.0, k=32, d=256, seed=1, kind="powerlaw", empty_frac=0.05):
    rng = np.random.default_rng(seed)
    indptr, idx = O.synth_graph(n, avg, seed=seed, kind=kind, empty_frac=empty_frac)
    val = rng.random(idx.shape[0], dtype=np.float32)
    x = rng.random((n, d), dtype=np.float32)
    sp_data, sp_index = O.topk_cbsr(x, k)
    w4 = O.worklist(indptr)
    return indptr, idx, val, x, sp_data, sp_index, w4


@pytest.mark.parametrize("k", [8, 16, 18, 32, 64])
def test_forward_equals_dense_twin(k):
    indptr, idx, val, x, sp_data, sp_index, w4 = _problem(k=k)
    y = O.spgemm_forward(w4, idx, val, sp_data, sp_index, 256)
    twin = O.cbsr_scatter(sp_data, sp_index, 256)
    np.testing.assert_array_equal(twin, O.maxk_dense(x, k))
    y2 = O.csr_spmm(indptr, idx, val, twin)
    np.testing.assert_allclose(y, y2, rtol=1e-12, atol=1e-12)
    y32 = O.spgemm_forward(w4, idx, val, sp_data, sp_index, 256, dtype=np.float32)
    np.testing.assert_allclose(y32, y, rtol=1e-5, atol=1e-6)
    # scipy third opinion
    import scipy.sparse as sp
    a = sp.csr_matrix((val.astype(np.float64), idx, indptr), shape=(x.shape[0], x.shape[0]))
    np.testing.assert_allclose(a @ twin.astype(np.float64), y, rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("k", [8, 16, 32, 64])
def test_backward_equals_sampled_transpose(k):
    indptr, idx, val, x, sp_data, sp_index, w4 = _problem(k=k, seed=2)
    rng = np.random.default_rng(5)
    dy = rng.random((x.shape[0], 256), dtype=np.float32)
    dxc = O.sspmm_backward(w4, idx, val, dy, sp_index)
    full = O.csr_spmm_t(indptr, idx, val, dy, x.shape[0])          # A^T dY
    want = np.take_along_axis(full, sp_index.astype(np.int64), axis=1)
    np.testing.assert_allclose(dxc, want, rtol=1e-12, atol=1e-12)
    dxc32 = O.sspmm_backward(w4, idx, val, dy, sp_index, dtype=np.float32)
    np.testing.assert_allclose(dxc32, dxc, rtol=1e-5, atol=1e-6)


def test_gradient_identity_against_autograd():
    """dX = mask * (A^T dY)  (utils/models.py:115-143 / SURVEY A.3) vs torch autograd on the
    dense formulation MaxK -> sparse.mm."""
    import torch
    indptr, idx, val, x, sp_data, sp_index, w4 = _problem(n=300, avg=12.0, k=16, seed=9)
    n = x.shape[0]
    xt = torch.tensor(x, dtype=torch.float64, requires_grad=True)
    a = torch.sparse_csr_tensor(torch.tensor(indptr, dtype=torch.int64),
                                torch.tensor(idx, dtype=torch.int64),
                                torch.tensor(val, dtype=torch.float64), size=(n, n))
    _, ti = torch.topk(xt, 16, dim=1)
    mask = torch.zeros_like(xt).scatter_(1, ti, 1.0)
    y = torch.sparse.mm(a, xt * mask)
    dy = torch.rand(n, 256, dtype=torch.float64, generator=torch.Generator().manual_seed(4))
    y.backward(dy)
    dxc = O.sspmm_backward(w4, idx, val, dy.numpy().astype(np.float32), sp_index)
    dx = np.zeros((n, 256))
    np.put_along_axis(dx, sp_index.astype(np.int64), dxc, axis=1)
    np.testing.assert_allclose(dx, xt.grad.numpy(), rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(O.spgemm_forward(w4, idx, val, sp_data, sp_index, 256),
                               y.detach().numpy(), rtol=1e-6, atol=1e-6)


def test_rectangular_partition_shapes():
    """Row partition: A_local is [n_local x N]; outputs have n_local rows, CBSR has N rows."""
    indptr, idx, val, x, sp_data, sp_index, w4 = _problem(n=500, avg=20.0, k=32, seed=4)
    lo, hi = 120, 310
    ip = (indptr[lo:hi + 1] - indptr[lo]).astype(np.int32)
    ix, vv = idx[indptr[lo]:indptr[hi]], val[indptr[lo]:indptr[hi]]
    w4l = O.worklist(ip)
    y_full = O.spgemm_forward(w4, idx, val, sp_data, sp_index, 256)
    y_loc = O.spgemm_forward(w4l, ix, vv, sp_data, sp_index, 256, n_out=hi - lo)
    np.testing.assert_allclose(y_loc, y_full[lo:hi], rtol=1e-12, atol=1e-12)


def test_q1_mask_known_answer():
    """debug.txt:106-128: k=18, only the last task is hit and every bad column has
    c % 32 >= 18; k=16 hits the last two CTAs' worth of tasks; k>=32 nothing."""
    w = 100
    m = O.q1_forward_dropped(w, 18)
    assert m[:-1].sum() == 0
    bad = np.nonzero(m[-1])[0]
    assert len(bad) > 0 and ((bad % 32) >= 18).all()
    assert set([29, 53, 82, 91, 92, 95, 115, 124, 147, 179, 253]).issubset(set(bad.tolist()))
    assert O.q1_forward_dropped(w, 32).sum() == 0 and O.q1_forward_dropped(w, 64).sum() == 0
    m16 = O.q1_forward_dropped(w, 16)
    hit = np.nonzero(m16.any(axis=1))[0]
    assert hit.min() >= w - 24


def test_cpu_path_port_matches_oracle():
    """oracle/cpu_path.py (torch.topk + CSR SpMM, the reference's CPU path restated) against
    the C oracle on tie-free data."""
    import torch
    from oracle.cpu_path import CpuAggregation
    indptr, idx, val, x, sp_data, sp_index, w4 = _problem(n=400, avg=25.0, k=32, seed=12)
    rng = np.random.default_rng(3)
    dy = rng.random((400, 256), dtype=np.float32)
    agg = CpuAggregation(torch.from_numpy(indptr), torch.from_numpy(idx), torch.from_numpy(val))
    h, mask = agg.prepare(torch.from_numpy(x), 32)
    y, dx = agg.step(h, torch.from_numpy(dy), mask)
    np.testing.assert_allclose(y.numpy(), O.spgemm_forward(w4, idx, val, sp_data, sp_index, 256),
                               rtol=1e-5, atol=1e-6)
    dxc = O.sspmm_backward(w4, idx, val, dy, sp_index)
    want = np.zeros((400, 256))
    np.put_along_axis(want, sp_index.astype(np.int64), dxc, axis=1)
    np.testing.assert_allclose(dx.numpy(), want, rtol=1e-5, atol=1e-6)


def test_warp4_file_round_trip(tmp_path, monkeypatch):
    """The on-disk work-list format: what save_warp4_metadata writes is byte for
    byte what kernels/generate_meta.py writes for the same indptr (np.int32 quadruples, row-major),
    and GraphDataLoader reads back the raw CSR files it saved."""
    import sys
    import torch
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "spgemm-new_b200"))
    import maxk_cuda_kernels as mk
    from graph_loader import GraphDataLoader
    monkeypatch.chdir(tmp_path)
    indptr, idx = O.synth_graph(300, 40.0, seed=9, empty_frac=0.1)
    w4 = O.worklist(indptr)
    path = mk.save_warp4_metadata(torch.from_numpy(w4), "toy")
    assert path.endswith(os.path.join("w12_nz64_warp_4", "toy.warp4"))
    raw = np.fromfile(path, dtype=np.int32)
    np.testing.assert_array_equal(raw, w4.reshape(-1))
    assert raw.reshape(-1, 4)[:, 3].max() == 0 and raw.reshape(-1, 4)[:, 2].max() <= 64
    g = {"indptr": torch.from_numpy(indptr), "indices": torch.from_numpy(idx)}
    GraphDataLoader.save_graph(g, str(tmp_path / "graphs"), "toy")
    loader = GraphDataLoader(str(tmp_path / "graphs"))
    assert loader.get_available_graphs() == ["toy"]
    back = loader.load_graph("toy")
    np.testing.assert_array_equal(back["indptr"].numpy(), indptr)
    np.testing.assert_array_equal(back["indices"].numpy(), idx)
    assert back["v_num"] == 300 and back["e_num"] == idx.shape[0]
