"""ctypes/numpy front-end of the CPU oracle (oracle/maxk_oracle.c).

TEST INFRASTRUCTURE ONLY -- see the header of maxk_oracle.c.  Imported by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs; never
by the product package.

All arrays are numpy, C-contiguous; dtypes follow the reference's boundary
(cuda_kernel_bindings.cpp:52-62): int32 CSR / work list, float32 values and
features, uint8 selectors.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libmaxk_oracle.so")
_lib = None

WARP_MAX_NZ = 64  # kernels/generate_meta.py:9


def build(force: bool = False) -> str:
    """Compile the C restatement (gcc + OpenMP).  Building the checker is not using it."""
    src = os.path.join(_HERE, "maxk_oracle.c")
    if (not force and os.path.exists(_LIB_PATH)
            and os.path.getmtime(_LIB_PATH) >= os.path.getmtime(src)):
        return _LIB_PATH
    subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "all"])
    return _LIB_PATH


def _load():
    global _lib
    if _lib is not None:
        return _lib
    build()
    lib = ctypes.CDLL(_LIB_PATH)
    i64, i32, vp = ctypes.c_int64, ctypes.c_int32, ctypes.c_void_p
    lib.oracle_worklist.restype = i64
    lib.oracle_worklist.argtypes = [vp, i64, i32, vp]
    lib.oracle_topk_cbsr_f32.argtypes = [vp, i64, i32, i32, vp, vp]
    lib.oracle_topk_cbsr_u8.argtypes = [vp, i64, i32, i32, vp, vp]
    lib.oracle_maxk_dense_f32.argtypes = [vp, i64, i32, i32, vp]
    lib.oracle_cbsr_scatter_f32.argtypes = [vp, vp, i64, i32, i32, vp]
    lib.oracle_cbsr_gather_f32.argtypes = [vp, vp, i64, i32, i32, vp]
    for name in ("oracle_spgemm_forward_f32", "oracle_spgemm_forward_f64",
                 "oracle_sspmm_backward_f32", "oracle_sspmm_backward_f64"):
        getattr(lib, name).argtypes = [vp, i64, vp, vp, vp, vp, i32, i32, i64, vp]
    lib.oracle_csr_spmm_f64.argtypes = [vp, vp, vp, i64, vp, i32, vp]
    lib.oracle_csr_spmm_t_f64.argtypes = [vp, vp, vp, i64, i64, vp, i32, vp]
    lib.oracle_q1_forward_dropped.argtypes = [i64, i32, i32, vp]
    lib.oracle_num_threads.restype = ctypes.c_int
    _lib = lib
    return lib


def _p(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)


def _c(a, dtype):
    a = np.ascontiguousarray(a, dtype=dtype)
    return a


def num_threads() -> int:
    return int(_load().oracle_num_threads())


# --------------------------------------------------------------------------- a4
def worklist(indptr, warp_max_nz: int = WARP_MAX_NZ) -> np.ndarray:
    """(row, loc, len, 0) int32 records, flat [4W] -- kernels/generate_meta.py:30-48."""
    lib = _load()
    indptr = _c(indptr, np.int32)
    n = indptr.shape[0] - 1
    w = lib.oracle_worklist(_p(indptr), n, warp_max_nz, None)
    out = np.empty(4 * w, dtype=np.int32)
    lib.oracle_worklist(_p(indptr), n, warp_max_nz, _p(out))
    return out


def worklist_numpy(indptr, warp_max_nz: int = WARP_MAX_NZ) -> np.ndarray:
    """Vectorised second opinion on the same algorithm (no shared code with the C one)."""
    indptr = np.asarray(indptr, dtype=np.int64)
    deg = np.diff(indptr)
    chunks = (deg + warp_max_nz - 1) // warp_max_nz
    w = int(chunks.sum())
    rows = np.repeat(np.arange(deg.shape[0], dtype=np.int64), chunks)
    first = np.cumsum(chunks) - chunks
    c = np.arange(w, dtype=np.int64) - np.repeat(first, chunks)
    loc = indptr[rows] + warp_max_nz * c
    ln = np.minimum(warp_max_nz, deg[rows] - warp_max_nz * c)
    out = np.zeros((w, 4), dtype=np.int32)
    out[:, 0], out[:, 1], out[:, 2] = rows, loc, ln
    return out.reshape(-1)


# --------------------------------------------------------------------------- a1
def topk_cbsr(x, k: int):
    """Exact top-k, lowest-column tie-break, ascending-column emission."""
    lib = _load()
    if x.dtype == np.uint8:
        x = _c(x, np.uint8)
        n, d = x.shape
        val = np.empty((n, k), np.uint8)
        ind = np.empty((n, k), np.uint8)
        lib.oracle_topk_cbsr_u8(_p(x), n, d, k, _p(val), _p(ind))
        return val, ind
    x = _c(x, np.float32)
    n, d = x.shape
    assert 1 <= k <= d <= 256
    val = np.empty((n, k), np.float32)
    ind = np.empty((n, k), np.uint8)
    lib.oracle_topk_cbsr_f32(_p(x), n, d, k, _p(val), _p(ind))
    return val, ind


def maxk_dense(x, k: int) -> np.ndarray:
    lib = _load()
    x = _c(x, np.float32)
    n, d = x.shape
    out = np.empty_like(x)
    lib.oracle_maxk_dense_f32(_p(x), n, d, k, _p(out))
    return out


# --------------------------------------------------------------------------- a5
def cbsr_scatter(sp_data, sp_index, d: int) -> np.ndarray:
    lib = _load()
    sp_data = _c(sp_data, np.float32)
    sp_index = _c(sp_index, np.uint8)
    n, k = sp_data.shape
    out = np.empty((n, d), np.float32)
    lib.oracle_cbsr_scatter_f32(_p(sp_data), _p(sp_index), n, k, d, _p(out))
    return out


def cbsr_gather(dense, sp_index) -> np.ndarray:
    lib = _load()
    dense = _c(dense, np.float32)
    sp_index = _c(sp_index, np.uint8)
    n, d = dense.shape
    k = sp_index.shape[1]
    out = np.empty((n, k), np.float32)
    lib.oracle_cbsr_gather_f32(_p(dense), _p(sp_index), n, k, d, _p(out))
    return out


# --------------------------------------------------------------------------- 8(f) rank 3
def cbsr_linear(sp_data, sp_index, weight, bias=None) -> np.ndarray:
    """fc_self on the MaxK-sparsified activations, the way the reference computes it: torch's dense
    Linear  y = h @ W^T + b  on the masked dense h = scatter(sp_data, sp_index)  (upstream
    utils/models_old.py:55-60, fork utils/models.py:242; W is [C, D] like nn.Linear.weight).  fp64."""
    w = np.asarray(weight, np.float64)
    h = cbsr_scatter(sp_data, sp_index, w.shape[1]).astype(np.float64)
    y = h @ w.T
    return y if bias is None else y + np.asarray(bias, np.float64)


def cbsr_linear_grad_values(g, sp_index, weight) -> np.ndarray:
    """d loss / d sp_data of cbsr_linear for g = d loss / d y:  gather(g @ W, sp_index)  -- autograd's
    Linear backward followed by the CBSR gather of TopK's backward (utils/models.py:55-59).  fp64."""
    gx = np.asarray(g, np.float64) @ np.asarray(weight, np.float64)
    return np.take_along_axis(gx, np.asarray(sp_index).astype(np.int64), axis=1)


# --------------------------------------------------------------------------- a2
def spgemm_forward(warp4, idx, val, sp_data, sp_index, d: int = 256, n_out=None,
                   dtype=np.float64) -> np.ndarray:
    lib = _load()
    warp4 = _c(warp4, np.int32).reshape(-1)
    idx = _c(idx, np.int32)
    val = _c(val, np.float32)
    sp_data = _c(sp_data, np.float32)
    sp_index = _c(sp_index, np.uint8)
    n, k = sp_data.shape
    n_out = n if n_out is None else int(n_out)
    w = warp4.shape[0] // 4
    y = np.empty((n_out, d), dtype)
    fn = lib.oracle_spgemm_forward_f64 if dtype == np.float64 else lib.oracle_spgemm_forward_f32
    fn(_p(warp4), w, _p(idx), _p(val), _p(sp_data), _p(sp_index), k, d, n_out, _p(y))
    return y


# --------------------------------------------------------------------------- a3
def sspmm_backward(warp4, idx, val, dy, sp_index, n_cbsr=None, dtype=np.float64) -> np.ndarray:
    lib = _load()
    warp4 = _c(warp4, np.int32).reshape(-1)
    idx = _c(idx, np.int32)
    val = _c(val, np.float32)
    dy = _c(dy, np.float32)
    sp_index = _c(sp_index, np.uint8)
    d = dy.shape[1]
    k = sp_index.shape[1]
    n_cbsr = sp_index.shape[0] if n_cbsr is None else int(n_cbsr)
    w = warp4.shape[0] // 4
    out = np.empty((n_cbsr, k), dtype)
    fn = lib.oracle_sspmm_backward_f64 if dtype == np.float64 else lib.oracle_sspmm_backward_f32
    fn(_p(warp4), w, _p(idx), _p(val), _p(dy), _p(sp_index), k, d, n_cbsr, _p(out))
    return out


# ----------------------------------------------------------- dense-twin cross-checks
def csr_spmm(indptr, idx, val, x_dense) -> np.ndarray:
    lib = _load()
    indptr, idx = _c(indptr, np.int32), _c(idx, np.int32)
    val, x_dense = _c(val, np.float32), _c(x_dense, np.float32)
    n_rows = indptr.shape[0] - 1
    d = x_dense.shape[1]
    y = np.empty((n_rows, d), np.float64)
    lib.oracle_csr_spmm_f64(_p(indptr), _p(idx), _p(val), n_rows, _p(x_dense), d, _p(y))
    return y


def csr_spmm_t(indptr, idx, val, dy, n_cols: int) -> np.ndarray:
    lib = _load()
    indptr, idx = _c(indptr, np.int32), _c(idx, np.int32)
    val, dy = _c(val, np.float32), _c(dy, np.float32)
    n_rows = indptr.shape[0] - 1
    d = dy.shape[1]
    dx = np.empty((n_cols, d), np.float64)
    lib.oracle_csr_spmm_t_f64(_p(indptr), _p(idx), _p(val), n_rows, n_cols, _p(dy), d, _p(dx))
    return dx


def q1_forward_dropped(w: int, k: int, d: int = 256) -> np.ndarray:
    """uint8 [W, d]: 1 where the reference forward kernel loses the column."""
    lib = _load()
    out = np.empty((w, d), np.uint8)
    lib.oracle_q1_forward_dropped(w, k, d, _p(out))
    return out


# ------------------------------------------------------------------ synthetic inputs
def synth_graph(n: int, avg_deg: float, seed: int = 0, kind: str = "powerlaw",
                n_cols=None, max_deg=None, empty_frac: float = 0.0):
    """Deterministic synthetic CSR with the degree shape of the BASELINE.json configs.

    kind: "powerlaw" (log-normal degrees, columns drawn proportionally to degree --
    Flickr/Reddit/products-like), "uniform" (near-constant dense degrees --
    proteins-like), "rmat" is provided by spgemm-new_b200's graph module for the
    big sweeps.  Columns are sorted within each row (DGL's CSR order).
    """
    rng = np.random.default_rng(seed)
    n_cols = n if n_cols is None else n_cols
    if kind == "uniform":
        deg = rng.integers(max(1, int(avg_deg * 0.75)), int(avg_deg * 1.25) + 2, size=n)
    else:
        sigma = 1.2
        raw = rng.lognormal(mean=0.0, sigma=sigma, size=n)
        deg = np.maximum(0, np.round(raw * (avg_deg / raw.mean()))).astype(np.int64)
    if max_deg is not None:
        deg = np.minimum(deg, max_deg)
    if empty_frac > 0:
        deg[rng.random(n) < empty_frac] = 0
    indptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(deg, out=indptr[1:])
    e = int(indptr[-1])
    assert e < 2**31
    if kind == "uniform":
        idx = rng.integers(0, n_cols, size=e, dtype=np.int64)
    else:
        popularity = rng.lognormal(mean=0.0, sigma=1.2, size=n_cols)
        cdf = np.cumsum(popularity)
        idx = np.searchsorted(cdf, rng.random(e) * cdf[-1]).astype(np.int64)
        idx = np.minimum(idx, n_cols - 1)
    rows = np.repeat(np.arange(n, dtype=np.int64), deg)
    order = np.lexsort((idx, rows))
    idx = idx[order]
    return indptr.astype(np.int32), idx.astype(np.int32)


def layernorm(x, gamma=None, beta=None, eps: float = 1e-5):
    """nn.LayerNorm over the last dimension (utils/models.py:207-213), fp64: (y, mean, rstd)."""
    x = np.asarray(x, np.float64)
    mean = x.mean(axis=1, keepdims=True)
    var = ((x - mean) ** 2).mean(axis=1, keepdims=True)
    rstd = 1.0 / np.sqrt(var + eps)
    y = (x - mean) * rstd
    if gamma is not None:
        y = y * np.asarray(gamma, np.float64)
    if beta is not None:
        y = y + np.asarray(beta, np.float64)
    return y, mean[:, 0], rstd[:, 0]


def layernorm_grad(g, x, gamma=None, eps: float = 1e-5):
    """Gradients of `layernorm`, fp64: (dx, dgamma, dbeta)."""
    g, x = np.asarray(g, np.float64), np.asarray(x, np.float64)
    _, mean, rstd = layernorm(x, None, None, eps)
    xhat = (x - mean[:, None]) * rstd[:, None]
    w = g if gamma is None else g * np.asarray(gamma, np.float64)
    dx = rstd[:, None] * (w - w.mean(axis=1, keepdims=True) - xhat * (w * xhat).mean(axis=1, keepdims=True))
    return dx, (g * xhat).sum(axis=0), g.sum(axis=0)
