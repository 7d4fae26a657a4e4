"""Shared helpers for the GPU parity tests: seeded problems on host (numpy, for the oracle)
and on the device (torch, for the C-ABI path)."""
import numpy as np
import torch

from oracle import oracle as O

RTOL, ATOL = 1e-5, 1e-6  # BASELINE.json north_star: fp32 rtol 1e-5 / atol 1e-6


class Problem:
    def __init__(self, n=600, avg=30.0, k=32, d=256, seed=1, kind="powerlaw", empty_frac=0.05,
                 max_deg=None, n_cols=None, selector="topk", warp_max_nz=64, values="uniform"):
        rng = np.random.default_rng(seed)
        self.n, self.k, self.d = n, k, d
        self.n_cols = n if n_cols is None else n_cols
        self.indptr, self.idx = O.synth_graph(n, avg, seed=seed, kind=kind, empty_frac=empty_frac,
                                              max_deg=max_deg, n_cols=self.n_cols)
        e = self.idx.shape[0]
        self.val = (rng.random(e, dtype=np.float32) if values == "uniform"
                    else np.ones(e, np.float32))
        self.x = rng.random((self.n_cols, d), dtype=np.float32)
        if selector == "topk":          # ascending columns (our TopK / the reference kernel)
            self.sp_data, self.sp_index = O.topk_cbsr(self.x, k)
        elif selector == "torch":       # descending values (torch.topk order, models.py:48)
            v, i = torch.topk(torch.from_numpy(self.x), k, dim=1)
            self.sp_data, self.sp_index = v.numpy().copy(), i.numpy().astype(np.uint8)
        else:                           # random distinct columns (main.cu:120-133)
            self.sp_index = np.argsort(rng.random((self.n_cols, d)), axis=1)[:, :k].astype(np.uint8)
            self.sp_data = rng.random((self.n_cols, k), dtype=np.float32)
        self.w4 = O.worklist(self.indptr, warp_max_nz)
        self.dy = rng.random((n, d), dtype=np.float32)

    def cuda(self):
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()  # noqa: E731
        return dict(indptr=t(self.indptr), indices=t(self.idx), values=t(self.val), x=t(self.x),
                    sp_data=t(self.sp_data), sp_index=t(self.sp_index), warp4=t(self.w4), dy=t(self.dy))

    def fwd_oracle(self):
        return O.spgemm_forward(self.w4, self.idx, self.val, self.sp_data, self.sp_index, self.d,
                                n_out=self.n)

    def bwd_oracle(self):
        return O.sspmm_backward(self.w4, self.idx, self.val, self.dy, self.sp_index)


def assert_close(got: torch.Tensor, want: np.ndarray, what=""):
    g = got.detach().cpu().numpy().astype(np.float64)
    np.testing.assert_allclose(g, want, rtol=RTOL, atol=ATOL, err_msg=what)


def assert_close_sum(got, want, abs_terms, n_terms, what=""):
    """fp32 sums of `n_terms` products against an fp64 oracle: the north-star tolerance (rtol 1e-5 / atol 1e-6),
    widened to the forward error bound of an fp32 sum, (n_terms + 2) * 2^-24 * sum|terms|, where a result is a
    cancellation of much larger terms (the bound fp32 summation in ANY order -- torch's GEMM included -- obeys)."""
    got = np.asarray(got, np.float64)
    allowed = np.maximum(ATOL + RTOL * np.abs(want), (n_terms + 2) * 2.0 ** -24 * np.asarray(abs_terms, np.float64))
    bad = np.abs(got - want) > allowed
    assert not bad.any(), (what, int(bad.sum()), float(np.abs(got - want).max()))


def topk_distribution_cases(n, seed):
    """Rows that exercise every way the fp32 TopK search can end (tests/test_gpu_parity.py and the SIMT mirror):
    smooth distributions (first probe from order statistics, finish by extraction), heavy tails, values on a coarse
    grid (ties inside the extraction window -> robust path), half-zero rows, denormals, signed zeros, NaN and
    infinities mixed in, values next to the fp32 maximum."""
    rng = np.random.default_rng(seed)
    shape = (n, 256)
    cases = {
        "normal": rng.standard_normal(shape),
        "uniform": rng.random(shape),
        "lognormal": np.exp(3 * rng.standard_normal(shape)),
        "grid16": np.round(rng.standard_normal(shape) * 16) / 16,
        "grid4": np.round(rng.standard_normal(shape) * 4) / 4,
        "relu": np.maximum(rng.standard_normal(shape), 0),
        "denormal": rng.standard_normal(shape) * 1e-41,
        "huge": np.clip(rng.standard_normal(shape) * 1e38, -3.3e38, 3.3e38),
    }
    z = rng.standard_normal(shape)
    z[rng.random(shape) < 0.3] = 0.0
    z[rng.random(shape) < 0.1] = -0.0
    cases["zeros"] = z
    s = rng.standard_normal(shape)
    s[rng.random(shape) < 0.05] = np.nan
    s[rng.random(shape) < 0.02] = np.inf
    s[rng.random(shape) < 0.02] = -np.inf
    cases["nan_inf"] = s
    return {name: np.ascontiguousarray(x, dtype=np.float32) for name, x in cases.items()}
