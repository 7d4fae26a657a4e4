"""CPU restatement of the reference's CPU path for this hot path -- the thing the reference
falls back to (and its upstream models always use) when the CUDA kernels are absent:

    h      = MaxK(X)              torch.topk + scatter_ mask       utils/models.py:48-52
    Y      = A @ h                DGL update_all(copy_u, sum|mean)  utils/models.py:100-111,
                                                                    utils/models_old.py:39
    dX     = mask * (A^T @ dY)    autograd through the two ops      utils/models.py:55-59

DGL is not installable here (no network), so the aggregation is restated with torch's CSR
SpMM (`torch.sparse.mm`), which is the same dense-feature CSR kernel DGL's CPU backend
runs; `values` carry the sum / mean / sym-norm weights.  TEST INFRASTRUCTURE ONLY (see
maxk_oracle.c): timed by bench.py's cpu_baseline leg and by `bench.py --impl reference`
on the GPU box's host cores; never imported by the product.  Parity of this port against
DGL itself is unpinned; it is checked against the C oracle in tests/.
"""
from __future__ import annotations

import os
import time

import torch


def set_threads(n=None) -> int:
    n = n or os.cpu_count() or 1
    torch.set_num_threads(n)
    return torch.get_num_threads()


class CpuAggregation:
    """Holds the CSR adjacency (built once, like the DGL graph object) and runs one forward +
    backward of the MaxK aggregation on the CPU."""

    def __init__(self, indptr, indices, values, n_cols=None):
        self.n_rows = indptr.numel() - 1
        self.n_cols = int(n_cols if n_cols is not None else self.n_rows)
        self.a = torch.sparse_csr_tensor(indptr.cpu().long(), indices.cpu().long(),
                                         values.cpu().float(), size=(self.n_rows, self.n_cols))
        # A^T as CSR for the backward (autograd of sparse.mm would build it every call)
        self.at = self.a.to_sparse_coo().t().coalesce().to_sparse_csr()

    @staticmethod
    def maxk(x, k):
        _, idx = x.topk(k, dim=1)
        mask = torch.zeros_like(x)
        mask.scatter_(1, idx, 1)
        return x * mask, mask

    def prepare(self, x, k):
        """MaxK on the CPU (outside the timed step, like the GPU arm whose step is
        SpGEMM + SSpMM on an already sparsified operand)."""
        return self.maxk(x, k)

    def forward(self, h):
        return torch.sparse.mm(self.a, h)

    def backward(self, dy, mask):
        return torch.sparse.mm(self.at, dy) * mask

    def step(self, h, dy, mask):
        """one forward aggregation + one backward (gradient w.r.t. the MaxK input)"""
        return self.forward(h), self.backward(dy, mask)

    def timed_step(self, h, dy, mask):
        t0 = time.perf_counter()
        out = self.step(h, dy, mask)
        return out, time.perf_counter() - t0
