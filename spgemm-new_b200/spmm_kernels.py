"""Drop-in for the reference's `spmm_kernels` module (Path A: kernels/spmm_bindings.cpp:209-262),
the names utils/models.py imports (models.py:20-26).

The reference implementation of these classes is non-functional (the selector never reaches
the kernel and the work list is read from a file that does not exist); these are
WORKING aliases over the same B200 kernels, with the work list generated on device and
cached per graph.  Conventions kept from the reference: the CALLER allocates the output
tensor and passes it in (models.py:75,121-122); run_kernel returns seconds.
"""
from __future__ import annotations

import time

import torch

from maxk_b200 import _cabi, ops
from maxk_b200.autograd import worklist_for

_cabi.load()


class _SpmmBase:
    def __init__(self, graph_name, indptr, indices, values, input_features, output_features):
        self._graph_name = graph_name
        self._indptr = indptr.int().contiguous()
        self._indices = indices.int().contiguous()
        self._values = values.float().contiguous()
        self._vin = input_features
        self._vout = output_features
        self._selector = None
        self._maxk = None
        self._warp4 = worklist_for(self._indptr)

    def update_input_output(self, input_features, output_features):
        self._vin, self._vout = input_features, output_features

    def set_sparse_params(self, sparse_selector, maxk):
        """sparse_selector: int32 (reference convention, spmm_bindings.cpp:70-77) or uint8."""
        if sparse_selector.dtype != torch.uint8:
            sparse_selector = sparse_selector.to(torch.uint8)
        self._selector = sparse_selector.contiguous()
        self._maxk = int(maxk)

    def get_graph_name(self):
        return self._graph_name

    def _check(self):
        if self._selector is None:
            raise RuntimeError("set_sparse_params() must be called before run_kernel()")

    def _timed(self, fn, timing):
        if not timing:
            fn()
            return 0.0
        for _ in range(4):  # spmm_base.h:56-74: 4 warm-up + 4 timed, wall clock incl. sync
            fn()
        torch.cuda.synchronize()
        t = 0.0
        for _ in range(4):
            t0 = time.perf_counter()
            fn()
            torch.cuda.synchronize()
            t += time.perf_counter() - t0
        return t / 4


class SpmmMaxK(_SpmmBase):
    """input_features = sparse_data f32[N,k]; output_features = f32[N,D] (written in place)."""

    def run_kernel(self, timing=False, dim=-1):
        self._check()
        d = self._vout.shape[1] if dim is None or dim < 0 else int(dim)
        return self._timed(lambda: ops.spgemm_forward(
            self._warp4, self._indices, self._values, self._vin.contiguous(), self._selector,
            dim_sparse=self._maxk, feat_dim=d, num_rows=self._vout.shape[0], out=self._vout), timing)


class SpmmMaxKBackward(_SpmmBase):
    """input_features = grad_output f32[N,D]; output_features = grad_sparse f32[N,k] (in place)."""

    def run_kernel(self, timing=False, dim=-1):
        self._check()
        return self._timed(lambda: ops.sspmm_backward(
            self._warp4, self._indices, self._values, self._vin.contiguous(), self._selector,
            dim_sparse=self._maxk, out=self._vout), timing)


def prepare_cbsr_format(features: torch.Tensor, maxk: int):
    """(sparse_data f32[N,k], sparse_selector int32[N,k]) (spmm_bindings.cpp:163-184)."""
    v, i = ops.topk_cbsr(features.contiguous(), maxk)
    return v, i.to(torch.int32)


def topk_nonlinearity(input: torch.Tensor, k: int) -> torch.Tensor:
    """Dense masked tensor f32[N,D] (spmm_bindings.cpp:189-204)."""
    return ops.topk_cbsr(input.contiguous(), k, want_dense=True, want_cbsr=False)
