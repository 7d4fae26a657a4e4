"""Drop-in for the reference's `maxk_cuda_kernels` extension module (Path B:
cuda_kernel_bindings.cpp:429-490).  Same function names, argument order and return
types; the work is done by libmaxk_b200.so (sm_100a) through its C ABI.

Extras beyond the reference surface are keyword-only or new names
(`generate_warp4_metadata`, feat_dim=, num_rows=).
"""
from __future__ import annotations

import os
from typing import List, Tuple

import numpy as np
import torch

from maxk_b200 import _cabi, ops

_cabi.load()  # fail at import time, like a missing compiled extension would

FULL_DIM = ops.FULL_DIM


def spmm_maxk_forward(warp4_metadata, indices, values, input_data, sparse_selector, num_warps,
                      dim_sparse, *, feat_dim: int = FULL_DIM, num_rows=None) -> torch.Tensor:
    """MaxK-GNN forward SPMM kernel (cuda_kernel_bindings.cpp:42-104) -> f32[N, 256]."""
    return ops.spgemm_forward(warp4_metadata, indices, values, input_data, sparse_selector,
                              num_warps, dim_sparse, feat_dim=feat_dim, num_rows=num_rows)


def spmm_maxk_backward(warp4_metadata, indices, values, grad_output, sparse_selector, num_warps,
                       dim_sparse) -> torch.Tensor:
    """MaxK-GNN backward SPMM kernel (cuda_kernel_bindings.cpp:106-161) -> f32[N, k]."""
    return ops.sspmm_backward(warp4_metadata, indices, values, grad_output, sparse_selector,
                              num_warps, dim_sparse)


def cuda_topk_maxk(input: torch.Tensor, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """uint8 TopK (cuda_kernel_bindings.cpp:164-201) -> (values u8[N,k], indices u8[N,k])."""
    return ops.topk_cbsr_u8(input, k)


def cuda_topk_maxk_float(input: torch.Tensor, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """TopK for float tensors (cuda_kernel_bindings.cpp:203-238) -> (values, indices int32).

    Native fp32 selection: values are bit copies of the inputs (the reference quantises to
    uint8 and back); indices ascend within each row."""
    if not isinstance(input, torch.Tensor) or not input.is_cuda:
        raise RuntimeError("Input must be on CUDA")
    if input.dim() != 2:
        raise RuntimeError("Input must be 2D tensor")
    if input.dtype == torch.float32:
        v, i = ops.topk_cbsr(input.contiguous(), k)
        return v, i.to(torch.int32)
    if input.dtype == torch.uint8:
        v, i = ops.topk_cbsr_u8(input.contiguous(), k)
        return v, i.to(torch.int32)
    raise RuntimeError("Input must be float32 or uint8")


def prepare_cbsr_format_maxk(features: torch.Tensor, maxk: int):
    """(sparse_data, sparse_indices) (cuda_kernel_bindings.cpp:240-251)."""
    if not isinstance(features, torch.Tensor) or not features.is_cuda:
        raise RuntimeError("Features must be on CUDA")
    if features.dim() != 2:
        raise RuntimeError("Features must be 2D tensor")
    if not (0 < maxk <= features.size(1)):
        raise RuntimeError("Invalid maxk value")
    return cuda_topk_maxk_float(features, maxk)


def warp4_path(graph_name: str, num_warps: int = 12, warp_max_nz: int = 64) -> str:
    return os.path.join("kernels", f"w{num_warps}_nz{warp_max_nz}_warp_4", graph_name + ".warp4")


def load_warp4_metadata(graph_name: str, num_warps: int = 12, warp_max_nz: int = 64) -> torch.Tensor:
    """Read kernels/w12_nz64_warp_4/<graph>.warp4 (raw int32) onto the GPU
    (cuda_kernel_bindings.cpp:287-317).  Prefer generate_warp4_metadata(indptr)."""
    path = warp4_path(graph_name, num_warps, warp_max_nz)
    if not os.path.exists(path):
        raise RuntimeError("Cannot open warp4 file: " + path)
    data = np.fromfile(path, dtype=np.int32)
    return torch.from_numpy(data).cuda()


def save_warp4_metadata(warp4: torch.Tensor, graph_name: str, num_warps: int = 12,
                        warp_max_nz: int = 64) -> str:
    """Write a work list in the reference's on-disk format (raw little-endian int32 quadruples,
    kernels/generate_meta.py:47-48) so that the reference's C++ bench (spmm_maxk.cu:117) and
    load_warp4_metadata read exactly what the device generator produced.  Returns the path."""
    path = warp4_path(graph_name, num_warps, warp_max_nz)
    os.makedirs(os.path.dirname(path), exist_ok=True)
    warp4.detach().cpu().numpy().astype(np.int32).tofile(path)
    return path


def generate_warp4_metadata(indptr: torch.Tensor, warp_max_nz: int = 64) -> torch.Tensor:
    """On-device replacement for kernels/generate_meta.py: int32[4W] on indptr's device."""
    return ops.build_worklist(indptr, warp_max_nz)


def generate_sparse_selector(num_v: int, dim_origin: int, dim_sparse: int) -> torch.Tensor:
    """k distinct random columns per row, uint8 [num_v, k] on the GPU, seed 123
    (cuda_kernel_bindings.cpp:320-340; one batched argsort instead of num_v randperms)."""
    if not (0 < dim_sparse <= dim_origin <= 256):
        raise RuntimeError("need 0 < dim_sparse <= dim_origin <= 256")
    g = torch.Generator(device="cuda")
    g.manual_seed(123)
    keys = torch.rand(num_v, dim_origin, device="cuda", generator=g)
    return keys.argsort(dim=1)[:, :dim_sparse].to(torch.uint8).contiguous()


class CudaTimer:
    """cudaEvent pair (cuda_kernel_bindings.cpp:343-369): start(), stop() -> ms."""

    def __init__(self):
        self._start = torch.cuda.Event(enable_timing=True)
        self._stop = torch.cuda.Event(enable_timing=True)

    def start(self):
        self._start.record()

    def stop(self) -> float:
        self._stop.record()
        self._stop.synchronize()
        return self._start.elapsed_time(self._stop)


def benchmark_spmm_maxk(warp4_metadata, indices, values, input_data, sparse_selector, num_warps,
                        dim_sparse, num_runs: int = 4) -> List[float]:
    """num_runs warm-up + num_runs timed launches, ms each (cuda_kernel_bindings.cpp:372-402)."""
    for _ in range(num_runs):
        spmm_maxk_forward(warp4_metadata, indices, values, input_data, sparse_selector, num_warps,
                          dim_sparse)
    torch.cuda.synchronize()
    times = []
    timer = CudaTimer()
    for _ in range(num_runs):
        timer.start()
        spmm_maxk_forward(warp4_metadata, indices, values, input_data, sparse_selector, num_warps,
                          dim_sparse)
        times.append(timer.stop())
    return times


def validate_spmm_maxk(warp4_metadata, indices, values, input_data, sparse_selector,
                       reference_output, num_warps, dim_sparse, tolerance: float = 0.001) -> bool:
    """mean |diff| < tolerance vs a caller-supplied reference (cuda_kernel_bindings.cpp:405-427)."""
    out = spmm_maxk_forward(warp4_metadata, indices, values, input_data, sparse_selector, num_warps,
                            dim_sparse, feat_dim=reference_output.shape[1])
    diff = (out - reference_output).abs()
    print(f"Validation - Max diff: {diff.max().item()}, Avg diff: {diff.mean().item()}")
    return diff.mean().item() < tolerance


def cusparse_spmm(indptr, indices, values, input_features, timing: bool = False) -> torch.Tensor:
    """Dense-X CSR SpMM comparator (cuda_kernel_bindings.cpp:253-284; kernels/spmm_cusparse.cu).
    A library call (torch's CSR SpMM -> cuSPARSE), used only to validate / restate the
    README's speed-up ratio -- never on the product path."""
    for name, t in (("indptr", indptr), ("indices", indices), ("values", values),
                    ("input_features", input_features)):
        if not (t.is_cuda and t.is_contiguous()):
            raise RuntimeError(f"{name} must be CUDA and contiguous")
    n = indptr.size(0) - 1
    a = torch.sparse_csr_tensor(indptr.long(), indices.long(), values,
                                size=(n, input_features.size(0)))
    return torch.sparse.mm(a, input_features)
