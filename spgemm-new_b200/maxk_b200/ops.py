"""Tensor-level operators over the C ABI (libmaxk_b200.so).

Each function validates its arguments the way the reference bindings do with
TORCH_CHECK (cuda_kernel_bindings.cpp:52-62 -> RuntimeError), allocates the outputs with
torch on the inputs' device, and launches on torch's current CUDA stream.  torch is used
for device memory and streams only; every kernel is in csrc/.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Tuple

import torch

from . import _cabi

FULL_DIM = 256  # cuda_kernel_bindings.cpp:70


def _req(cond: bool, msg: str) -> None:
    if not cond:
        raise RuntimeError(msg)


def _cuda(t: torch.Tensor, name: str, dtype=None, contiguous: bool = True) -> None:
    _req(isinstance(t, torch.Tensor), f"{name} must be a tensor")
    _req(t.is_cuda, f"{name} must be CUDA tensor")
    if dtype is not None:
        _req(t.dtype == dtype, f"{name} must be {str(dtype).replace('torch.', '')}")
    if contiguous:
        _req(t.is_contiguous(), f"{name} must be contiguous")


def _stream(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


# ------------------------------------------------------------------------------ a4
def build_worklist(indptr: torch.Tensor, warp_max_nz: int = _cabi.WARP_MAX_NZ) -> torch.Tensor:
    """On-device work list, flat int32[4W] (row, loc, len, 0) -- bit-identical to the
    .warp4 file kernels/generate_meta.py:30-48 writes for the same indptr."""
    lib = _cabi.load()
    _cuda(indptr, "indptr", torch.int32)
    _req(indptr.dim() == 1 and indptr.numel() >= 1, "indptr must be 1-D with N+1 entries")
    n = indptr.numel() - 1
    with torch.cuda.device(indptr.device):
        ws = torch.empty(int(lib.maxk_worklist_workspace_bytes(n)), dtype=torch.uint8,
                         device=indptr.device)
        w = ctypes.c_int64(0)
        _cabi.check(lib.maxk_worklist_count(indptr.data_ptr(), n, warp_max_nz, ws.data_ptr(),
                                            ctypes.byref(w), _stream(indptr)))
        out = torch.empty(4 * w.value, dtype=torch.int32, device=indptr.device)
        _cabi.check(lib.maxk_worklist_fill(indptr.data_ptr(), n, warp_max_nz, ws.data_ptr(),
                                           out.data_ptr(), w.value, _stream(indptr)))
    return out


# ------------------------------------------------------------------------------ a1
def topk_cbsr(x: torch.Tensor, k: int, want_dense: bool = False, want_cbsr: bool = True, out=None):
    """fp32 MaxK: (sp_data f32[N,k], sp_index u8[N,k][, dense f32[N,D]]).
    out=(sp_data, sp_index): write the CBSR pair into caller-provided contiguous tensors."""
    lib = _cabi.load()
    _cuda(x, "Input", None)
    _req(x.dim() == 2, "Input must be 2D tensor")
    _req(x.dtype == torch.float32, "Input must be float32")
    n, d = x.shape
    _req(0 < k <= d, "Invalid k value")
    sp_data = sp_index = dense = None
    with torch.cuda.device(x.device):
        if want_cbsr and out is not None:
            sp_data, sp_index = out
            _cuda(sp_data, "out[0]", torch.float32)
            _cuda(sp_index, "out[1]", torch.uint8)
            _req(tuple(sp_data.shape) == (n, k) and tuple(sp_index.shape) == (n, k),
                 "out tensors must be [N, k]")
        elif want_cbsr:
            sp_data = torch.empty((n, k), dtype=torch.float32, device=x.device)
            sp_index = torch.empty((n, k), dtype=torch.uint8, device=x.device)
        if want_dense:
            dense = torch.empty_like(x)
        _cabi.check(lib.maxk_topk_cbsr_f32(x.data_ptr(), n, d, k, _ptr(sp_data), _ptr(sp_index),
                                           _ptr(dense), _stream(x)))
    if want_cbsr and want_dense:
        return sp_data, sp_index, dense
    if want_dense:
        return dense
    return sp_data, sp_index


_LT_WORKSPACE = {}


def linear_topk_supported(x: torch.Tensor, weight: torch.Tensor) -> bool:
    """Shapes the fused tcgen05 kernel takes (csrc/linear_topk.cu): fp32, output width 256, in_dim % 4 == 0."""
    return (x.is_cuda and x.dtype == torch.float32 and weight.dtype == torch.float32 and x.dim() == 2
            and weight.dim() == 2 and weight.shape[0] == FULL_DIM and weight.shape[1] == x.shape[1]
            and x.shape[1] % 4 == 0 and x.shape[1] >= 4)


def linear_topk_cbsr(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], k: int,
                     want_linear: bool = False, want_cbsr: bool = True):
    """(sp_data, sp_index[, linear_out]) = TopK_k(x @ weight^T + bias) in one tcgen05 kernel: the dense [N, 256]
    activation never reaches HBM (include/maxk_b200.h: maxk_linear_topk_cbsr; ref utils/models.py:336-341)."""
    lib = _cabi.load()
    _cuda(x, "x", torch.float32)
    _cuda(weight, "weight", torch.float32)
    _req(x.dim() == 2 and weight.dim() == 2 and weight.shape[1] == x.shape[1], "x [N, in] and weight [D, in] expected")
    n, in_dim = x.shape
    d = weight.shape[0]
    _req(0 < k <= d, "Invalid k value")
    if bias is not None:
        _cuda(bias, "bias", torch.float32)
        _req(bias.numel() == d, "bias must have D elements")
    sp_data = sp_index = lin = None
    with torch.cuda.device(x.device):
        if want_cbsr:
            sp_data = torch.empty((n, k), dtype=torch.float32, device=x.device)
            sp_index = torch.empty((n, k), dtype=torch.uint8, device=x.device)
        if want_linear:
            lin = torch.empty((n, d), dtype=torch.float32, device=x.device)
        ws = _lt_workspace(lib, x, d, in_dim)
        _cabi.check(lib.maxk_linear_topk_cbsr(x.data_ptr(), n, in_dim, weight.data_ptr(), _ptr(bias), d, k,
                                              _ptr(sp_data), _ptr(sp_index), _ptr(lin), ws.data_ptr(), _stream(x)))
    out = tuple(t for t in (sp_data, sp_index, lin) if t is not None)
    return out if len(out) > 1 else out[0]


def _lt_workspace(lib, x, d, in_dim):
    key = (x.device.index, d, in_dim, torch.cuda.current_stream(x.device).cuda_stream)
    ws = _LT_WORKSPACE.get(key)
    if ws is None:   # one workspace per (device, shape, stream): the split weight is rewritten by every call
        ws = torch.empty(int(lib.maxk_linear_topk_workspace_bytes(d, in_dim)), dtype=torch.uint8, device=x.device)
        _LT_WORKSPACE[key] = ws
    return ws


def linear_f32(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor] = None,
               residual: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x @ weight^T + bias (+ residual [N, 256], added in the GEMM's epilogue) on the tensor cores at fp32 accuracy
    (split TF32 products; csrc/linear_topk.cu without the selection).  weight [256, in]."""
    lib = _cabi.load()
    _req(linear_topk_supported(x, weight), "linear_f32: fp32 CUDA tensors, x [N, in], weight [256, in], in % 4 == 0")
    x, weight = x.contiguous(), weight.contiguous()
    n, in_dim = x.shape
    with torch.cuda.device(x.device):
        out = torch.empty((n, FULL_DIM), dtype=torch.float32, device=x.device)
        if bias is not None:
            bias = bias.contiguous()
        if residual is not None:
            _cuda(residual, "residual", torch.float32)
            _req(tuple(residual.shape) == (n, FULL_DIM), "residual must be [N, 256]")
            residual = residual.contiguous()
        _cabi.check(lib.maxk_linear_f32(x.data_ptr(), n, in_dim, weight.data_ptr(), _ptr(bias), _ptr(residual), FULL_DIM,
                                        out.data_ptr(), _lt_workspace(lib, x, FULL_DIM, in_dim).data_ptr(), _stream(x)))
    return out


def linear_wgrad_supported(grad_out: torch.Tensor, x: torch.Tensor) -> bool:
    return (grad_out.is_cuda and grad_out.dtype == torch.float32 and x.dtype == torch.float32 and grad_out.dim() == 2
            and x.dim() == 2 and grad_out.shape[1] == FULL_DIM and x.shape[0] == grad_out.shape[0]
            and 4 <= x.shape[1] <= FULL_DIM and x.shape[1] % 4 == 0)


def linear_wgrad(grad_out: torch.Tensor, x: torch.Tensor, out: Optional[torch.Tensor] = None,
                 accumulate: bool = False) -> torch.Tensor:
    """grad_out^T @ x -> [256, in]: the weight gradient of a Linear, reduction over the rows on the tensor cores
    (csrc/linear_wgrad.cu)."""
    lib = _cabi.load()
    _req(linear_wgrad_supported(grad_out, x), "linear_wgrad: fp32 CUDA tensors, grad_out [N, 256], x [N, in <= 256], in % 4 == 0")
    grad_out, x = grad_out.contiguous(), x.contiguous()
    n, in_dim = x.shape
    with torch.cuda.device(x.device):
        if out is None:
            _req(not accumulate, "accumulate=True needs an `out` tensor")
            out = torch.empty((FULL_DIM, in_dim), dtype=torch.float32, device=x.device)
        _cuda(out, "out", torch.float32)
        _req(tuple(out.shape) == (FULL_DIM, in_dim), "out must be [256, in]")
        _cabi.check(lib.maxk_linear_wgrad(grad_out.data_ptr(), x.data_ptr(), n, FULL_DIM, in_dim, out.data_ptr(),
                                          1 if accumulate else 0, _stream(x)))
    return out


def layernorm_supported(x: torch.Tensor) -> bool:
    return x.is_cuda and x.dtype == torch.float32 and x.dim() == 2 and 4 <= x.shape[1] <= FULL_DIM and x.shape[1] % 4 == 0


def layernorm_forward(x: torch.Tensor, gamma: Optional[torch.Tensor], beta: Optional[torch.Tensor], eps: float = 1e-5):
    """-> (y, mean, rstd): nn.LayerNorm over the last dimension of x [N, d <= 256] (csrc/layernorm.cu)."""
    lib = _cabi.load()
    _req(layernorm_supported(x), "layernorm_forward: fp32 CUDA tensor [N, d], d a multiple of 4 in [4, 256]")
    x = x.contiguous()
    n, d = x.shape
    with torch.cuda.device(x.device):
        y = torch.empty_like(x)
        mean = torch.empty(n, dtype=torch.float32, device=x.device)
        rstd = torch.empty(n, dtype=torch.float32, device=x.device)
        gamma = None if gamma is None else gamma.contiguous()
        beta = None if beta is None else beta.contiguous()
        _cabi.check(lib.maxk_layernorm_forward(x.data_ptr(), n, d, _ptr(gamma), _ptr(beta), float(eps), y.data_ptr(),
                                               mean.data_ptr(), rstd.data_ptr(), _stream(x)))
    return y, mean, rstd


def layernorm_backward(g: torch.Tensor, x: torch.Tensor, mean: torch.Tensor, rstd: torch.Tensor,
                       gamma: Optional[torch.Tensor], want_affine_grads: bool = True):
    """-> (dx, dgamma, dbeta) of layernorm_forward."""
    lib = _cabi.load()
    g, x = g.contiguous(), x.contiguous()
    n, d = x.shape
    with torch.cuda.device(x.device):
        dx = torch.empty_like(x)
        dgamma = dbeta = None
        if want_affine_grads:
            dgamma = torch.empty(d, dtype=torch.float32, device=x.device)
            dbeta = torch.empty(d, dtype=torch.float32, device=x.device)
        gamma = None if gamma is None else gamma.contiguous()
        _cabi.check(lib.maxk_layernorm_backward(g.data_ptr(), x.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                                                _ptr(gamma), n, d, dx.data_ptr(), _ptr(dgamma), _ptr(dbeta), _stream(x)))
    return dx, dgamma, dbeta


def layernorm_topk_cbsr(x: torch.Tensor, gamma: Optional[torch.Tensor], beta: Optional[torch.Tensor], eps: float, k: int,
                        want_dense: bool = False):
    """-> (sp_data, sp_index, mean, rstd[, y]) = TopK_k(LayerNorm(x)) in one pass (csrc/topk_cbsr.cu)."""
    lib = _cabi.load()
    _req(layernorm_supported(x), "layernorm_topk_cbsr: fp32 CUDA tensor [N, d], d a multiple of 4 in [4, 256]")
    x = x.contiguous()
    n, d = x.shape
    _req(0 < k <= d, "Invalid k value")
    with torch.cuda.device(x.device):
        sp_data = torch.empty((n, k), dtype=torch.float32, device=x.device)
        sp_index = torch.empty((n, k), dtype=torch.uint8, device=x.device)
        mean = torch.empty(n, dtype=torch.float32, device=x.device)
        rstd = torch.empty(n, dtype=torch.float32, device=x.device)
        y = torch.empty_like(x) if want_dense else None
        gamma = None if gamma is None else gamma.contiguous()
        beta = None if beta is None else beta.contiguous()
        _cabi.check(lib.maxk_layernorm_topk_cbsr(x.data_ptr(), n, d, _ptr(gamma), _ptr(beta), float(eps), k,
                                                 sp_data.data_ptr(), sp_index.data_ptr(), mean.data_ptr(),
                                                 rstd.data_ptr(), _ptr(y), _stream(x)))
    return (sp_data, sp_index, mean, rstd, y) if want_dense else (sp_data, sp_index, mean, rstd)


def layernorm_backward_cbsr(g_data: torch.Tensor, sp_index: torch.Tensor, x: torch.Tensor, mean: torch.Tensor,
                            rstd: torch.Tensor, gamma: Optional[torch.Tensor], want_affine_grads: bool = True):
    """-> (dx, dgamma, dbeta): LayerNorm backward for an upstream gradient that is k-sparse per row (CBSR layout)."""
    lib = _cabi.load()
    g_data, x = g_data.contiguous(), x.contiguous()
    n, d = x.shape
    k = sp_index.shape[1]
    with torch.cuda.device(x.device):
        dx = torch.empty_like(x)
        dgamma = dbeta = None
        if want_affine_grads:
            dgamma = torch.empty(d, dtype=torch.float32, device=x.device)
            dbeta = torch.empty(d, dtype=torch.float32, device=x.device)
        gamma = None if gamma is None else gamma.contiguous()
        _cabi.check(lib.maxk_layernorm_backward_cbsr(g_data.data_ptr(), sp_index.data_ptr(), k, x.data_ptr(),
                                                     mean.data_ptr(), rstd.data_ptr(), _ptr(gamma), n, d, dx.data_ptr(),
                                                     _ptr(dgamma), _ptr(dbeta), _stream(x)))
    return dx, dgamma, dbeta


def topk_cbsr_u8(x: torch.Tensor, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
    lib = _cabi.load()
    _cuda(x, "Input", None)
    _req(x.dim() == 2, "Input must be 2D tensor")
    _req(x.dtype == torch.uint8, "Input must be uint8 tensor")
    n, d = x.shape
    _req(0 < k <= d, "Invalid k value")
    with torch.cuda.device(x.device):
        val = torch.empty((n, k), dtype=torch.uint8, device=x.device)
        ind = torch.empty((n, k), dtype=torch.uint8, device=x.device)
        _cabi.check(lib.maxk_topk_cbsr_u8(x.data_ptr(), n, d, k, val.data_ptr(), ind.data_ptr(),
                                          _stream(x)))
    return val, ind


# ------------------------------------------------------------------------------ a2
def spgemm_forward(warp4: torch.Tensor, indices: torch.Tensor, values: torch.Tensor,
                   sp_data: torch.Tensor, sp_index: torch.Tensor, num_warps: Optional[int] = None,
                   dim_sparse: Optional[int] = None, feat_dim: int = FULL_DIM,
                   num_rows: Optional[int] = None, out: Optional[torch.Tensor] = None,
                   accumulate: bool = False) -> torch.Tensor:
    lib = _cabi.load()
    _cuda(warp4, "warp4_metadata", torch.int32)
    _cuda(indices, "indices", torch.int32)
    _cuda(values, "values", torch.float32)
    _cuda(sp_data, "input_data", torch.float32)
    _cuda(sp_index, "sparse_selector", torch.uint8)
    _req(sp_data.dim() == 2 and sp_index.shape == sp_data.shape,
         "input_data and sparse_selector must both be [N, k]")
    k = sp_data.shape[1]
    _req(dim_sparse is None or dim_sparse == k, f"dim_sparse={dim_sparse} != input_data.size(1)={k}")
    w = warp4.numel() // 4
    _req(num_warps is None or 0 <= num_warps <= w, "num_warps exceeds the work list")
    w = w if num_warps is None else num_warps
    _req(indices.numel() == values.numel(), "indices and values differ in length")
    n_out = sp_data.shape[0] if num_rows is None else int(num_rows)  # bindings.cpp:65
    with torch.cuda.device(sp_data.device):
        if out is None:
            _req(not accumulate, "accumulate=True needs an `out` tensor")
            out = torch.empty((n_out, feat_dim), dtype=torch.float32, device=sp_data.device)
        else:
            _cuda(out, "out", torch.float32)
            _req(tuple(out.shape) == (n_out, feat_dim), "out has the wrong shape")
        _cabi.check(lib.maxk_spgemm_forward_acc(warp4.data_ptr(), w, indices.data_ptr(),
                                                values.data_ptr(), sp_data.data_ptr(),
                                                sp_index.data_ptr(), k, out.data_ptr(), n_out,
                                                feat_dim, 1 if accumulate else 0, _stream(sp_data)))
    return out


# ---------------------------------------------------- a2 through a bank plan (k = 32, 64)
L2_OPERAND_BYTES = 96 << 20   # a CBSR operand larger than this does not stay in the 126 MB L2


def use_bank_plan(k: int, num_tasks: int, num_rows: int, num_cbsr_rows: int) -> bool:
    """Whether the forward SpGEMM should go through the bank plan: k = 32 or 64, rows of several tasks
    (W >= 2N: the doubled flush of the two-home accumulator is amortised over >= 128 edges) and a CBSR
    operand that lives in the L2 (then the update's shared-memory bank conflicts are the bound; on a
    DRAM-bound operand they are hidden).  Measured (profiles/r02_variants.md): Reddit shape k = 32
    3.99 -> 3.1 ms, k = 64 7.1 -> 5.9 ms; products / Flickr shapes: no gain."""
    return (k in (32, 64) and num_tasks >= 2 * num_rows and num_cbsr_rows * 6 * k <= L2_OPERAND_BYTES)


def cbsr_bank_plan(sp_index: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """sp_index uint8 [N, k] -> sp_addr uint16 [N, k] (csrc/cbsr_plan.cu): the shared-memory accumulator
    word each entry uses in `spgemm_forward_planned`, one of the two homes of its column, chosen per row so
    that the 32 entries of one update spread over the 32 banks."""
    lib = _cabi.load()
    _cuda(sp_index, "sparse_selector", torch.uint8)
    _req(sp_index.dim() == 2, "sparse_selector must be [N, k]")
    n, k = sp_index.shape
    with torch.cuda.device(sp_index.device):
        if out is None:
            out = torch.empty((n, k), dtype=torch.uint16, device=sp_index.device)
        else:
            _cuda(out, "out", torch.uint16)
            _req(tuple(out.shape) == (n, k), "out has the wrong shape")
        _cabi.check(lib.maxk_cbsr_bank_plan(sp_index.data_ptr(), n, k, out.data_ptr(), _stream(sp_index)))
    return out


def spgemm_forward_planned(warp4: torch.Tensor, indices: torch.Tensor, values: torch.Tensor,
                           sp_data: torch.Tensor, sp_addr: torch.Tensor,
                           feat_dim: int = FULL_DIM, num_rows: Optional[int] = None,
                           out: Optional[torch.Tensor] = None, accumulate: bool = False) -> torch.Tensor:
    """`spgemm_forward` with the bank plan `sp_addr = cbsr_bank_plan(sp_index)` in place of the selectors
    (k = 32, 64).  Same output contract."""
    lib = _cabi.load()
    _cuda(warp4, "warp4_metadata", torch.int32)
    _cuda(indices, "indices", torch.int32)
    _cuda(values, "values", torch.float32)
    _cuda(sp_data, "input_data", torch.float32)
    _cuda(sp_addr, "sp_addr", torch.uint16)
    _req(sp_data.dim() == 2 and sp_addr.shape == sp_data.shape, "input_data and sp_addr must both be [N, k]")
    k = sp_data.shape[1]
    w = warp4.numel() // 4
    _req(indices.numel() == values.numel(), "indices and values differ in length")
    n_out = sp_data.shape[0] if num_rows is None else int(num_rows)
    with torch.cuda.device(sp_data.device):
        if out is None:
            _req(not accumulate, "accumulate=True needs an `out` tensor")
            out = torch.empty((n_out, feat_dim), dtype=torch.float32, device=sp_data.device)
        else:
            _cuda(out, "out", torch.float32)
            _req(tuple(out.shape) == (n_out, feat_dim), "out has the wrong shape")
        _cabi.check(lib.maxk_spgemm_forward_planned(warp4.data_ptr(), w, indices.data_ptr(), values.data_ptr(),
                                                    sp_data.data_ptr(), sp_addr.data_ptr(), k,
                                                    out.data_ptr(), n_out, feat_dim, 1 if accumulate else 0,
                                                    _stream(sp_data)))
    return out


def spgemm_forward_auto(warp4: torch.Tensor, indices: torch.Tensor, values: torch.Tensor,
                        sp_data: torch.Tensor, sp_index: torch.Tensor, feat_dim: int = FULL_DIM,
                        num_rows: Optional[int] = None, out: Optional[torch.Tensor] = None,
                        plan_buf: Optional[torch.Tensor] = None) -> torch.Tensor:
    """The forward SpGEMM as the layers call it: through the bank plan where `use_bank_plan` says it pays
    (the plan kernel runs here, in front of the SpGEMM -- it is part of the forward's cost), else directly."""
    k = sp_data.shape[1]
    n_out = sp_data.shape[0] if num_rows is None else int(num_rows)
    if sp_index.dtype == torch.uint8 and use_bank_plan(k, warp4.numel() // 4, n_out, sp_data.shape[0]):
        sp_addr = cbsr_bank_plan(sp_index, out=plan_buf)
        return spgemm_forward_planned(warp4, indices, values, sp_data, sp_addr, feat_dim=feat_dim,
                                      num_rows=num_rows, out=out)
    return spgemm_forward(warp4, indices, values, sp_data, sp_index, feat_dim=feat_dim, num_rows=num_rows, out=out)


def sspmm_backward_planned(warp4: torch.Tensor, indices: torch.Tensor, values: torch.Tensor,
                           grad_output: torch.Tensor, sp_addr: torch.Tensor,
                           out: Optional[torch.Tensor] = None, accumulate: bool = False) -> torch.Tensor:
    """`sspmm_backward` reading its columns out of the bank plan (the multi-GPU peer path ships the plan in
    place of the selectors).  Same output contract."""
    lib = _cabi.load()
    _cuda(warp4, "warp4_metadata", torch.int32)
    _cuda(indices, "indices", torch.int32)
    _cuda(values, "values", torch.float32)
    _cuda(grad_output, "grad_output", torch.float32)
    _cuda(sp_addr, "sp_addr", torch.uint16)
    _req(grad_output.dim() == 2 and sp_addr.dim() == 2, "grad_output and sp_addr must be 2-D")
    n_cbsr, k = sp_addr.shape
    d = grad_output.shape[1]
    w = warp4.numel() // 4
    with torch.cuda.device(grad_output.device):
        if out is None:
            _req(not accumulate, "accumulate=True needs an `out` tensor")
            out = torch.empty((n_cbsr, k), dtype=torch.float32, device=grad_output.device)
        else:
            _cuda(out, "out", torch.float32)
            _req(tuple(out.shape) == (n_cbsr, k), "out has the wrong shape")
        _cabi.check(lib.maxk_sspmm_backward_planned(warp4.data_ptr(), w, indices.data_ptr(), values.data_ptr(),
                                                    grad_output.data_ptr(), sp_addr.data_ptr(), k, d,
                                                    out.data_ptr(), n_cbsr, 1 if accumulate else 0,
                                                    _stream(grad_output)))
    return out


# ------------------------------------------------------------------------------ a3
def sspmm_backward(warp4: torch.Tensor, indices: torch.Tensor, values: torch.Tensor,
                   grad_output: torch.Tensor, sp_index: torch.Tensor,
                   num_warps: Optional[int] = None, dim_sparse: Optional[int] = None,
                   out: Optional[torch.Tensor] = None, accumulate: bool = False) -> torch.Tensor:
    lib = _cabi.load()
    _cuda(warp4, "warp4_metadata", torch.int32)
    _cuda(indices, "indices", torch.int32)
    _cuda(values, "values", torch.float32)
    _cuda(grad_output, "grad_output", torch.float32)
    _cuda(sp_index, "sparse_selector", torch.uint8)
    _req(grad_output.dim() == 2 and sp_index.dim() == 2, "grad_output and sparse_selector must be 2-D")
    k = sp_index.shape[1]
    _req(dim_sparse is None or dim_sparse == k, f"dim_sparse={dim_sparse} != sparse_selector.size(1)={k}")
    d = grad_output.shape[1]
    w = warp4.numel() // 4
    _req(num_warps is None or 0 <= num_warps <= w, "num_warps exceeds the work list")
    w = w if num_warps is None else num_warps
    n_cbsr = sp_index.shape[0]
    with torch.cuda.device(grad_output.device):
        if out is None:
            _req(not accumulate, "accumulate=True needs an `out` tensor")
            out = torch.empty((n_cbsr, k), dtype=torch.float32, device=grad_output.device)
        else:
            _cuda(out, "out", torch.float32)
            _req(tuple(out.shape) == (n_cbsr, k), "out has the wrong shape")
        _cabi.check(lib.maxk_sspmm_backward(warp4.data_ptr(), w, indices.data_ptr(),
                                            values.data_ptr(), grad_output.data_ptr(),
                                            sp_index.data_ptr(), k, d, out.data_ptr(), n_cbsr,
                                            1 if accumulate else 0, _stream(grad_output)))
    return out


# ------------------------------------------------------------------------------ a5
def cbsr_scatter(sp_data: torch.Tensor, sp_index: torch.Tensor, feat_dim: int = FULL_DIM) -> torch.Tensor:
    lib = _cabi.load()
    _cuda(sp_data, "sp_data", torch.float32)
    _cuda(sp_index, "sp_index", torch.uint8)
    _req(sp_data.shape == sp_index.shape and sp_data.dim() == 2, "sp_data/sp_index must be [N,k]")
    n, k = sp_data.shape
    with torch.cuda.device(sp_data.device):
        dense = torch.empty((n, feat_dim), dtype=torch.float32, device=sp_data.device)
        _cabi.check(lib.maxk_cbsr_scatter(sp_data.data_ptr(), sp_index.data_ptr(), n, k, feat_dim,
                                          dense.data_ptr(), _stream(sp_data)))
    return dense


def cbsr_gather(dense: torch.Tensor, sp_index: torch.Tensor) -> torch.Tensor:
    lib = _cabi.load()
    _cuda(dense, "dense", torch.float32)
    _cuda(sp_index, "sp_index", torch.uint8)
    n, d = dense.shape
    k = sp_index.shape[1]
    _req(sp_index.shape[0] == n, "row count mismatch")
    with torch.cuda.device(dense.device):
        out = torch.empty((n, k), dtype=torch.float32, device=dense.device)
        _cabi.check(lib.maxk_cbsr_gather(dense.data_ptr(), sp_index.data_ptr(), n, k, d,
                                         out.data_ptr(), _stream(dense)))
    return out


def cbsr_matmul(sp_data: torch.Tensor, sp_index: torch.Tensor, m: torch.Tensor,
                bias: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None,
                accumulate: bool = False) -> torch.Tensor:
    """out[N, C] (+)= bias + scatter(sp_data, sp_index) @ m without the dense
    twin; m float32 [D, C] row-major (a Linear's weight transposed), D <= 256, C % 4 == 0."""
    lib = _cabi.load()
    _cuda(sp_data, "sp_data", torch.float32)
    _cuda(sp_index, "sp_index", torch.uint8)
    _cuda(m, "m", torch.float32)
    n, k = sp_index.shape
    d, c = m.shape
    _req(tuple(sp_data.shape) == (n, k), "sp_data / sp_index shape mismatch")
    if bias is not None:
        _cuda(bias, "bias", torch.float32)
        _req(tuple(bias.shape) == (c,), "bias must be [C]")
    with torch.cuda.device(m.device):
        if out is None:
            _req(not accumulate, "accumulate needs out=")
            out = torch.empty((n, c), dtype=torch.float32, device=m.device)
        else:
            _cuda(out, "out", torch.float32)
            _req(tuple(out.shape) == (n, c), "out must be [N, C]")
        _cabi.check(lib.maxk_cbsr_matmul(sp_data.data_ptr(), sp_index.data_ptr(), n, k, m.data_ptr(), d, c,
                                         _ptr(bias), out.data_ptr(), 1 if accumulate else 0, _stream(m)))
    return out


def cbsr_axpy_(y: torch.Tensor, sp_data: torch.Tensor, sp_index: torch.Tensor, alpha: float = 1.0,
               alpha_dev: Optional[torch.Tensor] = None) -> torch.Tensor:
    """y[i, sp_index[i,j]] += (alpha + alpha_dev[0]) * sp_data[i,j], in place."""
    lib = _cabi.load()
    _cuda(y, "y", torch.float32)
    _cuda(sp_data, "sp_data", torch.float32)
    _cuda(sp_index, "sp_index", torch.uint8)
    n, k = sp_index.shape
    _req(tuple(sp_data.shape) == (n, k) and y.dim() == 2 and y.shape[0] == n, "shape mismatch")
    if alpha_dev is not None:
        _cuda(alpha_dev, "alpha_dev", torch.float32)
        _req(alpha_dev.numel() == 1, "alpha_dev must hold one float")
    with torch.cuda.device(y.device):
        _cabi.check(lib.maxk_cbsr_axpy(sp_data.data_ptr(), sp_index.data_ptr(), n, k, y.shape[1], _ptr(alpha_dev),
                                       float(alpha), y.data_ptr(), _stream(y)))
    return y


def cbsr_sampled_dot(g: torch.Tensor, m: torch.Tensor, sp_index: torch.Tensor,
                     out: Optional[torch.Tensor] = None, accumulate: bool = False) -> torch.Tensor:
    """out[i, j] (+)= <g[i, :], m[sp_index[i, j], :]>  (float32 [N, k]) --
    the gradient of cbsr_matmul with respect to sp_data."""
    lib = _cabi.load()
    _cuda(g, "g", torch.float32)
    _cuda(m, "m", torch.float32)
    _cuda(sp_index, "sp_index", torch.uint8)
    n, k = sp_index.shape
    d, c = m.shape
    _req(tuple(g.shape) == (n, c), "g must be [N, C]")
    with torch.cuda.device(m.device):
        if out is None:
            _req(not accumulate, "accumulate needs out=")
            out = torch.empty((n, k), dtype=torch.float32, device=m.device)
        else:
            _cuda(out, "out", torch.float32)
            _req(tuple(out.shape) == (n, k), "out must be [N, k]")
        _cabi.check(lib.maxk_cbsr_sampled_dot(g.data_ptr(), n, c, m.data_ptr(), d, sp_index.data_ptr(), k,
                                              out.data_ptr(), 1 if accumulate else 0, _stream(m)))
    return out


def cbsr_mask(dense: torch.Tensor, sp_index: torch.Tensor) -> torch.Tensor:
    """dense * mask(sp_index)  (MaxK.backward, utils/models.py:55-59)."""
    lib = _cabi.load()
    _cuda(dense, "dense", torch.float32)
    _cuda(sp_index, "sp_index", torch.uint8)
    n, d = dense.shape
    k = sp_index.shape[1]
    _req(sp_index.shape[0] == n, "row count mismatch")
    with torch.cuda.device(dense.device):
        out = torch.empty_like(dense)
        _cabi.check(lib.maxk_cbsr_mask(dense.data_ptr(), sp_index.data_ptr(), n, k, d,
                                       out.data_ptr(), _stream(dense)))
    return out


# -------------------------------------------------------------- multi-GPU wire format
def cbsr_record_bytes(k: int) -> int:
    return int(_cabi.load().maxk_cbsr_record_bytes(k))


def cbsr_pack(sp_data: torch.Tensor, sp_index: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    lib = _cabi.load()
    _cuda(sp_data, "sp_data", torch.float32)
    _cuda(sp_index, "sp_index", torch.uint8)
    n, k = sp_data.shape
    rb = cbsr_record_bytes(k)
    with torch.cuda.device(sp_data.device):
        if out is None:
            out = torch.empty((n, rb), dtype=torch.uint8, device=sp_data.device)
        _req(out.is_contiguous() and out.numel() >= n * rb, "records buffer too small")
        _cabi.check(lib.maxk_cbsr_pack(sp_data.data_ptr(), sp_index.data_ptr(), n, k,
                                       out.data_ptr(), _stream(sp_data)))
    return out


def cbsr_unpack(records: torch.Tensor, n: int, k: int):
    lib = _cabi.load()
    _cuda(records, "records", torch.uint8)
    _req(records.numel() >= n * cbsr_record_bytes(k), "records buffer too small")
    with torch.cuda.device(records.device):
        sp_data = torch.empty((n, k), dtype=torch.float32, device=records.device)
        sp_index = torch.empty((n, k), dtype=torch.uint8, device=records.device)
        _cabi.check(lib.maxk_cbsr_unpack(records.data_ptr(), n, k, sp_data.data_ptr(),
                                         sp_index.data_ptr(), _stream(records)))
    return sp_data, sp_index


def set_tuning(fwd_tasks_per_warp: int = 0, bwd_tasks_per_warp: int = 0) -> None:
    _cabi.check(_cabi.load().maxk_set_tuning(fwd_tasks_per_warp, bwd_tasks_per_warp))


def set_forward_depth(depth: int = 0) -> None:
    """Launch shape of the forward kernel (include/maxk_b200.h: maxk_set_forward_depth; 0 = automatic); bench/test hook."""
    _cabi.check(_cabi.load().maxk_set_forward_depth(int(depth)))


def set_topk_group(on: int = -1) -> None:
    """Grouped fp32 TopK kernel (several rows per warp, d = 128 / 64): -1 automatic, 0 off, 1 on; bench/test hook."""
    _cabi.check(_cabi.load().maxk_set_topk_group(int(on)))


def set_forward_copies(copies: int = 0) -> None:
    """Accumulator copies of the forward kernel (0 = automatic); bench/test hook."""
    _cabi.check(_cabi.load().maxk_set_forward_copies(int(copies)))
