"""Autograd Functions of the reference's model layer, on the B200 kernels.

    MaxK.apply(input, k)                                   utils/models.py:28-59
    SpGEMMFunction.apply(features, (indptr, indices, values), maxk)   utils/models.py:61-149

Same signatures and return shapes.  Differences, all deliberate:
  * no try/except fallbacks: a failing kernel raises (Q13);
  * SpGEMMFunction.backward returns the true gradient mask * (A^T dY), built on device by
    the SSpMM kernel + a scatter kernel instead of N*k .item() calls (models.py:136-141);
  * the work list is generated on device and cached per graph instead of being re-read
    from a file inside every call (Q10).
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Optional

import torch
from torch.autograd import Function

from . import ops

_WL_CACHE: "OrderedDict[tuple, tuple]" = OrderedDict()
_WL_CACHE_MAX = 16


def worklist_for(indptr: torch.Tensor) -> torch.Tensor:
    """Work list of a CSR graph, cached per indptr tensor (the cache holds a reference to the
    tensor, so its storage -- and therefore the key -- cannot be recycled while cached)."""
    # keyed on the CALLER's tensor: an int64 or strided indptr is converted once, not on every call
    key = (indptr.data_ptr(), indptr.numel(), str(indptr.device), indptr._version, indptr.dtype, indptr.stride())
    hit = _WL_CACHE.get(key)
    if hit is not None:
        _WL_CACHE.move_to_end(key)
        return hit[1]
    warp4 = ops.build_worklist(indptr.to(torch.int32).contiguous())
    _WL_CACHE[key] = (indptr, warp4)
    while len(_WL_CACHE) > _WL_CACHE_MAX:
        _WL_CACHE.popitem(last=False)
    return warp4


class MaxK(Function):
    """MaxK nonlinearity: keep the k largest entries of every row, zero the rest."""

    @staticmethod
    def forward(ctx, input, k=1):
        x = input.contiguous()
        _, sp_index, dense = ops.topk_cbsr(x, k, want_dense=True)
        ctx.save_for_backward(sp_index)
        return dense

    @staticmethod
    def backward(ctx, grad_output):
        (sp_index,) = ctx.saved_tensors
        return ops.cbsr_mask(grad_output.contiguous(), sp_index), None


class SpGEMMFunction(Function):
    """Y = A @ MaxK_k(features) through TopK->CBSR + forward SpGEMM; backward through the
    sampled SSpMM.  `values` are taken as given (1 = sum, 1/deg = mean, 1/sqrt(d_i d_j) = GCN;
    SURVEY Q11)."""

    @staticmethod
    def forward(ctx, features, graph_data, maxk):
        indptr, indices, values = graph_data
        x = features.contiguous()
        warp4 = worklist_for(indptr)
        indices = indices.int().contiguous() if indices.dtype != torch.int32 else indices.contiguous()
        values = values.float().contiguous() if values.dtype != torch.float32 else values.contiguous()
        sp_data, sp_index = ops.topk_cbsr(x, maxk)
        n_rows = indptr.numel() - 1
        out = ops.spgemm_forward_auto(warp4, indices, values, sp_data, sp_index,
                                      feat_dim=x.shape[1], num_rows=n_rows)
        ctx.save_for_backward(warp4, indices, values, sp_index)
        ctx.feat_dim = x.shape[1]
        return out

    @staticmethod
    def backward(ctx, grad_output):
        warp4, indices, values, sp_index = ctx.saved_tensors
        dxc = ops.sspmm_backward(warp4, indices, values, grad_output.contiguous(), sp_index)
        return ops.cbsr_scatter(dxc, sp_index, ctx.feat_dim), None, None


class CBSRAggregate(Function):
    """CBSR-native aggregation for layers that keep activations sparse end to end:
    Y = A @ scatter(sp_data, sp_index); gradient flows to sp_data only (CBSR layout).  The forward goes
    through the bank plan where that pays (ops.use_bank_plan: k = 32 / 64, long rows, L2-resident operand)."""

    @staticmethod
    def forward(ctx, sp_data, sp_index, warp4, indices, values, feat_dim, num_rows):
        out = ops.spgemm_forward_auto(warp4, indices, values, sp_data.contiguous(), sp_index,
                                      feat_dim=feat_dim, num_rows=num_rows)
        ctx.save_for_backward(warp4, indices, values, sp_index)
        return out

    @staticmethod
    def backward(ctx, grad_output):
        warp4, indices, values, sel = ctx.saved_tensors
        dxc = ops.sspmm_backward(warp4, indices, values, grad_output.contiguous(), sel)
        return dxc, None, None, None, None, None, None


def cbsr_native_enabled(k: int = 32) -> bool:
    """Whether the layers stay on the CBSR operand (csrc/cbsr_dense.cu): SAGE's fc_self without torch's dense
    Linear on the scattered twin, GIN's (1 + eps) * h added to the aggregate in place, the Linear in front of
    a MaxK with its dX computed on the k-sparse gradient.  Measured on the products shape (round 2,
    profiles/r02_variants.md): at k = 32 fc_self 3.8 vs 6.3 ms, gradient of the kept values 5.1 vs 6.2 ms,
    SAGE epoch 213 -> 196 ms; at k = 16 twice that gain; at k = 64 the sampled dot loses (7.9 vs 6.3 ms).
    Hence: on for k <= 32, off above; MAXK_CBSR_NATIVE=0 / 1 forces it off / on."""
    import os
    v = os.environ.get("MAXK_CBSR_NATIVE", "").strip()
    if v in ("0", "1"):
        return v == "1"
    return k <= 32


def linear_topk_fused_enabled() -> bool:
    """Whether LinearTopKCBSR's forward is the fused tcgen05 kernel (csrc/linear_topk.cu) where its shapes allow
    (output width 256, fp32).  MAXK_LINEAR_TOPK_FUSED=0 selects the library GEMM followed by the TopK kernel."""
    import os
    return os.environ.get("MAXK_LINEAR_TOPK_FUSED", "1").strip() != "0"


def tc_linear_enabled() -> bool:
    """Whether the 256-wide dense Linears next to the path run on the own tensor-core kernels (csrc/linear_topk.cu
    without the selection, csrc/linear_wgrad.cu: split-TF32 products at fp32 accuracy) instead of the library's fp32
    SGEMM.  MAXK_TC_LINEAR=0 restores torch.nn.functional.linear everywhere."""
    import os
    return os.environ.get("MAXK_TC_LINEAR", "1").strip() != "0"


def weight_grad(grad_out: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    """grad_out^T @ x, the weight gradient of a Linear: the tcgen05 kernel where its shapes fit, else the library."""
    if tc_linear_enabled() and ops.linear_wgrad_supported(grad_out, x):
        return ops.linear_wgrad(grad_out, x)
    return grad_out.t() @ x


class TensorCoreLinear(Function):
    """y = x @ weight^T + bias for a 256-wide Linear (fc_neigh of SAGE, utils/models.py:238-244) with all three
    products on the tensor cores at fp32 accuracy: forward and dX through csrc/linear_topk.cu (the GEMM without the
    selection), dW through csrc/linear_wgrad.cu."""

    @staticmethod
    def forward(ctx, x, weight, bias, residual=None):
        x = x.contiguous()
        y = ops.linear_f32(x, weight.detach(), None if bias is None else bias.detach(),
                           None if residual is None else residual.detach())   # + residual in the GEMM's epilogue
        ctx.save_for_backward(x, weight)
        ctx.has_bias = bias is not None
        ctx.has_residual = residual is not None
        return y

    @staticmethod
    def backward(ctx, g):
        x, weight = ctx.saved_tensors
        g = g.contiguous()
        d_x = d_w = d_b = None
        if ctx.needs_input_grad[0]:
            wt = weight.detach().t().contiguous()        # dX = g @ W = g @ (W^T)^T: "weight" [in, 256]
            d_x = ops.linear_f32(g, wt) if ops.linear_topk_supported(g, wt) else g @ weight
        if ctx.needs_input_grad[1]:
            d_w = weight_grad(g, x)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            d_b = g.sum(0)
        return d_x, d_w, d_b, (g if ctx.has_residual and ctx.needs_input_grad[3] else None)


class NarrowOutputLinear(Function):
    """y = x @ weight^T + bias for the classifier at the end of the models (`lin_out`: 256 -> classes, utils/models.py:
    216-218), whose BACKWARD runs on the tensor-core kernels with the roles of the operands exchanged: with the
    gradient padded to a multiple of 4 columns, dX = g_pad @ W_pad is `linear_f32` with in_dim = classes, and
    dW^T = x^T @ g_pad is `linear_wgrad` with x as the 256-wide operand (both far from the MMA floor: the narrow
    dimension is the MMA's N / K, so they run at the speed of reading x).  The forward stays the library's (its
    output is not 256 wide)."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        x = x.contiguous()
        ctx.save_for_backward(x, weight)
        ctx.has_bias = bias is not None
        return torch.nn.functional.linear(x, weight, bias)

    @staticmethod
    def backward(ctx, g):
        x, weight = ctx.saved_tensors
        c = weight.shape[0]
        cp = (c + 3) // 4 * 4
        g_pad = torch.nn.functional.pad(g, (0, cp - c)) if cp != c else g.contiguous()
        d_x = d_w = d_b = None
        if ctx.needs_input_grad[0]:
            w_pad = torch.nn.functional.pad(weight.detach(), (0, 0, 0, cp - c)) if cp != c else weight.detach()
            d_x = ops.linear_f32(g_pad, w_pad.t().contiguous())          # "weight" [256, cp]: out = g_pad @ W_pad
        if ctx.needs_input_grad[1]:
            d_w = ops.linear_wgrad(x, g_pad).t()[:c].contiguous()         # [256, cp]^T -> [c, 256]
        if ctx.has_bias and ctx.needs_input_grad[2]:
            d_b = g.sum(0)
        return d_x, d_w, d_b


def output_linear(lin: torch.nn.Linear, x: torch.Tensor) -> torch.Tensor:
    """`lin(x)` for a 256 -> classes layer with its backward on the tensor-core kernels (NarrowOutputLinear)."""
    if (tc_linear_enabled() and x.is_cuda and x.dtype == torch.float32 and x.dim() == 2
            and x.shape[1] == ops.FULL_DIM and lin.weight.shape[1] == ops.FULL_DIM and 4 <= lin.weight.shape[0] < ops.FULL_DIM):
        return NarrowOutputLinear.apply(x, lin.weight, lin.bias)
    return linear(lin, x)


def linear(lin: torch.nn.Linear, x: torch.Tensor, residual: Optional[torch.Tensor] = None) -> torch.Tensor:
    """`lin(x)` (+ residual) through TensorCoreLinear where the shapes allow (output width 256), else torch's own."""
    if tc_linear_enabled() and ops.linear_topk_supported(x, lin.weight):
        return TensorCoreLinear.apply(x, lin.weight, lin.bias, residual)
    y = lin(x)
    return y if residual is None else residual + y


class LayerNormFn(Function):
    """nn.LayerNorm(d, elementwise_affine) forward / backward on csrc/layernorm.cu (one read and one write of the
    activation each way; the library's kernels take 4.1 + 9.6 ms per layer on the products shape)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, eps):
        x = x.contiguous()
        y, mean, rstd = ops.layernorm_forward(x, None if gamma is None else gamma.detach(),
                                              None if beta is None else beta.detach(), eps)
        ctx.save_for_backward(x, mean, rstd, gamma)
        ctx.has_beta = beta is not None
        return y

    @staticmethod
    def backward(ctx, g):
        x, mean, rstd, gamma = ctx.saved_tensors
        affine = (gamma is not None and ctx.needs_input_grad[1]) or (ctx.has_beta and ctx.needs_input_grad[2])
        dx, dgamma, dbeta = ops.layernorm_backward(g, x, mean, rstd, None if gamma is None else gamma.detach(), affine)
        return (dx, dgamma if gamma is not None else None, dbeta if ctx.has_beta else None, None)


class LayerNormTopKCBSR(Function):
    """(sp_data, sp_index) = TopK_k(LayerNorm(x)): the LayerNorm that closes a SAGE layer together with the MaxK that
    opens the next one (utils/models.py:252 / 230).  The normalised activation is never written; the backward takes
    the k-sparse gradient of the kept values as it comes from the aggregation (csrc/layernorm.cu)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, eps, k):
        x = x.contiguous()
        sp_data, sp_index, mean, rstd = ops.layernorm_topk_cbsr(x, None if gamma is None else gamma.detach(),
                                                                None if beta is None else beta.detach(), eps, k)
        ctx.save_for_backward(x, mean, rstd, gamma, sp_index)
        ctx.has_beta = beta is not None
        ctx.mark_non_differentiable(sp_index)
        return sp_data, sp_index

    @staticmethod
    def backward(ctx, g_data, _g_index):
        x, mean, rstd, gamma, sp_index = ctx.saved_tensors
        affine = (gamma is not None and ctx.needs_input_grad[1]) or (ctx.has_beta and ctx.needs_input_grad[2])
        dx, dgamma, dbeta = ops.layernorm_backward_cbsr(g_data, sp_index, x, mean, rstd,
                                                        None if gamma is None else gamma.detach(), affine)
        return dx, (dgamma if gamma is not None else None), (dbeta if ctx.has_beta else None), None, None


def layer_norm_topk_supported(norm, x: torch.Tensor) -> bool:
    import os
    return (isinstance(norm, torch.nn.LayerNorm) and len(norm.normalized_shape) == 1 and ops.layernorm_supported(x)
            and norm.normalized_shape[0] == x.shape[1] and os.environ.get("MAXK_LAYERNORM", "1").strip() != "0"
            and os.environ.get("MAXK_LAYERNORM_TOPK", "1").strip() != "0")


def layer_norm(norm, x: torch.Tensor) -> torch.Tensor:
    """`norm(x)` for an nn.LayerNorm over the last dimension through LayerNormFn where the shapes allow (fp32, CUDA,
    width <= 256); anything else (nn.Identity, other widths) is called as it is.  MAXK_LAYERNORM=0 disables."""
    import os
    if (isinstance(norm, torch.nn.LayerNorm) and len(norm.normalized_shape) == 1 and ops.layernorm_supported(x)
            and norm.normalized_shape[0] == x.shape[1] and os.environ.get("MAXK_LAYERNORM", "1").strip() != "0"):
        return LayerNormFn.apply(x, norm.weight, norm.bias, norm.eps)
    return norm(x)


class CBSRLinear(Function):
    """y = scatter(sp_data, sp_index) @ weight^T + bias computed from the CBSR pair (N*k*C multiply-adds, no
    dense [N, D] twin in the forward or in the gradient of the values).  Same result as
    `Linear(scatter(...))` -- upstream's fc_self on the MaxK-sparsified h, utils/models_old.py:55-60 -- up to
    fp32 summation order.  The weight gradient is a dense library GEMM on the scattered twin."""

    @staticmethod
    def forward(ctx, sp_data, sp_index, weight, bias):
        m = weight.detach().t().contiguous()          # [D, C]: rows of m are what the selectors pick
        sp_data = sp_data.contiguous()
        out = ops.cbsr_matmul(sp_data, sp_index, m, None if bias is None else bias.detach().contiguous())
        ctx.save_for_backward(sp_data, sp_index, m)
        ctx.has_bias = bias is not None
        return out

    @staticmethod
    def backward(ctx, grad_output):
        sp_data, sp_index, m = ctx.saved_tensors
        g = grad_output.contiguous()
        d_sp = ops.cbsr_sampled_dot(g, m, sp_index) if ctx.needs_input_grad[0] else None
        d_w = None
        if ctx.needs_input_grad[2]:
            d_w = weight_grad(g, ops.cbsr_scatter(sp_data, sp_index, m.shape[0]))   # [C, D]
        d_b = g.sum(0) if ctx.has_bias and ctx.needs_input_grad[3] else None
        return d_sp, None, d_w, d_b


class CBSRAddScaled(Function):
    """y + (1 + eps) * scatter(sp_data, sp_index), computed in place on y (GINConv's self term,
    utils/models.py:361) without the dense twin.  `y` must be a tensor nobody else needs unchanged (the
    output of CBSRAggregate is: its backward saves inputs only).  eps: a one-element parameter."""

    @staticmethod
    def forward(ctx, y, sp_data, sp_index, eps):
        sp_data = sp_data.contiguous()
        ops.cbsr_axpy_(y, sp_data, sp_index, 1.0, eps.detach().reshape(1).contiguous())
        ctx.mark_dirty(y)
        ctx.save_for_backward(sp_data, sp_index, eps)
        return y

    @staticmethod
    def backward(ctx, grad_output):
        sp_data, sp_index, eps = ctx.saved_tensors
        picked = ops.cbsr_gather(grad_output.contiguous(), sp_index)          # dL/dy at the kept columns
        d_sp = picked * (1.0 + eps.detach().reshape(())) if ctx.needs_input_grad[1] else None
        d_eps = (picked * sp_data).sum().reshape(eps.shape) if ctx.needs_input_grad[3] else None
        return grad_output, d_sp, None, d_eps


class LinearTopKCBSR(Function):
    """(sp_data, sp_index) = TopK_k(x @ weight^T + bias): the Linear in front of the MaxK of the GCN / GIN layers
    (utils/models.py:318-322, 365-369) together with the selection.  The forward is the dense library GEMM
    followed by the TopK kernel -- or, at output width 256, ONE tcgen05 kernel with the selection in the GEMM's
    epilogue (csrc/linear_topk.cu); the backward: the gradient that reaches the Linear's output is
    k-sparse per row, so dX = scatter(g) @ weight is computed on the CBSR pair (N*k*in multiply-adds,
    csrc/cbsr_dense.cu) instead of as a dense [N, D] x [D, in] GEMM.  dW and db use the scattered twin."""

    @staticmethod
    def forward(ctx, x, weight, bias, k):
        if linear_topk_fused_enabled() and ops.linear_topk_supported(x, weight):
            # one tcgen05 kernel, selection in the GEMM's epilogue: no dense [N, 256] activation in HBM
            sp_data, sp_index = ops.linear_topk_cbsr(x.contiguous(), weight.detach().contiguous(),
                                                     None if bias is None else bias.detach().contiguous(), k)
        else:
            y = torch.nn.functional.linear(x, weight, bias)
            sp_data, sp_index = ops.topk_cbsr(y.contiguous(), k)
        ctx.save_for_backward(x, weight, sp_index)
        ctx.has_bias = bias is not None
        ctx.mark_non_differentiable(sp_index)
        return sp_data, sp_index

    @staticmethod
    def backward(ctx, g_data, _g_index):
        x, weight, sp_index = ctx.saved_tensors
        g_data = g_data.contiguous()
        d_x = d_w = d_b = None
        dense = None
        if ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2]):
            dense = ops.cbsr_scatter(g_data, sp_index, weight.shape[0])
        if ctx.needs_input_grad[0]:
            wt = weight.detach().t().contiguous()                      # dX = scatter(g) @ W = scatter(g) @ (W^T)^T
            if tc_linear_enabled() and g_data.is_cuda and tuple(weight.shape) == (ops.FULL_DIM, ops.FULL_DIM):
                # a 256 x 256 layer: with the tensor-core GEMM at 1.6 ms the dense product beats the k-sparse one
                # (cbsr_matmul: 3.8 ms at k = 32 on the products shape); the scattered twin exists for dW anyway
                if dense is None:
                    dense = ops.cbsr_scatter(g_data, sp_index, weight.shape[0])
                d_x = ops.linear_f32(dense, wt)
            elif weight.shape[1] % 4 == 0:
                d_x = ops.cbsr_matmul(g_data, sp_index, weight.detach().contiguous())   # rows of W picked by sel
            else:
                d_x = ops.cbsr_scatter(g_data, sp_index, weight.shape[0]) @ weight
        if dense is not None:
            if ctx.needs_input_grad[1]:
                d_w = weight_grad(dense, x)
            if ctx.has_bias and ctx.needs_input_grad[2]:
                d_b = dense.sum(0)
        return d_x, d_w, d_b, None


class TopKCBSR(Function):
    """x -> (sp_data, sp_index) with gradient dX = scatter(d sp_data, sp_index)."""

    @staticmethod
    def forward(ctx, x, k):
        sp_data, sp_index = ops.topk_cbsr(x.contiguous(), k)
        ctx.save_for_backward(sp_index)
        ctx.feat_dim = x.shape[1]
        ctx.mark_non_differentiable(sp_index)
        return sp_data, sp_index

    @staticmethod
    def backward(ctx, g_data, _g_index):
        (sp_index,) = ctx.saved_tensors
        return ops.cbsr_scatter(g_data.contiguous(), sp_index, ctx.feat_dim), None
