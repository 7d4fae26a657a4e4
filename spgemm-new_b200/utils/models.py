"""CBSR-native MaxK-GNN layers: SAGE / GCN / GIN with the constructor and `forward(g, x)`
signatures of the reference's utils/models.py:151-392 (and upstream utils/models_old.py),
re-expressed on (sp_data, sp_index) so the dense masked activation is never fed to the
aggregation:

    Linear -> TopK->CBSR -> dropout on the k CBSR values -> SpGEMM (values = 1, 1/deg, sym-norm)
           -> + (1+eps)*h for GIN / + fc_self(h) for SAGE -> LayerNorm

`g` is a maxk_b200.graph.Graph (one GPU) or DistGraph (row partition); both expose
`aggregate(sp_data, sp_index, kind, feat_dim)`.  Layer variants: SAGE applies
MaxK before BOTH branches like upstream (models_old.py:55-60); `fork_variant=True` gives
the fork's un-sparsified self branch with sum aggregation (models.py:227-253).
Dropout acts on the CBSR values: same distribution as torch's dense dropout on the masked
tensor, different RNG stream (parity tests run with p = 0 / eval()).
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F
import torch.nn.init as init
from torch.nn import Linear

from maxk_b200.autograd import (CBSRAddScaled, CBSRLinear, LinearTopKCBSR, MaxK, SpGEMMFunction, TopKCBSR,  # noqa: F401  (re-exported like the reference)
                                LayerNormTopKCBSR, cbsr_native_enabled, layer_norm as _layer_norm,
                                layer_norm_topk_supported as _ln_topk_ok, linear as _linear,
                                output_linear as _output_linear)
from maxk_b200 import ops
from torch.autograd import Function


class _Scatter(Function):
    """(sp_data, sp_index) -> dense [N, D]; gradient = gather at the selector."""

    @staticmethod
    def forward(ctx, sp_data, sp_index, feat_dim):
        ctx.save_for_backward(sp_index)
        return ops.cbsr_scatter(sp_data.contiguous(), sp_index, feat_dim)

    @staticmethod
    def backward(ctx, g):
        (sp_index,) = ctx.saved_tensors
        return ops.cbsr_gather(g.contiguous(), sp_index), None, None


def _tc_dense(sp_data, lin):
    """Whether a Linear on the scattered twin of a CBSR pair runs on the tensor-core GEMM (csrc/linear_topk.cu)."""
    from maxk_b200.autograd import tc_linear_enabled
    return (tc_linear_enabled() and sp_data.is_cuda and sp_data.dtype == torch.float32
            and tuple(lin.weight.shape) == (ops.FULL_DIM, ops.FULL_DIM))


def _sparsify(x, k, nonlinear, p, training):
    """MaxK (or ReLU) + dropout, returned as CBSR.  For ReLU the row is kept dense (k = D)."""
    if nonlinear == "maxk":
        sp_data, sp_index = TopKCBSR.apply(x, k)
    else:
        if nonlinear == "relu":
            x = F.relu(x)
        sp_data, sp_index = TopKCBSR.apply(x, x.shape[1])  # all columns, ascending
    if p > 0 and training:
        sp_data = F.dropout(sp_data, p, True)
    return sp_data, sp_index


def _linear_sparsify(lin, x, k, nonlinear, p, training):
    """Linear -> MaxK -> dropout as CBSR; for k <= 32 (autograd.cbsr_native_enabled) the pair is one autograd
    node whose backward computes dX on the CBSR gradient instead of as a dense GEMM."""
    if nonlinear == "maxk" and cbsr_native_enabled(k):
        sp_data, sp_index = LinearTopKCBSR.apply(x, lin.weight, lin.bias, k)
        if p > 0 and training:
            sp_data = F.dropout(sp_data, p, True)
        return sp_data, sp_index
    return _sparsify(lin(x), k, nonlinear, p, training)


class _Base(nn.Module):
    def __init__(self, in_size, hid_size, num_hid_layers, out_size, maxk, feat_drop, norm, nonlinear):
        super().__init__()
        assert hid_size <= 256 and hid_size % 4 == 0, "selectors are uint8: hidden <= 256 (SURVEY Q3)"
        self.num_layers, self.hid, self.maxk = num_hid_layers, hid_size, maxk
        self.p, self.norm, self.nonlinear = feat_drop, norm, nonlinear
        self.lin_in = Linear(in_size, hid_size)
        self.lin_out = Linear(hid_size, out_size)
        init.xavier_uniform_(self.lin_in.weight)
        init.xavier_uniform_(self.lin_out.weight)
        self.normlayers = nn.ModuleList(
            [nn.LayerNorm(hid_size, elementwise_affine=True) if norm else nn.Identity()
             for _ in range(num_hid_layers)])


class SAGE(_Base):
    """models_old.py:26-67 (upstream) / models.py:151-297 (fork)."""

    def __init__(self, in_size, hid_size, num_hid_layers, out_size, maxk=32, feat_drop=0.5, norm=False,
                 nonlinear="maxk", fork_variant=False):
        super().__init__(in_size, hid_size, num_hid_layers, out_size, maxk, feat_drop, norm, nonlinear)
        self.fork_variant = fork_variant
        self.fc_self_layers = nn.ModuleList([Linear(hid_size, hid_size, bias=not fork_variant)
                                             for _ in range(num_hid_layers)])
        self.fc_neigh_layers = nn.ModuleList([Linear(hid_size, hid_size, bias=False)
                                              for _ in range(num_hid_layers)])
        for i in range(num_hid_layers):
            init.xavier_uniform_(self.fc_self_layers[i].weight)
            init.xavier_uniform_(self.fc_neigh_layers[i].weight)

    def forward(self, g, x):
        # upstream variant: layer 0 only ever sees MaxK(lin_in(x)) -- the input Linear and the first selection are one
        # node (the fused tcgen05 kernel at width 256: the dense [N, hidden] activation of lin_in never exists)
        first = pending = None
        if not self.fork_variant and self.num_layers > 0 and self.nonlinear == "maxk":
            first = _linear_sparsify(self.lin_in, x, self.maxk, self.nonlinear, self.p, self.training)
        else:
            x = self.lin_in(x)
        for i in range(self.num_layers):
            if self.fork_variant:   # models.py:230-253: sum aggregation, self branch on raw x, dropout after
                sp_data, sp_index = _sparsify(x, self.maxk, self.nonlinear, 0.0, False)
                x = _linear(self.fc_self_layers[i], x) + _linear(self.fc_neigh_layers[i],
                                                                 g.aggregate(sp_data, sp_index, "sum", self.hid))
                x = F.dropout(x, self.p, self.training)
            else:                   # SAGEConv("mean", feat_drop): dropout(maxk(x)) feeds both branches
                if i == 0 and first is not None:
                    sp_data, sp_index = first
                elif pending is not None:   # LayerNorm of layer i - 1 and this layer's MaxK in one pass
                    sp_data, sp_index = LayerNormTopKCBSR.apply(x, pending.weight, pending.bias, pending.eps, self.maxk)
                    if self.p > 0 and self.training:
                        sp_data = F.dropout(sp_data, self.p, True)
                    pending = None
                else:
                    sp_data, sp_index = _sparsify(x, self.maxk, self.nonlinear, self.p, self.training)
                fc_self = self.fc_self_layers[i]
                if _tc_dense(sp_data, fc_self):
                    # 256-wide layer: the tensor-core GEMM on the scattered twin (0.7 + 1.6 ms on the products shape)
                    # beats the k-sparse product on CUDA cores (3.8 ms forward, 5.0 ms backward at k = 32)
                    own = _linear(fc_self, _Scatter.apply(sp_data, sp_index, self.hid))
                elif self.nonlinear == "maxk" and cbsr_native_enabled(self.maxk):   # fc_self on the CBSR operand
                    own = CBSRLinear.apply(sp_data, sp_index, fc_self.weight, fc_self.bias)
                else:
                    own = _linear(fc_self, _Scatter.apply(sp_data, sp_index, self.hid))
                x = _linear(self.fc_neigh_layers[i], g.aggregate(sp_data, sp_index, "mean", self.hid), residual=own)
            if (not self.fork_variant and self.nonlinear == "maxk" and i + 1 < self.num_layers
                    and _ln_topk_ok(self.normlayers[i], x)):
                pending = self.normlayers[i]   # its output feeds only the next layer's MaxK: fused there
            else:
                x = _layer_norm(self.normlayers[i], x)
        return _output_linear(self.lin_out, x)


class GCN(_Base):
    """models.py:300-345: lin_in->relu; per layer Linear -> MaxK -> dropout -> GraphConv(norm='both',
    weight=None, bias) -> LayerNorm."""

    def __init__(self, in_size, hid_size, num_hid_layers, out_size, maxk=32, feat_drop=0.5, norm=False,
                 nonlinear="maxk"):
        super().__init__(in_size, hid_size, num_hid_layers, out_size, maxk, feat_drop, norm, nonlinear)
        self.linlayers = nn.ModuleList([Linear(hid_size, hid_size) for _ in range(num_hid_layers)])
        for l in self.linlayers:
            init.xavier_uniform_(l.weight)
        self.conv_bias = nn.ParameterList([nn.Parameter(torch.zeros(hid_size)) for _ in range(num_hid_layers)])

    def forward(self, g, x):
        x = _linear(self.lin_in, x).relu()
        for i in range(self.num_layers):
            sp_data, sp_index = _linear_sparsify(self.linlayers[i], x, self.maxk, self.nonlinear, self.p, self.training)
            x = g.aggregate(sp_data, sp_index, "gcn", self.hid) + self.conv_bias[i]
            x = _layer_norm(self.normlayers[i], x)
        return _output_linear(self.lin_out, x)


class GIN(_Base):
    """models.py:347-392: per layer Linear -> MaxK -> dropout -> GINConv(learn_eps=True):
    (1 + eps) * h + sum_j h_j -> LayerNorm."""

    def __init__(self, in_size, hid_size, num_hid_layers, out_size, maxk=32, feat_drop=0.5, norm=False,
                 nonlinear="maxk"):
        super().__init__(in_size, hid_size, num_hid_layers, out_size, maxk, feat_drop, norm, nonlinear)
        self.linlayers = nn.ModuleList([Linear(hid_size, hid_size) for _ in range(num_hid_layers)])
        for l in self.linlayers:
            init.xavier_uniform_(l.weight)
        self.eps = nn.ParameterList([nn.Parameter(torch.zeros(1)) for _ in range(num_hid_layers)])

    def forward(self, g, x):
        x = _linear(self.lin_in, x).relu()
        for i in range(self.num_layers):
            sp_data, sp_index = _linear_sparsify(self.linlayers[i], x, self.maxk, self.nonlinear, self.p, self.training)
            if self.nonlinear == "maxk" and cbsr_native_enabled(self.maxk):   # self term added in place, no dense twin
                x = CBSRAddScaled.apply(g.aggregate(sp_data, sp_index, "sum", self.hid), sp_data, sp_index, self.eps[i])
            else:
                h = _Scatter.apply(sp_data, sp_index, self.hid)
                x = (1.0 + self.eps[i]) * h + g.aggregate(sp_data, sp_index, "sum", self.hid)
            x = _layer_norm(self.normlayers[i], x)
        return _output_linear(self.lin_out, x)
