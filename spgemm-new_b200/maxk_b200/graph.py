"""Graph container for the CBSR-native layers: the CSR adjacency resident on one GPU, its
on-device work list, and the edge weights of the three aggregation flavours the reference
models use:

    "sum"   values = 1                      GINConv            (utils/models.py:361)
    "mean"  values = 1/deg(row)             SAGEConv("mean")   (utils/models_old.py:39)
    "gcn"   values = 1/sqrt(d_row d_col)    GraphConv(norm="both")  (utils/models.py:314)

It stands where the DGL graph object `g` stands in the reference's `model(g, x)` calls and
offers the one DGL accessor the fork's SAGE uses (`adj_tensors('csr')`, models.py:217).
`DistGraph` is the row-partitioned multi-GPU flavour with the same `aggregate` method.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
from torch.autograd import Function

from . import ops
from .autograd import CBSRAggregate


def edge_values(indptr: torch.Tensor, indices: torch.Tensor, kind: str, n_cols: int,
                col_deg: Optional[torch.Tensor] = None) -> torch.Tensor:
    n = indptr.numel() - 1
    deg = (indptr[1:] - indptr[:-1]).long()
    if kind == "sum":
        return torch.ones(indices.numel(), dtype=torch.float32, device=indices.device)
    if kind == "mean":
        return torch.repeat_interleave(1.0 / deg.clamp(min=1).float(), deg)
    if kind == "gcn":
        if col_deg is None:
            col_deg = torch.bincount(indices.long(), minlength=n_cols)
        rows = torch.repeat_interleave(torch.arange(n, device=indices.device), deg)
        return torch.rsqrt(deg.clamp(min=1).float()[rows] * col_deg.clamp(min=1).float()[indices.long()])
    raise ValueError(f"unknown aggregation kind {kind!r}")


class Graph:
    def __init__(self, indptr: torch.Tensor, indices: torch.Tensor, n_cols: Optional[int] = None):
        assert indptr.is_cuda, "Graph lives on the GPU (no CPU fallback)"
        self.indptr = indptr.to(torch.int32).contiguous()
        self.indices = indices.to(torch.int32).contiguous()
        self.n_rows = self.indptr.numel() - 1
        self.n_cols = int(n_cols if n_cols is not None else self.n_rows)
        self.warp4 = ops.build_worklist(self.indptr)
        self._values: Dict[str, torch.Tensor] = {}

    # DGL-flavoured accessors used by the reference models
    def num_nodes(self) -> int:
        return self.n_rows

    def adj_tensors(self, fmt: str = "csr"):
        assert fmt == "csr"
        return self.indptr, self.indices, torch.arange(self.indices.numel(), device=self.indices.device)

    def values(self, kind: str) -> torch.Tensor:
        if kind not in self._values:
            self._values[kind] = edge_values(self.indptr, self.indices, kind, self.n_cols).contiguous()
        return self._values[kind]

    def aggregate(self, sp_data: torch.Tensor, sp_index: torch.Tensor, kind: str, feat_dim: int):
        """Y = A_kind @ scatter(sp_data, sp_index); differentiable w.r.t. sp_data."""
        return CBSRAggregate.apply(sp_data, sp_index, self.warp4, self.indices, self.values(kind),
                                   feat_dim, self.n_rows)


class _DistAggregate(Function):
    @staticmethod
    def forward(ctx, sp_data, sp_index, agg):
        g_data, g_index = agg.exchange(sp_data.contiguous(), sp_index)
        y = agg.forward(g_data, g_index).clone()
        ctx.agg = agg
        ctx.save_for_backward(g_index.clone())
        return y

    @staticmethod
    def backward(ctx, grad_output):
        (g_index,) = ctx.saved_tensors
        dxc = ctx.agg.backward(grad_output.contiguous(), g_index).clone()
        return dxc, None, None


class DistGraph:
    """Row block [bounds[rank], bounds[rank+1]) of a global CSR graph on this rank's GPU."""

    def __init__(self, indptr, indices, bounds: List[int], rank: int, group=None, split: bool = True,
                 comm: str = "nccl"):
        """comm: "nccl" (all-gather / reduce-scatter collectives, dist.DistAggregator) or "peer"
        (fused kernels over NVLink peer memory, peer.PeerAggregator; k in 8, 16, 32*m)."""
        self.bounds, self.rank, self.group, self.split = list(bounds), rank, group, split
        self.comm = comm
        lo, hi = bounds[rank], bounds[rank + 1]
        self.n_rows = hi - lo
        self.n_cols = indptr.numel() - 1
        e0, e1 = int(indptr[lo]), int(indptr[hi])
        self.indptr = (indptr[lo:hi + 1] - indptr[lo]).to(torch.int32).contiguous()
        self.indices = indices[e0:e1].to(torch.int32).contiguous()
        self._col_deg = torch.bincount(indices.long(), minlength=self.n_cols)  # global out-degrees
        self._aggs: Dict[tuple, object] = {}

    def num_nodes(self) -> int:
        return self.n_rows

    def aggregator(self, kind: str, k: int, feat_dim: int):
        from .dist import DistAggregator
        key = (kind, k, feat_dim)
        if key not in self._aggs:
            vals = edge_values(self.indptr, self.indices, kind, self.n_cols, self._col_deg)
            if self.comm == "peer":
                from .peer import PeerAggregator
                self._aggs[key] = PeerAggregator(self.indptr, self.indices, vals, self.bounds, self.rank, k,
                                                 feat_dim, group=self.group)
            else:
                self._aggs[key] = DistAggregator(self.indptr, self.indices, vals, self.bounds, self.rank, k,
                                                 feat_dim, group=self.group, split=self.split)
        return self._aggs[key]

    def aggregate(self, sp_data, sp_index, kind: str, feat_dim: int):
        return _DistAggregate.apply(sp_data, sp_index, self.aggregator(kind, sp_data.shape[1], feat_dim))

    def close(self):
        """Release the aggregators (IPC-exported buffers and the peers' mappings of them).  COLLECTIVE: every
        rank must call it, in the same order of keys (PeerAggregator.close synchronises the ranks so that no
        peer still has a kernel in flight on a buffer that is about to be freed)."""
        for key in sorted(self._aggs, key=repr):
            agg = self._aggs[key]
            if hasattr(agg, "close"):
                agg.close()
        self._aggs.clear()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False
