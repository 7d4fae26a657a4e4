"""Multi-GPU MaxK aggregation: 1-D row partition of the adjacency, CBSR exchange over NCCL.

The reference has no distributed code at all; this is the B200-native design
the north star asks for:

  * rows of A (destination nodes) are split into `world` contiguous blocks balanced by nnz;
    rank r keeps its CSR slice, its own work list (built on device from the local indptr)
    and its slice of X / Y / dY;
  * forward: every rank runs TopK on its rows, then ONLY the compact CBSR features
    (k fp32 + k uint8 per node) are all-gathered; column ids are remapped once, at set-up,
    to positions in the gathered (block-padded) buffer so the kernel gathers from it directly;
    the all-gather runs on NCCL's stream while the part of the aggregation whose columns are
    local (`split=True`) runs on the compute stream;
  * backward: each rank scatters into a full-height CBSR-shaped partial gradient, which is
    reduce-scattered (sum) so every rank ends with dXc for the nodes it owns.

One process per GPU (torchrun); `backend` is the operator set (maxk_b200.ops on CUDA).  The
host logic (partition, remap, padding, collectives) is backend-agnostic so that it can be
exercised with gloo on CPU in tests/ with a stand-in operator set.
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.distributed as dist


def row_cost_in_edges(k: int, feat_dim: int = 256) -> float:
    """What one destination row costs, in units of one edge, by the algorithmic byte model of
    SURVEY 8(d): a row is one write of Y and one read of dY (2 * 4 * D bytes), an edge is
    8 + 5k bytes in each direction.  6.1 for k=32, 11.6 for k=16 at D=256."""
    return (2 * 4 * feat_dim) / (2.0 * (8 + 5 * k))


def partition_rows_by_nnz(indptr: torch.Tensor, world: int, align: int = 1,
                          row_cost: float = 0.0) -> List[int]:
    """Contiguous row blocks with ~1/world of the work each, work = nnz + row_cost * rows.
    row_cost = 0 balances edges only (fine when rows are long: Reddit, proteins);
    `row_cost_in_edges(k)` also weighs the dense row traffic, which dominates on sparse
    skewed graphs (R-MAT at average degree 5: balancing edges alone hands one rank 44 % of all
    rows).  Returns world+1 row bounds; interior bounds are rounded down to a multiple of
    `align` (the peer-memory path moves blocks with 16-byte bulk copies: align=16)."""
    n = indptr.numel() - 1
    ip = indptr.long()
    if row_cost > 0.0:
        ip = ip + (torch.arange(n + 1, device=ip.device, dtype=torch.float64) * row_cost).long()
    e = int(ip[-1])
    bounds = [0]
    for r in range(1, world):
        target = (e * r) // world
        b = int(torch.searchsorted(ip, torch.tensor([target], device=ip.device, dtype=torch.long)))
        b = (b // align) * align
        b = min(max(b, bounds[-1]), n)
        bounds.append(b)
    bounds.append(n)
    return bounds


def remap_columns(indices: torch.Tensor, bounds: List[int], max_rows: int) -> torch.Tensor:
    """global column id -> row position in the block-padded gathered CBSR buffer."""
    b = torch.tensor(bounds, device=indices.device, dtype=torch.long)
    c = indices.long()
    owner = torch.searchsorted(b, c, right=True) - 1
    owner.clamp_(0, len(bounds) - 2)
    return (owner * max_rows + (c - b[owner])).to(torch.int32)


def _reduce_scatter(out: torch.Tensor, inp: torch.Tensor, rank: int, group) -> None:
    if dist.get_backend(group) == "nccl":
        dist.reduce_scatter_tensor(out, inp, group=group)
    else:  # gloo (CPU tests): no reduce_scatter_tensor
        tmp = inp.clone()
        dist.all_reduce(tmp, group=group)
        out.copy_(tmp.view(dist.get_world_size(group), -1)[rank].view_as(out))


class DistAggregator:
    def __init__(self, indptr_loc, indices_glob, values_loc, bounds: List[int], rank: int, k: int,
                 feat_dim: int = 256, group=None, backend=None, split: bool = True):
        if backend is None:
            from . import ops as backend  # CUDA operator set
        self.ops = backend
        self.group = group
        self.rank, self.world = rank, len(bounds) - 1
        self.bounds = list(bounds)
        self.k, self.d = int(k), int(feat_dim)
        self.n_local = bounds[rank + 1] - bounds[rank]
        self.max_rows = max(bounds[i + 1] - bounds[i] for i in range(self.world))
        self.n_pad = self.world * self.max_rows
        dev = indptr_loc.device
        self.device = dev
        indptr_loc = indptr_loc.to(torch.int32).contiguous()
        self.values = values_loc.to(torch.float32).contiguous()
        self.split = bool(split) and self.world > 1
        lo, hi = bounds[rank], bounds[rank + 1]
        if self.split:
            # local-column part reads this rank's own (un-gathered) CBSR block; the rest
            # reads the gathered buffer once the all-gather has landed
            c = indices_glob.long()
            is_loc = (c >= lo) & (c < hi)
            rows = torch.repeat_interleave(torch.arange(self.n_local, device=dev),
                                           (indptr_loc[1:] - indptr_loc[:-1]).long())
            self.part = []
            for sel, remap in ((is_loc, lambda cc: (cc - lo).to(torch.int32)),
                               (~is_loc, lambda cc: remap_columns(cc, bounds, self.max_rows))):
                r_sel = rows[sel]
                deg = torch.bincount(r_sel, minlength=self.n_local)
                ip = torch.zeros(self.n_local + 1, dtype=torch.int64, device=dev)
                torch.cumsum(deg, 0, out=ip[1:])
                ip = ip.to(torch.int32)
                self.part.append({"indptr": ip, "indices": remap(c[sel]).contiguous(),
                                  "values": self.values[sel].contiguous(),
                                  "warp4": self.ops.build_worklist(ip)})
            self.num_tasks = sum(p["warp4"].numel() // 4 for p in self.part)
            self.num_edges = int(indices_glob.numel())
        else:
            self.indices = remap_columns(indices_glob, bounds, self.max_rows).contiguous()
            self.warp4 = self.ops.build_worklist(indptr_loc)
            self.num_tasks = self.warp4.numel() // 4
            self.num_edges = int(indices_glob.numel())
        # persistent exchange buffers
        self.loc_data = torch.zeros((self.max_rows, self.k), dtype=torch.float32, device=dev)
        self.loc_index = torch.zeros((self.max_rows, self.k), dtype=torch.uint8, device=dev)
        self.g_data = torch.zeros((self.n_pad, self.k), dtype=torch.float32, device=dev)
        self.g_index = torch.zeros((self.n_pad, self.k), dtype=torch.uint8, device=dev)
        self.partial = torch.zeros((self.n_pad, self.k), dtype=torch.float32, device=dev)
        self.rs_out = torch.zeros((self.max_rows, self.k), dtype=torch.float32, device=dev)
        self.y = torch.zeros((self.n_local, self.d), dtype=torch.float32, device=dev)
        self._pending = []

    @classmethod
    def from_global_csr(cls, indptr, indices, values, bounds, rank, k, feat_dim=256, **kw):
        lo, hi = bounds[rank], bounds[rank + 1]
        e0, e1 = int(indptr[lo]), int(indptr[hi])
        ip = (indptr[lo:hi + 1] - indptr[lo]).to(torch.int32)
        return cls(ip, indices[e0:e1].contiguous(), values[e0:e1].contiguous(), bounds, rank, k,
                   feat_dim, **kw)

    # ------------------------------------------------------------------------- forward
    def topk_and_exchange(self, x_loc: torch.Tensor):
        """TopK on the local rows, then start the all-gather of the CBSR block.  Returns the
        gathered (sp_data, sp_index) buffers; `forward` waits for the exchange."""
        sd, si = self.ops.topk_cbsr(x_loc, self.k)
        return self.exchange(sd, si)

    def exchange(self, sp_data_loc: torch.Tensor, sp_index_loc: torch.Tensor):
        """Start the all-gather of this rank's CBSR block (k fp32 + k uint8 per node)."""
        self.loc_data[: self.n_local].copy_(sp_data_loc)
        self.loc_index[: self.n_local].copy_(sp_index_loc)
        if self.world > 1:
            self._pending = [dist.all_gather_into_tensor(self.g_data, self.loc_data, group=self.group,
                                                         async_op=True),
                             dist.all_gather_into_tensor(self.g_index, self.loc_index, group=self.group,
                                                         async_op=True)]
        else:
            self.g_data.copy_(self.loc_data)
            self.g_index.copy_(self.loc_index)
        return self.g_data, self.g_index

    def _wait_exchange(self):
        for h in self._pending:
            h.wait()
        self._pending = []

    def forward(self, g_data: torch.Tensor, g_index: torch.Tensor) -> torch.Tensor:
        if self.split:
            lp, rp = self.part
            # local columns: overlaps the all-gather (reads the un-gathered local block)
            self.ops.spgemm_forward(lp["warp4"], lp["indices"], lp["values"], self.loc_data,
                                    self.loc_index, feat_dim=self.d, num_rows=self.n_local, out=self.y)
            self._wait_exchange()
            self.ops.spgemm_forward(rp["warp4"], rp["indices"], rp["values"], g_data, g_index,
                                    feat_dim=self.d, num_rows=self.n_local, out=self.y,
                                    accumulate=True)
        else:
            self._wait_exchange()
            self.ops.spgemm_forward(self.warp4, self.indices, self.values, g_data, g_index,
                                    feat_dim=self.d, num_rows=self.n_local, out=self.y)
        return self.y

    # ------------------------------------------------------------------------ backward
    def backward(self, dy_loc: torch.Tensor, g_index: torch.Tensor) -> torch.Tensor:
        """-> dXc for the rows this rank owns, f32[n_local, k]."""
        self._wait_exchange()
        if self.split:
            lp, rp = self.part
            # remote columns first: their reduce-scatter then overlaps the local-column part
            self.ops.sspmm_backward(rp["warp4"], rp["indices"], rp["values"], dy_loc, g_index,
                                    out=self.partial)
            h = None
            if self.world > 1 and dist.get_backend(self.group) == "nccl":
                h = dist.reduce_scatter_tensor(self.rs_out, self.partial, group=self.group, async_op=True)
            else:
                _reduce_scatter(self.rs_out, self.partial, self.rank, self.group)
            own = g_index[self.rank * self.max_rows:(self.rank + 1) * self.max_rows]
            loc = self.ops.sspmm_backward(lp["warp4"], lp["indices"], lp["values"], dy_loc,
                                          own, out=None)
            if h is not None:
                h.wait()
            self.rs_out[: self.n_local].add_(loc[: self.n_local])
        else:
            self.ops.sspmm_backward(self.warp4, self.indices, self.values, dy_loc, g_index,
                                    out=self.partial)
            if self.world > 1:
                _reduce_scatter(self.rs_out, self.partial, self.rank, self.group)
            else:
                self.rs_out.copy_(self.partial[: self.max_rows])
        return self.rs_out[: self.n_local]
