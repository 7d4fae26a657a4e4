"""Host-buffer entry point: the call a CPU-side user of the reference path makes.

`HostAggregator` keeps the graph (CSR + work list) resident on the GPU -- the analogue of
the DGL graph object the reference builds once -- and runs one forward + backward of the
MaxK aggregation on HOST tensors:

    y_host, dxc_host = agg.step(x_host, dy_host)

Everything that can overlap does (PCIe is full duplex, copies have their own engines):

    H2D  | X c0 | X c1 | X c2 | X c3 |        dY          |
    SM          |topk0 |topk1 |topk2 |topk3| fwd r0 | fwd r1 | fwd r2 | fwd r3 |   bwd   |
    D2H                                             | Y r0   | Y r1   | Y r2   | Y r3 |dXc|

TopK is per row, so it runs chunk by chunk as X arrives; the forward needs the whole CBSR
operand, but it is launched per block of output rows (the work list is in row order, so a
row block is a contiguous task range) and every finished block of Y starts its download
while the next block is computed; the dY upload hides behind the forward, the Y download
behind the backward.  This is what bench.py's `e2e` times.
"""
from __future__ import annotations

import torch

from . import ops
from .autograd import worklist_for


class HostAggregator:
    def __init__(self, indptr, indices, values, k: int, feat_dim: int = 256, device="cuda",
                 n_cols=None, chunks: int = 4):
        self.device = torch.device(device)
        self.k, self.d = int(k), int(feat_dim)
        self.indptr = indptr.to(self.device, torch.int32).contiguous()
        self.indices = indices.to(self.device, torch.int32).contiguous()
        self.values = values.to(self.device, torch.float32).contiguous()
        self.n_rows = self.indptr.numel() - 1
        self.n_cols = int(n_cols if n_cols is not None else self.n_rows)
        self.warp4 = worklist_for(self.indptr)
        dev = self.device
        self.x_d = torch.empty((self.n_cols, self.d), dtype=torch.float32, device=dev)
        self.dy_d = torch.empty((self.n_rows, self.d), dtype=torch.float32, device=dev)
        self.y_d = torch.empty((self.n_rows, self.d), dtype=torch.float32, device=dev)
        self.dxc_d = torch.empty((self.n_cols, self.k), dtype=torch.float32, device=dev)
        self.sp_data = torch.empty((self.n_cols, self.k), dtype=torch.float32, device=dev)
        self.sp_index = torch.empty((self.n_cols, self.k), dtype=torch.uint8, device=dev)
        self.y_h = torch.empty((self.n_rows, self.d), dtype=torch.float32, pin_memory=True)
        self.dxc_h = torch.empty((self.n_cols, self.k), dtype=torch.float32, pin_memory=True)
        self.s_in = torch.cuda.Stream(dev)
        self.s_out = torch.cuda.Stream(dev)
        self.h2d_bytes = (self.x_d.numel() + self.dy_d.numel()) * 4
        self.d2h_bytes = (self.y_h.numel() + self.dxc_h.numel()) * 4
        # row blocks with ~equal nnz; a block of rows is a contiguous range of tasks
        chunks = max(1, min(int(chunks), self.n_rows))
        ip = self.indptr.long()
        e = int(ip[-1])
        targets = torch.tensor([(e * c) // chunks for c in range(1, chunks)], device=dev, dtype=torch.long)
        rb = [0] + torch.searchsorted(ip, targets).clamp_(max=self.n_rows).tolist() + [self.n_rows]
        self.row_bounds = sorted(set(rb))
        deg = ip[1:] - ip[:-1]
        tasks_before = torch.zeros(self.n_rows + 1, dtype=torch.long, device=dev)
        torch.cumsum((deg + 63) // 64, 0, out=tasks_before[1:])
        self.task_bounds = tasks_before[torch.tensor(self.row_bounds, device=dev)].tolist()
        self.col_bounds = [(self.n_cols * c) // chunks for c in range(chunks)] + [self.n_cols]

    def step(self, x_host: torch.Tensor, dy_host: torch.Tensor):
        """x_host f32[n_cols, d], dy_host f32[n_rows, d] (pinned for full speed) ->
        (y_host f32[n_rows, d], dxc_host f32[n_cols, k]); results are valid on return."""
        main = torch.cuda.current_stream(self.device)
        self.s_in.wait_stream(main)   # previous step's kernels are done with x_d / dy_d
        self.s_out.wait_stream(main)
        # --- X upload in column-node chunks, TopK chunk by chunk behind it
        x_events = []
        with torch.cuda.stream(self.s_in):
            for c0, c1 in zip(self.col_bounds[:-1], self.col_bounds[1:]):
                self.x_d[c0:c1].copy_(x_host[c0:c1], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self.s_in)
                x_events.append(ev)
            self.dy_d.copy_(dy_host, non_blocking=True)
            ev_dy = torch.cuda.Event()
            ev_dy.record(self.s_in)
        for (c0, c1), ev in zip(zip(self.col_bounds[:-1], self.col_bounds[1:]), x_events):
            main.wait_event(ev)
            if c1 > c0:
                ops.topk_cbsr(self.x_d[c0:c1], self.k, out=(self.sp_data[c0:c1], self.sp_index[c0:c1]))
        # --- forward per block of output rows, each block's Y downloaded as soon as it is done
        self.y_d.zero_()
        w4 = self.warp4.view(-1, 4)
        for (r0, r1), (t0, t1) in zip(zip(self.row_bounds[:-1], self.row_bounds[1:]),
                                       zip(self.task_bounds[:-1], self.task_bounds[1:])):
            if t1 > t0:
                ops.spgemm_forward(w4[t0:t1].reshape(-1), self.indices, self.values, self.sp_data,
                                   self.sp_index, feat_dim=self.d, num_rows=self.n_rows,
                                   out=self.y_d, accumulate=True)
            ev = torch.cuda.Event()
            ev.record(main)
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(ev)
                self.y_h[r0:r1].copy_(self.y_d[r0:r1], non_blocking=True)
        # --- backward behind the dY upload; dXc download last
        main.wait_event(ev_dy)
        ops.sspmm_backward(self.warp4, self.indices, self.values, self.dy_d, self.sp_index,
                           out=self.dxc_d)
        ev = torch.cuda.Event()
        ev.record(main)
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(ev)
            self.dxc_h.copy_(self.dxc_d, non_blocking=True)
        main.wait_stream(self.s_out)
        main.synchronize()
        return self.y_h, self.dxc_h
