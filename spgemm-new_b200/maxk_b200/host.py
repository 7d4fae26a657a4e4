"""Host-buffer entry point: the call a CPU-side user of the reference path makes.

`HostAggregator` keeps the graph (CSR + work list) resident on the GPU -- the analogue of
the DGL graph object the reference builds once -- and runs one forward + backward of the
MaxK aggregation on HOST tensors:

    y_host, dxc_host = agg.step(x_host, dy_host)

    H2D X -> TopK->CBSR -> forward SpGEMM -> D2H Y
    H2D dY ------------ -> backward SSpMM -> D2H dXc

Copies run on their own streams so that the dY upload overlaps TopK + SpGEMM and the Y
download overlaps the SSpMM (PCIe is full duplex).  This is what bench.py's `e2e` times.
"""
from __future__ import annotations

import torch

from . import ops
from .autograd import worklist_for


class HostAggregator:
    def __init__(self, indptr, indices, values, k: int, feat_dim: int = 256, device="cuda",
                 n_cols=None):
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
        self.y_h = torch.empty((self.n_rows, self.d), dtype=torch.float32, pin_memory=True)
        self.dxc_h = torch.empty((self.n_cols, self.k), dtype=torch.float32, pin_memory=True)
        self.s_in = torch.cuda.Stream(dev)
        self.s_out = torch.cuda.Stream(dev)
        self.h2d_bytes = (self.x_d.numel() + self.dy_d.numel()) * 4
        self.d2h_bytes = (self.y_h.numel() + self.dxc_h.numel()) * 4

    def step(self, x_host: torch.Tensor, dy_host: torch.Tensor):
        """x_host f32[n_cols, d], dy_host f32[n_rows, d] (pinned for full speed) ->
        (y_host f32[n_rows, d], dxc_host f32[n_cols, k]); results are valid on return."""
        main = torch.cuda.current_stream(self.device)
        ev_x, ev_dy = torch.cuda.Event(), torch.cuda.Event()
        ev_y, ev_dxc = torch.cuda.Event(), torch.cuda.Event()
        self.s_in.wait_stream(main)   # previous step's kernels are done with x_d / dy_d
        with torch.cuda.stream(self.s_in):
            self.x_d.copy_(x_host, non_blocking=True)
            ev_x.record(self.s_in)
            self.dy_d.copy_(dy_host, non_blocking=True)
            ev_dy.record(self.s_in)
        main.wait_event(ev_x)
        sp_data, sp_index = ops.topk_cbsr(self.x_d, self.k)
        ops.spgemm_forward(self.warp4, self.indices, self.values, sp_data, sp_index,
                           feat_dim=self.d, num_rows=self.n_rows, out=self.y_d)
        ev_y.record(main)
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(ev_y)
            self.y_h.copy_(self.y_d, non_blocking=True)
        main.wait_event(ev_dy)
        ops.sspmm_backward(self.warp4, self.indices, self.values, self.dy_d, sp_index, out=self.dxc_d)
        ev_dxc.record(main)
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(ev_dxc)
            self.dxc_h.copy_(self.dxc_d, non_blocking=True)
        main.wait_stream(self.s_out)
        main.synchronize()
        return self.y_h, self.dxc_h
