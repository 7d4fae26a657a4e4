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
behind the backward.  Across steps, `submit`/`result` overlap the next step's uploads with
this step's kernels and downloads.  This is what bench.py's `e2e` times.
"""
from __future__ import annotations

import torch

from . import ops
from .autograd import worklist_for


def bind_to_device_numa_node(device=None):
    """Restrict the calling process to the CPUs of the NUMA node the GPU hangs off (sysfs), so that pinned host
    buffers allocated AFTERWARDS are placed in that node's memory and the copy engines do not cross the socket
    interconnect.  One process per GPU: call it once, before the host buffers are created.  Returns the node, or
    None when the topology cannot be read (then nothing is changed)."""
    import os
    try:
        dev = torch.device(device if device is not None else torch.cuda.current_device())
        pr = torch.cuda.get_device_properties(dev)
        pci = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{pci}/numa_node").read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:  # noqa: BLE001  (no sysfs, no permission, old torch: leave the process where it is)
        return None


class _Slot:
    """Device inputs/outputs and pinned host outputs of one in-flight step."""

    def __init__(self, n_rows, n_cols, d, k, dev):
        self.x_d = torch.empty((n_cols, d), dtype=torch.float32, device=dev)
        self.dy_d = torch.empty((n_rows, d), dtype=torch.float32, device=dev)
        self.y_d = torch.empty((n_rows, d), dtype=torch.float32, device=dev)
        self.dxc_d = torch.empty((n_cols, k), dtype=torch.float32, device=dev)
        self.y_h = torch.empty((n_rows, d), dtype=torch.float32, pin_memory=True)
        self.dxc_h = torch.empty((n_cols, k), dtype=torch.float32, pin_memory=True)
        self.inputs_free = None   # event: the kernels that read x_d / dy_d have finished
        self.done = None          # event: y_h / dxc_h are complete


class HostAggregator:
    """`step(x, dy)` is the synchronous call.  `submit(x, dy)` / `result(ticket)` is the same
    work as a two-deep pipeline: with two slots of buffers the upload of step i+1 runs while
    step i computes and downloads, so a stream of steps is bound by the busier PCIe direction
    (X + dY going up) instead of by the sum of the stages."""

    def __init__(self, indptr, indices, values, k: int, feat_dim: int = 256, device="cuda",
                 n_cols=None, chunks: int = 4, slots: int = 2):
        self.device = torch.device(device)
        self.k, self.d = int(k), int(feat_dim)
        self.indptr = indptr.to(self.device, torch.int32).contiguous()
        self.indices = indices.to(self.device, torch.int32).contiguous()
        self.values = values.to(self.device, torch.float32).contiguous()
        self.n_rows = self.indptr.numel() - 1
        self.n_cols = int(n_cols if n_cols is not None else self.n_rows)
        self.warp4 = worklist_for(self.indptr)
        dev = self.device
        self.slots = [_Slot(self.n_rows, self.n_cols, self.d, self.k, dev) for _ in range(max(1, slots))]
        self._submitted = 0
        self.sp_data = torch.empty((self.n_cols, self.k), dtype=torch.float32, device=dev)
        self.sp_index = torch.empty((self.n_cols, self.k), dtype=torch.uint8, device=dev)
        self.s_in = torch.cuda.Stream(dev)
        self.s_main = torch.cuda.Stream(dev)
        self.s_out = torch.cuda.Stream(dev)
        self.h2d_bytes = (self.n_cols + self.n_rows) * self.d * 4
        self.d2h_bytes = (self.n_rows * self.d + self.n_cols * self.k) * 4
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

    def submit(self, x_host: torch.Tensor, dy_host: torch.Tensor) -> int:
        """Enqueue one step on host tensors x_host f32[n_cols, d], dy_host f32[n_rows, d]
        (pinned for full speed) and return a ticket for `result`.  At most len(slots) steps
        may be in flight: the ticket's host buffers are reused by the submit after next."""
        slot = self.slots[self._submitted % len(self.slots)]
        ticket = self._submitted
        self._submitted += 1
        main, s_in, s_out = self.s_main, self.s_in, self.s_out
        main.wait_stream(torch.cuda.current_stream(self.device))
        if slot.inputs_free is not None:
            s_in.wait_event(slot.inputs_free)  # this slot's previous kernels are done with x_d / dy_d
        if slot.done is not None:
            main.wait_event(slot.done)         # ... and its previous downloads with y_d / dxc_d
        col_chunks = list(zip(self.col_bounds[:-1], self.col_bounds[1:]))
        # --- X upload in column-node chunks, TopK chunk by chunk behind it; then dY
        x_events = []
        with torch.cuda.stream(s_in):
            for c0, c1 in col_chunks:
                slot.x_d[c0:c1].copy_(x_host[c0:c1], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(s_in)
                x_events.append(ev)
            slot.dy_d.copy_(dy_host, non_blocking=True)
            ev_dy = torch.cuda.Event()
            ev_dy.record(s_in)
        with torch.cuda.stream(main):
            for (c0, c1), ev in zip(col_chunks, x_events):
                main.wait_event(ev)
                if c1 > c0:
                    ops.topk_cbsr(slot.x_d[c0:c1], self.k, out=(self.sp_data[c0:c1], self.sp_index[c0:c1]))
            # --- forward per block of output rows, each block's Y downloaded as soon as it is done
            slot.y_d.zero_()
            w4 = self.warp4.view(-1, 4)
            for (r0, r1), (t0, t1) in zip(zip(self.row_bounds[:-1], self.row_bounds[1:]),
                                           zip(self.task_bounds[:-1], self.task_bounds[1:])):
                if t1 > t0:
                    ops.spgemm_forward(w4[t0:t1].reshape(-1), self.indices, self.values, self.sp_data,
                                       self.sp_index, feat_dim=self.d, num_rows=self.n_rows,
                                       out=slot.y_d, accumulate=True)
                ev = torch.cuda.Event()
                ev.record(main)
                with torch.cuda.stream(s_out):
                    s_out.wait_event(ev)
                    slot.y_h[r0:r1].copy_(slot.y_d[r0:r1], non_blocking=True)
            # --- backward behind the dY upload; dXc download last
            main.wait_event(ev_dy)
            ops.sspmm_backward(self.warp4, self.indices, self.values, slot.dy_d, self.sp_index,
                               out=slot.dxc_d)
            slot.inputs_free = torch.cuda.Event()
            slot.inputs_free.record(main)
        with torch.cuda.stream(s_out):
            s_out.wait_event(slot.inputs_free)
            slot.dxc_h.copy_(slot.dxc_d, non_blocking=True)
            slot.done = torch.cuda.Event()
            slot.done.record(s_out)
        return ticket

    def result(self, ticket: int):
        """Block until the step behind `ticket` is complete; returns (y_host f32[n_rows, d],
        dxc_host f32[n_cols, k]) -- pinned buffers owned by the aggregator, valid until the
        submit after next reuses them."""
        if not (self._submitted - len(self.slots) <= ticket < self._submitted):
            raise RuntimeError(f"HostAggregator.result: ticket {ticket} is no longer in flight")
        slot = self.slots[ticket % len(self.slots)]
        slot.done.synchronize()
        return slot.y_h, slot.dxc_h

    def step(self, x_host: torch.Tensor, dy_host: torch.Tensor):
        """Synchronous form: results are valid on return."""
        return self.result(self.submit(x_host, dy_host))


class _ShardSlot:
    """Device inputs / outputs and pinned host outputs of one in-flight step of a rank."""

    def __init__(self, n_loc, d, k, dev, pin):
        self.x_d = torch.empty((n_loc, d), dtype=torch.float32, device=dev)
        self.dy_d = torch.empty((n_loc, d), dtype=torch.float32, device=dev)
        self.y_d = torch.empty((n_loc, d), dtype=torch.float32, device=dev)
        self.dxc_d = torch.empty((max(n_loc, 1), k), dtype=torch.float32, device=dev)
        self.y_h = torch.empty((n_loc, d), dtype=torch.float32, pin_memory=pin)
        self.dxc_h = torch.empty((n_loc, k), dtype=torch.float32, pin_memory=pin)
        self.inputs_free = None   # event: the kernels that read x_d / dy_d have finished
        self.done = None          # event: y_h / dxc_h are complete


class ShardHostAggregator:
    """Host-buffer entry point of ONE RANK of a row-partitioned job (one process per GPU): wraps a
    `peer.PeerAggregator` or `dist.DistAggregator` and runs one forward + backward of the MaxK aggregation
    on this rank's HOST rows:

        y_host_loc, dxc_host_loc = shard.step(x_host_loc, dy_host_loc)            # synchronous
        t = shard.submit(x_host_loc, dy_host_loc); ...; y, dxc = shard.result(t)  # two steps in flight

    Per step: upload of the rank's X and dY rows (pinned buffers for full speed), TopK -> CBSR in place,
    exchange with the other ranks, forward, download of Y behind the backward, backward + reduce-scatter,
    download of dXc.  Every GPU uses its own PCIe link, so the host traffic of the job (the same bytes as
    on one GPU) is spread over all of them.  `submit` / `result` is the same work as a two-deep pipeline --
    the same shape as `HostAggregator` on one GPU: with two slots of buffers the uploads of step i+1 run on a
    copy stream while step i computes, and its downloads on another one while step i+1 computes, so a stream
    of steps is bound by the slowest of (upload, kernels + exchange, download) instead of by their sum.
    EVERY rank must submit the same sequence of steps (each step contains the exchange with the peers).
    This is what bench.py reports as `e2e` at N > 1.
    overlap (one synchronous step): the dY upload runs behind the forward and the Y download behind the
    backward; off = everything on the caller's stream."""

    def __init__(self, agg, overlap: bool = True, slots: int = 2):
        self.agg = agg
        dev = agg.y.device
        self.device = dev
        n_loc, d, k = int(agg.n_local), int(agg.d), int(agg.k)
        pin = dev.type == "cuda"
        self.pipelined = pin and bool(getattr(agg, "supports_out", False))
        self.slots = [_ShardSlot(n_loc, d, k, dev, pin) for _ in range(max(1, int(slots)) if self.pipelined else 1)]
        s0 = self.slots[0]
        self.x_d, self.dy_d, self.y_h, self.dxc_h = s0.x_d, s0.dy_d, s0.y_h, s0.dxc_h
        self.h2d_bytes = 2 * n_loc * d * 4
        self.d2h_bytes = n_loc * (d + k) * 4
        self.overlap = bool(overlap) and pin
        self._submitted = 0
        if pin:
            self.s_in, self.s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    # ------------------------------------------------------------------ two steps in flight
    def submit(self, x_host: torch.Tensor, dy_host: torch.Tensor) -> int:
        """Enqueue one step on this rank's host rows and return a ticket for `result`.  At most len(slots)
        steps may be in flight; the kernels run on the caller's current stream."""
        if not self.pipelined:
            raise RuntimeError("ShardHostAggregator.submit needs a CUDA aggregator with caller-owned outputs "
                               "(peer.PeerAggregator); use step()")
        slot = self.slots[self._submitted % len(self.slots)]
        ticket = self._submitted
        self._submitted += 1
        agg, cur, s_in, s_out = self.agg, torch.cuda.current_stream(self.device), self.s_in, self.s_out
        if slot.inputs_free is not None:
            s_in.wait_event(slot.inputs_free)   # this slot's previous kernels are done with x_d / dy_d
        if slot.done is not None:
            cur.wait_event(slot.done)           # ... and its previous downloads with y_d / dxc_d
        with torch.cuda.stream(s_in):
            slot.x_d.copy_(x_host, non_blocking=True)
            ev_x = torch.cuda.Event()
            ev_x.record(s_in)
            slot.dy_d.copy_(dy_host, non_blocking=True)
            ev_dy = torch.cuda.Event()
            ev_dy.record(s_in)
        cur.wait_event(ev_x)
        g_data, g_index = agg.topk_and_exchange(slot.x_d)
        y = agg.forward(g_data, g_index, out=slot.y_d)
        ev_y = torch.cuda.Event()
        ev_y.record(cur)
        with torch.cuda.stream(s_out):          # Y goes home while the backward runs
            s_out.wait_event(ev_y)
            slot.y_h.copy_(y, non_blocking=True)
        cur.wait_event(ev_dy)
        dxc = agg.backward(slot.dy_d, g_index, out=slot.dxc_d)
        slot.inputs_free = torch.cuda.Event()
        slot.inputs_free.record(cur)
        with torch.cuda.stream(s_out):
            s_out.wait_event(slot.inputs_free)
            slot.dxc_h.copy_(dxc, non_blocking=True)
            slot.done = torch.cuda.Event()
            slot.done.record(s_out)
        return ticket

    def result(self, ticket: int):
        """Block until the step behind `ticket` is complete; returns (y f32[n_local, d], dxc f32[n_local, k]) --
        pinned host buffers owned by the object, valid until the submit after next reuses them."""
        if not (self._submitted - len(self.slots) <= ticket < self._submitted):
            raise RuntimeError(f"ShardHostAggregator.result: ticket {ticket} is no longer in flight")
        slot = self.slots[ticket % len(self.slots)]
        slot.done.synchronize()
        return slot.y_h, slot.dxc_h

    # ------------------------------------------------------------------ one step at a time
    def _step_overlapped(self, x_host, dy_host):
        cur = torch.cuda.current_stream(self.device)
        self.s_in.wait_stream(cur)   # whatever read x_d / dy_d before (it ran on the caller's stream) is done
        with torch.cuda.stream(self.s_in):
            self.x_d.copy_(x_host, non_blocking=True)
            ev_x = torch.cuda.Event()
            ev_x.record(self.s_in)
            self.dy_d.copy_(dy_host, non_blocking=True)
            ev_dy = torch.cuda.Event()
            ev_dy.record(self.s_in)
        cur.wait_event(ev_x)
        g_data, g_index = self.agg.topk_and_exchange(self.x_d)
        y = self.agg.forward(g_data, g_index)
        ev_y = torch.cuda.Event()
        ev_y.record(cur)
        with torch.cuda.stream(self.s_out):   # Y goes home while the backward runs (the backward does not touch y)
            self.s_out.wait_event(ev_y)
            self.y_h.copy_(y, non_blocking=True)
        cur.wait_event(ev_dy)
        dxc = self.agg.backward(self.dy_d, g_index)
        self.dxc_h.copy_(dxc, non_blocking=True)
        cur.synchronize()
        self.s_out.synchronize()
        return self.y_h, self.dxc_h

    def step(self, x_host: torch.Tensor, dy_host: torch.Tensor):
        """x_host, dy_host: f32[n_local, d] host tensors (this rank's rows).  Returns (y f32[n_local, d],
        dxc f32[n_local, k]) in pinned host buffers owned by the object, valid until the next step."""
        if self._submitted:
            for s in self.slots:            # steps in flight use the same buffers: drain them first
                if s.done is not None:
                    s.done.synchronize()
        if self.overlap:
            return self._step_overlapped(x_host, dy_host)
        self.x_d.copy_(x_host, non_blocking=True)
        self.dy_d.copy_(dy_host, non_blocking=True)
        g_data, g_index = self.agg.topk_and_exchange(self.x_d)
        y = self.agg.forward(g_data, g_index)
        self.y_h.copy_(y, non_blocking=True)
        dxc = self.agg.backward(self.dy_d, g_index)
        self.dxc_h.copy_(dxc, non_blocking=True)
        if self.device.type == "cuda":
            torch.cuda.current_stream(self.device).synchronize()
        return self.y_h, self.dxc_h
