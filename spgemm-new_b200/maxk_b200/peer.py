"""Multi-GPU MaxK aggregation over NVLink peer memory: the collectives are OUR kernels.

Same partition and the same three calls as `dist.DistAggregator` (exchange / forward /
backward), but nothing on the data path goes through NCCL:

  * forward: ONE kernel (`maxk_spgemm_forward_fused`).  Its first CTAs push this rank's CBSR
    block (k fp32 + k uint8 per node) straight into every peer's gathered buffer with 128-bit
    stores over NVLink/NVSwitch, while the remaining CTAs already aggregate the edges whose
    columns this rank owns; the tasks that read other ranks' columns start as soon as the
    peers' blocks have arrived.  No separate all-gather launch, no second stream;
  * backward: one SSpMM over all columns into a full-height CBSR-shaped partial gradient,
    then a pull reduce-scatter kernel: every rank reads its own block out of all partials
    (again over NVLink) and adds them in rank order -- deterministic, unlike a ring;
  * per (sender, receiver) pair either the whole block travels or -- "halo" mode, for sparse
    graphs where an all-gather of every node's CBSR row would cost more than the aggregation --
    only the rows the receiver's edges reference (row lists exchanged once at set-up).  The
    buffers use the global node numbering, blocks need not be padded to equal size.

Buffers that peers read are cudaMalloc'ed by the C library, exported as CUDA IPC handles and
exchanged once through torch.distributed (host side only).  Step numbers ("epochs") published
in flag words order producers and consumers; two sets of buffers alternate by epoch parity
(see csrc/peer.cuh for the protocol and why no further handshake is needed).

One process per GPU, all GPUs of one NVLink domain (<= 16 ranks).
"""
from __future__ import annotations

import ctypes
import os
from typing import List, Optional

import torch
import torch.distributed as dist

from . import _cabi, ops
from .dist import partition_rows_by_nnz, remap_columns  # noqa: F401  (re-exported)

PEER_MAX = 16


class PeerCtx(ctypes.Structure):
    """Mirror of `maxk_peer_ctx` (include/maxk_b200.h)."""
    _fields_ = [("world", ctypes.c_int32), ("rank", ctypes.c_int32), ("k", ctypes.c_int32),
                ("reserved", ctypes.c_int32), ("n_total", ctypes.c_int64),
                ("row_start", ctypes.c_int64 * PEER_MAX), ("rows_of", ctypes.c_int64 * PEER_MAX),
                ("g_data", ctypes.c_void_p * PEER_MAX), ("g_index", ctypes.c_void_p * PEER_MAX),
                ("partial", ctypes.c_void_p * PEER_MAX), ("flags", ctypes.c_void_p * PEER_MAX),
                ("send_rows", ctypes.c_void_p * PEER_MAX), ("send_count", ctypes.c_int64 * PEER_MAX),
                ("recv_rows", ctypes.c_void_p * PEER_MAX), ("recv_count", ctypes.c_int64 * PEER_MAX),
                ("local_sync", ctypes.c_void_p)]


class _RawCuda:
    """A cudaMalloc'ed range seen through __cuda_array_interface__ (so torch can wrap it)."""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr,
                                         "data": (int(ptr), False), "version": 3}


def _wrap(ptr: int, shape, dtype: torch.dtype, device) -> torch.Tensor:
    typestr = {torch.float32: "<f4", torch.uint8: "|u1", torch.int32: "<i4", torch.uint16: "<u2"}[dtype]
    return torch.as_tensor(_RawCuda(ptr, shape, typestr), device=device)


class _Shared:
    """One IPC-exported allocation of this rank plus the peers' mappings of theirs."""

    def __init__(self, nbytes: int, rank: int, world: int, group):
        lib = _cabi.load()
        self.own = ctypes.c_void_p()
        handle = (ctypes.c_ubyte * 64)()
        _cabi.check(lib.maxk_peer_alloc(int(nbytes), ctypes.byref(self.own), handle))
        handles: List[bytes] = [b""] * world
        dist.all_gather_object(handles, bytes(handle), group=group)
        self.ptrs: List[int] = []
        self._opened: List[int] = []
        for r in range(world):
            if r == rank:
                self.ptrs.append(int(self.own.value))
                continue
            p = ctypes.c_void_p()
            buf = (ctypes.c_ubyte * 64).from_buffer_copy(handles[r])
            _cabi.check(lib.maxk_peer_open(buf, ctypes.byref(p)))
            self.ptrs.append(int(p.value))
            self._opened.append(int(p.value))

    def release(self):
        lib = _cabi.load()
        for p in self._opened:
            lib.maxk_peer_close(ctypes.c_void_p(p))
        self._opened = []
        if self.own:
            lib.maxk_peer_free(self.own)
            self.own = ctypes.c_void_p()


class PeerAggregator:
    def __init__(self, indptr_loc, indices_glob, values_loc, bounds: List[int], rank: int, k: int,
                 feat_dim: int = 256, group=None, copy_ctas: int = 16, halo_threshold: float = 0.3,
                 homes="auto"):
        """halo_threshold: a (sender, receiver) pair exchanges only the rows the receiver's edges
        reference when those are fewer than this fraction of the sender's block, the whole block
        otherwise (0 = always the whole block, > 1 = always row lists).
        homes: 2 = the bank plan of csrc/cbsr_plan.cu (uint16 per entry) travels in place of the selectors,
        the forward runs the planned kernel and the backward decodes its columns from the plan; 0 = the
        selectors travel; "auto" = 2 where ops.use_bank_plan says the plan pays on EVERY rank (k = 32 / 64,
        long rows, L2-resident operand), decided collectively."""
        self.group = group
        assert indptr_loc.is_cuda, "the peer-memory path is CUDA only"
        self.homes = 0
        self._plan(indptr_loc, indices_glob, values_loc, bounds, rank, k, feat_dim, copy_ctas,
                   halo_threshold)
        world, dev = self.world, self.device
        if homes == "auto":
            mine = torch.tensor([1 if ops.use_bank_plan(self.k, self.num_tasks, max(self.n_local, 1), self.n_total)
                                 else 0], device=dev)
            if world > 1:
                dist.all_reduce(mine, op=dist.ReduceOp.MIN, group=group)
            homes = 2 if int(mine) else 0
        if homes not in (0, 2) or (homes and self.k not in (32, 64)):
            raise RuntimeError(f"PeerAggregator: homes={homes!r} must be 0, 2 or 'auto' (2: k = 32 or 64, k={self.k})")
        self.homes = int(homes)
        send = self._exchange_lists(group)
        # ---- peer-visible buffers (CUDA IPC), two sets alternating by epoch parity -----------
        nb_data = max(self.n_total * self.k * 4, 16)
        nb_index = max(self.n_total * self.k * (2 if self.homes else 1), 16)
        flags = _Shared(256, rank, world, group)
        self._shared = [flags]
        tables = []
        for _ in range(2):
            sd = _Shared(nb_data, rank, world, group)
            si = _Shared(nb_index, rank, world, group)
            sp = _Shared(nb_data, rank, world, group)
            self._shared += [sd, si, sp]
            tables.append((sd.ptrs, si.ptrs, sp.ptrs))
        self._wire(send, flags.ptrs, tables)
        torch.cuda.synchronize(dev)
        dist.barrier(group=group)  # every mapping exists before anybody launches

    # -------------------------------------------------------------------------- set-up
    def _exchange_lists(self, group):
        """Tell every owner which of its rows this rank's edges reference.  Returns send[p] = the
        rows of THIS rank's block that rank p references (ascending global ids, int32)."""
        recv, rank, world, dev = self._recv, self.rank, self.world, self.device
        if world == 1:
            return [recv[0]]
        recv_counts = torch.tensor([r.numel() for r in recv], device=dev, dtype=torch.int64)
        all_counts = [torch.zeros_like(recv_counts) for _ in range(world)]
        dist.all_gather(all_counts, recv_counts, group=group)
        send_counts = [int(all_counts[p][rank]) for p in range(world)]  # what p needs from me
        send_flat = torch.empty(sum(send_counts), dtype=torch.int32, device=dev)
        dist.all_to_all_single(send_flat, torch.cat(recv), send_counts,
                               [int(x) for x in recv_counts.tolist()], group=group)
        return list(torch.split(send_flat, send_counts))

    def _plan(self, indptr_loc, indices_glob, values_loc, bounds, rank, k, feat_dim, copy_ctas,
              halo_threshold, backend=None):
        """Everything a rank can work out alone: task lists and the rows it references.
        `backend`: operator set providing build_worklist (maxk_b200.ops; a CPU stand-in in the
        gloo tests of this host logic)."""
        if backend is None:
            backend = ops
        self.rank, self.world = rank, len(bounds) - 1
        assert self.world <= PEER_MAX, f"at most {PEER_MAX} ranks"
        self.bounds = list(bounds)
        self.k, self.d = int(k), int(feat_dim)
        if self.k not in (8, 16, 32, 64, 96, 128, 256):   # on every rank, before anything is launched
            raise ValueError(f"the fused peer-memory forward is built for k = 8, 16, 32, 64, 96, 128, 256 (k = {self.k}): "
                             "use comm='nccl' for other widths")
        if any((b * self.k) % 16 for b in bounds[:-1]):
            raise ValueError("row block starts must satisfy start * k % 16 == 0: use "
                             "partition_rows_by_nnz(indptr, world, align=16)")
        self.copy_ctas = int(copy_ctas)
        # measurement knob: copy CTAs of the fused forward's push (one SM's TMA unit moves ~15 GB/s over NVLink,
        # so a large block wants more of them than the default)
        env = os.environ.get("MAXK_PEER_COPY_CTAS", "").strip()
        if env:
            self.copy_ctas = int(env)
        self.halo_threshold = float(halo_threshold)
        self.n_local = bounds[rank + 1] - bounds[rank]
        self.n_total = int(bounds[-1])
        dev = indptr_loc.device
        self.device = dev
        world = self.world
        indptr_loc = indptr_loc.to(torch.int32).contiguous()
        values_loc = values_loc.to(torch.float32).contiguous()
        lo, hi = bounds[rank], bounds[rank + 1]
        self.rows_of = [bounds[i + 1] - bounds[i] for i in range(world)]
        c = indices_glob.long()
        # the buffers use the global node numbering, so column ids need no remapping
        self.b_indices = indices_glob.to(torch.int32).contiguous()
        self.b_values = values_loc
        self.b_warp4 = backend.build_worklist(indptr_loc)  # backward: one row-major list, all columns
        # forward: [tasks on own columns | tasks on other ranks' columns] over a concatenated CSR
        is_loc = (c >= lo) & (c < hi)
        rows = torch.repeat_interleave(torch.arange(self.n_local, device=dev),
                                       (indptr_loc[1:] - indptr_loc[:-1]).long())
        idx_parts, val_parts, w4_parts = [], [], []
        edge_off = 0
        for sel in (is_loc, ~is_loc):
            deg = torch.bincount(rows[sel], minlength=self.n_local)
            ip = torch.zeros(self.n_local + 1, dtype=torch.int64, device=dev)
            torch.cumsum(deg, 0, out=ip[1:])
            w4 = backend.build_worklist(ip.to(torch.int32)).view(-1, 4).clone()
            w4[:, 1] += edge_off
            idx_parts.append(self.b_indices[sel])
            val_parts.append(values_loc[sel])
            w4_parts.append(w4)
            edge_off += int(ip[-1])
        self.f_indices = torch.cat(idx_parts).contiguous()
        self.f_values = torch.cat(val_parts).contiguous()
        self.f_warp4 = torch.cat(w4_parts).reshape(-1).contiguous()
        self.wait_from_task = w4_parts[0].shape[0]
        self.num_tasks = self.f_warp4.numel() // 4
        self.num_edges = int(indices_glob.numel())
        # who needs which rows: recv[p] = rows of p's block this rank's edges reference
        bt = torch.tensor(bounds, device=dev, dtype=torch.long)
        uniq = torch.unique(c)  # sorted
        owner = torch.searchsorted(bt, uniq, right=True) - 1
        recv = [uniq[owner == p].to(torch.int32).contiguous() for p in range(world)]
        recv[rank] = recv[rank][:0]
        self._recv = recv

    def _wire(self, send, flag_ptrs, tables):
        """send[p]: rows of this rank's block that rank p references; flag_ptrs[r]: rank r's flag
        words; tables: per epoch parity (g_data, g_index, partial) pointer lists indexed by rank --
        all as seen from THIS process."""
        rank, world, dev, bounds, rows_of = self.rank, self.world, self.device, self.bounds, self.rows_of
        recv = self._recv
        send_counts = [int(s.numel()) for s in send]
        thr = self.halo_threshold
        # same numbers on both sides of a pair -> same decision
        self.send_dense = [p != rank and send_counts[p] >= thr * max(rows_of[rank], 1) for p in range(world)]
        self.recv_dense = [p != rank and recv[p].numel() >= thr * max(rows_of[p], 1) for p in range(world)]
        self._dummy = torch.zeros(4, dtype=torch.int32, device=dev)
        self._send = [s.contiguous() for s in send]
        self.exchange_rows = sum(rows_of[rank] if self.send_dense[p] else send_counts[p]
                                 for p in range(world) if p != rank)
        self.local_sync = torch.zeros(4, dtype=torch.int32, device=dev)
        self._ctx, self.g_data, self.g_index, self.partial = [], [], [], []
        for d_ptrs, i_ptrs, p_ptrs in tables:
            ctx = PeerCtx()
            ctx.world, ctx.rank, ctx.k, ctx.n_total = world, rank, self.k, self.n_total
            for r in range(world):
                ctx.row_start[r], ctx.rows_of[r] = bounds[r], rows_of[r]
                ctx.g_data[r], ctx.g_index[r] = d_ptrs[r], i_ptrs[r]
                ctx.partial[r], ctx.flags[r] = p_ptrs[r], flag_ptrs[r]
                if r != rank:
                    ctx.send_count[r], ctx.recv_count[r] = send_counts[r], recv[r].numel()
                    ctx.send_rows[r] = None if self.send_dense[r] else (
                        self._send[r].data_ptr() if send_counts[r] else self._dummy.data_ptr())
                    ctx.recv_rows[r] = None if self.recv_dense[r] else (
                        recv[r].data_ptr() if recv[r].numel() else self._dummy.data_ptr())
            ctx.local_sync = self.local_sync.data_ptr()
            self._ctx.append(ctx)
            self.g_data.append(_wrap(d_ptrs[rank], (self.n_total, self.k), torch.float32, dev))
            self.g_index.append(_wrap(i_ptrs[rank], (self.n_total, self.k),
                                      torch.uint16 if getattr(self, "homes", 0) else torch.uint8, dev))
            self.partial.append(_wrap(p_ptrs[rank], (self.n_total, self.k), torch.float32, dev))
        self._flag_ptr = flag_ptrs[rank]
        self.y = torch.zeros((self.n_local, self.d), dtype=torch.float32, device=dev)
        self.rs_out = torch.zeros((max(self.n_local, 1), self.k), dtype=torch.float32, device=dev)
        self._fwd_epoch = 0
        self._bwd_epoch = 0

    @classmethod
    def local_group(cls, indptr, indices, values, bounds, k, feat_dim=256, copy_ctas=2,
                    halo_threshold=0.3, homes=0):
        """All ranks of a job inside ONE process on ONE GPU (plain device tensors instead of IPC
        mappings, no torch.distributed): the kernels, flags and row lists are exactly those of
        the multi-process job.  For tests and debugging on a single-GPU box -- the ranks' kernels
        wait for each other, so every step must be ENQUEUED for all ranks (one stream each)
        before anything synchronises, and the grids must be small enough to be co-resident."""
        world = len(bounds) - 1
        aggs = []
        for r in range(world):
            lo, hi = bounds[r], bounds[r + 1]
            e0, e1 = int(indptr[lo]), int(indptr[hi])
            a = cls.__new__(cls)
            a.group = None
            a._shared = []
            a.homes = int(homes)
            assert indptr.is_cuda, "the peer-memory path is CUDA only"
            a._plan((indptr[lo:hi + 1] - indptr[lo]).to(torch.int32), indices[e0:e1].contiguous(),
                    values[e0:e1].contiguous(), list(bounds), r, k, feat_dim, copy_ctas, halo_threshold)
            aggs.append(a)
        dev = aggs[0].device
        n_total = aggs[0].n_total
        keep = []

        def alloc(shape, dtype):
            ts = [torch.zeros(shape, dtype=dtype, device=dev) for _ in range(world)]
            keep.append(ts)
            return [t.data_ptr() for t in ts]
        flag_ptrs = alloc((64,), torch.int32)
        tables = [(alloc((max(n_total, 1), k), torch.float32),
                   alloc((max(n_total, 1), k), torch.uint16 if homes else torch.uint8),
                   alloc((max(n_total, 1), k), torch.float32)) for _ in range(2)]
        for r, a in enumerate(aggs):
            a._keepalive = keep
            a._wire([aggs[p]._recv[r] for p in range(world)], flag_ptrs, tables)
        torch.cuda.synchronize(dev)
        return aggs

    @classmethod
    def from_global_csr(cls, indptr, indices, values, bounds, rank, k, feat_dim=256, **kw):
        lo, hi = bounds[rank], bounds[rank + 1]
        e0, e1 = int(indptr[lo]), int(indptr[hi])
        ip = (indptr[lo:hi + 1] - indptr[lo]).to(torch.int32)
        return cls(ip, indices[e0:e1].contiguous(), values[e0:e1].contiguous(), bounds, rank, k,
                   feat_dim, **kw)

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def own_block(self, parity: int):
        """Views of this rank's rows inside the gathered buffers (TopK can write them in place)."""
        r0 = self.bounds[self.rank]
        return (self.g_data[parity][r0:r0 + self.n_local], self.g_index[parity][r0:r0 + self.n_local])

    # ------------------------------------------------------------------------- forward
    def topk_and_exchange(self, x_loc: torch.Tensor):
        """TopK straight into the peer-visible buffer (no staging copy), then publish."""
        par = (self._fwd_epoch + 1) & 1
        if self.homes:
            d_own, a_own = self.own_block(par)
            if getattr(self, "_sel_tmp", None) is None:
                self._sel_tmp = torch.empty((self.n_local, self.k), dtype=torch.uint8, device=self.device)
            ops.topk_cbsr(x_loc, self.k, out=(d_own, self._sel_tmp))
            ops.cbsr_bank_plan(self._sel_tmp, out=a_own)
        else:
            ops.topk_cbsr(x_loc, self.k, out=self.own_block(par))
        return self._publish_forward()

    def exchange(self, sp_data_loc: torch.Tensor, sp_index_loc: torch.Tensor, sp_addr_loc=None):
        """Place this rank's CBSR block where the peers will pull it from and publish it.
        Returns the gathered buffers of this step; the other ranks' blocks are only valid once
        `forward` has run."""
        par = (self._fwd_epoch + 1) & 1
        d_own, i_own = self.own_block(par)
        d_own.copy_(sp_data_loc)
        if self.homes and sp_addr_loc is not None:   # plan made by the caller (next to its TopK)
            i_own.copy_(sp_addr_loc)
        elif self.homes:   # the plan of this rank's rows travels in place of their selectors
            ops.cbsr_bank_plan(sp_index_loc.contiguous(), out=i_own)
        else:
            i_own.copy_(sp_index_loc)
        return self._publish_forward()

    def _publish_forward(self):
        # nothing to signal here: the fused forward kernel pushes the block itself and stream
        # order already puts the TopK / copy above in front of it
        self._fwd_epoch += 1
        par = self._fwd_epoch & 1
        return self.g_data[par], self.g_index[par]

    supports_out = True   # forward / backward take caller-owned output buffers (host.ShardHostAggregator's pipeline)

    def forward(self, g_data=None, g_index=None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Y for this rank's rows, f32[n_local, d]: into `out` if given, else into the aggregator's own buffer
        (valid until the next forward)."""
        par = self._fwd_epoch & 1
        y = self.y if out is None else out
        assert y.is_contiguous() and tuple(y.shape) == (self.n_local, self.d) and y.dtype == torch.float32
        with torch.cuda.device(self.device):
            if self.homes:
                _cabi.check(_cabi.load().maxk_spgemm_forward_fused_planned(
                    ctypes.byref(self._ctx[par]), self._fwd_epoch, self.f_warp4.data_ptr(),
                    self.num_tasks, self.wait_from_task, self.f_indices.data_ptr(), self.f_values.data_ptr(),
                    y.data_ptr(), self.n_local, self.d, self.copy_ctas, self._stream()))
            else:
                _cabi.check(_cabi.load().maxk_spgemm_forward_fused(
                    ctypes.byref(self._ctx[par]), self._fwd_epoch, self.f_warp4.data_ptr(), self.num_tasks,
                    self.wait_from_task, self.f_indices.data_ptr(), self.f_values.data_ptr(),
                    y.data_ptr(), self.n_local, self.d, self.copy_ctas, self._stream()))
        return y

    # ------------------------------------------------------------------------ backward
    def backward(self, dy_loc: torch.Tensor, g_index: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """-> dXc for the rows this rank owns, f32[n_local, k] (into `out` if given: contiguous, 16-byte aligned)."""
        rs = self.rs_out if out is None else out
        assert rs.is_contiguous() and rs.shape[0] >= self.n_local and rs.shape[1] == self.k and rs.dtype == torch.float32
        self._bwd_epoch += 1
        par = self._bwd_epoch & 1
        lib = _cabi.load()
        with torch.cuda.device(self.device):
            _cabi.check(lib.maxk_peer_zero_partial(ctypes.byref(self._ctx[par]), self._stream()))
        # Tasks on OTHER ranks' columns first: their partial gradients are what the peers wait for, so they are
        # published before the tasks on this rank's own columns run (those only touch the own block, which nobody
        # else reads) -- the peers' pulls and this rank's wait for their flags overlap with that second kernel.
        bwd = ops.sspmm_backward_planned if self.homes else ops.sspmm_backward
        w4 = self.f_warp4.view(-1, 4)
        split = int(self.wait_from_task)
        bwd(w4[split:].reshape(-1), self.f_indices, self.f_values, dy_loc, g_index, out=self.partial[par], accumulate=True)
        with torch.cuda.device(self.device):
            _cabi.check(lib.maxk_peer_publish(ctypes.c_void_p(self._flag_ptr + 4 * 1), self._bwd_epoch,
                                              self._stream()))
        bwd(w4[:split].reshape(-1), self.f_indices, self.f_values, dy_loc, g_index, out=self.partial[par], accumulate=True)
        with torch.cuda.device(self.device):
            _cabi.check(lib.maxk_reduce_scatter_pull(ctypes.byref(self._ctx[par]), self._bwd_epoch,
                                                     rs.data_ptr(), self._stream()))
        return rs[: self.n_local]

    def close(self):
        torch.cuda.synchronize(self.device)
        if self._shared:
            dist.barrier(group=self.group)  # nobody still reads our memory
        self.g_data, self.g_index, self.partial = [], [], []
        for s in self._shared:
            s.release()
        self._shared = []
