#!/usr/bin/env python3
"""bench.py -- headline benchmark of the MaxK-GNN sparse-aggregation hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl own|reference] [--k 32]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Metric (BASELINE.json): SpGEMM + SSpMM algorithmic GB/s on the synthetic Reddit-shaped
graph (N = 232 965, E = 114 615 892, hidden 256, k = 32): one STEP = one forward row-wise
SpGEMM  Y = A * X_cbsr  plus one backward sampled SSpMM  dXc = sample(A^T dY).
   value    = (B_fwd + B_bwd) / device time, inputs resident in HBM   
   e2e      = same bytes / time of HostAggregator.step() on pinned HOST buffers
              (H2D X and dY, TopK, both kernels, D2H Y and dXc inside the timed region); at N > 1 every rank
              does that for its own rows through ShardHostAggregator.step() (max over ranks)
   roofline = dominant kernel's algorithmic GB/s against MEASURED_PEAKS.json's HBM figure
   cpu_baseline = the reference's CPU path (torch.topk + CSR SpMM, oracle/cpu_path.py) on a
              bounded row-sample of the same graph, on this box's host cores
N > 1: the adjacency is 1-D row-partitioned (balanced by nnz) over the ranks; every step
exchanges the CBSR features and reduce-scatters the CBSR-shaped partial gradients -- with our own
kernels over NVLink peer memory (--comm peer, default) or NCCL collectives (--comm nccl); strong
scaling: the graph is fixed.
   epoch    = (extra leg, after the timed region) full-graph training epochs/s of GraphSAGE MaxK k=32, 3 x 256,
              LayerNorm, Adam on the ogbn-products-shaped graph (BASELINE.json configs[2]) at the N of the run --
              the second half of BASELINE.json's metric; maxk_b200.train.run_epochs
   parity_ok = (N > 1, outside the timed region) every rank checks its rows of Y and dXc of one step against the
              single-GPU kernels run on the gathered operands
`--impl reference` times the reference's CPU implementation of the same step on a bounded
sample (rank 0 only) and, on a GPU box, the reference's own CUDA kernels (oracle/_ref, kernels/*.cu compiled
unmodified for sm_100a) on the full graph.  One JSON line on stdout; everything else goes to stderr.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "spgemm-new_b200"))

import torch  # noqa: E402

METRIC = "spgemm_sspmm_algorithmic_GBps"
UNIT = "GB/s"
FEAT_DIM = 256
HBM_FALLBACK_GBS = 6650.0  # B200_PROFILING.md fallback ("of fallback")


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:  # pragma: no cover
            pass
    return HBM_FALLBACK_GBS, "fallback"


def dram_traffic(kernel: str, k: int):
    """dram bytes per launch from the committed ncu --set full capture (profiles/), or None."""
    p = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    try:
        return json.load(open(p)).get(f"{kernel}_k{k}")
    except Exception:
        return None


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index: int, period_s: float = 0.02):
        super().__init__(daemon=True)
        self.index, self.period = index, period_s
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_ev = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:  # pragma: no cover
            log("clock sampling unavailable:", e)

    NAMES = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap",
             0x8: "hw_slowdown", 0x10: "sync_boost", 0x20: "sw_thermal_slowdown",
             0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def run(self):
        if not self.ok:
            return
        while not self._stop_ev.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                r = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in self.NAMES.items():
                    if r & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:  # pragma: no cover
                pass
            time.sleep(self.period)

    def stop(self):
        self._stop_ev.set()
        if self.is_alive():
            self.join(timeout=1.0)
        return {"sm_mhz": (statistics.median(self.samples) if self.samples else None),
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def physical_gpu_index(local_rank: int) -> int:
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_rank])
        except Exception:
            return local_rank
    return local_rank


# --------------------------------------------------------------------------------------
def make_reddit(device, seed=0):
    from maxk_b200 import graphs
    scale = float(os.environ.get("MAXK_BENCH_SCALE", "1.0"))  # < 1 only for smoke runs
    t0 = time.time()
    g = graphs.make_graph("reddit", seed=seed, device=device, values="uniform", scale=scale)
    log(f"[graph] reddit-shaped N={g['v_num']} E={g['e_num']} on {device} in {time.time() - t0:.1f}s")
    return g, scale


def cpu_sample(g, frac_rows: float):
    """First rows of the graph (a contiguous row block keeps the degree mix) as a CPU problem."""
    n = g["v_num"]
    rows = max(64, int(n * frac_rows))
    indptr = g["indptr"][: rows + 1].cpu()
    e = int(indptr[-1])
    return {"indptr": indptr, "indices": g["indices"][:e].cpu(), "values": g["values"][:e].cpu(),
            "rows": rows, "e": e, "n_cols": n}


def run_cpu_reference(g, k, steps, warmup, target_step_s=2.0):
    """The reference's CPU path on a bounded sample; returns (GB/s, info dict)."""
    from maxk_b200 import graphs
    from oracle import cpu_path
    cores = cpu_path.set_threads()
    n = g["v_num"]
    gen = torch.Generator().manual_seed(123)
    x = torch.rand(n, FEAT_DIM, generator=gen)
    h, mask = cpu_path.CpuAggregation.maxk(x, k)   # TopK is outside the step, as on the GPU
    # pilot on 1/64 of the rows to size the sample for ~target_step_s per step
    frac = 1.0 / 64
    s = cpu_sample(g, frac)
    agg = cpu_path.CpuAggregation(s["indptr"], s["indices"], s["values"], n_cols=n)
    dy = torch.rand(s["rows"], FEAT_DIM, generator=gen)
    agg.step(h, dy, mask)
    _, t = agg.timed_step(h, dy, mask)
    frac = min(1.0, max(frac, frac * target_step_s / max(t, 1e-3)))
    s = cpu_sample(g, frac)
    agg = cpu_path.CpuAggregation(s["indptr"], s["indices"], s["values"], n_cols=n)
    dy = torch.rand(s["rows"], FEAT_DIM, generator=gen)
    deg = (s["indptr"][1:] - s["indptr"][:-1]).long()
    w = int(((deg + 63) // 64).sum())
    b = graphs.algorithmic_bytes(s["rows"], s["e"], w, k, FEAT_DIM, n_cbsr=n)
    step_bytes = b["fwd"] + b["bwd"]
    for _ in range(warmup):
        agg.step(h, dy, mask)
    times = []
    for _ in range(steps):
        _, t = agg.timed_step(h, dy, mask)
        times.append(t)
    ms = 1e3 * sum(times) / len(times)
    value = step_bytes / (ms * 1e-3) / 1e9
    info = {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"first {s['rows']} of {n} rows ({s['e']} of {g['e_num']} edges) of the same "
                      f"Reddit-shaped graph, k={k}, MaxK-masked dense features, torch.sparse.mm CSR (fwd) / A^T CSR * mask (bwd), "
                      f"fp32, {cores} threads",
            "ms_per_step": ms, "edges_per_s": s["e"] * 2 / (ms * 1e-3)}
    return value, ms, info


def timeit_events(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def time_reference_kernels(g, k, dev):
    """The reference's CUDA kernels (oracle/_ref: kernels/spmm_maxk.cu, spmm_maxk_backward.cu compiled unmodified
    for sm_100a, reference launch shape 12 warps / CTA) on the full graph of the bench, reference protocol
    (4 + 4 launches, outputs not re-zeroed: spmm_base.h:56-74).  None if there is no GPU or no oracle/_ref."""
    if not torch.cuda.is_available():
        return None
    try:
        from maxk_b200 import ops
        from oracle import ref_kernels as R
        if not R.available():
            return None
        n = g["v_num"]
        warp4 = ops.build_worklist(g["indptr"])
        gen = torch.Generator(dev).manual_seed(123)
        x = torch.rand(n, FEAT_DIM, device=dev, generator=gen)
        dy = torch.rand(n, FEAT_DIM, device=dev, generator=gen)
        sp_data, sp_index = ops.topk_cbsr(x, k)
        yr = torch.zeros(n, FEAT_DIM, device=dev)
        dr = torch.zeros(n, k, device=dev)
        tf = timeit_events(lambda: R.spmm_forward(warp4, g["indices"], g["values"], sp_data, sp_index, FEAT_DIM, out=yr), 4, 4)
        tb = timeit_events(lambda: R.spmm_backward(warp4, g["indices"], g["values"], dy, sp_index, out=dr), 4, 4)
        return {"fwd_ms": tf, "bwd_ms": tb, "ms_per_step": tf + tb,
                "note": "the reference's kernels/*.cu compiled unmodified for sm_100a (oracle/_ref), reference launch "
                        "shape, full graph, outputs not re-zeroed (reference protocol, spmm_base.h:56-74)"}
    except Exception as exc:  # noqa: BLE001
        log("reference kernels on this GPU: skipped:", exc)
        return None


def multi_gpu_parity(agg, g, bounds, rank, world, k, dev, sd_loc, si_loc, dy_loc, dist, ops):
    """One more step of the N-GPU job, outside the timed region, checked on EVERY rank against the single-GPU
    kernels: the ranks' CBSR blocks and dY rows are gathered with plain broadcasts, each rank runs the one-GPU
    forward and backward on the full graph and compares its own rows.  Tolerance: north-star rtol 1e-5 / atol 1e-6,
    widened to the fp32 summation bound (fan-in + 2) * 2^-24 * sum|terms| where a row has thousands of terms (all
    terms are non-negative here, so sum|terms| = |reference|)."""
    n = g["v_num"]
    lo, hi = bounds[rank], bounds[rank + 1]
    gsd = torch.empty((n, k), dtype=torch.float32, device=dev)
    gsi = torch.empty((n, k), dtype=torch.uint8, device=dev)
    gdy = torch.empty((n, FEAT_DIM), dtype=torch.float32, device=dev)
    gsd[lo:hi], gsi[lo:hi], gdy[lo:hi] = sd_loc, si_loc, dy_loc
    for r in range(world):
        a, b = bounds[r], bounds[r + 1]
        if b > a:
            for t in (gsd, gsi, gdy):
                dist.broadcast(t[a:b], src=r)
    gd, gi = agg.exchange(sd_loc, si_loc)
    y = agg.forward(gd, gi).clone()
    dxc = agg.backward(dy_loc, gi).clone()
    warp4 = ops.build_worklist(g["indptr"])
    y_ref = ops.spgemm_forward(warp4, g["indices"], g["values"], gsd, gsi, num_rows=n)[lo:hi]
    dxc_ref = ops.sspmm_backward(warp4, g["indices"], g["values"], gdy, gsi)[lo:hi]
    deg = (g["indptr"][1:] - g["indptr"][:-1])[lo:hi].double()
    col_deg = torch.bincount(g["indices"].long(), minlength=n)[lo:hi].double()

    def excess(got, ref, fan):
        got, ref = got.double(), ref.double()
        allowed = torch.maximum(1e-6 + 1e-5 * ref.abs(), (fan[:, None] + 2.0) * 2.0 ** -24 * ref.abs())
        err = (got - ref).abs()
        strict = (err <= 1e-6 + 1e-5 * ref.abs()).double().mean()
        return float((err - allowed).max()), float(err.max()), float(strict)
    ey, ay, sy = excess(y, y_ref, deg)
    eg, ag, sg = excess(dxc, dxc_ref, col_deg)
    t = torch.tensor([max(ey, eg), ay, ag, -sy, -sg], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return {"parity_ok": bool(float(t[0]) <= 0.0), "max_abs_err_y": float(t[1]), "max_abs_err_dxc": float(t[2]),
            "frac_within_rtol1e-5_atol1e-6": {"y": -float(t[3]), "dxc": -float(t[4])},
            "checked": f"every rank: its rows of Y and dXc of one extra step (fused push all-gather + SpGEMM, SSpMM + pull "
                       f"reduce-scatter at N={world}) against the single-GPU kernels on the gathered operands; tolerance "
                       "rtol 1e-5 / atol 1e-6 widened to the fp32 summation bound (fan-in + 2) * 2^-24 * |ref|"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--k", type=int, default=32)
    ap.add_argument("--sweep", action="store_true", help="also time k=8/16/64, TopK and the work list")
    ap.add_argument("--no-plan", action="store_true",
                    help="measurement knob: the forward without the bank plan (csrc/cbsr_plan.cu) even where "
                         "ops.use_bank_plan says it pays")
    ap.add_argument("--epochs", type=int, default=5,
                    help="epochs of the training leg (products-shaped graph, SAGE MaxK k=32; 0 = skip the leg)")
    ap.add_argument("--no-parity", action="store_true", help="N>1: skip the parity check of one extra step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--comm", choices=("peer", "nccl"), default=None,
                    help="N>1: our fused kernels over NVLink peer memory (default) or NCCL collectives")
    ap.add_argument("--ref-kernels", action="store_true",
                    help="also time the reference CUDA kernels (oracle/_ref) on the same inputs")
    args = ap.parse_args()
    if args.impl == "own" and args.warmup < 3:
        log(f"[bench] --warmup {args.warmup} raised to 3 (timing rules: at least 3 warm-up steps); the line reports 3")
        args.warmup = 3

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        if rank != 0:
            return 0
        dev = "cuda" if torch.cuda.is_available() else "cpu"
        g, scale = make_reddit(dev)
        value, ms, info = run_cpu_reference(g, args.k, max(1, args.steps), max(0, args.warmup))
        ref_gpu = time_reference_kernels(g, args.k, dev)
        if ref_gpu is not None:
            from maxk_b200 import graphs as _graphs
            deg_all = (g["indptr"][1:] - g["indptr"][:-1]).long()
            w_all = int(((deg_all + 63) // 64).sum())
            b_all = _graphs.algorithmic_bytes(g["v_num"], g["e_num"], w_all, args.k, FEAT_DIM)
            ref_gpu["GBps"] = (b_all["fwd"] + b_all["bwd"]) / (ref_gpu["ms_per_step"] * 1e-3) / 1e9
            log(f"[ref kernels on this GPU] fwd {ref_gpu['fwd_ms']:.3f} ms  bwd {ref_gpu['bwd_ms']:.3f} ms  {ref_gpu['GBps']:.0f} GB/s")
        line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic",
                # the own arm's workload, word for word; what the CPU actually ran per step is in `sample`
                "config": {"workload": f"reddit-shaped graph N={g['v_num']} E={g['e_num']}: forward SpGEMM + backward "
                                       f"SSpMM, k={args.k}, hidden {FEAT_DIM}",
                           "k": args.k, "feat_dim": FEAT_DIM, "graph_scale": scale,
                           "sample": "bounded row sample of that graph on the host cores: " + info["sample"]},
                "cpu_baseline": info,
                "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        line["config"]["note"] = ("the CPU step times forward + backward only (its TopK runs before the timed region); "
                                  "the own arm's e2e additionally has TopK and the host<->device copies inside")
        if ref_gpu is not None:
            line["reference_kernels_on_this_gpu"] = ref_gpu
        print(json.dumps(line), flush=True)
        return 0

    # ----------------------------------------------------------------------------- own arm
    if not torch.cuda.is_available():
        log("bench.py --impl own needs a CUDA device (no CPU fallback).")
        return 2
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from maxk_b200 import _cabi, graphs, ops
    from maxk_b200.host import HostAggregator, bind_to_device_numa_node
    numa_node = bind_to_device_numa_node(dev)   # pinned host buffers of the e2e leg land next to this rank's GPU

    g, scale = make_reddit(dev)          # same seed on every rank -> identical graph
    n, e, k = g["v_num"], g["e_num"], args.k
    peak, peak_src = measured_peak()

    if world > 1:
        from maxk_b200.dist import DistAggregator, partition_rows_by_nnz
        bounds = partition_rows_by_nnz(g["indptr"], world, align=16)
        comm = args.comm or "peer"
        agg = None
        if comm == "peer":
            # fused push all-gather + SpGEMM kernel and pull reduce-scatter over NVLink peer
            # memory (maxk_b200/peer.py).  If CUDA IPC cannot be set up on this box every rank
            # agrees to use the NCCL collectives instead -- a transport choice, still our kernels.
            from maxk_b200.peer import PeerAggregator
            try:
                agg = PeerAggregator.from_global_csr(g["indptr"], g["indices"], g["values"], bounds,
                                                     rank, k, FEAT_DIM, copy_ctas=16,
                                                     homes=0 if args.no_plan else "auto")
                ok = 1
            except Exception as exc:  # noqa: BLE001
                log(f"[rank {rank}] peer-memory path unavailable: {exc}")
                ok = 0
            flag = torch.tensor([ok], device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if int(flag) == 0:
                agg, comm = None, "nccl"
        if agg is None:
            agg = DistAggregator.from_global_csr(g["indptr"], g["indices"], g["values"], bounds, rank, k,
                                                 FEAT_DIM)
        n_loc = agg.n_local
        gen = torch.Generator(dev).manual_seed(123 + rank)
        x_loc = torch.rand(n_loc, FEAT_DIM, device=dev, generator=gen)
        dy_loc = torch.rand(n_loc, FEAT_DIM, device=dev, generator=gen)
        w_total = torch.tensor([agg.num_tasks], device=dev, dtype=torch.int64)
        dist.all_reduce(w_total)
        w = int(w_total)

        sd_loc, si_loc = ops.topk_cbsr(x_loc, k)   # like N=1: TopK is outside the step
        planned_peer = comm == "peer" and bool(getattr(agg, "homes", 0))

        def step():
            gd, gi = agg.exchange(sd_loc, si_loc)    # own block into the peer-visible buffer (+ its bank plan)
            agg.forward(gd, gi)                      # local columns overlap the all-gather
            agg.backward(dy_loc, gi)                 # + reduce-scatter of the partial gradients
        fwd_ms = bwd_ms = None
    else:
        warp4 = ops.build_worklist(g["indptr"])
        w = warp4.numel() // 4
        gen = torch.Generator(dev).manual_seed(123)
        x = torch.rand(n, FEAT_DIM, device=dev, generator=gen)
        dy = torch.rand(n, FEAT_DIM, device=dev, generator=gen)
        sp_data, sp_index = ops.topk_cbsr(x, k)
        use_plan = (not args.no_plan) and ops.use_bank_plan(k, w, n, n)
        sp_addr = torch.empty((n, k), dtype=torch.uint16, device=dev) if use_plan else None
        y = torch.empty(n, FEAT_DIM, device=dev)
        dxc = torch.empty(n, k, device=dev)
        ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
        step_no = [-1]

        def step():
            i = step_no[0]
            if i >= 0:
                ev[i][0].record()
            if use_plan:   # the plan is made from sp_index INSIDE the step: it is part of the forward's cost
                ops.cbsr_bank_plan(sp_index, out=sp_addr)
                ops.spgemm_forward_planned(warp4, g["indices"], g["values"], sp_data, sp_addr, out=y, num_rows=n)
            else:
                ops.spgemm_forward(warp4, g["indices"], g["values"], sp_data, sp_index, out=y, num_rows=n)
            if i >= 0:
                ev[i][1].record()
            ops.sspmm_backward(warp4, g["indices"], g["values"], dy, sp_index, out=dxc)
            if i >= 0:
                ev[i][2].record()

    bytes_ = graphs.algorithmic_bytes(n, e, w, k, FEAT_DIM)
    step_bytes = bytes_["fwd"] + bytes_["bwd"]

    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(physical_gpu_index(local_rank))
    sampler.start()  # before the barrier: starting the sampler must not skew the ranks
    launches0 = _cabi.launch_count()
    t_start, t_stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    t_start.record()
    for i in range(args.steps):
        if world == 1:
            step_no[0] = i
        step()
    t_stop.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    total_ms = t_start.elapsed_time(t_stop)
    launches = _cabi.launch_count() - launches0
    clocks = sampler.stop()
    if world > 1:
        t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t)
        lt = torch.tensor([launches], device=dev, dtype=torch.int64)
        dist.all_reduce(lt)
        launches = int(lt)
    ms_per_step = total_ms / args.steps
    value = step_bytes / (ms_per_step * 1e-3) / 1e9

    extra = {}
    roofline = None
    if world == 1:
        step_no[0] = -1
        fwd_ms = sum(a[0].elapsed_time(a[1]) for a in ev) / args.steps
        bwd_ms = sum(a[1].elapsed_time(a[2]) for a in ev) / args.steps
        dom, dom_ms, dom_bytes = (("spgemm_fwd", fwd_ms, bytes_["fwd"]) if fwd_ms >= bwd_ms
                                  else ("sspmm_bwd", bwd_ms, bytes_["bwd"]))
        ach = dom_bytes / (dom_ms * 1e-3) / 1e9
        roofline = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s",
                    "frac": ach / peak, "peak_source": f"of {peak_src}",
                    "traffic": dram_traffic(dom, k), "algorithmic_bytes": dom_bytes, "ms": dom_ms}
        extra["kernels"] = {
            "spgemm_fwd": {"ms": fwd_ms, "GBps": bytes_["fwd"] / fwd_ms / 1e6,
                           "frac_of_hbm_peak": bytes_["fwd"] / fwd_ms / 1e6 / peak,
                           "edges_per_s": e / (fwd_ms * 1e-3)},
            "sspmm_bwd": {"ms": bwd_ms, "GBps": bytes_["bwd"] / bwd_ms / 1e6,
                          "frac_of_hbm_peak": bytes_["bwd"] / bwd_ms / 1e6 / peak,
                          "edges_per_s": e / (bwd_ms * 1e-3)}}
        roofline["includes"] = ("CUDA events around the forward of every timed step: "
                                + ("bank-plan kernel (cbsr_plan.cu) + " if use_plan else "")
                                + "row-clearing pre-kernel + spgemm_fwd_kernel") if dom == "spgemm_fwd" else \
                               "CUDA events around the backward of every timed step: memset of dXc + sspmm_bwd_kernel"
        # the same kernels with the L2 flushed in front of every launch (a 512 MB buffer is
        # overwritten, outside the event pair) -- the CBSR operand then comes from HBM once per launch
        flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
        cold_f = cold_b = cold_p = 0.0
        reps = 5
        ce = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        for _ in range(reps):
            flush.zero_()
            ce[0].record()
            if use_plan:
                ops.cbsr_bank_plan(sp_index, out=sp_addr)
                ce[1].record()
                ops.spgemm_forward_planned(warp4, g["indices"], g["values"], sp_data, sp_addr, out=y, num_rows=n)
            else:
                ce[1].record()
                ops.spgemm_forward(warp4, g["indices"], g["values"], sp_data, sp_index, out=y, num_rows=n)
            ce[2].record()
            flush.zero_()
            ce[3].record()
            ops.sspmm_backward(warp4, g["indices"], g["values"], dy, sp_index, out=dxc)
            ce[4].record()
            torch.cuda.synchronize()
            cold_p += ce[0].elapsed_time(ce[1]) / reps
            cold_f += ce[0].elapsed_time(ce[2]) / reps
            cold_b += ce[3].elapsed_time(ce[4]) / reps
        del flush
        extra["kernels"]["spgemm_fwd"].update({"ms_cold_l2": cold_f, "frac_of_hbm_peak_cold_l2": bytes_["fwd"] / cold_f / 1e6 / peak,
                                               "bank_plan": bool(use_plan), "bank_plan_ms_cold_l2": cold_p if use_plan else None})
        extra["kernels"]["sspmm_bwd"].update({"ms_cold_l2": cold_b, "frac_of_hbm_peak_cold_l2": bytes_["bwd"] / cold_b / 1e6 / peak})
        log(f"[own] fwd {fwd_ms:.3f} ms  bwd {bwd_ms:.3f} ms  step {ms_per_step:.3f} ms  "
            f"{value:.0f} GB/s  roofline {dom} {roofline['frac']:.3f} of {peak_src} peak;  L2 flushed before every "
            f"launch: fwd {cold_f:.3f} ms (plan {cold_p:.3f})  bwd {cold_b:.3f} ms")
    elif not args.no_e2e:
        # N > 1: forward and backward of a few extra steps timed apart (CUDA events on every rank, max over
        # ranks); the roofline is PER GPU -- each rank's share of the algorithmic bytes against one GPU's peak
        # (a host-side slip here must not cost the line: the same exception on every rank skips the block)
        try:
            reps, f_ms, b_ms = 5, 0.0, 0.0
            evs = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            for _ in range(reps):
                gd, gi = agg.exchange(sd_loc, si_loc)
                evs[0].record()
                agg.forward(gd, gi)
                evs[1].record()
                agg.backward(dy_loc, gi)
                evs[2].record()
                torch.cuda.synchronize()
                f_ms += evs[0].elapsed_time(evs[1]) / reps
                b_ms += evs[1].elapsed_time(evs[2]) / reps
            tt = torch.tensor([f_ms, b_ms], device=dev, dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            fwd_ms, bwd_ms = float(tt[0]), float(tt[1])
            dom, dom_ms, dom_bytes = (("spgemm_fwd_fused" if comm == "peer" else "spgemm_fwd", fwd_ms, bytes_["fwd"] / world)
                                      if fwd_ms >= bwd_ms else ("sspmm_bwd+reduce_scatter", bwd_ms, bytes_["bwd"] / world))
            ach = dom_bytes / (dom_ms * 1e-3) / 1e9
            roofline = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                        "peak_source": f"of {peak_src}", "traffic": None, "algorithmic_bytes": dom_bytes, "ms": dom_ms,
                        "per_gpu": True, "note": "per GPU: 1/N of the job's algorithmic bytes over the slowest rank's phase "
                                                 "(exchange waits included), against one GPU's peak"}
            extra["kernels"] = {"forward_phase_ms": fwd_ms, "backward_phase_ms": bwd_ms}
            log(f"[own] N={world}: forward phase {fwd_ms:.3f} ms  backward phase {bwd_ms:.3f} ms  (max over ranks)")
        except Exception as exc:  # noqa: BLE001
            log(f"[rank {rank}] N > 1 roofline leg skipped: {type(exc).__name__}: {exc}")
            roofline = None
    extra["edges_per_s"] = 2 * e / (ms_per_step * 1e-3)
    extra["num_tasks"] = w

    # ------------------------------------------------------------------ optional sweeps
    def timeit(fn, iters=10, warm=3):
        for _ in range(warm):
            fn()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        for _ in range(iters):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / iters

    if world == 1 and args.sweep:
        sweep = {}
        for kk in (8, 16, 32, 64):
            sd, si = ops.topk_cbsr(x, kk)
            dx2 = torch.empty(n, kk, device=dev)
            bb = graphs.algorithmic_bytes(n, e, w, kk, FEAT_DIM)
            tf = timeit(lambda: ops.spgemm_forward_auto(warp4, g["indices"], g["values"], sd, si, out=y, num_rows=n))
            tb = timeit(lambda: ops.sspmm_backward(warp4, g["indices"], g["values"], dy, si, out=dx2))
            tt = timeit(lambda: ops.topk_cbsr(x, kk))
            sweep[f"k{kk}"] = {"fwd_ms": tf, "bwd_ms": tb, "topk_ms": tt,
                               "fwd_frac": bb["fwd"] / tf / 1e6 / peak, "bwd_frac": bb["bwd"] / tb / 1e6 / peak,
                               "topk_frac": bb["topk"] / tt / 1e6 / peak}
            log(f"[sweep] k={kk}: fwd {tf:.3f} ms ({sweep[f'k{kk}']['fwd_frac']:.2f}) "
                f"bwd {tb:.3f} ms ({sweep[f'k{kk}']['bwd_frac']:.2f}) topk {tt:.3f} ms ({sweep[f'k{kk}']['topk_frac']:.2f})")
        tm = timeit(lambda: ops.build_worklist(g["indptr"]))
        sweep["worklist_ms"] = tm
        sweep["worklist_frac"] = bytes_["meta"] / tm / 1e6 / peak
        extra["sweep"] = sweep
    if world == 1 and args.ref_kernels:   # (the driver gets this from `--impl reference`; here for one-call comparisons)
        rk = time_reference_kernels(g, k, dev)
        if rk is not None:
            extra["reference_kernels_on_this_gpu"] = rk
            log(f"[ref kernels] fwd {rk['fwd_ms']:.3f} ms  bwd {rk['bwd_ms']:.3f} ms")

    # ------------------------------------------------- N > 1: did the timed step compute the right thing?
    parity = None
    if world > 1 and not args.no_parity:
        ok = 1
        try:
            parity = multi_gpu_parity(agg, g, bounds, rank, world, k, dev, sd_loc, si_loc, dy_loc, dist, ops)
        except Exception as exc:  # noqa: BLE001
            log(f"[rank {rank}] parity leg failed: {type(exc).__name__}: {exc}")
            ok = 0
        flag = torch.tensor([ok], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag) == 0:
            parity = {"parity_ok": False, "error": "the parity leg raised on at least one rank (see stderr)"}
        if rank == 0:
            log(f"[parity] N={world}: {parity}")
    if world > 1:
        g.pop("indices", None)
        g.pop("values", None)
        torch.cuda.empty_cache()

    # ------------------------------------------------------------------------------ e2e
    e2e = None
    if world == 1 and not args.no_e2e:
        host = HostAggregator(g["indptr"], g["indices"], g["values"], k, FEAT_DIM, dev)
        xh = x.cpu().pin_memory()
        dyh = dy.cpu().pin_memory()
        e_steps = max(3, min(args.steps, 10))
        for _ in range(2):
            host.step(xh, dyh)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(e_steps):  # one step at a time: upload, kernels, download, return
            host.step(xh, dyh)
        torch.cuda.synchronize()
        sync_ms = (time.perf_counter() - t0) * 1e3 / e_steps
        # a stream of steps through the same public API, two in flight: submit(i+1) before
        # result(i).  Every step still uploads its own X and dY and downloads its own Y, dXc.
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        prev = None
        for _ in range(e_steps):
            t = host.submit(xh, dyh)
            if prev is not None:
                host.result(prev)
            prev = t
        host.result(prev)
        torch.cuda.synchronize()
        e_ms = (time.perf_counter() - t0) * 1e3 / e_steps
        e2e = {"value": step_bytes / (e_ms * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": e_ms,
               "h2d_bytes_per_step": host.h2d_bytes, "d2h_bytes_per_step": host.d2h_bytes,
               "steps": e_steps, "ms_per_step_synchronous": sync_ms, "numa_node_of_rank0": numa_node,
               "includes": "per step: H2D X+dY (pinned), TopK->CBSR, SpGEMM, SSpMM, D2H Y+dXc; "
                           "HostAggregator.submit/result with two steps in flight "
                           "(ms_per_step_synchronous = HostAggregator.step, one at a time)"}
        log(f"[e2e] pipelined {e_ms:.2f} ms/step  {e2e['value']:.0f} GB/s   synchronous {sync_ms:.2f} ms/step")
        del host, xh, dyh
    elif world > 1 and not args.no_e2e:
        # every rank: its own rows from pinned host memory in, its own rows of Y and dXc back out, through
        # maxk_b200.host.ShardHostAggregator (upload, TopK in place, exchange, forward, backward + reduce-scatter,
        # download; results valid on return).  Set-up is agreed on by all ranks before anybody enters a step.
        ok = 1
        try:
            from maxk_b200.host import ShardHostAggregator
            shard = ShardHostAggregator(agg)
            xh = x_loc.cpu().pin_memory()
            dyh = dy_loc.cpu().pin_memory()
        except Exception as exc:  # noqa: BLE001
            log(f"[rank {rank}] e2e set-up failed: {exc}")
            ok = 0
        flag = torch.tensor([ok], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag) == 1:
            try:
                e_steps = max(3, min(args.steps, 10))
                for _ in range(2):
                    shard.step(xh, dyh)
                torch.cuda.synchronize()
                dist.barrier()
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for _ in range(e_steps):   # one step at a time: upload, kernels + exchange, download, return
                    shard.step(xh, dyh)
                torch.cuda.synchronize()
                sync_ms = (time.perf_counter() - t0) * 1e3 / e_steps
                pipelined = shard.pipelined
                e_ms = sync_ms
                if pipelined:
                    # a stream of steps through the same public API, two in flight: submit(i+1) before result(i).
                    # Every step still uploads its own X and dY rows and downloads its own Y and dXc rows.
                    dist.barrier()
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    prev = None
                    for _ in range(e_steps):
                        t = shard.submit(xh, dyh)
                        if prev is not None:
                            shard.result(prev)
                        prev = t
                    shard.result(prev)
                    torch.cuda.synchronize()
                    e_ms = (time.perf_counter() - t0) * 1e3 / e_steps
                tmax = torch.tensor([e_ms, sync_ms], device=dev, dtype=torch.float64)
                dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
                nbytes = torch.tensor([shard.h2d_bytes, shard.d2h_bytes], device=dev, dtype=torch.int64)
                dist.all_reduce(nbytes)
                e_ms, sync_ms = float(tmax[0]), float(tmax[1])
                e2e = {"value": step_bytes / (e_ms * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": e_ms,
                       "h2d_bytes_per_step": int(nbytes[0]), "d2h_bytes_per_step": int(nbytes[1]), "steps": e_steps,
                       "ms_per_step_synchronous": sync_ms, "numa_node_of_rank0": numa_node,
                       "includes": f"per step and rank (max over {world} ranks, wall clock between synchronisations): H2D of "
                                   "the rank's X and dY rows (pinned), TopK->CBSR in place, exchange, forward, backward + "
                                   "reduce-scatter, D2H of its Y and dXc rows; "
                                   + ("ShardHostAggregator.submit/result with two steps in flight "
                                      "(ms_per_step_synchronous = ShardHostAggregator.step, one at a time)" if pipelined
                                      else "ShardHostAggregator.step, one step at a time")
                                   + "; bytes are summed over the ranks"}
                log(f"[e2e] {e_ms:.2f} ms/step  {e2e['value']:.0f} GB/s over {world} ranks (synchronous {sync_ms:.2f} ms/step)")
                del shard, xh, dyh
            except Exception as exc:  # noqa: BLE001
                log(f"[rank {rank}] N > 1 e2e leg skipped: {type(exc).__name__}: {exc}")
                e2e = None

    # ------------------------------------------------------------------ CPU baseline leg
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            _, _, cpu_baseline = run_cpu_reference(g, k, steps=3, warmup=1, target_step_s=3.0)
            log(f"[cpu] {cpu_baseline['value']:.2f} GB/s on {cpu_baseline['cores']} threads")
        except Exception as ex:  # pragma: no cover
            log("cpu baseline failed:", ex)

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": f"reddit-shaped graph N={n} E={e}: forward SpGEMM + backward SSpMM, "
                                       f"k={k}, hidden {FEAT_DIM}",
                           "k": k, "feat_dim": FEAT_DIM, "num_tasks": w, "graph_scale": scale,
                           "parallelism": ("single GPU" if world == 1 else
                                           (f"1-D row partition x{world}, fused push all-gather + SpGEMM kernel and pull "
                                            "reduce-scatter over NVLink peer memory (own kernels, no NCCL on the data path)"
                                            if comm == "peer" else
                                            f"1-D row partition x{world}, CBSR all-gather + reduce-scatter (NCCL)")),
                           "l2": "inputs larger than L2 (CSR stream 917 MB per kernel); CBSR operand "
                                 "(37 MB) is L2-resident by design",
                           "step_algorithmic_bytes": step_bytes},
                "clocks": clocks, "gpu_launches": launches}
        line["config"]["timed_region"] = ("value: forward SpGEMM (incl. its bank-plan kernel where used) + backward SSpMM of "
                                          "every step; TopK->CBSR and the work list are made once, before the timed region "
                                          "(their times: --sweep / profiles/r02_sweep.json); e2e has TopK inside")
        if world == 1:
            line["config"]["forward"] = ("through the two-home bank plan (cbsr_plan.cu), plan kernel inside the step"
                                         if use_plan else "default kernel (no bank plan)")
        elif comm == "peer":
            line["config"]["forward"] = ("bank plan made inside the step, travels over NVLink in place of the selectors; the "
                                         "backward decodes its columns from it" if planned_peer else "selectors travel (no bank plan)")
        if roofline:
            line["roofline"] = roofline
        if cpu_baseline:
            line["cpu_baseline"] = cpu_baseline
        if e2e:
            line["e2e"] = e2e
        if parity:
            line.update({"parity_ok": parity.get("parity_ok"), "parity": parity})
        line.update(extra)
    else:
        line = None

    # ------------------------------------------------------------------ epoch leg (second half of the metric)
    # Full-graph training epochs/s on the products-shaped graph at the N of this run, through the same package
    # (maxk_b200.train).  Extra key `epoch`, after the timed region.  A watchdog prints the line without it if the
    # leg does not finish (it contains collectives): the bench line must not be lost to its extra leg.
    if args.epochs > 0:
        def give_up():
            if rank == 0:
                line["epoch"] = {"error": "the epoch leg did not finish within 300 s"}
                print(json.dumps(line), flush=True)
            os._exit(0)
        dog = threading.Timer(300.0, give_up)
        dog.daemon = True
        dog.start()
        try:
            if world > 1:
                getattr(agg, "close", lambda: None)()   # collective: the Reddit job's peer buffers are released on every rank
                del agg
            else:
                del warp4, x, dy, sp_data, sp_index, y, dxc
            g.clear()
            torch.cuda.empty_cache()
            from maxk_b200 import train
            # transport of the epoch: the one asked for with --comm, else by block size (train.Trainer: "auto")
            rec = train.run_epochs("products", epochs=args.epochs, warmup=2, device=dev, rank=rank, world=world,
                                   comm=(args.comm or "auto") if world > 1 else "peer")
            if rank == 0:
                line["epoch"] = rec
                log(f"[epoch] products-shaped SAGE k=32: {rec['ms_per_epoch']:.2f} ms/epoch = {rec['value']:.3f} epochs/s at N={world}")
        except Exception as exc:  # noqa: BLE001
            log(f"[rank {rank}] epoch leg failed: {type(exc).__name__}: {exc}")
            if rank == 0:
                line["epoch"] = {"error": f"{type(exc).__name__}: {exc}"[:300]}
        # -------------------------------------------------------------- fused Linear + TopK leg
        # The Linear -> MaxK pair of the GCN / GIN layers on the products-sized activation: the own tcgen05 kernel
        # (TopK in the GEMM's epilogue, no dense [N, 256] tensor in HBM) against torch's fp32 Linear + the TopK kernel.
        if world == 1:
            try:
                torch.cuda.empty_cache()
                gen = torch.Generator(device=dev).manual_seed(7)
                n_lt = 2_449_029
                xl = torch.randn(n_lt, 256, device=dev, generator=gen)
                wl = torch.randn(256, 256, device=dev, generator=gen) / 16
                bl = torch.randn(256, device=dev, generator=gen) * 0.1
                launches0 = _cabi.launch_count()
                t_f = timeit_events(lambda: ops.linear_topk_cbsr(xl, wl, bl, 32), iters=5, warm=3)
                fused_launches = (_cabi.launch_count() - launches0) // 8
                t_l = timeit_events(lambda: torch.nn.functional.linear(xl, wl, bl), iters=5, warm=3)
                yl = torch.nn.functional.linear(xl, wl, bl)
                t_t = timeit_events(lambda: ops.topk_cbsr(yl, 32), iters=5, warm=3)
                line["linear_topk"] = {
                    "what": "TopK_32(X W^T + b), X [2449029, 256] fp32 (products-sized), W [256, 256]: own tcgen05 kernel "
                            "(3 x TF32 operand split, TMEM accumulators, TopK in the epilogue) vs torch fp32 Linear + TopK kernel",
                    "fused_ms": t_f, "torch_linear_fp32_ms": t_l, "topk_kernel_ms": t_t,
                    "speedup": (t_l + t_t) / t_f, "fp32_equiv_TFLOPs": 2.0 * n_lt * 256 * 256 / t_f / 1e9,
                    "hbm_bytes_avoided": 8 * n_lt * 256, "launches_per_call": fused_launches}
                log(f"[linear_topk] fused {t_f:.3f} ms vs Linear {t_l:.3f} + TopK {t_t:.3f} ms")
                del xl, yl
            except Exception as exc:  # noqa: BLE001
                log(f"linear_topk leg failed: {type(exc).__name__}: {exc}")
                line["linear_topk"] = {"error": f"{type(exc).__name__}: {exc}"[:300]}
        dog.cancel()
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
