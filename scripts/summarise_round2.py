#!/usr/bin/env python3
"""Read what scripts/round2_first_run.sh left in gpurun_out/ and print, per (shape, k), the fastest
variant of every kernel next to the default -- the table the round-2 prune/promote decisions are taken from.

    python scripts/summarise_round2.py [gpurun_out]
"""
import glob
import json
import os
import sys

out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out"


def rows(path):
    res = []
    if not os.path.exists(path):
        return res
    for ln in open(path):
        ln = ln.strip()
        if ln.startswith("{"):
            try:
                res.append(json.loads(ln))
            except ValueError:
                pass
    return res


def label(r):
    keys = ("depth", "copies", "tpw", "homes", "reductions", "row_prefetch")
    return ",".join(f"{k}={r[k]}" for k in keys if k in r and r[k] not in (0, None, False, -1))


print("== parity ==")
for f in ("pytest_default.log", "pytest_experimental.log"):
    p = os.path.join(out, f)
    print(f"{f}: {open(p).read().strip().splitlines()[-1] if os.path.exists(p) and os.path.getsize(p) else 'missing'}")

print("\n== kernels: default vs best variant (ms) ==")
for path in sorted(glob.glob(os.path.join(out, "sweep_db_*.jsonl"))):
    by = {}
    for r in rows(path):
        fam = "fwd" if r["kernel"].startswith("fwd") else "bwd"
        by.setdefault((r["shape"], r["k"], fam), []).append(r)
    for (shape, k, fam), rs in sorted(by.items()):
        def total(r):
            return r["ms"] + (r.get("plan_ms", 0.0) if fam == "fwd" else 0.0)
        default = [r for r in rs if r["kernel"] == fam and not label(r)]
        best = min(rs, key=total)
        d_ms = min(r["ms"] for r in default) if default else float("nan")
        print(f"{shape:9s} k={k:3d} {fam}: default {d_ms:7.3f}   best {total(best):7.3f} "
              f"({best['kernel']} {label(best) or 'default'}; kernel alone {best['ms']:.3f}) "
              f"{'<-- promote?' if total(best) < 0.95 * d_ms else ''}")

print("\n== TopK (us) ==")
p = os.path.join(out, "topk_probe_round2.log")
if os.path.exists(p):
    sys.stdout.write(open(p).read())

print("\n== CBSR x dense products vs torch fp32 (ms) ==")
for r in rows(os.path.join(out, "cbsr_linear_probe.jsonl")):
    print(f"k={r['k']:3d}: fc_self {r['cbsr_matmul_ms']:.3f} vs scatter+mm {r['scatter_plus_dense_mm_ms']:.3f} | "
          f"d values {r['cbsr_sampled_dot_ms']:.3f} vs mm+gather {r['dense_mm_bwd_plus_gather_ms']:.3f} | "
          f"model {r['smem_bytes_model_ms']:.3f} | max diff {r['max_abs_diff_fwd']:.2e} / {r['max_abs_diff_bwd']:.2e}")

print("\n== epochs (ms) ==")
for path in sorted(glob.glob(os.path.join(out, "epoch_*.json"))):
    for r in rows(path):
        print(f"{os.path.basename(path):40s} {r.get('ms_per_epoch', float('nan')):8.2f}")

print("\n== bench step with the planned kernels ==")
for path in sorted(glob.glob(os.path.join(out, "bench_planned_*.json"))):
    for r in rows(path):
        print(f"{os.path.basename(path)}: {r['ms_per_step']:.3f} ms/step (default 6.571), plan {r.get('plan_ms', float('nan')):.3f} ms, "
              f"fwd {r['kernels']['spgemm_fwd']['ms']:.3f} bwd {r['kernels']['sspmm_bwd']['ms']:.3f}")
