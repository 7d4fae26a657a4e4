#!/bin/bash
# Round 2, call 9 (1 GPU): fused Linear + TopK -- Linear-only timing, ncu capture, GCN epoch with / without it; TopK <128,2>, <64,1>.
O=gpurun_out; mkdir -p $O
step() { local t=$1 name=$2; shift 2; timeout "$t" "$@" > "$O/$name" 2> "$O/$name.err"; echo "[$name] rc=$? $(tail -c 400 "$O/$name" | tail -2 | cut -c1-300)"; }
step 200 r9_lt_probe.jsonl python scripts/linear_topk_probe.py
TOPK_PROBE_SHORT=1 step 300 r9_topk_probe.log python scripts/topk_probe.py
MAXK_LINEAR_TOPK_FUSED=0 step 400 r9_epoch_gcn_unfused.json python scripts/train_epoch.py --shape products --model gcn --k 32 --epochs 5 --warmup 2
MAXK_LINEAR_TOPK_FUSED=1 step 400 r9_epoch_gcn_fused.json python scripts/train_epoch.py --shape products --model gcn --k 32 --epochs 5 --warmup 2
MAXK_LINEAR_TOPK_FUSED=0 step 400 r9_epoch_gin_unfused.json python scripts/train_epoch.py --shape products --model gin --k 32 --epochs 5 --warmup 2
MAXK_LINEAR_TOPK_FUSED=1 step 400 r9_epoch_gin_fused.json python scripts/train_epoch.py --shape products --model gin --k 32 --epochs 5 --warmup 2
step 500 r9_ncu_lt.log ncu --set full --clock-control none --import-source on -k regex:linear_topk_kernel -s 1 -c 1 -f -o $O/r9_prof_lt python scripts/linear_topk_probe.py quick
