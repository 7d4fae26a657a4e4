#!/bin/bash
# Round 2, call 30 (1 GPU): full GPU suite on the TopK changes (order-statistics seed for k <= 64, extraction finish,
# level snapping in the robust search), TopK timings on four distributions, the bench line.
O=gpurun_out; mkdir -p $O
step() { local t=$1 name=$2; shift 2; timeout "$t" "$@" > "$O/$name" 2> "$O/$name.err"; echo "[$name] rc=$? $(tail -c 300 "$O/$name" | tail -1 | cut -c1-200)"; }
step 600 r30_pytest_gpu.log python -m pytest tests -x -q -m gpu
step 200 r30_topk_probe.log python scripts/topk_probe.py
step 600 r30_bench_n1.json python bench.py
