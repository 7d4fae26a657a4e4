#!/bin/bash
# Round 2, call 29 (1 GPU): TopK with the order-statistics seed and the extraction finish -- parity, then timings.
O=gpurun_out; mkdir -p $O
step() { local t=$1 name=$2; shift 2; timeout "$t" "$@" > "$O/$name" 2> "$O/$name.err"; echo "[$name] rc=$? $(tail -c 300 "$O/$name" | tail -1 | cut -c1-200)"; }
step 300 r29_pytest_topk.log python -m pytest tests/test_gpu_parity.py tests/test_gpu_layernorm.py -x -q -k "topk or layernorm"
step 200 r29_topk_probe.log python scripts/topk_probe.py
