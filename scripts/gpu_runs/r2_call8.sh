#!/bin/bash
# Round 2, call 8 (1 GPU): first run of the fused tcgen05 Linear + TopK kernel (parity, timing); grouped TopK variants.
O=gpurun_out; mkdir -p $O
step() { local t=$1 name=$2; shift 2; timeout "$t" "$@" > "$O/$name" 2> "$O/$name.err"; echo "[$name] rc=$? $(tail -c 400 "$O/$name" | tail -2 | cut -c1-300)"; }
step 300 r8_lt_pytest.log python -m pytest tests/test_gpu_linear_topk.py -x -q
step 200 r8_lt_probe.jsonl python scripts/linear_topk_probe.py
step 300 r8_topk_probe.log python scripts/topk_probe.py
step 600 r8_pytest.log python -m pytest tests -m gpu -q -x
