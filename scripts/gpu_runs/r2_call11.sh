#!/bin/bash
# Round 2, call 11 (1 GPU): fused Linear + TopK v3 (v1 search, 128-bit copy-out, relaxed producer wait, 4-stage K=16 ring); whole suite.
O=gpurun_out; mkdir -p $O
step() { local t=$1 name=$2; shift 2; timeout "$t" "$@" > "$O/$name" 2> "$O/$name.err"; echo "[$name] rc=$? $(tail -c 400 "$O/$name" | tail -2 | cut -c1-300)"; }
step 600 r11_pytest.log python -m pytest tests -m gpu -x -q
step 200 r11_lt_probe.jsonl python scripts/linear_topk_probe.py
step 300 r11_topk_probe.log python scripts/topk_probe.py
step 500 r11_ncu_lt.log ncu --set full --clock-control none --import-source on -k regex:linear_topk_kernel -s 1 -c 1 -f -o $O/r11_prof_lt python scripts/linear_topk_probe.py quick
