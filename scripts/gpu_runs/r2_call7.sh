#!/bin/bash
# Round 2, call 7 (1 GPU): grouped TopK (4 / 8 lanes per row) against the one-row-per-warp kernel: parity, timings, ncu.
O=gpurun_out; mkdir -p $O
step() { local t=$1 name=$2; shift 2; timeout "$t" "$@" > "$O/$name" 2> "$O/$name.err"; echo "[$name] rc=$? $(tail -c 300 "$O/$name" | tail -1 | cut -c1-300)"; }
step 900 r7_pytest.log python -m pytest tests -m gpu -q -x
step 300 r7_topk_probe.log python scripts/topk_probe.py
step 400 r7_ncu_topk.log ncu --set full --clock-control none --import-source on -k regex:topk_cbsr_f32 -s 3 -c 3 -f -o $O/r7_prof_topk python scripts/topk_probe.py quick
