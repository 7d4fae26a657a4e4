#!/bin/bash
# Round 2, call 32 (1 GPU): records of the final TopK -- shape / k sweep, ncu capture, four-distribution timings,
# bench line, smoke.
O=gpurun_out; mkdir -p $O
step() { local t=$1 name=$2; shift 2; timeout "$t" "$@" > "$O/$name" 2> "$O/$name.err"; echo "[$name] rc=$? $(tail -c 300 "$O/$name" | tail -1 | cut -c1-200)"; }
step 200 r32_topk_probe.log python scripts/topk_probe.py
step 600 r32_sweep.json python scripts/sweep_shapes.py
step 600 r32_bench_n1.json python bench.py
step 400 r32_ncu_topk.log ncu --set full --clock-control none --import-source on -k regex:topk_cbsr_f32 -s 3 -c 3 -f -o $O/r32_prof_topk python scripts/topk_probe.py quick
python __graft_entry__.py smoke > $O/r32_smoke.log 2>&1; echo "smoke rc=$? $(tail -2 $O/r32_smoke.log | cut -c1-200)"
