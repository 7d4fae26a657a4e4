#!/bin/bash
# Round 2, call 28 (1 GPU): final records -- bench line, reference arm, ncu launch list of the bench command.
O=gpurun_out; mkdir -p $O
step() { local t=$1 name=$2; shift 2; timeout "$t" "$@" > "$O/$name" 2> "$O/$name.err"; echo "[$name] rc=$? $(tail -c 300 "$O/$name" | tail -1 | cut -c1-200)"; }
step 600 r28_bench_n1.json python bench.py
step 600 r28_bench_ref.json python bench.py --impl reference
step 600 r28_ncu_list.log ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r28_launches_bench.csv python bench.py --steps 2 --warmup 3 --epochs 0 --no-cpu-baseline --no-e2e
python __graft_entry__.py smoke > $O/r28_smoke.log 2>&1; echo "smoke rc=$? $(tail -2 $O/r28_smoke.log | cut -c1-200)"
