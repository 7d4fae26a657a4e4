#!/bin/bash
# Round 2, call 36 (1 GPU): final records -- full GPU suite, smoke, bench line, ncu launch list of the bench command,
# ncu capture of the fused Linear + TopK kernel with the seeded search.
O=gpurun_out; mkdir -p $O
step() { local t=$1 name=$2; shift 2; timeout "$t" "$@" > "$O/$name" 2> "$O/$name.err"; echo "[$name] rc=$? $(tail -c 300 "$O/$name" | tail -1 | cut -c1-200)"; }
step 600 r36_pytest_gpu.log python -m pytest tests -x -q -m gpu
python __graft_entry__.py smoke > $O/r36_smoke.log 2>&1; echo "smoke rc=$? $(tail -2 $O/r36_smoke.log | cut -c1-200)"
step 600 r36_bench_n1.json python bench.py
step 300 r36_ncu_list.log ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r36_launches_bench.csv python bench.py --steps 2 --warmup 3 --epochs 0 --no-cpu-baseline --no-e2e
step 300 r36_ncu_lt.log ncu --set full --clock-control none --import-source on -k regex:linear_topk_kernel -s 1 -c 1 -f -o $O/r36_prof_lt python scripts/linear_topk_probe.py quick
