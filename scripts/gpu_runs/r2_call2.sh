#!/bin/bash
# Round 2, call 2 (1 GPU): everything after the prune -- suite, bench line (with the epoch leg), reference arm, sweep,
# TopK probe, then ncu (launch list of the bench step + full captures of the planned forward, the plan kernel and the
# ring forward on the products shape).
O=gpurun_out; mkdir -p $O
step() { local t=$1 name=$2; shift 2; timeout "$t" "$@" > "$O/$name" 2> "$O/$name.err"; echo "[$name] rc=$? $(tail -c 300 "$O/$name" | tail -1 | cut -c1-300)"; }
step 900 r2_pytest.log python -m pytest tests -m gpu -q -x
step 600 r2_bench_n1.json python bench.py
step 400 r2_bench_ref.json python bench.py --impl reference --steps 3 --warmup 1
step 600 r2_sweep.json python scripts/sweep_shapes.py
step 200 r2_topk_probe.log python scripts/topk_probe.py
step 300 r2_ncu_list.log ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $O/r2_launches_bench.csv \
     python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --epochs 0
step 400 r2_ncu_fwd.log ncu --set full --clock-control none --import-source on -k regex:spgemm_fwd_kernel -s 2 -c 1 -f -o $O/r2_prof_fwd_planned python scripts/profile_step.py --iters 3
step 400 r2_ncu_plan.log ncu --set full --clock-control none --import-source on -k regex:cbsr_bank_plan -s 1 -c 1 -f -o $O/r2_prof_plan python scripts/profile_step.py --iters 2
step 400 r2_ncu_ring.log ncu --set full --clock-control none --import-source on -k regex:spgemm_fwd_kernel -s 1 -c 1 -f -o $O/r2_prof_fwd_products python scripts/profile_step.py --shape products --iters 2
step 400 r2_ncu_topk.log ncu --set full --clock-control none --import-source on -k regex:topk_cbsr_f32 -s 1 -c 1 -f -o $O/r2_prof_topk python scripts/profile_step.py --iters 2
