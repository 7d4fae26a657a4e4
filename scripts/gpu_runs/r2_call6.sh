#!/bin/bash
# Round 2, call 6 (1 GPU): suite, bench line, decode probe, R-MAT one-GPU rows, ncu of the final forward / TopK.
O=gpurun_out; mkdir -p $O
step() { local t=$1 name=$2; shift 2; timeout "$t" "$@" > "$O/$name" 2> "$O/$name.err"; echo "[$name] rc=$? $(tail -c 300 "$O/$name" | tail -1 | cut -c1-300)"; }
step 900 r6_pytest.log python -m pytest tests -m gpu -q -x
step 600 r6_bench_n1.json python bench.py
step 200 r6_bwd_decode.jsonl python scripts/bwd_decode_probe.py
step 200 r6_topk_probe.log python scripts/topk_probe.py
step 600 r6_rmat_n1.jsonl python scripts/rmat_sweep.py
step 400 r6_ncu_topk.log ncu --set full --clock-control none --import-source on -k regex:topk_cbsr_f32 -s 1 -c 1 -f -o $O/r6_prof_topk python scripts/profile_step.py --iters 2
