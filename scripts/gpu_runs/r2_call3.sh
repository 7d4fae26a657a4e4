#!/bin/bash
# Round 2, call 3 (1 GPU): the rewritten ring path, the TopK variants, the suite again.
O=gpurun_out; mkdir -p $O
step() { local t=$1 name=$2; shift 2; timeout "$t" "$@" > "$O/$name" 2> "$O/$name.err"; echo "[$name] rc=$? $(tail -c 300 "$O/$name" | tail -1 | cut -c1-300)"; }
step 900 r3_pytest.log python -m pytest tests -m gpu -q -x
step 200 r3_topk_probe.log python scripts/topk_probe.py
step 400 r3_sweep_products.json python scripts/sweep_shapes.py --shapes products,flickr
step 300 r3_sweep_fwd_products.jsonl python scripts/sweep_fwd.py --shape products --ks 32,16,64 --copies 0 --tpw=-1 --depth 0,2,8,12
step 300 r3_sweep_fwd_tpw.jsonl python scripts/sweep_fwd.py --shape products --ks 32 --copies 0 --tpw=1,2,4,8,16 --depth 8
step 400 r3_ncu_ring.log ncu --set full --clock-control none --import-source on -k regex:spgemm_fwd_kernel -s 1 -c 1 -f -o $O/r3_prof_fwd_products python scripts/profile_step.py --shape products --iters 2
step 400 r3_ncu_topk.log ncu --set full --clock-control none --import-source on -k regex:topk_cbsr_f32 -s 1 -c 1 -f -o $O/r3_prof_topk_pair python scripts/profile_step.py --iters 2
