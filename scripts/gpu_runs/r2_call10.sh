#!/bin/bash
# Round 2, call 10 (1 GPU): fused Linear + TopK v2 (exact finish, 128-bit copy-out, 4-stage K=16 ring) -- parity, timing,
# ncu; the same with the 2-stage K=32 ring (library rebuilt on the box).
O=gpurun_out; mkdir -p $O
step() { local t=$1 name=$2; shift 2; timeout "$t" "$@" > "$O/$name" 2> "$O/$name.err"; echo "[$name] rc=$? $(tail -c 400 "$O/$name" | tail -2 | cut -c1-300)"; }
step 300 r10_lt_pytest.log python -m pytest tests/test_gpu_linear_topk.py -x -q
step 200 r10_lt_probe_k16.jsonl python scripts/linear_topk_probe.py
step 500 r10_ncu_lt.log ncu --set full --clock-control none --import-source on -k regex:linear_topk_kernel -s 1 -c 1 -f -o $O/r10_prof_lt python scripts/linear_topk_probe.py quick
step 600 r10_rebuild.log python -c "import sys; sys.path.insert(0,'spgemm-new_b200'); import build; print(build.build(force=True, extra=['-DMAXK_LT_BK=32']))"
step 300 r10_lt_pytest_k32.log python -m pytest tests/test_gpu_linear_topk.py -x -q
step 200 r10_lt_probe_k32.jsonl python scripts/linear_topk_probe.py
