#!/bin/bash
# Round 2, call 21 (1 GPU): L2 prefetch of the next rows + separate X / W arrival barriers in the tcgen05 kernels.
O=gpurun_out; mkdir -p $O
step() { local t=$1 name=$2; shift 2; timeout "$t" "$@" > "$O/$name" 2> "$O/$name.err"; echo "[$name] rc=$? $(tail -c 400 "$O/$name" | tail -2 | cut -c1-300)"; }
step 300 r21_lt_pytest.log python -m pytest tests/test_gpu_linear_topk.py -q
step 300 r21_lt_probe.jsonl python scripts/linear_topk_probe.py
step 400 r21_epoch_sage.json python scripts/train_epoch.py --shape products --model sage --k 32 --epochs 5 --warmup 2
