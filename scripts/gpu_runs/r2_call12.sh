#!/bin/bash
# Round 2, call 12 (1 GPU): bench line with the linear_topk leg; SAGE epoch with / without the fused input layer.
O=gpurun_out; mkdir -p $O
step() { local t=$1 name=$2; shift 2; timeout "$t" "$@" > "$O/$name" 2> "$O/$name.err"; echo "[$name] rc=$? $(tail -c 400 "$O/$name" | tail -2 | cut -c1-300)"; }
step 600 r12_bench_n1.json python bench.py
MAXK_LINEAR_TOPK_FUSED=0 step 400 r12_epoch_sage_unfused.json python scripts/train_epoch.py --shape products --model sage --k 32 --epochs 5 --warmup 2
MAXK_LINEAR_TOPK_FUSED=1 step 400 r12_epoch_sage_fused.json python scripts/train_epoch.py --shape products --model sage --k 32 --epochs 5 --warmup 2
step 300 r12_pytest_models.log python -m pytest tests/test_gpu_models.py tests/test_gpu_train.py -x -q
