#!/bin/bash
# Round 2, call 14 (1 GPU): weight-gradient kernel with the MN-major TF32 layout (128-byte swizzle, 32-byte base).
O=gpurun_out; mkdir -p $O
step() { local t=$1 name=$2; shift 2; timeout "$t" "$@" > "$O/$name" 2> "$O/$name.err"; echo "[$name] rc=$? $(tail -c 400 "$O/$name" | tail -2 | cut -c1-300)"; }
step 300 r14_lt_pytest.log python -m pytest tests/test_gpu_linear_topk.py -q
step 200 r14_lt_probe.jsonl python scripts/linear_topk_probe.py
for m in sage gcn; do
  step 400 r14_epoch_${m}_tc.json python scripts/train_epoch.py --shape products --model $m --k 32 --epochs 5 --warmup 2
done
step 300 r14_profile_sage_tc.log python scripts/epoch_profile.py --model sage
step 600 r14_pytest.log python -m pytest tests -m gpu -x -q
