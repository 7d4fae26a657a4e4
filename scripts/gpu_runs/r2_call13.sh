#!/bin/bash
# Round 2, call 13 (1 GPU): Linear alone (coalesced epilogue) + tcgen05 weight gradient: parity, timing; epochs; epoch kernel profile.
O=gpurun_out; mkdir -p $O
step() { local t=$1 name=$2; shift 2; timeout "$t" "$@" > "$O/$name" 2> "$O/$name.err"; echo "[$name] rc=$? $(tail -c 400 "$O/$name" | tail -2 | cut -c1-300)"; }
step 300 r13_lt_pytest.log python -m pytest tests/test_gpu_linear_topk.py -x -q
step 200 r13_lt_probe.jsonl python scripts/linear_topk_probe.py
MAXK_TC_LINEAR=0 step 300 r13_profile_sage_lib.log python scripts/epoch_profile.py --model sage
step 300 r13_profile_sage_tc.log python scripts/epoch_profile.py --model sage
for m in sage gcn gin; do
  MAXK_TC_LINEAR=0 step 400 r13_epoch_${m}_lib.json python scripts/train_epoch.py --shape products --model $m --k 32 --epochs 5 --warmup 2
  step 400 r13_epoch_${m}_tc.json python scripts/train_epoch.py --shape products --model $m --k 32 --epochs 5 --warmup 2
done
step 600 r13_pytest.log python -m pytest tests -m gpu -x -q
