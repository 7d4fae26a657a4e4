#!/bin/bash
# Round 2, call 27 (1 GPU): the classifier's backward on the tensor-core kernels: suite, epochs.
O=gpurun_out; mkdir -p $O
step() { local t=$1 name=$2; shift 2; timeout "$t" "$@" > "$O/$name" 2> "$O/$name.err"; echo "[$name] rc=$? $(tail -c 400 "$O/$name" | tail -2 | cut -c1-300)"; }
step 600 r27_pytest.log python -m pytest tests -m gpu -q
for m in sage gcn gin; do
  step 400 r27_epoch_${m}.json python scripts/train_epoch.py --shape products --model $m --k 32 --epochs 5 --warmup 2
done
