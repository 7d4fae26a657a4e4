#!/bin/bash
# Round 2, call 16 (1 GPU): dense tensor-core route instead of the k-sparse products at width 256: epochs, profile, suite.
O=gpurun_out; mkdir -p $O
step() { local t=$1 name=$2; shift 2; timeout "$t" "$@" > "$O/$name" 2> "$O/$name.err"; echo "[$name] rc=$? $(tail -c 400 "$O/$name" | tail -2 | cut -c1-300)"; }
for m in sage gcn gin; do
  step 400 r16_epoch_${m}.json python scripts/train_epoch.py --shape products --model $m --k 32 --epochs 5 --warmup 2
done
step 300 r16_profile_sage.log python scripts/epoch_profile.py --model sage
step 300 r16_profile_gcn.log python scripts/epoch_profile.py --model gcn
step 600 r16_pytest.log python -m pytest tests -m gpu -q
