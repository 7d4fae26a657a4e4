#!/bin/bash
# Round 2, third multi-GPU call:  gpurun --gpus N -- 'bash scripts/gpu_runs/r2_multi3.sh N'   (kernel profile of an epoch at N)
N=${1:-8}
O=gpurun_out; mkdir -p $O
run() { local t=$1 port=$2; shift 2; timeout "$t" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N \
        --master-addr 127.0.0.1 --master-port "$port" "$@"; }
run 400 29706 scripts/epoch_profile.py --model sage --rows 40 > $O/r5_profile_sage_n$N.log 2> $O/r5_profile_sage_n$N.err
echo "profile rc=$?"; head -30 $O/r5_profile_sage_n$N.log | cut -c1-150
run 400 29707 scripts/train_epoch.py --shape products --model sage --k 32 --epochs 5 --warmup 2 > $O/r5_epoch_sage_n$N.json 2> $O/r5_epoch_sage_n$N.err
echo "sage epoch rc=$? $(cut -c1-200 $O/r5_epoch_sage_n$N.json)"
