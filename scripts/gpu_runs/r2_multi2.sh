#!/bin/bash
# Round 2, second multi-GPU call:  gpurun --gpus N -- 'bash scripts/gpu_runs/r2_multi2.sh N'
# The bench line at N exactly as the driver launches it (parity_ok, per-GPU roofline, pipelined e2e, epoch leg) with the
# tensor-core Linears / LayerNorm in the epoch, then GCN epochs at N.
N=${1:-2}
O=gpurun_out; mkdir -p $O
run() { local t=$1 port=$2; shift 2; timeout "$t" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N \
        --master-addr 127.0.0.1 --master-port "$port" "$@"; }
run 600 29702 bench.py --gpus $N --steps 40 --warmup 5 > $O/r5_scale_n$N.json 2> $O/r5_bench_n$N.err
echo "bench rc=$? $(cut -c1-150 $O/r5_scale_n$N.json)"
tail -4 $O/r5_bench_n$N.err | cut -c1-300
run 400 29705 scripts/train_epoch.py --shape products --model gcn --k 32 --epochs 5 --warmup 2 > $O/r5_epoch_gcn_n$N.json 2> $O/r5_epoch_gcn_n$N.err
echo "gcn epoch rc=$? $(cut -c1-200 $O/r5_epoch_gcn_n$N.json)"
