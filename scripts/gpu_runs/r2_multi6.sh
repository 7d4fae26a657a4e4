#!/bin/bash
# Round 2, final multi-GPU calls:  gpurun --gpus N -- 'bash scripts/gpu_runs/r2_multi6.sh N'  (the bench line at N, as the driver launches it)
N=${1:-2}
O=gpurun_out; mkdir -p $O
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29731 \
    bench.py --gpus $N --steps 40 --warmup 5 > $O/r6_scale_n$N.json 2> $O/r6_bench_n$N.err
echo "bench rc=$? $(cut -c1-150 $O/r6_scale_n$N.json)"; tail -3 $O/r6_bench_n$N.err | cut -c1-300
