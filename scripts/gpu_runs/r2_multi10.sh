#!/bin/bash
# Round 2: the NCCL transport's two collectives at the products sizes on N GPUs under a few NCCL settings.
N=${1:-8}
O=gpurun_out; mkdir -p $O
run() { timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2978$N scripts/nccl_exchange_probe.py >> $O/r10_nccl_probe_n$N.jsonl 2>> $O/r10_nccl_probe_n$N.err; echo "rc=$? $(tail -1 $O/r10_nccl_probe_n$N.jsonl | cut -c1-400)"; }
run
NCCL_MIN_NCHANNELS=32 run
NCCL_MIN_NCHANNELS=32 NCCL_BUFFSIZE=16777216 run
NCCL_ALGO=Ring run
