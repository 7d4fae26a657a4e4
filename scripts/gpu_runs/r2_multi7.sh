#!/bin/bash
# Round 2: per-phase serial costs of the aggregation step on the PRODUCTS shape at N GPUs (NCCL transport pieces).
N=${1:-8}
O=gpurun_out; mkdir -p $O
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29741 \
    scripts/dist_phases.py products 32 > $O/r7_phases_products_n$N.json 2> $O/r7_phases_products_n$N.err
echo "rc=$?"; cat $O/r7_phases_products_n$N.json | cut -c1-1500; tail -3 $O/r7_phases_products_n$N.err
