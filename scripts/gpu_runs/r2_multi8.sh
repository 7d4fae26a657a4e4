#!/bin/bash
# Round 2: products SAGE epoch at N with the NCCL transport against the peer-memory transport.
N=${1:-4}
O=gpurun_out; mkdir -p $O
run() { local t=$1 port=$2; shift 2; timeout "$t" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N \
        --master-addr 127.0.0.1 --master-port "$port" "$@"; }
for c in nccl peer; do
run 300 2975$N scripts/train_epoch.py --shape products --model sage --k 32 --epochs 5 --warmup 2 --comm $c > $O/r8_epoch_sage_n${N}_$c.json 2> $O/r8_epoch_sage_n${N}_$c.err
echo "$c rc=$? $(python -c "import json,sys; d=json.loads(open('$O/r8_epoch_sage_n${N}_$c.json').read().strip().splitlines()[-1]); print(d['ms_per_epoch'], d['loss'])")"
done
