#!/bin/bash
# Round 2, fourth multi-GPU call: copy CTAs of the fused forward's push on the products epoch (16 = default, 48, 96).
N=${1:-8}
O=gpurun_out; mkdir -p $O
run() { local t=$1 port=$2; shift 2; timeout "$t" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N \
        --master-addr 127.0.0.1 --master-port "$port" "$@"; }
p=29710
for cc in 48 96 16; do
  p=$((p+1))
  MAXK_PEER_COPY_CTAS=$cc run 300 $p scripts/train_epoch.py --shape products --model sage --k 32 --epochs 5 --warmup 2 > $O/r5_epoch_sage_n${N}_cc$cc.json 2> $O/r5_epoch_sage_n${N}_cc$cc.err
  echo "copy_ctas=$cc rc=$? $(python -c "import json,sys; d=json.loads(open('$O/r5_epoch_sage_n${N}_cc$cc.json').read().strip().splitlines()[-1]); print(d['ms_per_epoch'], d['loss'])")"
done
