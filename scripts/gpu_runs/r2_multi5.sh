#!/bin/bash
# Round 2, fifth multi-GPU call: products SAGE epoch at N with the NCCL transport and with the peer transport (LayerNorm + TopK fused).
N=${1:-8}
O=gpurun_out; mkdir -p $O
run() { local t=$1 port=$2; shift 2; timeout "$t" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N \
        --master-addr 127.0.0.1 --master-port "$port" "$@"; }
run 300 29721 scripts/train_epoch.py --shape products --model sage --k 32 --epochs 5 --warmup 2 --comm nccl > $O/r5_epoch_sage_n${N}_nccl.json 2> $O/r5_epoch_sage_n${N}_nccl.err
echo "nccl rc=$? $(python -c "import json,sys; d=json.loads(open('$O/r5_epoch_sage_n${N}_nccl.json').read().strip().splitlines()[-1]); print(d['ms_per_epoch'], d['loss'])")"
run 300 29722 scripts/train_epoch.py --shape products --model sage --k 32 --epochs 5 --warmup 2 > $O/r5_epoch_sage_n${N}_peer2.json 2> $O/r5_epoch_sage_n${N}_peer2.err
echo "peer rc=$? $(python -c "import json,sys; d=json.loads(open('$O/r5_epoch_sage_n${N}_peer2.json').read().strip().splitlines()[-1]); print(d['ms_per_epoch'], d['loss'])")"
run 400 29723 scripts/epoch_profile.py --model sage --rows 12 > $O/r5_profile_sage_n${N}_v2.log 2> $O/r5_profile_sage_n${N}_v2.err
head -14 $O/r5_profile_sage_n${N}_v2.log | cut -c1-150
