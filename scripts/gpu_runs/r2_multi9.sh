#!/bin/bash
# Round 2: the bench line at N GPUs exactly as the driver launches it (final code of the round).
N=${1:-2}
O=gpurun_out; mkdir -p $O
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2977$N \
    bench.py --gpus $N --steps 20 --warmup 5 > $O/r9_bench_n$N.json 2> $O/r9_bench_n$N.err
echo "rc=$?"; python - <<PY
import json
d=json.loads(open('$O/r9_bench_n$N.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','n_gpus','parity_ok')}, d.get('parity'), 'e2e', d['e2e'].get('ms_per_step'), 'epoch', d.get('epoch',{}).get('ms_per_epoch'), d.get('epoch',{}).get('comm'))
PY
tail -4 $O/r9_bench_n$N.err
