#!/bin/bash
O=gpurun_out; mkdir -p $O
for sc in 0.02 0.125; do
timeout 300 python scripts/train_epoch.py --shape products --model sage --k 32 --epochs 10 --warmup 3 --scale $sc > $O/r26_epoch_sage_scale$sc.json 2> $O/r26.err
python -c "import json; d=json.loads(open('$O/r26_epoch_sage_scale$sc.json').read().strip().splitlines()[-1]); print('scale $sc nodes', d['nodes'], 'ms/epoch', d['ms_per_epoch'], 'wall', d['ms_per_epoch_wall'])"
done
