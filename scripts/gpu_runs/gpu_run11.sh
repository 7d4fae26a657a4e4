mkdir -p gpurun_out
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 600 $TR scripts/train_epoch.py --shape products --model sage --epochs 5 --warmup 2 2>&1 | grep -v -E "Warn|OMP|\*\*\*" | tail -3
timeout 600 $TR scripts/train_epoch.py --shape products --model sage --epochs 5 --warmup 2 --no-split 2>&1 | grep -v -E "Warn|OMP|\*\*\*" | tail -1
timeout 600 $TR scripts/train_epoch.py --shape products --model gcn --epochs 5 --warmup 2 2>&1 | grep -v -E "Warn|OMP|\*\*\*" | tail -1
