mkdir -p gpurun_out
run() { N=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 295$N "$@" 2> gpurun_out/err_$N.log | grep -E "^\{|world=" ; }
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29588 scripts/dist_check.py 2>/dev/null | grep world=
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null > gpurun_out/scale_n1.json; cat gpurun_out/scale_n1.json | cut -c1-200
for N in 2 4 8; do timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2960$N bench.py --gpus $N --steps 20 --warmup 3 2> gpurun_out/bench_err_$N.log > gpurun_out/scale_n$N.json; cut -c1-200 gpurun_out/scale_n$N.json; done
for N in 4 8; do timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2970$N scripts/train_epoch.py --shape products --model sage --epochs 5 --warmup 2 2> gpurun_out/train_err_$N.log | grep "^{" | tee gpurun_out/epoch_n$N.json; done
