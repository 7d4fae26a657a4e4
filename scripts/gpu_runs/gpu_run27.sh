mkdir -p gpurun_out
run() { N=$1; shift; timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT "$@"; }
PORT=29701 run 2 bench.py --gpus 2 --steps 40 --warmup 5 2> gpurun_out/bench_peer_2.err > gpurun_out/scale_peer_n2.json; cut -c1-230 gpurun_out/scale_peer_n2.json; grep -i "unavailable" gpurun_out/bench_peer_2.err | head -2
PORT=29702 run 2 scripts/rmat_sweep.py --cases 22:5 --partition nnz --iters 3 2>/dev/null | cut -c1-700 | tee gpurun_out/rmat_n2_nnz.jsonl
PORT=29703 run 2 scripts/rmat_sweep.py --cases 22:5 --partition cost --iters 3 2>/dev/null | cut -c1-700 | tee gpurun_out/rmat_n2_cost.jsonl
PORT=29704 run 2 scripts/train_epoch.py --shape products --model sage --epochs 4 --warmup 2 --comm peer --row-cost 0 2>/dev/null | grep "^{" | cut -c1-200 | tee gpurun_out/epoch_n2_nnz.json
PORT=29705 run 2 scripts/train_epoch.py --shape products --model sage --epochs 4 --warmup 2 --comm peer 2>/dev/null | grep "^{" | cut -c1-200 | tee gpurun_out/epoch_n2_rowcost.json
