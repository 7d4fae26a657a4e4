mkdir -p gpurun_out
run() { N=$1; shift; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT "$@"; }
PORT=29651 run 8 bench.py --gpus 8 --steps 40 --warmup 5 --comm peer 2> gpurun_out/bench_peer_8.err > gpurun_out/scale_peer_n8.json; cut -c1-400 gpurun_out/scale_peer_n8.json; grep -i "unavailable\|error" gpurun_out/bench_peer_8.err | head -3
PORT=29652 run 8 bench.py --gpus 8 --steps 40 --warmup 5 --comm nccl 2> gpurun_out/bench_nccl_8.err > gpurun_out/scale_nccl_n8.json; cut -c1-200 gpurun_out/scale_nccl_n8.json
