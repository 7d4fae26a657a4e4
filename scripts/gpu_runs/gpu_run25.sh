mkdir -p gpurun_out
run() { N=$1; shift; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT "$@"; }
PORT=29681 run 8 scripts/peer_check.py --quick > gpurun_out/peer_check_n8.log 2> gpurun_out/peer_check_n8.err; echo "peer_check rc=$? ok=$(grep -c 'all-ranks True' gpurun_out/peer_check_n8.log)"; grep "False\|halo_threshold" gpurun_out/peer_check_n8.log | head
PORT=29682 run 8 scripts/rmat_sweep.py --cases 20:100,22:20,24:5,24:20,25:5 --comm peer 2> gpurun_out/rmat_n8.err | tee gpurun_out/rmat_n8_halo.jsonl | cut -c1-330
PORT=29683 run 8 bench.py --gpus 8 --steps 40 --warmup 5 --comm peer 2> gpurun_out/bench_peer_8.err > gpurun_out/scale_peer_n8.json; cut -c1-260 gpurun_out/scale_peer_n8.json
