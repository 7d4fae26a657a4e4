mkdir -p gpurun_out
run() { N=$1; shift; timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT "$@"; }
PORT=29691 run 8 scripts/rmat_sweep.py --cases 20:100,24:5,25:5 --comm peer --iters 3 2> gpurun_out/rmat_n8.err | tee gpurun_out/rmat_n8_halo2.jsonl | cut -c1-900
