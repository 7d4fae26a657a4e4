mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR scripts/dist_check.py > gpurun_out/dist_check.log 2>&1; echo "dist_check rc=$?" >> gpurun_out/dist_check.log
tail -12 gpurun_out/dist_check.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench2 rc=$?" >> gpurun_out/bench_n2.err
tail -5 gpurun_out/bench_n1.err gpurun_out/bench_n2.err
cat gpurun_out/bench_n2.json
