mkdir -p gpurun_out
( time timeout 900 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err ) 2> gpurun_out/bench_ref.time
( time timeout 900 python bench.py > gpurun_out/bench_own.json 2> gpurun_out/bench_own.err ) 2> gpurun_out/bench_own.time
B="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
$B > gpurun_out/plain_bench.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:spgemm_fwd|sspmm_bwd|topk_cbsr|wl_' -c 400 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/ncu_bench.log 2>&1
cat gpurun_out/bench_ref.json; cat gpurun_out/bench_ref.time gpurun_out/bench_own.time; grep -v Warn gpurun_out/bench_own.err | tail -8; head -30 gpurun_out/launches.csv
