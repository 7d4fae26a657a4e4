mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
tail -3 gpurun_out/pytest.log
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -1 | tee gpurun_out/smoke.log
timeout 400 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -4 gpurun_out/bench_final.err | cut -c1-200
B="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
$B > gpurun_out/plain_bench.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"maxk|spgemm|sspmm|topk|worklist" -c 400 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/ncu_bench.log 2>&1
P="python scripts/profile_step.py --k 32 --iters 2"
$P > gpurun_out/plain_profile.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:spgemm_fwd_kernel -s 1 -c 1 -f -o gpurun_out/prof_fwd_k32_v3 $P > gpurun_out/ncu_fwd.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:topk_cbsr_f32 -s 1 -c 1 -f -o gpurun_out/prof_topk_k32_v5 $P > gpurun_out/ncu_topk.log 2>&1
tail -2 gpurun_out/ncu_fwd.log gpurun_out/ncu_topk.log; wc -l gpurun_out/launches.csv
