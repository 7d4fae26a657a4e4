mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
tail -15 gpurun_out/pytest.log
B="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
$B > gpurun_out/plain_bench.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -k regex:maxk -c 400 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/ncu_bench.log 2>&1
P="python scripts/profile_step.py --k 32 --iters 2"
$P > gpurun_out/plain_profile.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:spgemm_fwd -s 1 -c 1 -f -o gpurun_out/prof_fwd_k32 $P > gpurun_out/ncu_fwd.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:sspmm_bwd -s 1 -c 1 -f -o gpurun_out/prof_bwd_k32 $P > gpurun_out/ncu_bwd.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:topk_cbsr -s 1 -c 1 -f -o gpurun_out/prof_topk_k32 $P > gpurun_out/ncu_topk.log 2>&1
tail -3 gpurun_out/plain_profile.log gpurun_out/ncu_fwd.log gpurun_out/ncu_topk.log
ls -la gpurun_out
