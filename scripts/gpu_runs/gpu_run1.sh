mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 -x > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 300 python tests/golden/make_golden_gpu.py gpurun_out/ref_kernels.npz > gpurun_out/golden.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 3 --sweep --ref-kernels > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
tail -5 gpurun_out/pytest.log; tail -3 gpurun_out/smoke.log; tail -15 gpurun_out/bench.err
