mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q -k "multi_rank_on_one_gpu" > gpurun_out/pytest_mr.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_mr.log
grep -E "passed|failed|FAILED|rc=|timed out" gpurun_out/pytest_mr.log | tail -8 | cut -c1-200
