#!/bin/bash
# Round 2, call 35 (1 GPU): fused Linear + TopK with the order-statistics seed in its epilogue -- parity, timings.
O=gpurun_out; mkdir -p $O
timeout 200 python -m pytest tests/test_gpu_linear_topk.py tests/test_gpu_models.py -x -q > $O/r35_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 $O/r35_pytest.log)"
timeout 200 python scripts/linear_topk_quick.py > $O/r35_linear_topk.json 2> $O/r35_linear_topk.err; echo "probe rc=$?"; cat $O/r35_linear_topk.json
