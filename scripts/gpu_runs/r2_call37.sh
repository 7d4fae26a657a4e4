#!/bin/bash
# Round 2, call 37 (1 GPU): TopK with its outputs / k <= 32 known at compile time -- full GPU suite, smoke, timings.
O=gpurun_out; mkdir -p $O
timeout 300 python -m pytest tests -x -q -m gpu > $O/r37_pytest_gpu.log 2>&1; echo "pytest rc=$? $(tail -1 $O/r37_pytest_gpu.log)"
python __graft_entry__.py smoke > $O/r37_smoke.log 2>&1; echo "smoke rc=$? $(tail -1 $O/r37_smoke.log | cut -c1-120)"
TOPK_PROBE_SHORT=1 timeout 100 python scripts/topk_probe.py > $O/r37_topk_probe.log 2>&1; head -6 $O/r37_topk_probe.log
