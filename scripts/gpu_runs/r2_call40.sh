#!/bin/bash
# Round 2, call 40 (1 GPU): TopK / LayerNorm + TopK with the seed configuration as a template parameter -- full suite,
# smoke, timings.  (The last call of the round's budget.)
O=gpurun_out; mkdir -p $O
timeout 60 python -m pytest tests -x -q -m gpu > $O/r40_pytest_gpu.log 2>&1; echo "pytest rc=$? $(tail -1 $O/r40_pytest_gpu.log)"
python __graft_entry__.py smoke > $O/r40_smoke.log 2>&1; echo "smoke rc=$? $(tail -1 $O/r40_smoke.log | cut -c1-100)"
timeout 30 python scripts/topk_occupancy_probe.py > $O/r40_topk_lntopk.json 2> $O/r40_topk_lntopk.err; cat $O/r40_topk_lntopk.json
