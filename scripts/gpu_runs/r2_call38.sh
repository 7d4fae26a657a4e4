#!/bin/bash
# Round 2, call 38 (1 GPU): LayerNorm + TopK specialised for d = 256 / k <= 32 -- full GPU suite, smoke, its timing, and
# the shape / k sweep of the final build.
O=gpurun_out; mkdir -p $O
timeout 300 python -m pytest tests -x -q -m gpu > $O/r38_pytest_gpu.log 2>&1; echo "pytest rc=$? $(tail -1 $O/r38_pytest_gpu.log)"
python __graft_entry__.py smoke > $O/r38_smoke.log 2>&1; echo "smoke rc=$? $(tail -1 $O/r38_smoke.log | cut -c1-120)"
timeout 100 python scripts/topk_occupancy_probe.py > $O/r38_topk_lntopk.json 2> $O/r38_topk_lntopk.err; cat $O/r38_topk_lntopk.json
timeout 200 python scripts/sweep_shapes.py > $O/r38_sweep.json 2> $O/r38_sweep.err; echo "sweep rc=$?"
