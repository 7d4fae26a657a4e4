#!/bin/bash
# Round 2, call 31 (1 GPU): TopK / LayerNorm+TopK occupancy: 5 / 6 / 8 CTAs per SM (45-48 / 40 / 32 registers).
O=gpurun_out; mkdir -p $O
for mb in 5 6 8; do
  MAXK_TOPK_MINB=$mb MAXK_LNTOPK_MINB=$mb timeout 200 python scripts/topk_occupancy_probe.py >> $O/r31_topk_occupancy.jsonl 2>> $O/r31_topk_occupancy.err
  echo "minb=$mb rc=$?"; tail -1 $O/r31_topk_occupancy.jsonl | cut -c1-600
done
timeout 200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_layernorm.py -x -q -k "topk or layernorm" > $O/r31_pytest_mb8.log 2>&1; echo "pytest rc=$?"
