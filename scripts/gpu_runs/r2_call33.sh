#!/bin/bash
# Round 2, call 33 (1 GPU): TopK CTA shape (8 / 4 / 2 warps) and a persistent grid-stride form (rows of a warp in a loop).
O=gpurun_out; mkdir -p $O
for v in w8 w4 w2 p8 p4 p2; do
  MAXK_TOPK_VARIANT=$v timeout 200 python scripts/topk_occupancy_probe.py >> $O/r33_topk_variants.jsonl 2>> $O/r33_topk_variants.err
  echo "variant=$v rc=$?"; tail -1 $O/r33_topk_variants.jsonl | cut -c1-420
done
for v in p4 w2; do
MAXK_TOPK_VARIANT=$v timeout 200 python -m pytest tests/test_gpu_parity.py -x -q -k "topk" > $O/r33_pytest_$v.log 2>&1; echo "pytest $v rc=$? $(tail -1 $O/r33_pytest_$v.log)"
done
