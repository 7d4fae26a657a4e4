mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
tail -4 gpurun_out/pytest.log
python scripts/topk_probe.py 2>&1 | tee gpurun_out/topk_probe.log
P="python scripts/profile_step.py --k 32 --iters 1"
$P --shape products > gpurun_out/plain_profile.log 2>&1 && ncu --set full --import-source on --clock-control none -k regex:"spgemm_fwd|sspmm_bwd" -c 2 -f -o gpurun_out/prof_products_k32 $P --shape products > gpurun_out/ncu_products.log 2>&1
$P --scale 0.5 > gpurun_out/plain_profile2.log 2>&1 && ncu --set full --import-source on --clock-control none -k regex:topk_cbsr -c 1 -f -o gpurun_out/prof_topk_v5 $P --scale 0.5 > gpurun_out/ncu_topk5.log 2>&1
tail -2 gpurun_out/ncu_products.log gpurun_out/ncu_topk5.log
