mkdir -p gpurun_out
P="python scripts/profile_step.py --k 32 --iters 1 --scale 0.2"
$P > gpurun_out/plain_profile.log 2>&1 && ncu --section SourceCounters --import-source on --clock-control none -k regex:topk_cbsr -c 1 -f -o gpurun_out/prof_topk_v3 $P > gpurun_out/ncu_topk3.log 2>&1
tail -2 gpurun_out/ncu_topk3.log
