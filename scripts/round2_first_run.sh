#!/bin/bash
# First gpurun call of round 2 (1 GPU, ~10 minutes): validate and measure what round 1 wrote after
# its GPU budget was spent -- the double-buffered (launch shapes 4-7) and cp.async-ring (8-11)
# forward variants and the bank-planned forward (csrc/cbsr_plan.cu), all already checked under the SIMT interpreter (tests/test_simt_kernels.py) -- and
# re-measure the one-GPU epoch that the scaling table still quotes from the first session.
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash scripts/round2_first_run.sh'
mkdir -p gpurun_out
MAXK_EXPERIMENTAL=1 timeout 600 python -m pytest tests -m gpu -q -k "experimental" > gpurun_out/pytest_experimental.log 2>&1
echo "experimental shapes: rc=$? $(tail -1 gpurun_out/pytest_experimental.log)"
for shape in products reddit flickr; do
  python scripts/sweep_fwd.py --shape $shape --ks 32,16,64 --copies 0 --tpw=-1 --depth 0,1,2,4,5,6,7,8,9,10,11,12 --planned 2>/dev/null | tee gpurun_out/sweep_db_$shape.jsonl
done
python scripts/train_epoch.py --shape products --model sage --epochs 5 --warmup 2 2>/dev/null | grep "^{" | tee gpurun_out/epoch_n1_round2.json
python scripts/topk_probe.py 2>/dev/null | tee gpurun_out/topk_probe_round2.log   # default and 2/4/8 rows per warp
# fc_self on the CBSR operand vs torch's dense fp32 Linear (products-sized), then the epoch with it
python scripts/cbsr_linear_probe.py 2>/dev/null | tee gpurun_out/cbsr_linear_probe.jsonl
MAXK_CBSR_NATIVE=1 python scripts/train_epoch.py --shape products --model sage --epochs 5 --warmup 2 2>/dev/null | grep "^{" | tee gpurun_out/epoch_n1_cbsr_native_sage.json
for v in 0 1; do
  MAXK_CBSR_NATIVE=$v python scripts/train_epoch.py --shape products --model gin --epochs 5 --warmup 2 2>/dev/null | grep "^{" | tee gpurun_out/epoch_n1_gin_native$v.json
done
# hardware questions the models had to guess (scalar vs vector reductions at the L2, conflict cost,
# LDS vs SHFL broadcast, cp.async / cp.async.bulk gathers): one short probe, bounded
mkdir -p scripts/probes/_bin
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a scripts/probes/pipe_probe.cu -o scripts/probes/_bin/pipe_probe \
  && timeout 180 scripts/probes/_bin/pipe_probe | tee gpurun_out/pipe_probe.log
# planned kernels end to end: the bench step and a dense-degree epoch (GIN k=64 on the proteins shape)
for h in 2 4 8; do
  python bench.py --planned $h --steps 20 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | tee gpurun_out/bench_planned_$h.json
done
python scripts/train_epoch.py --shape proteins --model gin --k 64 --epochs 5 --warmup 2 2>/dev/null | grep "^{" | tee gpurun_out/epoch_proteins_default.json
MAXK_PLANNED_HOMES=8 python scripts/train_epoch.py --shape proteins --model gin --k 64 --epochs 5 --warmup 2 2>/dev/null | grep "^{" | tee gpurun_out/epoch_proteins_planned8.json
MAXK_PLANNED_HOMES=auto python scripts/train_epoch.py --shape proteins --model gin --k 64 --epochs 5 --warmup 2 2>/dev/null | grep "^{" | tee gpurun_out/epoch_proteins_auto.json
# (with 2 or 8 GPUs: `torchrun ... bench.py --gpus N --planned 8` runs the planned kernels on the peer transport)
# ncu of the planned kernels (only after the runs above exited 0 without ncu): launch list + one full capture each
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_planned.csv \
  python scripts/profile_step.py --planned 8 --iters 2 > gpurun_out/ncu_planned_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:spgemm_fwd_kernel -s 3 -c 1 -o gpurun_out/prof_fwd_planned8 \
  python scripts/profile_step.py --planned 8 --iters 2 > gpurun_out/ncu_fwd_planned.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sspmm_bwd_planned -s 1 -c 1 -o gpurun_out/prof_bwd_planned8 \
  python scripts/profile_step.py --planned 8 --iters 2 > gpurun_out/ncu_bwd_planned.log 2>&1
