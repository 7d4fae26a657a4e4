#!/bin/bash
# First gpurun call of round 2 (1 GPU, ~12 minutes; every step has its own time limit, so one hang costs
# that step only): validate and measure what round 1 wrote after its GPU budget was spent -- all of it
# already checked under the SIMT interpreter (tests/test_simt_kernels.py) -- and re-measure the one-GPU
# epoch that the scaling table still quotes from the first session.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash scripts/round2_first_run.sh'
# Then: python scripts/summarise_round2.py   (reads gpurun_out/, prints what beat the defaults)
mkdir -p gpurun_out
O=gpurun_out
step() {  # step <seconds> <name> <command...>: stdout -> $O/<name>, stderr -> $O/<name>.err
  local t=$1 name=$2; shift 2
  timeout "$t" "$@" > "$O/$name" 2> "$O/$name.err"
  echo "[$name] rc=$? $(tail -c 300 "$O/$name" | tail -1)"
}
# 0. the default suite first: the build that round 1 left must still be green
step 900 pytest_default.log python -m pytest tests -m gpu -q -x
# 1. parity of every experimental variant on the GPU
MAXK_EXPERIMENTAL=1 step 900 pytest_experimental.log python -m pytest tests -m gpu -q -k "experimental or planned"
# 2. forward launch shapes (double-buffered 4-7, cp.async ring 8-12), planned forward / backward, dY prefetch
for shape in products reddit flickr; do
  step 300 sweep_db_$shape.jsonl python scripts/sweep_fwd.py --shape $shape --ks 32,16,64 --copies 0 --tpw=-1 \
       --depth 0,1,2,4,5,6,7,8,9,10,11,12 --planned
done
# 3. TopK: default and 2 / 4 / 8 rows per warp on four value distributions
step 120 topk_probe_round2.log python scripts/topk_probe.py
# 4. SURVEY 8(f) rank 3: fc_self / GIN self term on the CBSR operand vs torch's dense fp32 path
step 180 cbsr_linear_probe.jsonl python scripts/cbsr_linear_probe.py
for model in sage; do
  for v in 0 1; do
    MAXK_CBSR_NATIVE=$v step 240 epoch_n1_${model}_native$v.json python scripts/train_epoch.py --shape products \
        --model $model --epochs 5 --warmup 2
  done
done
# 5. hardware questions the models had to guess (scalar vs vector reductions at the L2, conflict cost,
#    LDS vs SHFL broadcast, cp.async / cp.async.bulk gathers)
mkdir -p scripts/probes/_bin
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a scripts/probes/pipe_probe.cu -o scripts/probes/_bin/pipe_probe \
  && step 180 pipe_probe.log scripts/probes/_bin/pipe_probe
# 6. planned kernels end to end: the bench step and a dense-degree epoch (GIN k=64 on the proteins shape)
for h in 2 4 8; do
  step 240 bench_planned_$h.json python bench.py --planned $h --steps 20 --warmup 3 --no-cpu-baseline --no-e2e
done
step 240 epoch_proteins_default.json python scripts/train_epoch.py --shape proteins --model gin --k 64 --epochs 5 --warmup 2
MAXK_PLANNED_HOMES=8 step 240 epoch_proteins_planned8.json python scripts/train_epoch.py --shape proteins --model gin --k 64 --epochs 5 --warmup 2
# (with 2 or 8 GPUs: `torchrun ... bench.py --gpus N --planned 8` runs the planned kernels on the peer transport;
#  `MAXK_SHARD_OVERLAP=1 torchrun ... bench.py --gpus N` times the e2e leg with the overlapped ShardHostAggregator)
# 7. ncu of the planned kernels (only after the runs above exited without ncu): launch list + one full capture each
step 300 ncu_planned_list.log ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv \
     --log-file $O/launches_planned.csv python scripts/profile_step.py --planned 8 --iters 2
step 400 ncu_fwd_planned.log ncu --set full --clock-control none --import-source on -k regex:spgemm_fwd_kernel -s 3 -c 1 \
     -f -o $O/prof_fwd_planned8 python scripts/profile_step.py --planned 8 --iters 2
step 400 ncu_bwd_planned.log ncu --set full --clock-control none --import-source on -k regex:sspmm_bwd_planned -s 1 -c 1 \
     -f -o $O/prof_bwd_planned8 python scripts/profile_step.py --planned 8 --iters 2
# ... and of the other new kernels: the CBSR x dense products and the rows-per-warp TopK
step 400 ncu_cbsr_matmul.log ncu --set full --clock-control none --import-source on -k regex:cbsr_matmul_kernel -s 2 -c 1 \
     -f -o $O/prof_cbsr_matmul python scripts/cbsr_linear_probe.py --ks 32 --iters 2
step 400 ncu_cbsr_sampled_dot.log ncu --set full --clock-control none --import-source on -k regex:cbsr_sampled_dot_kernel -s 2 -c 1 \
     -f -o $O/prof_cbsr_sampled_dot python scripts/cbsr_linear_probe.py --ks 32 --iters 2
TOPK_RPW=4 step 300 ncu_topk_rows.log ncu --set full --clock-control none --import-source on -k regex:topk_cbsr_f32_rows_kernel -s 1 -c 1 \
     -f -o $O/prof_topk_rows4 python scripts/topk_probe.py quick
