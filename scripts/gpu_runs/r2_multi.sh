#!/bin/bash
# Round 2, multi-GPU call:  gpurun --gpus N -- 'bash scripts/gpu_runs/r2_multi.sh N [quick]'
# Parity of the peer-memory path on real GPUs (peer_check: plan travelling, pipelined host entry point), then the
# bench line at N exactly as the driver launches it (parity_ok, per-GPU roofline, pipelined e2e, epoch leg).
N=${1:-2}
O=gpurun_out; mkdir -p $O
run() { local t=$1 port=$2; shift 2; timeout "$t" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N \
        --master-addr 127.0.0.1 --master-port "$port" "$@"; }
if [ "$2" != "benchonly" ]; then
run 500 29701 scripts/peer_check.py --quick > $O/r4_peer_check_n$N.log 2> $O/r4_peer_check_n$N.err
echo "peer_check rc=$? true=$(grep -c 'all-ranks True' $O/r4_peer_check_n$N.log) false=$(grep -c 'all-ranks False' $O/r4_peer_check_n$N.log)"
fi
run 600 29702 bench.py --gpus $N --steps 40 --warmup 5 > $O/r4_scale_n$N.json 2> $O/r4_bench_n$N.err
echo "bench rc=$? $(cut -c1-150 $O/r4_scale_n$N.json)"
tail -5 $O/r4_bench_n$N.err | cut -c1-300
if [ "$2" == "phases" ]; then
run 300 29703 scripts/dist_phases.py > $O/r4_phases_n$N.json 2> $O/r4_phases_n$N.err
echo "phases rc=$? $(cut -c1-600 $O/r4_phases_n$N.json)"
fi
if [ "$2" == "rmat" ] || [ "$3" == "rmat" ]; then
run 900 29704 scripts/rmat_sweep.py --cases 20:20,20:100,20:500,24:5,24:20,25:5,26:5 > $O/r4_rmat_n$N.jsonl 2> $O/r4_rmat_n$N.err
echo "rmat rc=$? lines=$(grep -c '^{' $O/r4_rmat_n$N.jsonl)"
fi
