#!/bin/bash
# Multi-GPU call of round 2 (N = 2 while developing, 8 to confirm):
#   /usr/local/graft/bin/gpurun --gpus 2 --timeout 900 -- 'bash scripts/round2_multi_gpu.sh 2'
# Parity of the peer-memory path incl. the per-rank host entry point (ShardHostAggregator, one stream and
# overlapped), then the bench line at N with its new e2e / per-GPU roofline legs, with the overlapped e2e,
# over NCCL, and with the planned kernels on the peer transport.
N=${1:-2}
O=gpurun_out
mkdir -p $O
run() { local t=$1 port=$2; shift 2; timeout "$t" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N \
        --master-addr 127.0.0.1 --master-port "$port" "$@"; }
run 400 29701 scripts/peer_check.py --quick > $O/peer_check_n$N.log 2> $O/peer_check_n$N.err
echo "peer_check rc=$? true=$(grep -c 'all-ranks True' $O/peer_check_n$N.log) false=$(grep -c 'all-ranks False' $O/peer_check_n$N.log)"
run 300 29702 bench.py --gpus $N --steps 40 --warmup 5 > $O/scale_peer_n$N.json 2> $O/bench_peer_$N.err
echo "bench peer rc=$? $(cut -c1-200 $O/scale_peer_n$N.json)"
MAXK_SHARD_OVERLAP=1 run 300 29703 bench.py --gpus $N --steps 40 --warmup 5 > $O/scale_peer_overlap_n$N.json 2> $O/bench_peer_overlap_$N.err
echo "bench peer, overlapped e2e rc=$?"
run 300 29704 bench.py --gpus $N --steps 40 --warmup 5 --comm nccl > $O/scale_nccl_n$N.json 2> $O/bench_nccl_$N.err
echo "bench nccl rc=$?"
run 300 29705 bench.py --gpus $N --steps 40 --warmup 5 --planned 8 > $O/scale_peer_planned8_n$N.json 2> $O/bench_planned_$N.err
echo "bench peer planned rc=$?"
for v in 0 1; do
  MAXK_CBSR_NATIVE=$v run 400 2971$v scripts/train_epoch.py --shape products --model sage --epochs 5 --warmup 2 --comm peer \
      2> $O/epoch_n${N}_native$v.err | grep "^{" > $O/epoch_n${N}_sage_native$v.json
  echo "epoch N=$N MAXK_CBSR_NATIVE=$v: $(cut -c1-160 $O/epoch_n${N}_sage_native$v.json)"
done
python - <<'PY'
import glob, json
for p in sorted(glob.glob("gpurun_out/scale_*_n*.json")):
    try:
        d = json.loads(open(p).read().strip().splitlines()[-1])
    except Exception:  # noqa: BLE001
        print(p, "no line"); continue
    e, r = d.get("e2e") or {}, d.get("roofline") or {}
    print(f"{p}: {d['ms_per_step']:.3f} ms/step  e2e {e.get('ms_per_step')} ms  roofline {r.get('kernel')} {r.get('frac')}")
PY
