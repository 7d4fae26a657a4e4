mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
tail -4 gpurun_out/pytest.log
python scripts/sweep_fwd.py --shape flickr --ks 32 --copies 0 --tpw=-1 --depth 0 --bwd-tpw=-1,16 2>&1 | tee gpurun_out/sweep_auto_flickr.jsonl
python scripts/sweep_fwd.py --shape reddit --ks 16,8 --copies 0 --tpw=-1 --depth 0,2 2>&1 | tee gpurun_out/sweep_auto_reddit.jsonl
