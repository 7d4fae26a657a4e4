mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
tail -4 gpurun_out/pytest.log
python scripts/sweep_fwd.py --shape products --ks 32,16,64 --copies 0 --tpw 16,64 --depth 1,2 2>&1 | tee gpurun_out/sweep_roll_products.jsonl
python scripts/sweep_fwd.py --shape reddit --ks 32,16,8,64 --copies 0 --tpw 16 --depth 1,2 2>&1 | tee gpurun_out/sweep_roll_reddit.jsonl
python scripts/sweep_fwd.py --shape proteins --ks 64,32 --copies 0 --tpw 16 --depth 1 2>&1 | tee gpurun_out/sweep_roll_proteins.jsonl
python scripts/sweep_fwd.py --shape flickr --ks 32 --copies 0 --tpw 4,16 --depth 1 2>&1 | tee gpurun_out/sweep_roll_flickr.jsonl
