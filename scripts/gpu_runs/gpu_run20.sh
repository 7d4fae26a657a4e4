mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "forward or fwd or spgemm or dropin or reference" > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
tail -3 gpurun_out/pytest.log
python scripts/sweep_fwd.py --shape products --ks 32 --copies 0 --tpw 4,16,64 --depth 1,2,3,4,5,6,7 2>&1 | tee gpurun_out/sweep_occ_products.jsonl
python scripts/sweep_fwd.py --shape reddit --ks 32 --copies 0 --tpw 16 --depth 1,2,3,4,5,6,7 2>&1 | tee gpurun_out/sweep_occ_reddit.jsonl
python scripts/sweep_fwd.py --shape flickr --ks 32 --copies 0 --tpw 2,4,16 --depth 1,2,5 2>&1 | tee gpurun_out/sweep_occ_flickr.jsonl
