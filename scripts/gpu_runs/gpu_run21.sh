mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
tail -4 gpurun_out/pytest.log
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r21.json 2> gpurun_out/bench_r21.err; tail -5 gpurun_out/bench_r21.err; cat gpurun_out/bench_r21.json | cut -c1-1500
python scripts/sweep_fwd.py --shape products --ks 32,16,64 --copies 0 --tpw=-1 --depth 0 --bwd-tpw=-1 2>&1 | tee gpurun_out/sweep_auto_products.jsonl
python scripts/sweep_fwd.py --shape flickr --ks 32 --copies 0 --tpw=-1 --depth 0 --bwd-tpw=-1,16 2>&1 | tee gpurun_out/sweep_auto_flickr.jsonl
