mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
tail -4 gpurun_out/pytest.log
timeout 900 python scripts/sweep_fwd.py --shape reddit --ks 32,16,8,64 --copies 0,1,2,4,8,16 --tpw 16 --bwd-tpw 8,16,32 > gpurun_out/sweep_reddit.jsonl 2> gpurun_out/sweep_reddit.err
timeout 900 python scripts/sweep_fwd.py --shape reddit --ks 32 --copies 8,16 --tpw 8,32,64 > gpurun_out/sweep_reddit_tpw.jsonl 2>> gpurun_out/sweep_reddit.err
timeout 900 python scripts/sweep_fwd.py --shape products --ks 32 --copies 0,1,2,4,8 --tpw 16,64 --bwd-tpw 16,64 > gpurun_out/sweep_products.jsonl 2> gpurun_out/sweep_products.err
timeout 900 python scripts/sweep_fwd.py --shape proteins --ks 64,32 --copies 0,1,4,8,16 --tpw 16 --bwd-tpw 16 > gpurun_out/sweep_proteins.jsonl 2> gpurun_out/sweep_proteins.err
cat gpurun_out/sweep_reddit.jsonl gpurun_out/sweep_reddit_tpw.jsonl gpurun_out/sweep_products.jsonl gpurun_out/sweep_proteins.jsonl
tail -3 gpurun_out/sweep_*.err
