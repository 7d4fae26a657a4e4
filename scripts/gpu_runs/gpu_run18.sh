mkdir -p gpurun_out
python scripts/sweep_fwd.py --shape products --ks 32,16,64 --copies 0 --tpw 16,64 --depth 1,2,4 2>&1 | tee gpurun_out/sweep_depth_products.jsonl
python scripts/sweep_fwd.py --shape reddit --ks 32,16,64 --copies 0 --tpw 16 --depth 1,2,4 2>&1 | tee gpurun_out/sweep_depth_reddit.jsonl
python scripts/sweep_fwd.py --shape proteins --ks 64,32 --copies 0 --tpw 16 --depth 1,2,4 2>&1 | tee gpurun_out/sweep_depth_proteins.jsonl
