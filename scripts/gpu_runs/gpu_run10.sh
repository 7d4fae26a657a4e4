mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_models.py -m gpu -q --maxfail=10 > gpurun_out/pytest_models.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_models.log
tail -30 gpurun_out/pytest_models.log
for m in sage gcn gin; do timeout 600 python scripts/train_epoch.py --shape products --model $m --epochs 5 --warmup 2 2>&1 | tail -1; done
timeout 600 python scripts/train_epoch.py --shape products --model sage --epochs 5 --warmup 2 --tf32 2>&1 | tail -1
timeout 600 python scripts/train_epoch.py --shape proteins --model gin --k 64 --epochs 5 --warmup 2 2>&1 | tail -1
