mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 -k "topk or smoke or dropin or autograd or full_size" > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
tail -4 gpurun_out/pytest.log
python - <<'PY'
import sys, torch
sys.path.insert(0, "spgemm-new_b200")
from maxk_b200 import ops
def timeit(fn, iters=20, warm=3):
    for _ in range(warm): fn()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); s.record()
    for _ in range(iters): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / iters
n = 232965
for name, x in (("uniform", torch.rand(n, 256, device="cuda")), ("normal", torch.randn(n, 256, device="cuda")),
                ("relu_normal", torch.randn(n, 256, device="cuda").relu()),
                ("quantised", (torch.rand(n, 256, device="cuda") * 8).floor())):
    for k in (8, 32, 64):
        t = timeit(lambda: ops.topk_cbsr(x, k))
        b = 4 * n * 256 + 5 * n * k
        tt = timeit(lambda: torch.topk(x, k, dim=1))
        print(f"topk {name:12s} k={k:3d}: {t*1e3:7.1f} us  {b/t/1e6:7.0f} GB/s  frac {b/t/1e6/6543.7:.3f}   torch.topk {tt*1e3:7.1f} us", flush=True)
x = torch.rand(n, 256, device="cuda")
t = timeit(lambda: ops.topk_cbsr(x, 32, want_dense=True))
print(f"topk+dense k=32: {t*1e3:.1f} us")
PY
