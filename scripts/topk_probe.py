#!/usr/bin/env python3
"""TopK timing / instruction-count driver (several value distributions)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "spgemm-new_b200"))
import torch
from maxk_b200 import ops
n = 232965
def timeit(fn, iters=20, warm=3):
    for _ in range(warm): fn()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); s.record()
    for _ in range(iters): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / iters
quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
g = torch.Generator("cuda").manual_seed(1)
dists = (("uniform", torch.rand(n, 256, device="cuda", generator=g)),
         ("normal", torch.randn(n, 256, device="cuda", generator=g)),
         ("relu_normal", torch.randn(n, 256, device="cuda", generator=g).relu()),
         ("quantised", (torch.rand(n, 256, device="cuda", generator=g) * 8).floor()))
lanes_list = (0,)
for name, x in dists[:2] if os.environ.get('TOPK_PROBE_SHORT') else dists:
    for k in (8, 32, 64):
        for lanes in lanes_list:
            ops.set_topk_group(lanes)
            if quick:   # for ncu: one launch per (distribution, k, lanes per row)
                ops.topk_cbsr(x, k); continue
            b = 4 * n * 256 + 5 * n * k
            t = timeit(lambda: ops.topk_cbsr(x, k))
            print(f"topk {name:12s} k={k:3d} {'grouped' if lanes else 'one row / warp':14s}: {t*1e3:7.1f} us  {b/t/1e6:7.0f} GB/s  frac {b/t/1e6/6543.7:.3f}", flush=True)
ops.set_topk_group(-1)
for d, k in ((128, 16), (128, 32), (64, 16)):
    x = torch.randn(n * 256 // d, d, device="cuda", generator=g)
    for lanes in (0, -1):
        ops.set_topk_group(lanes)
        b = x.numel() * 4 + 5 * x.shape[0] * k
        t = timeit(lambda: ops.topk_cbsr(x, k))
        print(f"topk normal d={d} k={k:3d} {'grouped' if lanes else 'one row / warp':14s}: {t*1e3:7.1f} us  {b/t/1e6:7.0f} GB/s  frac {b/t/1e6/6543.7:.3f}", flush=True)
ops.set_topk_group(-1)
torch.cuda.synchronize()
