#!/usr/bin/env python3
"""TopK / LayerNorm+TopK at one launch-bound setting (env MAXK_TOPK_MINB / MAXK_LNTOPK_MINB): products-sized and
Reddit-sized activations, uniform and normal values."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "spgemm-new_b200"))
import torch
from maxk_b200 import ops

def timeit(fn, iters=20, warm=3):
    for _ in range(warm): fn()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); s.record()
    for _ in range(iters): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / iters

g = torch.Generator("cuda").manual_seed(1)
out = {"variant": os.environ.get("MAXK_TOPK_VARIANT")}
for n, tag in ((232965, "reddit"), (2449029, "products")):
    for name, x in (("uniform", torch.rand(n, 256, device="cuda", generator=g)),
                    ("normal", torch.randn(n, 256, device="cuda", generator=g))):
        for k in (32, 64):
            out[f"topk_{tag}_{name}_k{k}_us"] = round(timeit(lambda: ops.topk_cbsr(x, k)) * 1e3, 1)
        if name == "normal":
            gamma = torch.rand(256, device="cuda", generator=g) + 0.5
            beta = torch.randn(256, device="cuda", generator=g)
            out[f"lntopk_{tag}_k32_us"] = round(timeit(lambda: ops.layernorm_topk_cbsr(x, gamma, beta, 1e-5, 32)) * 1e3, 1)
        del x
print(json.dumps(out), flush=True)
