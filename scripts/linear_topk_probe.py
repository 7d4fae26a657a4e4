#!/usr/bin/env python3
"""Fused Linear + TopK (csrc/linear_topk.cu) against torch's fp32 Linear followed by the TopK kernel: timings."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "spgemm-new_b200"))
import torch
from maxk_b200 import ops

def timeit(fn, iters=10, warm=3):
    for _ in range(warm): fn()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); s.record()
    for _ in range(iters): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / iters

quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
g = torch.Generator("cuda").manual_seed(1)
for name, n in (("reddit", 232965), ("products", 2449029)):
    for in_dim in (256,):
        x = torch.randn(n, in_dim, device="cuda", generator=g)
        w = torch.randn(256, in_dim, device="cuda", generator=g) / in_dim ** 0.5
        b = torch.randn(256, device="cuda", generator=g) * 0.1
        for k in (32,) if quick else (16, 32):
            if quick:
                ops.linear_topk_cbsr(x, w, b, k); continue
            t_fused = timeit(lambda: ops.linear_topk_cbsr(x, w, b, k))
            t_gemm_only = timeit(lambda: ops.linear_topk_cbsr(x, w, b, k, want_linear=True, want_cbsr=False))
            t_lin = timeit(lambda: torch.nn.functional.linear(x, w, b))
            y = torch.nn.functional.linear(x, w, b)
            t_topk = timeit(lambda: ops.topk_cbsr(y, k))
            torch.backends.cuda.matmul.allow_tf32 = True
            t_lin_tf32 = timeit(lambda: torch.nn.functional.linear(x, w, b))
            torch.backends.cuda.matmul.allow_tf32 = False
            gy = torch.randn(n, 256, device="cuda", generator=g)
            t_wg = timeit(lambda: ops.linear_wgrad(gy, x))
            t_wg_lib = timeit(lambda: gy.t() @ x)
            del gy
            norm = torch.nn.LayerNorm(256).cuda()
            yy = torch.randn(n, 256, device="cuda", generator=g)
            t_ln_lib = timeit(lambda: norm(yy))
            t_ln = timeit(lambda: ops.layernorm_forward(yy, norm.weight, norm.bias))
            o_, mean_, rstd_ = ops.layernorm_forward(yy, norm.weight, norm.bias)
            t_lnb = timeit(lambda: ops.layernorm_backward(o_, yy, mean_, rstd_, norm.weight))
            yr = yy.clone().requires_grad_(True)
            def lib_bwd():
                norm.zero_grad(); yr.grad = None
                norm(yr).backward(o_)
            t_ln_lib_fb = timeit(lib_bwd)
            del yy, yr, o_
            flops = 2.0 * n * in_dim * 256
            print(json.dumps({"shape": name, "n": n, "in_dim": in_dim, "k": k, "fused_ms": round(t_fused, 4),
                              "fused_linear_only_ms_incl_dense_write": round(t_gemm_only, 4), "torch_linear_fp32_ms": round(t_lin, 4), "topk_ms": round(t_topk, 4),
                              "torch_linear_tf32_ms": round(t_lin_tf32, 4), "wgrad_ms": round(t_wg, 4), "layernorm_fwd_ms": round(t_ln, 4), "torch_layernorm_fwd_ms": round(t_ln_lib, 4), "layernorm_bwd_ms": round(t_lnb, 4), "torch_layernorm_fwd_plus_bwd_ms": round(t_ln_lib_fb, 4), "torch_wgrad_fp32_ms": round(t_wg_lib, 4),
                              "speedup_vs_fp32_linear_plus_topk": round((t_lin + t_topk) / t_fused, 2),
                              "fused_TFLOPs_fp32_equiv": round(flops / t_fused / 1e9, 1),
                              "hbm_bytes_saved": 8 * n * 256}), flush=True)
        del x
torch.cuda.synchronize()
