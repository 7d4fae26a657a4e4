#!/usr/bin/env python3
"""Produce tests/golden/ref_kernels.npz: outputs of the REFERENCE CUDA kernels (compiled
unmodified into oracle/_ref/libmaxk_ref.so in the build container) on small seeded inputs.
Needs a GPU: run under gpurun, e.g.
    gpurun -- 'python tests/golden/make_golden_gpu.py gpurun_out/ref_kernels.npz'
and copy the file to tests/golden/.  The CPU test tests/test_oracle_vs_reference.py then
pins the oracle against these vectors without a GPU."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from helpers import Problem  # noqa: E402
from oracle import ref_kernels as R  # noqa: E402


def main(out_path):
    blob = {}
    for tag, kw in (("k32", dict(n=96, avg=20.0, k=32, seed=41)),
                    ("k64", dict(n=64, avg=30.0, k=64, seed=42)),
                    ("k16", dict(n=96, avg=20.0, k=16, seed=43)),
                    ("k8", dict(n=96, avg=20.0, k=8, seed=44)),
                    ("hub32", dict(n=12, avg=400.0, k=32, seed=45, empty_frac=0.0))):
        p = Problem(**kw)
        g = p.cuda()
        fwd = R.spmm_forward(g["warp4"], g["indices"], g["values"], g["sp_data"], g["sp_index"], 256)
        bwd = R.spmm_backward(g["warp4"], g["indices"], g["values"], g["dy"], g["sp_index"])
        torch.cuda.synchronize()
        for name, arr in (("indptr", p.indptr), ("idx", p.idx), ("val", p.val), ("sp_data", p.sp_data),
                          ("sp_index", p.sp_index), ("w4", p.w4), ("dy", p.dy),
                          ("ref_fwd", fwd.cpu().numpy()), ("ref_bwd", bwd.cpu().numpy())):
            blob[f"{tag}__{name}"] = arr
    # the reference uint8 topk kernel at the one configuration where its layout is right
    # (k = 32, N % 16 == 0) -- recorded as evidence, defects and all
    x = torch.randint(0, 256, (64, 256), dtype=torch.uint8, device="cuda",
                      generator=torch.Generator("cuda").manual_seed(5))
    v, i = R.topk_u8(x, 32)
    torch.cuda.synchronize()
    blob["topk_u8__x"], blob["topk_u8__val"], blob["topk_u8__idx"] = (x.cpu().numpy(), v.cpu().numpy(),
                                                                       i.cpu().numpy())
    np.savez_compressed(out_path, **blob)
    print("wrote", out_path, len(blob), "arrays")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "ref_kernels.npz"))
