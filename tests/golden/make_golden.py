#!/usr/bin/env python3
"""Generate the committed golden vectors under tests/golden/.

Run in the BUILD container only (it reads /root/reference, which does not exist on the
GPU box).  Two families are produced here on the CPU:

  worklist_cases.npz   -- the reference's own kernels/generate_meta.py executed
                          unmodified (subprocess, cwd = a temp dir holding graphs/*.indptr
                          and *.indices in the raw int32 format of SURVEY Appendix B) on
                          small CSR graphs covering the edge cases; its .warp4 output is
                          stored next to the inputs.
  topk_cases.npz       -- torch.topk (the arithmetic utils/models.py:48 and
                          kernels/spmm_bindings.cpp:179 call) on tie-free fp32 rows, stored
                          as the column-sorted index set + values.

The third family (ref_kernels_*.npz: outputs of the reference CUDA kernels compiled
unmodified) needs a GPU and is produced by tests/golden/make_golden_gpu.py under gpurun.
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("MAXK_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(HERE, "..", ".."))


def worklist_inputs():
    rng = np.random.default_rng(7)
    cases = {}

    def from_deg(name, deg):
        deg = np.asarray(deg, dtype=np.int64)
        indptr = np.zeros(deg.shape[0] + 1, dtype=np.int32)
        indptr[1:] = np.cumsum(deg)
        n = deg.shape[0]
        indices = rng.integers(0, max(n, 1), size=int(indptr[-1]), dtype=np.int32)
        cases[name] = (indptr, indices)

    from_deg("boundaries", [0, 1, 63, 64, 65, 127, 128, 129, 0, 0, 640, 641, 1, 0])
    from_deg("all_empty_but_one", [0, 0, 0, 5, 0, 0])
    from_deg("single_big", [5000])
    from_deg("ones", [1] * 100)
    from_deg("lognormal", np.round(rng.lognormal(3.0, 1.3, size=400)).astype(np.int64))
    from_deg("dense_uniform", rng.integers(250, 350, size=60))
    from_deg("trailing_empty", [70, 3, 0, 0, 0])
    return cases


def run_reference_generate_meta(cases):
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        os.makedirs(os.path.join(tmp, "graphs"))
        for name, (indptr, indices) in cases.items():
            indptr.tofile(os.path.join(tmp, "graphs", name + ".indptr"))
            indices.tofile(os.path.join(tmp, "graphs", name + ".indices"))
        subprocess.check_call([sys.executable, os.path.join(REF, "kernels", "generate_meta.py")],
                              cwd=tmp, stdout=subprocess.DEVNULL)
        for name in cases:
            out[name] = np.fromfile(os.path.join(tmp, "w12_nz64_warp_4", name + ".warp4"),
                                    dtype=np.int32)
    return out


def main():
    cases = worklist_inputs()
    warp4 = run_reference_generate_meta(cases)
    blob = {}
    for name, (indptr, indices) in cases.items():
        blob[name + "__indptr"] = indptr
        blob[name + "__warp4"] = warp4[name]
    np.savez_compressed(os.path.join(HERE, "worklist_cases.npz"), **blob)
    print("worklist_cases.npz:", {k: int(v.shape[0]) // 4 for k, v in warp4.items()})

    import torch
    torch.manual_seed(123)  # the reference's RNG recipe (main.cu:74-97, direct_kernel_interface.py:320)
    blob = {}
    for d, n in ((256, 48), (128, 16), (64, 8)):
        x = torch.rand(n, d, dtype=torch.float32)
        # tie-free by construction check
        assert all(torch.unique(r).numel() == d for r in x)
        blob[f"x_d{d}"] = x.numpy()
        for k in (1, 8, 16, 18, 32, 64):
            if k > d:
                continue
            v, i = torch.topk(x, k, dim=1)
            order = torch.argsort(i, dim=1)
            blob[f"idx_d{d}_k{k}"] = torch.gather(i, 1, order).numpy().astype(np.uint8)
            blob[f"val_d{d}_k{k}"] = torch.gather(v, 1, order).numpy()
    # signed values (negative activations happen after Linear)
    x = torch.randn(32, 256, dtype=torch.float32)
    assert all(torch.unique(r).numel() == 256 for r in x)
    blob["xneg_d256"] = x.numpy()
    for k in (8, 32, 64):
        v, i = torch.topk(x, k, dim=1)
        order = torch.argsort(i, dim=1)
        blob[f"negidx_d256_k{k}"] = torch.gather(i, 1, order).numpy().astype(np.uint8)
        blob[f"negval_d256_k{k}"] = torch.gather(v, 1, order).numpy()
    np.savez_compressed(os.path.join(HERE, "topk_cases.npz"), **blob)
    print("topk_cases.npz:", len(blob), "arrays")


if __name__ == "__main__":
    main()
