#!/usr/bin/env python3
"""Build libmaxk_b200.so (the C-ABI CUDA library) in-tree for sm_100a.

    python spgemm-new_b200/build.py [--force] [--verbose]

One nvcc invocation, no torch headers: the library's ABI is plain C (include/maxk_b200.h).
The .so lands in spgemm-new_b200/lib/ (git-ignored, travels to the GPU box with gpurun).
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIB_DIR, "libmaxk_b200.so")
SOURCES = ["api.cu", "worklist.cu", "topk_cbsr.cu", "spgemm_forward.cu", "sspmm_backward.cu",
           "cbsr_layout.cu", "cbsr_plan.cu", "cbsr_dense.cu", "peer.cu", "linear_topk.cu", "linear_wgrad.cu", "layernorm.cu"]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def up_to_date() -> bool:
    if not os.path.exists(LIB):
        return False
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(ROOT, "include", "maxk_b200.h")]
    return all(os.path.getmtime(d) <= t for d in deps)


def build(force: bool = False, verbose: bool = False, extra=()) -> str:
    if not force and up_to_date():
        return LIB
    os.makedirs(LIB_DIR, exist_ok=True)
    cmd = [nvcc_path(), "-O3", "-std=c++17", *ARCH, "-lineinfo", "--threads", "0", "-shared", "-Xcompiler", "-fPIC",
           "-Xcompiler", "-fvisibility=hidden", "-I", os.path.join(ROOT, "include"), "-I", CSRC,
           "-DMAXK_BUILDING_LIB", *extra]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += [os.path.join(CSRC, s) for s in SOURCES] + ["-o", LIB]
    env = dict(os.environ)
    # the image exports CC/CXX wrappers; nvcc only needs a host g++ on PATH
    subprocess.check_call(cmd, env=env)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
