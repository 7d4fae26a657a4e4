#!/usr/bin/env python3
"""TEST INFRASTRUCTURE: build tests/simt/_build/libmaxk_simt.so -- the kernel sources of
spgemm-new_b200/csrc translated (translate.py) and compiled with g++ against the SIMT test
interpreter.  Exports the same C ABI as lib/libmaxk_b200.so (include/maxk_b200.h) with host
pointers standing in for device pointers.  Loaded by tests/simt_lib.py only.

    python tests/simt/build.py [--force]
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
# SIMT_CSRC / SIMT_OUT: build from a (mutated) copy of the sources -- tests/simt/mutation_check.py
CSRC = os.environ.get("SIMT_CSRC") or os.path.join(ROOT, "spgemm-new_b200", "csrc")
OUT = os.environ.get("SIMT_OUT") or os.path.join(HERE, "_build")
GEN = os.path.join(OUT, "gen")
LIB = os.path.join(OUT, "libmaxk_simt.so")

sys.path.insert(0, HERE)
import translate  # noqa: E402

CXXFLAGS = ["-std=c++17", "-O1", "-g", "-fPIC", "-fvisibility=hidden", "-fno-omit-frame-pointer",
            "-DMAXK_SIMT", "-DMAXK_BUILDING_LIB", "-Wall", "-Wno-unused-variable", "-Wno-unused-function",
            "-Wno-unknown-pragmas", "-Wno-maybe-uninitialized", "-Wno-uninitialized", "-Wno-attributes",
            "-I", os.path.join(HERE, "include"), "-I", os.path.join(ROOT, "include"), "-I", GEN]


def _deps():
    d = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    d += [os.path.join(HERE, f) for f in ("translate.py", "build.py", "simt_runtime.cpp", "tensor_core_stubs.cpp")]
    d += [os.path.join(HERE, "include", "cuda_runtime.h"), os.path.join(ROOT, "include", "maxk_b200.h")]
    return d


def up_to_date():
    if not os.path.exists(LIB):
        return False
    t = os.path.getmtime(LIB)
    return all(os.path.getmtime(p) <= t for p in _deps())


def build(force=False):
    if not force and up_to_date():
        return LIB
    os.makedirs(GEN, exist_ok=True)
    gen = translate.translate_tree(CSRC, GEN)
    srcs = [p for p in gen if p.endswith(".cpp")] + [os.path.join(HERE, f) for f in ("simt_runtime.cpp", "tensor_core_stubs.cpp")]
    cxx = os.environ.get("SIMT_CXX", "g++")

    def compile_one(src):
        obj = os.path.join(OUT, os.path.basename(src) + ".o")
        subprocess.check_call([cxx, *CXXFLAGS, "-c", src, "-o", obj])
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, srcs))
    subprocess.check_call([cxx, "-shared", "-o", LIB, *objs])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
