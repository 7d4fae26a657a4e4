"""ctypes front-end of oracle/_ref/libmaxk_ref.so: the REFERENCE CUDA kernels compiled
unmodified (see oracle/Makefile `ref`, oracle/ref_harness.cu).

TEST INFRASTRUCTURE ONLY.  Comparator for the GPU parity tests and the "reference kernels on
the same B200" bar in bench.py.  The library is prebuilt in the build container (the GPU box
has no /root/reference); `available()` says whether it travelled.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libmaxk_ref.so")
_lib = None


def build(reference_root: str = "/root/reference") -> bool:
    """Compile the reference kernels where they lie; returns False when the tree is absent."""
    if not os.path.isdir(os.path.join(reference_root, "kernels")):
        return os.path.exists(LIB_PATH)
    subprocess.check_call(["make", "-C", _HERE, "-s", "ref", f"REF={reference_root}"])
    return True


def available() -> bool:
    return os.path.exists(LIB_PATH)


def _load():
    global _lib
    if _lib is None:
        lib = ctypes.CDLL(LIB_PATH)
        vp, ci = ctypes.c_void_p, ctypes.c_int
        lib.ref_spmm_forward.argtypes = [vp, vp, vp, vp, vp, vp, ci, ci, ci, ci, ci]
        lib.ref_spmm_backward.argtypes = [vp, vp, vp, vp, vp, vp, ci, ci, ci, ci, ci]
        lib.ref_topk_u8.argtypes = [vp, vp, vp, ctypes.c_uint, ci, ci]
        _lib = lib
    return _lib


def _check(rc, what):
    if rc != 0:
        raise RuntimeError(f"reference {what} failed: cudaError {rc}")


def spmm_forward(warp4, indices, values, sp_data, sp_index, feat_dim=256, out=None):
    """Reference forward kernel on the legacy default stream; `out` must be zero on entry."""
    lib = _load()
    n, k = sp_data.shape
    if out is None:
        out = torch.zeros((n, feat_dim), dtype=torch.float32, device=sp_data.device)
    _check(lib.ref_spmm_forward(warp4.data_ptr(), indices.data_ptr(), values.data_ptr(),
                                sp_data.data_ptr(), sp_index.data_ptr(), out.data_ptr(), n,
                                indices.numel(), feat_dim, k, warp4.numel() // 4), "forward")
    return out


def spmm_backward(warp4, indices, values, dy, sp_index, out=None):
    lib = _load()
    n, d = dy.shape
    k = sp_index.shape[1]
    if out is None:
        out = torch.zeros((sp_index.shape[0], k), dtype=torch.float32, device=dy.device)
    _check(lib.ref_spmm_backward(warp4.data_ptr(), indices.data_ptr(), values.data_ptr(),
                                 dy.data_ptr(), sp_index.data_ptr(), out.data_ptr(), n,
                                 indices.numel(), d, k, warp4.numel() // 4), "backward")
    return out


def topk_u8(x, k):
    """Reference uint8 topk kernel (only meaningful for k=32, N%16==0: SURVEY Q5-Q8)."""
    lib = _load()
    n, d = x.shape
    val = torch.zeros((n, k), dtype=torch.uint8, device=x.device)
    ind = torch.zeros((n, k), dtype=torch.uint8, device=x.device)
    _check(lib.ref_topk_u8(x.data_ptr(), val.data_ptr(), ind.data_ptr(), k, n, d), "topk")
    return val, ind


# ---- zero-edit route: the reference's OWN binding file, unmodified, on our library ----
ZEROEDIT_PATH = os.path.join(_HERE, "_ref", "zeroedit", "maxk_cuda_kernels.so")


def build_zeroedit(reference_root: str = "/root/reference") -> bool:
    """Compile /root/reference/cuda_kernel_bindings.cpp + kernels/spmm_cusparse.cu where they lie and link them
    against spgemm-new_b200/lib/libmaxk_b200.so (oracle/Makefile `zeroedit`); False when the tree is absent."""
    if not os.path.isfile(os.path.join(reference_root, "cuda_kernel_bindings.cpp")):
        return os.path.exists(ZEROEDIT_PATH)
    import sys
    subprocess.check_call(["make", "-C", _HERE, "-s", "zeroedit", f"REF={reference_root}", f"PYTHON={sys.executable}"])
    return True


def zeroedit_available() -> bool:
    return os.path.exists(ZEROEDIT_PATH)


def load_zeroedit():
    """The reference's pybind module `maxk_cuda_kernels` (cuda_kernel_bindings.cpp:429-490) built on our kernels.
    Loaded under a private name so that it does not shadow the package's Python mirror of the same name."""
    import importlib.util
    import torch  # noqa: F401  (libtorch must be loaded first)
    spec = importlib.util.spec_from_file_location("maxk_cuda_kernels", ZEROEDIT_PATH)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod
