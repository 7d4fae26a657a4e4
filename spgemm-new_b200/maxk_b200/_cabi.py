"""ctypes binding of libmaxk_b200.so -- the C ABI declared in include/maxk_b200.h.

There is NO fallback: if the CUDA library has not been built the import of any op fails
loudly (north star: "no CPU fallback").  Build it with `python spgemm-new_b200/build.py`
or `__graft_entry__.build()`.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "lib", "libmaxk_b200.so")

MAXK_OK = 0
ABI_VERSION = 1
WARP_MAX_NZ = 64
MAX_DIM = 256

_lib = None

_i64, _i32, _vp = ctypes.c_int64, ctypes.c_int32, ctypes.c_void_p

# name -> (restype, argtypes); mirrors include/maxk_b200.h one to one
PROTOTYPES = {
    "maxk_abi_version": (ctypes.c_int, []),
    "maxk_last_error": (ctypes.c_char_p, []),
    "maxk_launch_count": (_i64, []),
    "maxk_set_tuning": (ctypes.c_int, [ctypes.c_int, ctypes.c_int]),
    "maxk_set_forward_copies": (ctypes.c_int, [ctypes.c_int]),
    "maxk_set_forward_depth": (ctypes.c_int, [ctypes.c_int]),
    "maxk_set_topk_group": (ctypes.c_int, [ctypes.c_int]),
    "maxk_linear_topk_workspace_bytes": (_i64, [_i32, _i32]),
    "maxk_linear_topk_cbsr": (ctypes.c_int, [_vp, _i64, _i32, _vp, _vp, _i32, _i32, _vp, _vp, _vp, _vp, _vp]),
    "maxk_linear_f32": (ctypes.c_int, [_vp, _i64, _i32, _vp, _vp, _vp, _i32, _vp, _vp, _vp]),
    "maxk_linear_wgrad": (ctypes.c_int, [_vp, _vp, _i64, _i32, _i32, _vp, _i32, _vp]),
    "maxk_layernorm_forward": (ctypes.c_int, [_vp, _i64, _i32, _vp, _vp, ctypes.c_float, _vp, _vp, _vp, _vp]),
    "maxk_layernorm_backward": (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, _i64, _i32, _vp, _vp, _vp, _vp]),
    "maxk_layernorm_topk_cbsr": (ctypes.c_int, [_vp, _i64, _i32, _vp, _vp, ctypes.c_float, _i32, _vp, _vp, _vp, _vp, _vp, _vp]),
    "maxk_layernorm_backward_cbsr": (ctypes.c_int, [_vp, _vp, _i32, _vp, _vp, _vp, _vp, _i64, _i32, _vp, _vp, _vp, _vp]),
    "maxk_worklist_workspace_bytes": (_i64, [_i64]),
    "maxk_worklist_count": (ctypes.c_int, [_vp, _i64, _i32, _vp, ctypes.POINTER(_i64), _vp]),
    "maxk_worklist_fill": (ctypes.c_int, [_vp, _i64, _i32, _vp, _vp, _i64, _vp]),
    "maxk_topk_cbsr_f32": (ctypes.c_int, [_vp, _i64, _i32, _i32, _vp, _vp, _vp, _vp]),
    "maxk_topk_cbsr_u8": (ctypes.c_int, [_vp, _i64, _i32, _i32, _vp, _vp, _vp]),
    "maxk_spgemm_forward": (ctypes.c_int, [_vp, _i64, _vp, _vp, _vp, _vp, _i32, _vp, _i64, _i32, _vp]),
    "maxk_spgemm_forward_acc": (ctypes.c_int, [_vp, _i64, _vp, _vp, _vp, _vp, _i32, _vp, _i64, _i32, _i32, _vp]),
    "maxk_cbsr_bank_plan": (ctypes.c_int, [_vp, _i64, _i32, _vp, _vp]),
    "maxk_spgemm_forward_planned": (ctypes.c_int, [_vp, _i64, _vp, _vp, _vp, _vp, _i32, _vp, _i64, _i32, _i32, _vp]),
    "maxk_cbsr_matmul": (ctypes.c_int, [_vp, _vp, _i64, _i32, _vp, _i32, _i32, _vp, _vp, _i32, _vp]),
    "maxk_cbsr_axpy": (ctypes.c_int, [_vp, _vp, _i64, _i32, _i32, _vp, ctypes.c_float, _vp, _vp]),
    "maxk_cbsr_sampled_dot": (ctypes.c_int, [_vp, _i64, _i32, _vp, _i32, _vp, _i32, _vp, _i32, _vp]),
    "maxk_sspmm_backward_planned": (ctypes.c_int, [_vp, _i64, _vp, _vp, _vp, _vp, _i32, _i32, _vp, _i64, _i32, _vp]),
    "maxk_sspmm_backward": (ctypes.c_int, [_vp, _i64, _vp, _vp, _vp, _vp, _i32, _i32, _vp, _i64, _i32, _vp]),
    "maxk_cbsr_scatter": (ctypes.c_int, [_vp, _vp, _i64, _i32, _i32, _vp, _vp]),
    "maxk_cbsr_gather": (ctypes.c_int, [_vp, _vp, _i64, _i32, _i32, _vp, _vp]),
    "maxk_cbsr_mask": (ctypes.c_int, [_vp, _vp, _i64, _i32, _i32, _vp, _vp]),
    "maxk_cbsr_record_bytes": (_i32, [_i32]),
    "maxk_cbsr_pack": (ctypes.c_int, [_vp, _vp, _i64, _i32, _vp, _vp]),
    "maxk_cbsr_unpack": (ctypes.c_int, [_vp, _i64, _i32, _vp, _vp, _vp]),
    "maxk_set_peer_timeout_ms": (ctypes.c_int, [_i64]),
    "maxk_peer_alloc": (ctypes.c_int, [_i64, ctypes.POINTER(_vp), _vp]),
    "maxk_peer_free": (ctypes.c_int, [_vp]),
    "maxk_peer_open": (ctypes.c_int, [_vp, ctypes.POINTER(_vp)]),
    "maxk_peer_close": (ctypes.c_int, [_vp]),
    "maxk_peer_publish": (ctypes.c_int, [_vp, ctypes.c_uint32, _vp]),
    "maxk_spgemm_forward_fused": (ctypes.c_int, [_vp, ctypes.c_uint32, _vp, _i64, _i64, _vp, _vp, _vp,
                                                 _i64, _i32, _i32, _vp]),
    "maxk_spgemm_forward_fused_planned": (ctypes.c_int, [_vp, ctypes.c_uint32, _vp, _i64, _i64, _vp, _vp, _vp,
                                                         _i64, _i32, _i32, _vp]),
    "maxk_peer_zero_partial": (ctypes.c_int, [_vp, _vp]),
    "maxk_reduce_scatter_pull": (ctypes.c_int, [_vp, ctypes.c_uint32, _vp, _vp]),
    # reference-named shims (cuda_kernel_wrappers.cu:38-91); dim3 by value
    "spmm_kernel_opt2_sparse_v3_wrapper": (None, None),
    "spmm_kernel_opt2_sparse_backward_v3_wrapper": (None, None),
    "topk_kernel_wrapper": (None, None),
}


class MaxKLibraryMissing(ImportError):
    pass


def load():
    """Load (once) and return the ctypes handle; raises if the library is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MaxKLibraryMissing(
            f"{LIB_PATH} not found: build the CUDA library first "
            "(python spgemm-new_b200/build.py). There is no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if a declared symbol is not exported
        if args is not None:
            fn.restype = res
            fn.argtypes = args
    if lib.maxk_abi_version() != ABI_VERSION:
        raise ImportError(f"libmaxk_b200 ABI {lib.maxk_abi_version()} != expected {ABI_VERSION}")
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    """Translate a MAXK_ERR_* code into the RuntimeError TORCH_CHECK would have raised."""
    if rc != MAXK_OK:
        msg = load().maxk_last_error().decode("utf-8", "replace")
        raise RuntimeError(msg or f"{what} failed with code {rc}")


def launch_count() -> int:
    return int(load().maxk_launch_count())
