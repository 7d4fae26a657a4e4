"""CPU-side checks of the drop-in boundary: the C-ABI library loads without a GPU and
exports every symbol include/maxk_b200.h declares; host-side validation raises the
RuntimeErrors the reference's TORCH_CHECKs would."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "maxk_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    # every exported function is declared as `MAXK_API <type> name(`
    names = re.findall(r"MAXK_API[^;(]*?\b([A-Za-z_][A-Za-z0-9_]*)\s*\(", src)
    return sorted(set(names) - {"__attribute__"})  # the MAXK_API macro definition itself


@pytest.fixture(scope="module")
def lib():
    import sys
    sys.path.insert(0, os.path.join(ROOT, "spgemm-new_b200"))
    import build as _build  # spgemm-new_b200/build.py
    path = _build.build()
    return ctypes.CDLL(path)


def test_header_symbols_are_exported(lib):
    names = declared_symbols()
    assert "maxk_spgemm_forward" in names and "topk_kernel_wrapper" in names
    assert len(names) >= 20
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_binding_table_matches_header(lib):
    from maxk_b200 import _cabi
    assert sorted(_cabi.PROTOTYPES) == declared_symbols()
    h = _cabi.load()
    assert h.maxk_abi_version() == _cabi.ABI_VERSION
    assert h.maxk_cbsr_record_bytes(32) == 160
    assert h.maxk_cbsr_record_bytes(8) == 48
    assert h.maxk_cbsr_record_bytes(18) == 96
    assert h.maxk_worklist_workspace_bytes(2048) >= 12


def test_argument_errors_without_gpu(lib):
    """Invalid arguments are rejected on the host before anything touches the device."""
    from maxk_b200 import _cabi
    h = _cabi.load()
    rc = h.maxk_topk_cbsr_f32(None, 4, 256, 0, None, None, None, None)
    assert rc == 1 and b"Invalid k" in h.maxk_last_error()
    rc = h.maxk_spgemm_forward(None, 1, None, None, None, None, 32, None, 4, 250, None)
    assert rc == 1 and b"multiple of 4" in h.maxk_last_error()
    rc = h.maxk_sspmm_backward(None, 1, None, None, None, None, 300, 256, None, 4, 0, None)
    assert rc == 1
    with pytest.raises(RuntimeError):
        _cabi.check(rc)
    # the CBSR x dense products (experimental): shapes are checked before anything is launched
    rc = h.maxk_cbsr_matmul(None, None, 4, 32, None, 256, 250, None, None, 0, None)
    assert rc == 1 and b"multiple of 4" in h.maxk_last_error()
    rc = h.maxk_cbsr_matmul(None, None, 4, 32, None, 256, 256, None, None, 0, None)
    assert rc == 1 and b"NULL pointer" in h.maxk_last_error()
    rc = h.maxk_cbsr_sampled_dot(None, 4, 256, None, 300, None, 32, None, 0, None)
    assert rc == 1 and b"d=300" in h.maxk_last_error()
    rc = h.maxk_cbsr_axpy(None, None, 4, 0, 256, None, ctypes.c_float(1.0), None, None)
    assert rc == 1 and b"k=0" in h.maxk_last_error()
    assert h.maxk_cbsr_matmul(None, None, 0, 32, None, 256, 256, None, None, 0, None) == 0   # n = 0: nothing to do
    with pytest.raises(RuntimeError, match="CUDA"):
        from maxk_b200 import ops
        ops.cbsr_matmul(torch.zeros(2, 4), torch.zeros(2, 4, dtype=torch.uint8), torch.zeros(8, 4))


def test_peer_ctx_struct_matches_header(tmp_path):
    """The ctypes mirror of `maxk_peer_ctx` has the size and field offsets the C compiler gives
    the struct in include/maxk_b200.h (the multi-GPU part of the ABI is a struct by pointer)."""
    import subprocess
    from maxk_b200.peer import PEER_MAX, PeerCtx
    fields = [name for name, _ in PeerCtx._fields_]
    prog = ['#include <stdio.h>', '#include <stddef.h>', '#include "maxk_b200.h"', 'int main(void) {',
            '  printf("%d %zu\\n", MAXK_PEER_MAX, sizeof(maxk_peer_ctx));']
    prog += [f'  printf("{f} %zu\\n", offsetof(maxk_peer_ctx, {f}));' for f in fields]
    prog += ['  return 0; }']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(prog))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    lines = subprocess.check_output([str(exe)], text=True).split("\n")
    peer_max, size = (int(v) for v in lines[0].split())
    assert peer_max == PEER_MAX and size == ctypes.sizeof(PeerCtx)
    for line in lines[1:]:
        if line:
            name, off = line.split()
            assert getattr(PeerCtx, name).offset == int(off), name


def test_peer_argument_errors_without_gpu(lib):
    from maxk_b200 import _cabi
    from maxk_b200.peer import PeerCtx
    h = _cabi.load()
    ctx = PeerCtx()
    ctx.world, ctx.rank, ctx.k, ctx.n_total = 99, 0, 32, 10
    assert h.maxk_reduce_scatter_pull(ctypes.byref(ctx), 1, None, None) == 1
    assert b"out of range" in h.maxk_last_error()
    ctx.world, ctx.rows_of[0], ctx.row_start[0] = 1, 10, 1     # start * k not 16-byte aligned... 32 is
    ctx.k = 4                                                   # ... with k = 4 it is not
    assert h.maxk_peer_zero_partial(ctypes.byref(ctx), None) == 1
    assert b"multiple of 16" in h.maxk_last_error() or b"outside" in h.maxk_last_error()
    assert h.maxk_spgemm_forward_fused(None, 1, None, 0, 0, None, None, None, 0, 256, 1, None) == 1
    assert h.maxk_peer_publish(None, 1, None) == 1


def test_python_surface_rejects_cpu_tensors(lib):
    from maxk_b200 import ops
    x = torch.rand(4, 256)
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.topk_cbsr(x, 32)
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.build_worklist(torch.zeros(5, dtype=torch.int32))
    import maxk_cuda_kernels as mk
    with pytest.raises(RuntimeError, match="Input must be on CUDA"):
        mk.cuda_topk_maxk_float(x, 32)
    with pytest.raises(RuntimeError, match="warp4_metadata must be CUDA tensor"):
        mk.spmm_maxk_forward(torch.zeros(4, dtype=torch.int32), torch.zeros(1, dtype=torch.int32),
                             torch.zeros(1), torch.zeros(1, 32), torch.zeros(1, 32, dtype=torch.uint8), 1, 32)
    with pytest.raises(RuntimeError, match="Cannot open warp4 file"):
        mk.load_warp4_metadata("no_such_graph")


def test_reference_surface_names():
    """Every name the reference's Path-B module and Path-A module export exists."""
    import maxk_cuda_kernels as mk
    import spmm_kernels as sk
    for n in ("spmm_maxk_forward", "spmm_maxk_backward", "cuda_topk_maxk", "cuda_topk_maxk_float",
              "prepare_cbsr_format_maxk", "load_warp4_metadata", "generate_sparse_selector",
              "benchmark_spmm_maxk", "validate_spmm_maxk", "cusparse_spmm", "CudaTimer"):
        assert hasattr(mk, n), n
    for n in ("SpmmMaxK", "SpmmMaxKBackward", "prepare_cbsr_format", "topk_nonlinearity"):
        assert hasattr(sk, n), n
    from direct_kernel_interface import DirectMaxKKernels
    for n in ("load_warp4_metadata", "generate_maxk_sparse_data", "run_forward_kernel",
              "run_backward_kernel", "validate_against_cusparse", "benchmark_all_k_values"):
        assert hasattr(DirectMaxKKernels, n), n
    from maxk_b200 import MaxK, SpGEMMFunction  # noqa: F401


def test_synthetic_graph_shapes():
    from maxk_b200 import graphs
    g = graphs.make_graph("flickr", seed=0, scale=0.05)
    n, e = g["v_num"], g["e_num"]
    assert g["indptr"].dtype == torch.int32 and g["indices"].dtype == torch.int32
    assert int(g["indptr"][-1]) == e == g["values"].numel()
    assert int(g["indices"].max()) < n and int(g["indices"].min()) >= 0
    deg = (g["indptr"][1:] - g["indptr"][:-1])
    assert (deg >= 1).all()  # self loops
    g2 = graphs.make_graph("flickr", seed=0, scale=0.05)
    assert torch.equal(g["indices"], g2["indices"]) and torch.equal(g["values"], g2["values"])
    r = graphs.make_rmat(10, 8.0, seed=1)
    assert r["v_num"] == 1024 and r["e_num"] == 8192
    b = graphs.algorithmic_bytes(232965, 114615892, 1907000, 32)
    assert abs(b["fwd"] / 1e9 - 19.53) < 0.05
