"""INTEGRATION.md section 1, exercised: the reference's own binding file (cuda_kernel_bindings.cpp, with
kernels/spmm_cusparse.cu) compiled UNMODIFIED and linked against libmaxk_b200.so in place of the reference's
kernels/*.cu + cuda_kernel_wrappers.cu (oracle/Makefile `zeroedit`; built in the build container, where
/root/reference exists, and carried to the GPU box under oracle/_ref/).  The module it yields is the reference's
`maxk_cuda_kernels` torch extension -- same pybind signatures, TORCH_CHECKs and allocation behaviour, because it IS
the reference's host code -- running on this repository's kernels through the three reference-named launch shims."""
import numpy as np
import pytest
import torch

from oracle import ref_kernels as R

needs_ext = pytest.mark.skipif(not R.zeroedit_available(), reason="oracle/_ref/zeroedit not built (no reference tree)")


@needs_ext
def test_reference_binding_file_links_against_the_library_and_imports():
    m = R.load_zeroedit()
    for name in ("spmm_maxk_forward", "spmm_maxk_backward", "cuda_topk_maxk", "cuda_topk_maxk_float",
                 "prepare_cbsr_format_maxk", "load_warp4_metadata", "generate_sparse_selector", "benchmark_spmm_maxk",
                 "validate_spmm_maxk", "cusparse_spmm", "CudaTimer"):
        assert hasattr(m, name), name
    with pytest.raises(RuntimeError, match="CUDA"):   # the reference's own TORCH_CHECK (cuda_kernel_bindings.cpp:52)
        z = torch.zeros(4, dtype=torch.int32)
        m.spmm_maxk_forward(z, z, torch.zeros(4), torch.zeros(1, 4), torch.zeros(1, 4, dtype=torch.uint8), 1, 4)


@needs_ext
@pytest.mark.gpu
@pytest.mark.parametrize("k", [32, 64, 16])
def test_reference_extension_runs_on_our_kernels(k):
    from helpers import Problem, assert_close
    from maxk_b200 import _cabi
    m = R.load_zeroedit()
    p = Problem(n=700, avg=40.0, k=k, seed=11 + k, empty_frac=0.05)
    g = p.cuda()
    nw = g["warp4"].numel() // 4
    n0 = _cabi.launch_count()
    y = m.spmm_maxk_forward(g["warp4"], g["indices"], g["values"], g["sp_data"], g["sp_index"], nw, k)
    assert tuple(y.shape) == (p.n, 256)                    # FULL_DIM hard-coded by the reference (bindings.cpp:70)
    assert_close(y, p.fwd_oracle(), "spmm_maxk_forward through the reference's bindings")
    dxc = m.spmm_maxk_backward(g["warp4"], g["indices"], g["values"], g["dy"], g["sp_index"], nw, k)
    assert_close(dxc, p.bwd_oracle(), "spmm_maxk_backward through the reference's bindings")
    assert _cabi.launch_count() - n0 >= 2, "the reference's host code must have launched OUR kernels"
    assert m.validate_spmm_maxk(g["warp4"], g["indices"], g["values"], g["sp_data"], g["sp_index"],
                                torch.from_numpy(p.fwd_oracle().astype(np.float32)).cuda(), nw, k, 1e-3)
    times = m.benchmark_spmm_maxk(g["warp4"], g["indices"], g["values"], g["sp_data"], g["sp_index"], nw, k, 2)
    assert len(times) == 2 and all(t > 0 for t in times)


@needs_ext
@pytest.mark.gpu
def test_reference_extension_topk_and_cusparse():
    from oracle import oracle as O
    m = R.load_zeroedit()
    rng = np.random.default_rng(3)
    x = rng.integers(0, 256, size=(333, 256), dtype=np.uint8)
    v, i = m.cuda_topk_maxk(torch.from_numpy(x).cuda(), 32)      # the reference's uint8 entry point, our exact TopK
    want_v, want_i = O.topk_cbsr_u8(x, 32) if hasattr(O, "topk_cbsr_u8") else (None, None)
    vi = i.cpu().numpy().astype(np.int64)
    picked = np.take_along_axis(x, vi, axis=1)
    np.testing.assert_array_equal(v.cpu().numpy(), picked)                       # values are the picked bytes
    assert (np.diff(vi, axis=1) > 0).all()                                       # ascending distinct columns
    kth = np.sort(x, axis=1)[:, -32]
    assert (picked.min(axis=1) == kth).all()                                     # and they are the 32 largest
    if want_i is not None:
        np.testing.assert_array_equal(vi, want_i.astype(np.int64))
    # the cuSPARSE comparator the reference binds (kernels/spmm_cusparse.cu) against a dense product
    from helpers import Problem
    p = Problem(n=300, avg=12.0, k=8, seed=4)
    g = p.cuda()
    feats = torch.rand(p.n, 256, device="cuda")
    y = m.cusparse_spmm(g["indptr"], g["indices"], g["values"], feats, False)
    a = torch.sparse_csr_tensor(g["indptr"].long(), g["indices"].long(), g["values"], size=(p.n, p.n)).to_dense()
    torch.testing.assert_close(y, a @ feats, rtol=1e-4, atol=1e-4)
