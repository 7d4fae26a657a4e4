"""Drop-in for the reference's direct_kernel_interface.py: class DirectMaxKKernels with the
same method names, arguments and return values (direct_kernel_interface.py:24-382), on
libmaxk_b200.so.

Deliberate differences:
  * no silent fallbacks: the torch.topk fallback on exception (l.74-79) is gone -- errors raise;
    `use_cuda_topk=False` still selects torch.topk explicitly, as in the reference;
  * load_warp4_metadata can build the work list on the GPU from `indptr` instead of needing
    the file generate_meta.py writes;
  * timing uses CUDA events on the current stream with the output buffer allocated once.
"""
from __future__ import annotations

import numpy as np
import torch

import maxk_cuda_kernels
from graph_loader import GraphDataLoader  # noqa: F401  (same import as the reference, l.12)

DIRECT_KERNELS_AVAILABLE = True  # import of maxk_cuda_kernels raises if the library is missing


class DirectMaxKKernels:
    """Direct interface to the MaxK-GNN CUDA kernels."""

    def __init__(self, graph_name=""):
        self.graph_name = graph_name
        self.warp4_metadata = None
        self.num_warps = 0

    # -- work list -----------------------------------------------------------------------
    def load_warp4_metadata(self, graph_name=None, num_warps=12, warp_max_nz=64, indptr=None):
        """Reference signature (l.35) + optional `indptr`: when given, the list is generated on
        the device (replacement for kernels/generate_meta.py); otherwise the .warp4 file is read.
        Returns True/False like the reference."""
        if graph_name is None:
            graph_name = self.graph_name
        try:
            if indptr is not None:
                self.warp4_metadata = maxk_cuda_kernels.generate_warp4_metadata(indptr, warp_max_nz)
            else:
                self.warp4_metadata = maxk_cuda_kernels.load_warp4_metadata(graph_name, num_warps,
                                                                            warp_max_nz)
        except RuntimeError as e:
            print(f"Failed to load warp4 metadata: {e}")
            return False
        self.num_warps = self.warp4_metadata.size(0) // 4
        return True

    # -- TopK ----------------------------------------------------------------------------
    def generate_maxk_sparse_data(self, input_features, dim_k, use_cuda_topk=True):
        """(sparse_data f32[N,k], sparse_selector u8[N,k])  (l.58-85)."""
        if use_cuda_topk:
            topk_values, topk_indices = maxk_cuda_kernels.cuda_topk_maxk_float(input_features, dim_k)
        else:
            topk_values, topk_indices = torch.topk(input_features, dim_k, dim=1)
        return topk_values.contiguous(), topk_indices.to(torch.uint8).contiguous()

    def _require_ready(self):
        if self.warp4_metadata is None:
            raise RuntimeError("Warp4 metadata not loaded. Call load_warp4_metadata() first")

    # -- forward -------------------------------------------------------------------------
    def run_forward_kernel(self, graph_data, input_features, dim_k, timing=True, use_cuda_topk=True):
        """-> (output f32[N,256], mean ms over 4 timed runs or 0.0)  (l.87-153)."""
        self._require_ready()
        sparse_data, sparse_selector = self.generate_maxk_sparse_data(input_features, dim_k, use_cuda_topk)
        avg_time = 0.0
        if timing:
            times = maxk_cuda_kernels.benchmark_spmm_maxk(
                self.warp4_metadata, graph_data['indices'], graph_data['values'], sparse_data,
                sparse_selector, self.num_warps, dim_k, num_runs=4)
            avg_time = float(np.mean(times))
        output = maxk_cuda_kernels.spmm_maxk_forward(
            self.warp4_metadata, graph_data['indices'], graph_data['values'], sparse_data,
            sparse_selector, self.num_warps, dim_k, feat_dim=input_features.shape[1],
            num_rows=self._num_rows(graph_data, input_features))
        return output, avg_time

    @staticmethod
    def _num_rows(graph_data, features):
        if 'indptr' in graph_data and graph_data['indptr'] is not None:
            return graph_data['indptr'].numel() - 1
        return features.shape[0]

    # -- backward ------------------------------------------------------------------------
    def run_backward_kernel(self, graph_data, grad_output, dim_k, timing=True, use_cuda_topk=True,
                            sparse_selector=None):
        """-> (grad_input f32[N,k], mean ms)  (l.155-213).  As in the reference the selector is
        derived from grad_output itself unless `sparse_selector` is passed."""
        self._require_ready()
        if sparse_selector is None:
            _, sparse_selector = self.generate_maxk_sparse_data(grad_output, dim_k, use_cuda_topk)
        run = lambda: maxk_cuda_kernels.spmm_maxk_backward(  # noqa: E731
            self.warp4_metadata, graph_data['indices'], graph_data['values'], grad_output,
            sparse_selector, self.num_warps, dim_k)
        avg_time = 0.0
        if timing:
            timer = maxk_cuda_kernels.CudaTimer()
            times = []
            for i in range(8):  # 4 warm-up + 4 timed (l.184-198)
                timer.start()
                run()
                t = timer.stop()
                if i >= 4:
                    times.append(t)
            avg_time = float(np.mean(times))
        return run(), avg_time

    # -- validation ----------------------------------------------------------------------
    def validate_against_cusparse(self, graph_data, input_features, dim_k, tolerance=0.001,
                                  use_cuda_topk=True):
        """Compare the SpGEMM output with a dense-X CSR SpMM on the scatter-to-dense twin of
        the SAME TopK (l.215-298): max |diff| at the positions where the input was non-zero."""
        self._require_ready()
        if use_cuda_topk:
            topk_values, topk_indices = maxk_cuda_kernels.cuda_topk_maxk_float(input_features, dim_k)
            topk_indices_int64 = topk_indices.long()
        else:
            topk_values, topk_indices_int64 = torch.topk(input_features, dim_k, dim=1)
        sparse_input = torch.zeros_like(input_features)
        sparse_input.scatter_(1, topk_indices_int64, topk_values)
        maxk_output = maxk_cuda_kernels.spmm_maxk_forward(
            self.warp4_metadata, graph_data['indices'], graph_data['values'],
            topk_values.contiguous(), topk_indices_int64.to(torch.uint8).contiguous(),
            self.num_warps, dim_k, feat_dim=input_features.shape[1],
            num_rows=self._num_rows(graph_data, input_features))
        cusparse_output = maxk_cuda_kernels.cusparse_spmm(
            graph_data['indptr'], graph_data['indices'], graph_data['values'], sparse_input,
            timing=False)
        if maxk_output.shape != cusparse_output.shape:
            print(f"Shape mismatch! MaxK: {maxk_output.shape}, cuSPARSE: {cusparse_output.shape}")
            return False
        mask = sparse_input != 0
        diff = (maxk_output - cusparse_output).abs()
        rel = diff[mask] if mask.shape == diff.shape else diff.flatten()
        max_error = rel.max().item() if rel.numel() > 0 else 0.0
        self.last_validation = {"max_error": max_error,
                                "max_error_all": diff.max().item() if diff.numel() else 0.0}
        return max_error < tolerance

    # -- benchmark -----------------------------------------------------------------------
    def benchmark_all_k_values(self, graph_data, dim_origin=256, k_values=[16, 32, 64], num_runs=4,
                               use_cuda_topk=True):
        """{k: {'forward_time': ms, 'backward_time': ms}}  (l.300-382), inputs ~ U[0,1) seed 123."""
        v_num = graph_data['indptr'].size(0) - 1
        torch.manual_seed(123)
        input_features = torch.rand(v_num, dim_origin, device='cuda', dtype=torch.float32)
        results = {}
        print("num graph dim_origin dim_k kernel time(ms)")
        for dim_k in k_values:
            if dim_k > dim_origin:
                continue
            _, tf = self.run_forward_kernel(graph_data, input_features, dim_k, timing=True,
                                            use_cuda_topk=use_cuda_topk)
            grad_output = torch.rand_like(input_features)
            _, tb = self.run_backward_kernel(graph_data, grad_output, dim_k, timing=True,
                                             use_cuda_topk=use_cuda_topk)
            results[dim_k] = {'forward_time': tf, 'backward_time': tb}
            print(f"1/1 {self.graph_name} {dim_origin} {dim_k} maxk {tf:.3f}")
            print(f"1/1 {self.graph_name} {dim_origin} {dim_k} maxk_backward {tb:.3f}")
        return results


def test_direct_kernels(shape="flickr", scale=0.1):
    """Smoke run on a synthetic graph (the reference's version needs shipped graph files)."""
    loader = GraphDataLoader()
    g = loader.to_cuda_tensors(loader.synthetic(shape, scale=scale))
    kernels = DirectMaxKKernels(shape)
    if not kernels.load_warp4_metadata(indptr=g['indptr']):
        return False
    x = torch.rand(g['v_num'], 256, device='cuda')
    ok = kernels.validate_against_cusparse(g, x, dim_k=32, use_cuda_topk=True)
    ok = kernels.validate_against_cusparse(g, x, dim_k=32, use_cuda_topk=False) and ok
    kernels.benchmark_all_k_values(g, k_values=[16, 32], num_runs=2)
    return ok


if __name__ == "__main__":
    print("validation:", test_direct_kernels())
