// ref_harness.cu -- thin C entry points around the REFERENCE CUDA kernels, which the
// Makefile compiles UNMODIFIED from /root/reference/kernels (spmm_maxk.cu,
// spmm_maxk_backward.cu, maxk_kernel.cu) together with the reference's own launch shims
// (cuda_kernel_wrappers.cu).  TEST INFRASTRUCTURE ONLY (see maxk_oracle.c): used by the
// GPU parity tests as the comparator and by bench.py as the "reference kernels on the same
// B200" bar.  Launch geometry is the reference's: grid = ceil(W/12), block = 384,
// smem = 12*feat_in*4 (spmm_maxk.cu:108-125, cuda_kernel_bindings.cpp:77-85); TopK:
// grid = ceil(N/16), block = 512, smem = 16*(dim+2k+1) (cuda_kernel_bindings.cpp:181-187).
// Outputs must be zeroed by the caller (the reference kernels accumulate).
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

// spmm_maxk.cu:9 declares these extern; main.cu (not compiled here) defines them.
std::string base_dir = "", graph = "";

extern "C" {
void spmm_kernel_opt2_sparse_v3_wrapper(const int *warp4, const int *idx, const float *val,
                                        const float *vin_data, const uint8_t *vin_selector,
                                        float *vout, const int num_v, const int num_e,
                                        const int feat_in, const int dim_sparse,
                                        const int num_warps, dim3 grid, dim3 block, int shared_size);
void spmm_kernel_opt2_sparse_backward_v3_wrapper(const int *warp4, const int *idx, const float *val,
                                                 const float *vin_data, const uint8_t *vin_selector,
                                                 float *vout, const int num_v, const int num_e,
                                                 const int feat_in, const int dim_sparse,
                                                 const int num_warps, dim3 grid, dim3 block,
                                                 int shared_size);
void topk_kernel_wrapper(uint8_t *data, uint8_t *value, uint8_t *index, unsigned int k, int N,
                         int dim_origin, dim3 grid, dim3 block, int shared_size);

#define REF_API __attribute__((visibility("default")))

REF_API int ref_spmm_forward(const int *warp4, const int *idx, const float *val, const float *sp_data,
                             const uint8_t *sel, float *vout, int num_v, int num_e, int feat_in,
                             int dim_sparse, int num_warps)
{
    if (num_warps <= 0) return 0;
    dim3 grid((num_warps + 11) / 12), block(12 * 32);
    spmm_kernel_opt2_sparse_v3_wrapper(warp4, idx, val, sp_data, sel, vout, num_v, num_e, feat_in,
                                       dim_sparse, num_warps, grid, block,
                                       12 * feat_in * (int)sizeof(float));
    return (int)cudaGetLastError();
}

REF_API int ref_spmm_backward(const int *warp4, const int *idx, const float *val, const float *dy,
                              const uint8_t *sel, float *vout, int num_v, int num_e, int feat_in,
                              int dim_sparse, int num_warps)
{
    if (num_warps <= 0) return 0;
    dim3 grid((num_warps + 11) / 12), block(12 * 32);
    spmm_kernel_opt2_sparse_backward_v3_wrapper(warp4, idx, val, dy, sel, vout, num_v, num_e,
                                                feat_in, dim_sparse, num_warps, grid, block,
                                                12 * feat_in * (int)sizeof(float));
    return (int)cudaGetLastError();
}

REF_API int ref_topk_u8(uint8_t *data, uint8_t *value, uint8_t *index, unsigned k, int N,
                        int dim_origin)
{
    dim3 grid((N + 15) / 16), block(16 * 32);
    topk_kernel_wrapper(data, value, index, k, N, dim_origin, grid, block,
                        16 * (dim_origin + 2 * (int)k + 1));
    return (int)cudaGetLastError();
}
}
