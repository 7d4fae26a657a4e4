// api.cu -- error plumbing, launch accounting, tuning knobs and the reference-named
// launch shims of libmaxk_b200 (ref: cuda_kernel_wrappers.cu:36-93).
#include <atomic>
#include <cstdarg>
#include <cstring>
#include <mutex>
#include <set>
#include <utility>

#include "common.cuh"

namespace maxk {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

extern int g_fwd_tasks_per_warp;
extern int g_bwd_tasks_per_warp;
extern int g_fwd_copies;
extern int g_fwd_depth;

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char *what)
{
    set_error("CUDA error in %s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
    return MAXK_ERR_CUDA;
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// cudaFuncAttributeMaxDynamicSharedMemorySize is a property of (kernel, device): a process that drives
// several GPUs has to raise it on each of them, once.
int opt_in_dynamic_smem(const void *kernel, const char *what)
{
    static std::mutex mu;
    static std::set<std::pair<const void *, int>> done;
    int dev = 0;
    MAXK_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(mu);
    if (done.count({kernel, dev})) return MAXK_OK;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return cuda_fail(e, what);
    done.insert({kernel, dev});
    return MAXK_OK;
}

}  // namespace maxk

using namespace maxk;

extern "C" int maxk_abi_version(void) { return MAXK_ABI_VERSION; }
extern "C" const char *maxk_last_error(void) { return g_err; }
extern "C" int64_t maxk_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

// Bench/test hook: run lengths (tasks per warp) of the forward / backward kernels.
// 0 leaves a knob unchanged, a negative value returns it to automatic.
// Not part of the reference surface.
extern "C" int maxk_set_tuning(int fwd_tasks_per_warp, int bwd_tasks_per_warp)
{
    if (fwd_tasks_per_warp != 0) g_fwd_tasks_per_warp = fwd_tasks_per_warp < 0 ? 0 : fwd_tasks_per_warp;
    if (bwd_tasks_per_warp != 0) g_bwd_tasks_per_warp = bwd_tasks_per_warp < 0 ? 0 : bwd_tasks_per_warp;
    return MAXK_OK;
}

// Bench/test hook: launch shape of the forward kernel (see spgemm_forward.cu; 0 = automatic).
extern "C" int maxk_set_forward_depth(int depth)
{
    g_fwd_depth = depth < 0 ? 0 : depth;
    return MAXK_OK;
}

// Bench/test hook: number of accumulator copies of the forward kernel (0 = automatic).
extern "C" int maxk_set_forward_copies(int copies)
{
    g_fwd_copies = copies < 0 ? 0 : copies;
    return MAXK_OK;
}

// ---- reference-named shims -----------------------------------------------------------------
// Argument lists are those of cuda_kernel_wrappers.cu:38-91 (dim3 by value is layout-
// compatible with maxk_dim3).  Launch geometry arguments are ignored.
static void shim_report(int rc, const char *who)
{
    if (rc != MAXK_OK) printf("%s failed: %s\n", who, maxk_last_error());
}

extern "C" void spmm_kernel_opt2_sparse_v3_wrapper(const int *warp4, const int *idx, const float *val,
                                                   const float *vin_data, const uint8_t *vin_selector,
                                                   float *vout, const int num_v, const int num_e,
                                                   const int feat_in, const int dim_sparse,
                                                   const int num_warps, maxk_dim3, maxk_dim3, int)
{
    (void)num_e;
    shim_report(maxk_spgemm_forward(warp4, num_warps, idx, val, vin_data, vin_selector, dim_sparse,
                                    vout, num_v, feat_in, nullptr),
                "CUDA kernel launch");
}

extern "C" void spmm_kernel_opt2_sparse_backward_v3_wrapper(
    const int *warp4, const int *idx, const float *val, const float *vin_data,
    const uint8_t *vin_selector, float *vout, const int num_v, const int num_e, const int feat_in,
    const int dim_sparse, const int num_warps, maxk_dim3, maxk_dim3, int)
{
    (void)num_e;
    shim_report(maxk_sspmm_backward(warp4, num_warps, idx, val, vin_data, vin_selector, dim_sparse,
                                    feat_in, vout, num_v, /*accumulate=*/0, nullptr),
                "CUDA kernel launch");
}

extern "C" void topk_kernel_wrapper(uint8_t *data, uint8_t *value, uint8_t *index, unsigned int k,
                                    int N, int dim_origin, maxk_dim3, maxk_dim3, int)
{
    shim_report(maxk_topk_cbsr_u8(data, N, dim_origin, (int)k, value, index, nullptr),
                "TopK kernel launch");
}
