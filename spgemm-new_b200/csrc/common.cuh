// common.cuh -- shared device/host helpers for libmaxk_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "maxk_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libmaxk_b200 is written for sm_100a (B200) only"
#endif

namespace maxk {

constexpr int kWarp = 32;
constexpr int kNumSMs = 148;  // B200

// ---- host-side error plumbing (thread local) -------------------------------------------
void set_error(const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what);
void count_launch(int n = 1);
int opt_in_dynamic_smem(const void *kernel, const char *what);  // up to 200 KB, once per (kernel, device)

#define MAXK_REQUIRE(cond, ...)                    \
    do {                                           \
        if (!(cond)) {                             \
            ::maxk::set_error(__VA_ARGS__);        \
            return MAXK_ERR_INVALID_ARGUMENT;      \
        }                                          \
    } while (0)

#define MAXK_CUDA(expr)                                         \
    do {                                                        \
        cudaError_t _e = (expr);                                \
        if (_e != cudaSuccess) return ::maxk::cuda_fail(_e, #expr); \
    } while (0)

#define MAXK_LAUNCH_CHECK(name)                                   \
    do {                                                          \
        ::maxk::count_launch();                                   \
        cudaError_t _e = cudaGetLastError();                      \
        if (_e != cudaSuccess) return ::maxk::cuda_fail(_e, name); \
    } while (0)

// ---- cache-hinted memory access ----------------------------------------------------------
// Streams that are read exactly once (CSR indices/values, work list, dY): evict-first so
// they do not push the L2-resident CBSR (5k bytes/node) out of the 126 MB L2.
__device__ __forceinline__ int ld_stream_s32(const int *p) { return __ldcs(p); }
__device__ __forceinline__ float ld_stream_f32(const float *p) { return __ldcs(p); }
__device__ __forceinline__ int4 ld_stream_int4(const int4 *p) { return __ldcs(p); }
__device__ __forceinline__ float4 ld_stream_f4(const float4 *p) { return __ldcs(p); }

// Gathers that are re-read by many CTAs (CBSR rows): read-only path, normal allocation.
__device__ __forceinline__ float ld_gather_f32(const float *p) { return __ldg(p); }
__device__ __forceinline__ float4 ld_gather_f4(const float4 *p) { return __ldg(p); }
__device__ __forceinline__ unsigned ld_gather_u32(const unsigned *p) { return __ldg(p); }
__device__ __forceinline__ unsigned char ld_gather_u8(const unsigned char *p) { return __ldg(p); }

__device__ __forceinline__ void st_stream_f4(float4 *p, float4 v) { __stcs(p, v); }

// 128-bit vector reduction (sm_90+): one REDG.E.ADD.F32x4 instead of four scalar REDGs.
__device__ __forceinline__ void red_add_f4(float *addr, float a, float b, float c, float d)
{
    asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a),
                 "f"(b), "f"(c), "f"(d)
                 : "memory");
}
__device__ __forceinline__ void red_add_f1(float *addr, float a)
{
    asm volatile("red.relaxed.gpu.global.add.f32 [%0], %1;" ::"l"(addr), "f"(a) : "memory");
}

// Asynchronous 16-byte global -> shared copy (LDGSTS, L2 only: gathered rows are not re-read
// through the L1).  Loads in flight hold no registers.  Completion: cp_async_wait<N>() (all but
// the N most recent groups of the calling thread) followed by a barrier among the readers.
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"l"(__cvta_generic_to_shared(smem_dst)),
                 "l"(gmem_src)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait()
{
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// ---- bank plan (cbsr_plan.cu, planned forward): the two homes of an output column -------------------
// The accumulator row of the planned forward kernel has PLAN_HOMES = 2 copies of 256 words:
//   home 0 of column c: word c                       bank c mod 32
//   home 1 of column c: word 256 + (c mod 8)*32 + c/8   bank c / 8     (the 8 x 32 transpose of home 0)
// Two columns that share a bank in one home (same c mod 32) never share it in the other (their c/8
// differ), which is what gives the plan its choice.  (Round 2 measured 2, 4 and 8 homes with an exact
// matching: 3.09 / 2.93 / 3.46 ms on the Reddit shape at k = 32 against 3.99 ms without a plan -- two
// homes keep 4 CTAs per SM and halve the flush work, profiles/r02_variants.md.)
constexpr int PLAN_HOMES = 2;
__host__ __device__ constexpr int plan_s(int) { return 0; }
__host__ __device__ constexpr int plan_t(int) { return 0; }
__host__ __device__ __forceinline__ unsigned plan_home_word(unsigned c, int h)
{
    return h == 0 ? c : 256u + (c & 7u) * 32u + (c >> 3);
}
// the column of an accumulator word (inverse of plan_home_word for either home)
__host__ __device__ __forceinline__ unsigned plan_word_column(unsigned w)
{
    return w < 256u ? w : (((w & 31u) << 3) | ((w >> 5) & 7u));
}

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

template <typename T>
__host__ __device__ constexpr T ceil_div(T a, T b)
{
    return (a + b - 1) / b;
}

// Run length (tasks per warp) of the forward / backward kernels when the caller did not fix
// it: 16 on big graphs, shorter on small ones so that the grid still fills 148 SMs x 32
// warps several times over (measured on the Flickr shape: 0.078 ms at 2-4 vs 0.119 ms at 16).
inline int auto_tasks_per_warp(int64_t num_tasks)
{
    const int64_t t = num_tasks / (int64_t)(kNumSMs * 32 * 8);
    return (int)(t < 1 ? 1 : (t > 16 ? 16 : t));
}

}  // namespace maxk
