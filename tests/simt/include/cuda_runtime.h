// tests/simt/include/cuda_runtime.h -- TEST INFRASTRUCTURE, never part of the product.
//
// A host stand-in for the CUDA toolkit header, used ONLY to build tests/simt/_build/libmaxk_simt.so:
// the kernel sources of spgemm-new_b200/csrc are compiled unmodified (after tests/simt/translate.py
// has rewritten the <<<...>>> launches, the `extern __shared__` declarations and the inline PTX)
// with g++ and run under a small SIMT interpreter (simt_runtime.cpp): one fibre per CUDA thread,
// warp collectives and barriers as rendezvous points between fibres, CTAs one after another.
//
// Purpose: check kernel LOGIC (indexing, masks, barriers, tails, ragged and empty inputs) and
// memory safety (guard pages around every buffer, tests/simt_lib.py) in a container without a
// GPU, and let new kernel variants be debugged before GPU minutes are spent on them.  It says
// nothing about performance, memory ordering between CTAs, or what the hardware does with a
// race.  The product library (lib/libmaxk_b200.so) never includes or links any of this, and
// nothing outside tests/ loads libmaxk_simt.so.
#pragma once
#ifndef MAXK_SIMT
#error "this header is the SIMT test shim; product builds use the CUDA toolkit's cuda_runtime.h"
#endif

#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <functional>

// ---- qualifiers -----------------------------------------------------------------------------
#define __global__
#define __device__
#define __host__
#define __constant__ static
#define __forceinline__ inline __attribute__((always_inline))
#define SIMT_NOINLINE __attribute__((noinline))
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))
#define __shared__ static thread_local  // one CTA at a time per host thread: one copy per kernel and thread

// ---- vector types ---------------------------------------------------------------------------
struct uint3 { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    constexpr dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct __attribute__((aligned(8))) int2 { int x, y; };
struct __attribute__((aligned(8))) uint2 { unsigned x, y; };
struct __attribute__((aligned(8))) float2 { float x, y; };
struct __attribute__((aligned(16))) int4 { int x, y, z, w; };
struct __attribute__((aligned(16))) uint4 { unsigned x, y, z, w; };
struct __attribute__((aligned(16))) float4 { float x, y, z, w; };
inline int2 make_int2(int x, int y) { return int2{x, y}; }
inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
inline float2 make_float2(float x, float y) { return float2{x, y}; }
inline int4 make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }
inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }

// ---- runtime API (host memory stands in for device memory) -------------------------------------
typedef int cudaError_t;
enum : int {
    cudaSuccess = 0,
    cudaErrorInvalidValue = 1,
    cudaErrorMemoryAllocation = 2,
    cudaErrorInvalidConfiguration = 9,
    cudaErrorLaunchFailure = 719,
    cudaErrorNotSupported = 801,
};
typedef struct simt_stream_st *cudaStream_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
struct cudaIpcMemHandle_t { char reserved[64]; };
enum : unsigned { cudaIpcMemLazyEnablePeerAccess = 1 };

namespace simt {

struct LaunchCfg {
    dim3 grid, block;
    size_t smem;
    cudaStream_t stream;
    LaunchCfg(dim3 g, dim3 b, size_t s = 0, cudaStream_t st = nullptr) : grid(g), block(b), smem(s), stream(st) {}
};

struct Thread {  // what a fibre knows about itself
    uint3 tid, bid;
    dim3 bdim, gdim;
    int lane, warp;
};
extern thread_local Thread *self_;  // the running fibre of this host thread
inline Thread &self() { return *self_; }

void launch(const void *kernel, const LaunchCfg &cfg, const std::function<void()> &body);
void *dynamic_smem();
void set_func_max_dynamic_smem(const void *kernel, int bytes);

// rendezvous points
void syncthreads();
void syncwarp(unsigned mask);
enum Op { OP_SHFL, OP_SHFL_UP, OP_SHFL_DOWN, OP_SHFL_XOR, OP_ADD, OP_MIN, OP_MAX, OP_UMIN, OP_UMAX, OP_AND, OP_OR, OP_BALLOT };
uint64_t collective(unsigned mask, Op op, uint64_t value, int arg, int width);

// cp.async (LDGSTS): the copy is performed as LATE as the PTX model allows -- when the issuing
// thread waits for its group (or returns) -- so a missing wait or barrier reads stale shared memory
void cp_async(void *smem_dst, const void *gmem_src, int bytes);
void cp_async_commit();
void cp_async_wait(int all_but_newest);

// 32-bit shared-memory addresses (what `cvta.to.shared` gives the PTX of the TMA path): a token is
// the low 32 bits of the host address; the pages around the object are remembered per host thread
// so that token arithmetic (base + offset) resolves back to host addresses
size_t smem_token(const void *p);
void *smem_resolve(size_t token);
// mbarrier + cp.async.bulk (TMA unit), as late as PTX allows: a bulk load lands when its mbarrier
// is waited for, a bulk store when its bulk group is waited for (or the thread returns)
void mbar_init(size_t mbar, int count);
void mbar_expect_tx(size_t mbar, unsigned bytes);
void mbar_wait(size_t mbar, unsigned parity);
void bulk_load(size_t smem_dst, const void *gmem_src, unsigned bytes, size_t mbar);
void bulk_store(void *gmem_dst, size_t smem_src, unsigned bytes);
void bulk_wait();

void yield_spin();  // a polite spin-wait iteration (__nanosleep)
long long clock_ticks();
[[noreturn]] void trap(const char *why);
[[noreturn]] void unsupported_asm(const char *ptx);

template <typename T>
inline uint64_t to_bits(T v)
{
    static_assert(sizeof(T) <= 8, "collectives move at most 64 bits");
    uint64_t b = 0;
    memcpy(&b, &v, sizeof(T));
    return b;
}
template <typename T>
inline T from_bits(uint64_t b)
{
    T v;
    memcpy(&v, &b, sizeof(T));
    return v;
}

}  // namespace simt

#define threadIdx (::simt::self().tid)
#define blockIdx (::simt::self().bid)
#define blockDim (::simt::self().bdim)
#define gridDim (::simt::self().gdim)
constexpr int warpSize = 32;

cudaError_t cudaMalloc(void **p, size_t bytes);
cudaError_t cudaFree(void *p);
cudaError_t cudaMemset(void *p, int v, size_t bytes);
cudaError_t cudaMemsetAsync(void *p, int v, size_t bytes, cudaStream_t st = nullptr);
cudaError_t cudaMemcpyAsync(void *dst, const void *src, size_t bytes, cudaMemcpyKind kind, cudaStream_t st = nullptr);
cudaError_t cudaMemcpy(void *dst, const void *src, size_t bytes, cudaMemcpyKind kind);
cudaError_t cudaStreamSynchronize(cudaStream_t st);
cudaError_t cudaDeviceSynchronize();
cudaError_t cudaGetLastError();
const char *cudaGetErrorString(cudaError_t e);
const char *cudaGetErrorName(cudaError_t e);
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t *h, void *p);
cudaError_t cudaIpcOpenMemHandle(void **p, cudaIpcMemHandle_t h, unsigned flags);
cudaError_t cudaIpcCloseMemHandle(void *p);
inline cudaError_t cudaGetDevice(int *dev) { *dev = 0; return cudaSuccess; }
template <typename F>
inline cudaError_t cudaFuncSetAttribute(F *kernel, cudaFuncAttribute attr, int value)
{
    if (attr != cudaFuncAttributeMaxDynamicSharedMemorySize) return cudaErrorInvalidValue;
    if (value > 227 * 1024) return cudaErrorInvalidValue;  // sm_100: 227 KB per CTA
    ::simt::set_func_max_dynamic_smem(reinterpret_cast<const void *>(kernel), value);
    return cudaSuccess;
}

// ---- device intrinsics ----------------------------------------------------------------------
using std::max;
using std::min;
inline long long max(long long a, int b) { return a > b ? a : (long long)b; }
inline long long min(long long a, int b) { return a < b ? a : (long long)b; }
inline long long min(int a, long long b) { return a < b ? (long long)a : b; }
inline long min(long a, long long b) { return a < b ? a : (long)b; }
inline long min(long long a, long b) { return a < b ? (long)a : b; }
inline unsigned max(unsigned a, int b) { return a > (unsigned)b ? a : (unsigned)b; }
inline unsigned min(unsigned a, int b) { return a < (unsigned)b ? a : (unsigned)b; }

template <typename T> inline T __ldg(const T *p) { return *p; }
template <typename T> inline T __ldcs(const T *p) { return *p; }
template <typename T> inline T __ldcg(const T *p) { return *p; }
template <typename T> inline T __ldca(const T *p) { return *p; }
template <typename T> inline void __stcs(T *p, T v) { *p = v; }
template <typename T> inline void __stcg(T *p, T v) { *p = v; }
template <typename T> inline void __stwt(T *p, T v) { *p = v; }

inline void __syncthreads() { ::simt::syncthreads(); }
inline void __syncwarp(unsigned mask = 0xffffffffu) { ::simt::syncwarp(mask); }
inline void __threadfence() {}
inline void __threadfence_block() {}
inline void __threadfence_system() {}
inline void __nanosleep(unsigned) { ::simt::yield_spin(); }
inline long long clock64() { return ::simt::clock_ticks(); }
inline void __trap() { ::simt::trap("__trap()"); }

template <typename T> inline T __shfl_sync(unsigned m, T v, int src, int width = 32)
{ return ::simt::from_bits<T>(::simt::collective(m, ::simt::OP_SHFL, ::simt::to_bits(v), src, width)); }
template <typename T> inline T __shfl_up_sync(unsigned m, T v, unsigned delta, int width = 32)
{ return ::simt::from_bits<T>(::simt::collective(m, ::simt::OP_SHFL_UP, ::simt::to_bits(v), (int)delta, width)); }
template <typename T> inline T __shfl_down_sync(unsigned m, T v, unsigned delta, int width = 32)
{ return ::simt::from_bits<T>(::simt::collective(m, ::simt::OP_SHFL_DOWN, ::simt::to_bits(v), (int)delta, width)); }
template <typename T> inline T __shfl_xor_sync(unsigned m, T v, int lanemask, int width = 32)
{ return ::simt::from_bits<T>(::simt::collective(m, ::simt::OP_SHFL_XOR, ::simt::to_bits(v), lanemask, width)); }
inline unsigned __ballot_sync(unsigned m, int pred) { return (unsigned)::simt::collective(m, ::simt::OP_BALLOT, pred ? 1 : 0, 0, 32); }
inline int __any_sync(unsigned m, int pred) { return (__ballot_sync(m, pred) & m) != 0; }
inline int __all_sync(unsigned m, int pred) { return (__ballot_sync(m, !pred) & m) == 0; }
inline unsigned __activemask() { return 0xffffffffu; }  // only meaningful for converged warps
inline int __reduce_add_sync(unsigned m, int v) { return (int)(uint32_t)::simt::collective(m, ::simt::OP_ADD, (uint32_t)v, 0, 32); }
inline unsigned __reduce_add_sync(unsigned m, unsigned v) { return (unsigned)::simt::collective(m, ::simt::OP_ADD, v, 0, 32); }
inline int __reduce_min_sync(unsigned m, int v) { return (int)(uint32_t)::simt::collective(m, ::simt::OP_MIN, (uint32_t)v, 0, 32); }
inline int __reduce_max_sync(unsigned m, int v) { return (int)(uint32_t)::simt::collective(m, ::simt::OP_MAX, (uint32_t)v, 0, 32); }
inline unsigned __reduce_min_sync(unsigned m, unsigned v) { return (unsigned)::simt::collective(m, ::simt::OP_UMIN, v, 0, 32); }
inline unsigned __reduce_max_sync(unsigned m, unsigned v) { return (unsigned)::simt::collective(m, ::simt::OP_UMAX, v, 0, 32); }
inline unsigned __reduce_and_sync(unsigned m, unsigned v) { return (unsigned)::simt::collective(m, ::simt::OP_AND, v, 0, 32); }
inline unsigned __reduce_or_sync(unsigned m, unsigned v) { return (unsigned)::simt::collective(m, ::simt::OP_OR, v, 0, 32); }

// redux.sync.{min,max}.f32 (sm_100a): NaN operands are skipped, an all-NaN warp yields NaN.  Done on order-preserving
// keys through the integer collective (key 0 / 0xffffffff = "no value").
inline float simt_redux_f32(float v, bool want_max)
{
    uint32_t b;
    __builtin_memcpy(&b, &v, 4);
    const bool nan = (b & 0x7fffffffu) > 0x7f800000u;
    uint32_t key = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
    if (nan) key = want_max ? 0u : 0xffffffffu;
    const uint32_t r = (uint32_t)::simt::collective(0xffffffffu, want_max ? ::simt::OP_UMAX : ::simt::OP_UMIN, key, 0, 32);
    if (r == (want_max ? 0u : 0xffffffffu)) return __builtin_nanf("");
    const uint32_t rb = (r & 0x80000000u) ? (r & 0x7fffffffu) : ~r;
    float out;
    __builtin_memcpy(&out, &rb, 4);
    return out;
}

inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
inline int __clz(int v) { return v == 0 ? 32 : __builtin_clz((unsigned)v); }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline unsigned __brev(unsigned v)
{
    unsigned r = 0;
    for (int i = 0; i < 32; ++i) r |= ((v >> i) & 1u) << (31 - i);
    return r;
}
inline float __uint_as_float(unsigned v) { return ::simt::from_bits<float>(v); }
inline float __int_as_float(int v) { return ::simt::from_bits<float>((uint32_t)v); }
inline unsigned __float_as_uint(float v) { return (unsigned)::simt::to_bits(v); }
inline int __float_as_int(float v) { return (int)(uint32_t)::simt::to_bits(v); }
inline float __frcp_rn(float v) { return 1.0f / v; }
inline float rsqrtf(float v) { return 1.0f / sqrtf(v); }
inline float __fdividef(float a, float b) { return a / b; }
inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
inline size_t __cvta_generic_to_shared(const void *p) { return ::simt::smem_token(p); }

// atomics: fibres never preempt each other, so plain read-modify-write is atomic here
template <typename T> inline T atomicAdd(T *p, T v) { T o = *p; *p = o + v; return o; }
template <typename T> inline T atomicSub(T *p, T v) { T o = *p; *p = o - v; return o; }
template <typename T> inline T atomicMax(T *p, T v) { T o = *p; *p = o > v ? o : v; return o; }
template <typename T> inline T atomicMin(T *p, T v) { T o = *p; *p = o < v ? o : v; return o; }
template <typename T> inline T atomicExch(T *p, T v) { T o = *p; *p = v; return o; }
template <typename T> inline T atomicOr(T *p, T v) { T o = *p; *p = o | v; return o; }
template <typename T> inline T atomicAnd(T *p, T v) { T o = *p; *p = o & v; return o; }
template <typename T> inline T atomicCAS(T *p, T cmp, T v) { T o = *p; if (o == cmp) *p = v; return o; }
inline float4 atomicAdd(float4 *p, float4 v)
{
    float4 o = *p;
    p->x += v.x; p->y += v.y; p->z += v.z; p->w += v.w;
    return o;
}
