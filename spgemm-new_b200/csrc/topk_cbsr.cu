// topk_cbsr.cu -- MaxK nonlinearity: exact per-row top-k of d <= 256 features, emitted as
// CBSR (k values + k uint8 column ids, ascending column order, ties -> lowest column).
//
// Replaces `topk` (ref: kernels/maxk_kernel.cu:23-96) and its float facade
// (cuda_kernel_bindings.cpp:203-238).  The reference idea -- bisect for a threshold with
// warp votes, then compact with ballot/popc -- is kept, but done on order-preserving
// 32-bit keys of the fp32 values themselves (no uint8 quantisation), with a
// warp-uniform threshold (Q5/Q6), for any k and N (Q7) and with an exact tie rule (Q8):
//   * one warp per row, 8 elements per lane in registers (two 128-bit loads per lane);
//   * MSB-first radix select over only the bit positions that actually vary in the row
//     (warp AND/OR reduction), one REDUX per bit, early exit when the count hits k;
//   * emission positions from one packed warp scan; optional dense x*mask output
//     (MaxK.forward, utils/models.py:48-52) from the same registers.
// Pure streaming kernel: 4*d bytes read + 5*k bytes written per row -> HBM-bound.
#include "common.cuh"

namespace maxk {

constexpr int TOPK_WPB = 8;

__device__ __forceinline__ unsigned f32_order_key(float f)
{
    unsigned b = __float_as_uint(f);
    if ((b & 0x7fffffffu) > 0x7f800000u) return 0xffffffffu;  // NaN ranks largest (torch rule)
    if (b == 0x80000000u) b = 0u;                              // -0 == +0
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

__device__ __forceinline__ unsigned warp_incl_scan(unsigned v, int lane)
{
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        unsigned n = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += n;
    }
    return v;
}

// Element i of a lane sits at column  (i < 4 ? 4*lane + i : 128 + 4*lane + (i-4)):
// group 0 of every lane precedes group 1 of every lane in column order.
template <bool U8>
__global__ void __launch_bounds__(TOPK_WPB * 32)
topk_cbsr_kernel(const void *__restrict__ x_, int64_t n, int d, int k, void *__restrict__ out_val_,
                 unsigned char *__restrict__ out_idx, float *__restrict__ dense_out)
{
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * TOPK_WPB + (threadIdx.x >> 5);
    if (row >= n) return;

    const int c0 = 4 * lane, c1 = 128 + 4 * lane;
    const bool ok0 = c0 < d, ok1 = c1 < d;

    unsigned key[8];
    unsigned raw[8];  // fp32 bit patterns or byte values (bit-copied to the output)
    if constexpr (U8) {
        const unsigned char *x = static_cast<const unsigned char *>(x_) + row * d;
        unsigned wa = ok0 ? __ldcs(reinterpret_cast<const unsigned *>(x + c0)) : 0u;
        unsigned wb = ok1 ? __ldcs(reinterpret_cast<const unsigned *>(x + c1)) : 0u;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            raw[i] = (wa >> (8 * i)) & 0xff;
            raw[4 + i] = (wb >> (8 * i)) & 0xff;
            key[i] = ok0 ? raw[i] + 1u : 0u;  // 0 is reserved for "no element"
            key[4 + i] = ok1 ? raw[4 + i] + 1u : 0u;
        }
    } else {
        const float *x = static_cast<const float *>(x_) + row * d;
        float4 a = make_float4(0, 0, 0, 0), b = make_float4(0, 0, 0, 0);
        if (ok0) a = ld_stream_f4(reinterpret_cast<const float4 *>(x + c0));
        if (ok1) b = ld_stream_f4(reinterpret_cast<const float4 *>(x + c1));
        const float f[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            raw[i] = __float_as_uint(f[i]);
            key[i] = ((i < 4) ? ok0 : ok1) ? f32_order_key(f[i]) : 0u;  // valid keys are >= 0x007fffff
        }
    }

    // ---- bit positions that vary among the row's valid keys ----------------------------
    unsigned land = 0xffffffffu, lor = 0u;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const bool ok = (i < 4) ? ok0 : ok1;
        if (ok) {
            land &= key[i];
            lor |= key[i];
        }
    }
    land = __reduce_and_sync(0xffffffffu, land);
    lor = __reduce_or_sync(0xffffffffu, lor);
    const unsigned common = land;        // bits set in every key
    unsigned vary = land ^ lor;          // bits that differ somewhere

    // ---- MSB-first radix select of the k-th largest key --------------------------------
    unsigned decided = 0u;  // accepted varying bits
    unsigned thr = 0u;      // selection threshold
    bool exact = false;     // count(key >= thr) == k
    while (vary) {
        const int b = 31 - __clz(vary);
        vary &= ~(1u << b);
        const unsigned himask = ~((1u << b) - 1u);
        const unsigned cand = (decided | common | (1u << b)) & himask;
        int cnt = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) cnt += (key[i] >= cand) ? 1 : 0;
        cnt = __reduce_add_sync(0xffffffffu, cnt);
        if (cnt >= k) {
            decided |= (1u << b);
            if (cnt == k) {
                thr = cand;
                exact = true;
                break;
            }
        }
    }
    unsigned sel = 0u;  // bit i: element i selected
    if (exact) {
#pragma unroll
        for (int i = 0; i < 8; ++i) sel |= (key[i] >= thr) ? (1u << i) : 0u;
    } else {
        thr = decided | common;  // exact k-th largest key; ties at thr are possible
        unsigned gt = 0u, eq = 0u;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            gt |= (key[i] > thr) ? (1u << i) : 0u;
            eq |= (key[i] == thr) ? (1u << i) : 0u;
        }
        const int n_gt = __reduce_add_sync(0xffffffffu, __popc(gt));
        const int need = k - n_gt;  // ties taken in ascending column order
        const unsigned e0 = __popc(eq & 0xfu), e1 = __popc(eq >> 4);
        const unsigned inc = warp_incl_scan(e0 | (e1 << 16), lane);
        const unsigned tot0 = __shfl_sync(0xffffffffu, inc, 31) & 0xffffu;
        int r0 = (int)(inc & 0xffffu) - (int)e0;
        int r1 = (int)tot0 + (int)(inc >> 16) - (int)e1;
        sel = gt;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (eq & (1u << i)) {
                if (r0 < need) sel |= (1u << i);
                ++r0;
            }
            if (eq & (16u << i)) {
                if (r1 < need) sel |= (16u << i);
                ++r1;
            }
        }
    }

    // ---- emission: ascending column order ----------------------------------------------
    if (out_idx != nullptr) {
        const unsigned n0 = __popc(sel & 0xfu), n1 = __popc(sel >> 4);
        const unsigned inc = warp_incl_scan(n0 | (n1 << 16), lane);
        const unsigned tot0 = __shfl_sync(0xffffffffu, inc, 31) & 0xffffu;
        int p0 = (int)(inc & 0xffffu) - (int)n0;
        int p1 = (int)tot0 + (int)(inc >> 16) - (int)n1;
        unsigned char *oi = out_idx + row * k;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (sel & (1u << i)) {
                const int p = (i < 4) ? p0++ : p1++;
                const int col = (i < 4) ? (c0 + i) : (c1 + i - 4);
                if constexpr (U8)
                    (static_cast<unsigned char *>(out_val_) + row * k)[p] = (unsigned char)raw[i];
                else
                    (static_cast<unsigned *>(out_val_) + row * k)[p] = raw[i];
                oi[p] = (unsigned char)col;
            }
        }
    }
    if constexpr (!U8) {
        if (dense_out != nullptr) {
            float *o = dense_out + row * d;
            float4 a, b;
            a.x = (sel & 1u) ? __uint_as_float(raw[0]) : 0.f;
            a.y = (sel & 2u) ? __uint_as_float(raw[1]) : 0.f;
            a.z = (sel & 4u) ? __uint_as_float(raw[2]) : 0.f;
            a.w = (sel & 8u) ? __uint_as_float(raw[3]) : 0.f;
            b.x = (sel & 16u) ? __uint_as_float(raw[4]) : 0.f;
            b.y = (sel & 32u) ? __uint_as_float(raw[5]) : 0.f;
            b.z = (sel & 64u) ? __uint_as_float(raw[6]) : 0.f;
            b.w = (sel & 128u) ? __uint_as_float(raw[7]) : 0.f;
            if (ok0) st_stream_f4(reinterpret_cast<float4 *>(o + c0), a);
            if (ok1) st_stream_f4(reinterpret_cast<float4 *>(o + c1), b);
        }
    }
}

}  // namespace maxk

using namespace maxk;

extern "C" int maxk_topk_cbsr_f32(const float *x, int64_t n, int32_t d, int32_t k, float *sp_data,
                                  uint8_t *sp_index, float *dense_out, void *stream)
{
    MAXK_REQUIRE(n >= 0, "maxk_topk_cbsr_f32: negative n");
    MAXK_REQUIRE(d > 0 && d <= MAXK_MAX_DIM && d % 4 == 0,
                 "maxk_topk_cbsr_f32: d=%d must be a multiple of 4 in [4,256]", d);
    MAXK_REQUIRE(k >= 1 && k <= d, "maxk_topk_cbsr_f32: Invalid k value (k=%d, d=%d)", k, d);
    if (n == 0) return MAXK_OK;
    MAXK_REQUIRE((sp_data == nullptr) == (sp_index == nullptr),
                 "maxk_topk_cbsr_f32: sp_data and sp_index must both be given or both NULL");
    MAXK_REQUIRE(sp_index != nullptr || dense_out != nullptr, "maxk_topk_cbsr_f32: no output");
    MAXK_REQUIRE(x != nullptr, "maxk_topk_cbsr_f32: x is NULL");
    MAXK_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 &&
                     (reinterpret_cast<uintptr_t>(dense_out) & 15) == 0,
                 "maxk_topk_cbsr_f32: x and dense_out must be 16-byte aligned");
    const int64_t grid = ceil_div<int64_t>(n, TOPK_WPB);
    MAXK_REQUIRE(grid <= 0x7fffffffLL, "maxk_topk_cbsr_f32: n too large");
    topk_cbsr_kernel<false><<<(unsigned)grid, TOPK_WPB * 32, 0, static_cast<cudaStream_t>(stream)>>>(
        x, n, d, k, sp_data, sp_index, dense_out);
    MAXK_LAUNCH_CHECK("topk_cbsr_kernel<f32>");
    return MAXK_OK;
}

extern "C" int maxk_topk_cbsr_u8(const uint8_t *x, int64_t n, int32_t d, int32_t k, uint8_t *value,
                                 uint8_t *index, void *stream)
{
    MAXK_REQUIRE(n >= 0, "maxk_topk_cbsr_u8: negative n");
    MAXK_REQUIRE(d > 0 && d <= MAXK_MAX_DIM && d % 4 == 0,
                 "maxk_topk_cbsr_u8: d=%d must be a multiple of 4 in [4,256]", d);
    MAXK_REQUIRE(k >= 1 && k <= d, "maxk_topk_cbsr_u8: Invalid k value (k=%d, d=%d)", k, d);
    if (n == 0) return MAXK_OK;
    MAXK_REQUIRE(value != nullptr && index != nullptr, "maxk_topk_cbsr_u8: NULL output");
    MAXK_REQUIRE(x != nullptr && (reinterpret_cast<uintptr_t>(x) & 3) == 0,
                 "maxk_topk_cbsr_u8: x must be non-NULL and 4-byte aligned");
    const int64_t grid = ceil_div<int64_t>(n, TOPK_WPB);
    MAXK_REQUIRE(grid <= 0x7fffffffLL, "maxk_topk_cbsr_u8: n too large");
    topk_cbsr_kernel<true><<<(unsigned)grid, TOPK_WPB * 32, 0, static_cast<cudaStream_t>(stream)>>>(
        x, n, d, k, value, index, nullptr);
    MAXK_LAUNCH_CHECK("topk_cbsr_kernel<u8>");
    return MAXK_OK;
}
