// topk_cbsr.cu -- MaxK nonlinearity: exact per-row top-k of d <= 256 features, emitted as
// CBSR (k values + k uint8 column ids, ascending column order, ties -> lowest column).
//
// Replaces `topk` (ref: kernels/maxk_kernel.cu:23-96) and its float facade
// (cuda_kernel_bindings.cpp:203-238).  The reference idea -- bisect for a threshold with
// warp votes, then compact with ballot/popc -- is kept, but done on order-preserving
// 32-bit keys of the fp32 values themselves (no uint8 quantisation), with a
// warp-uniform threshold (Q5/Q6), for any k and N (Q7) and with an exact tie rule (Q8):
//   * one warp per row, 8 elements per lane in registers (two 128-bit loads per lane);
//   * bracketing search for the k-th largest key: one REDUX-counted probe per step,
//     alternating value-space interpolation and bisection, ending as soon as a probe
//     separates exactly k elements (v1 walked the key bit by bit: ~20 probes per row and
//     785 warp instructions, issue-bound at 0.23 of the HBM peak);
//   * emission positions from one packed warp scan; optional dense x*mask output
//     (MaxK.forward, utils/models.py:48-52) from the same registers.
// Pure streaming kernel: 4*d bytes read + 5*k bytes written per row -> HBM-bound.
#include "common.cuh"

namespace maxk {

constexpr int TOPK_WPB = 8;

// order-preserving key of an fp32 bit pattern, branch-free:
// NaN ranks largest (torch rule), -0 == +0, otherwise unsigned order == float order
__device__ __forceinline__ unsigned f32_order_key(unsigned b)
{
    const unsigned a = b & 0x7fffffffu;
    b = (a == 0u) ? 0u : b;
    const unsigned key = b ^ ((unsigned)((int)b >> 31) | 0x80000000u);
    return (a > 0x7f800000u) ? 0xffffffffu : key;
}
// inverse (canonical NaN / -0 aside)
__device__ __forceinline__ float key_to_f32(unsigned key)
{
    return __uint_as_float((key & 0x80000000u) ? (key ^ 0x80000000u) : ~key);
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

__device__ __forceinline__ int warp_count_ge(const unsigned (&key)[8], unsigned m)
{
    int c = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) c += (key[i] >= m) ? 1 : 0;
    return __reduce_add_sync(0xffffffffu, c);
}

// Element i of a lane sits at column  (i < 4 ? 4*lane + i : 128 + 4*lane + (i-4)):
// group 0 of every lane precedes group 1 of every lane in column order.
template <bool U8>
__global__ void __launch_bounds__(TOPK_WPB * 32)
topk_cbsr_kernel(const void *__restrict__ x_, int64_t n, int d, int k, void *__restrict__ out_val_,
                 unsigned char *__restrict__ out_idx, float *__restrict__ dense_out)
{
    __shared__ unsigned s_val[TOPK_WPB][MAXK_MAX_DIM];
    __shared__ unsigned char s_idx[TOPK_WPB][MAXK_MAX_DIM];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t row = (int64_t)blockIdx.x * TOPK_WPB + warp;
    if (row >= n) return;  // warp-uniform; no CTA barrier below

    const int c0 = 4 * lane, c1 = 128 + 4 * lane;
    const bool ok0 = c0 < d, ok1 = c1 < d;

    unsigned key[8];
    unsigned raw[8];  // fp32 bit patterns or byte values (bit-copied to the output)
    if constexpr (U8) {
        const unsigned char *x = static_cast<const unsigned char *>(x_) + row * d;
        const unsigned wa = ok0 ? __ldcs(reinterpret_cast<const unsigned *>(x + c0)) : 0u;
        const unsigned wb = ok1 ? __ldcs(reinterpret_cast<const unsigned *>(x + c1)) : 0u;
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
            // valid keys are >= 0x007fffff (-inf); 0 marks "no element"
            key[i] = ((i < 4) ? ok0 : ok1) ? f32_order_key(raw[i]) : 0u;
        }
    }

    // ---- k-th largest key by bracketing search ------------------------------------------
    // Invariant: count(key >= L) = cL >= k  and  count(key >= H) = cH < k,  L < H.
    // A probe is one REDUX-counted compare of the 8 register keys.  Probes interpolate on
    // the counts -- in VALUE space for fp32, where activations are far closer to
    // piecewise-linear than in the exponent|mantissa key space -- and fall back to key-space
    // bisection whenever the same end point moved twice in a row, so smooth data needs a
    // handful of probes and adversarial data at most O(32) bisections.  The search ends on a
    // probe that separates exactly k elements (no tie at the threshold) or with H == L + 1,
    // i.e. L is the k-th largest key and ties at L are resolved by column order.
    unsigned lmin = 0xffffffffu, lmax = 0u;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        lmax = max(lmax, key[i]);
        lmin = min(lmin, key[i] == 0u ? 0xffffffffu : key[i]);
    }
    unsigned L = __reduce_min_sync(0xffffffffu, lmin);
    unsigned H = __reduce_max_sync(0xffffffffu, lmax);
    int cL = d;
    int cH = warp_count_ge(key, H);
    unsigned thr;        // selection threshold
    bool exact = false;  // count(key >= thr) == k
    if (cH >= k) {       // the maximum itself is (tied for) the k-th largest
        thr = H;
        exact = (cH == k);
    } else if (cL == k) {  // k == d: everything is kept
        thr = L;
        exact = true;
    } else {
        int stuck = 0;  // +n: L moved n times in a row, -n: H did
        while (true) {
            if (H - L == 1u) {  // cL > k > cH  ->  ties at L
                thr = L;
                break;
            }
            unsigned M = L + ((H - L) >> 1);
            if (stuck < 2 && stuck > -2) {
                const float frac = ((float)(cL - k) - 0.5f) * __frcp_rn((float)(cL - cH));
                if constexpr (U8) {  // byte keys are already linear
                    const unsigned g = L + (unsigned)((float)(H - L) * frac);
                    if (g > L && g < H) M = g;
                } else {
                    const float fl = key_to_f32(L), fh = key_to_f32(H);
                    const float fm = fmaf(fh - fl, frac, fl);
                    if (fm > fl && fm < fh) {  // false for inf/NaN end points
                        const unsigned g = f32_order_key(__float_as_uint(fm));
                        if (g > L && g < H) M = g;
                    }
                }
            } else {
                stuck = 0;
            }
            const int c = warp_count_ge(key, M);
            if (c == k) {
                thr = M;
                exact = true;
                break;
            }
            if (c > k) {
                L = M;
                cL = c;
                stuck = stuck > 0 ? stuck + 1 : 1;
            } else {
                H = M;
                cH = c;
                stuck = stuck < 0 ? stuck - 1 : -1;
            }
        }
    }

    unsigned sel = 0u;  // bit i: element i selected
    if (exact) {
#pragma unroll
        for (int i = 0; i < 8; ++i) sel |= (key[i] >= thr) ? (1u << i) : 0u;
    } else {
        // thr is the exact k-th largest key; ties at thr are taken in ascending column order
        unsigned gt = 0u, eq = 0u;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            gt |= (key[i] > thr) ? (1u << i) : 0u;
            eq |= (key[i] == thr) ? (1u << i) : 0u;
        }
        const int n_gt = __reduce_add_sync(0xffffffffu, __popc(gt));
        const int need = k - n_gt;
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

    // ---- emission: ascending column order, staged in shared memory so that the global
    // stores are coalesced (one 4k-byte value line + one k-byte selector line per row) ------
    if (out_idx != nullptr) {
        const unsigned n0 = __popc(sel & 0xfu), n1 = __popc(sel >> 4);
        const unsigned inc = warp_incl_scan(n0 | (n1 << 16), lane);
        const unsigned tot0 = __shfl_sync(0xffffffffu, inc, 31) & 0xffffu;
        int p0 = (int)(inc & 0xffffu) - (int)n0;
        int p1 = (int)tot0 + (int)(inc >> 16) - (int)n1;
        unsigned *sv = s_val[warp];
        unsigned char *si = s_idx[warp];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (sel & (1u << i)) {
                sv[p0] = raw[i];
                si[p0] = (unsigned char)(c0 + i);
                ++p0;
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (sel & (16u << i)) {
                sv[p1] = raw[4 + i];
                si[p1] = (unsigned char)(c1 + i);
                ++p1;
            }
        }
        __syncwarp();
        unsigned char *oi = out_idx + row * k;
        for (int j = lane; j < k; j += 32) {
            if constexpr (U8)
                (static_cast<unsigned char *>(out_val_) + row * k)[j] = (unsigned char)sv[j];
            else
                (static_cast<unsigned *>(out_val_) + row * k)[j] = sv[j];
            oi[j] = si[j];
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
