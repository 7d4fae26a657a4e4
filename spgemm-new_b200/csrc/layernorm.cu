// layernorm.cu -- the LayerNorm that closes every MaxK-GNN layer (`self.normlayers[i]`, utils/models.py:207-213,
// 252, 343, 390: nn.LayerNorm(hidden, elementwise_affine=True)), forward and backward, for rows of d <= 256 floats.
//
// Why an own kernel: on the products shape (2.45 M rows x 256) the library's forward takes 4.1 ms and its backward
// 7.5 ms (gamma / beta) + 2.1 ms (input) per layer -- 42 ms of a 125 ms epoch (scripts/epoch_profile.py) -- for work
// that is one read and one write of the activation: 5 GB forward (0.8 ms at the HBM roofline), 7.5 GB backward.
//   * forward: one warp per row, the row in registers (two float4 per lane, the layout of the TopK kernel), mean and
//     variance in two passes over the registers (no cancellation), y = (x - mean) * rstd * gamma + beta with 128-bit
//     streaming loads / stores; mean and rstd are kept for the backward (8 bytes per row);
//   * backward: dx = rstd * (g*gamma - mean_d(g*gamma) - xhat * mean_d(g*gamma*xhat)), again one warp per row; the
//     column sums dgamma = sum_rows g*xhat, dbeta = sum_rows g accumulate in registers over all the rows a warp
//     visits (persistent grid), are combined across the warps of a CTA in shared memory and leave as one
//     red.global.add per column and CTA.
// Both are pure streaming kernels: HBM-bound.
#include "common.cuh"

namespace maxk {

constexpr int LN_WPB = 8;

__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// lane element i sits at column (i < 4 ? 4 * lane + i : 128 + 4 * lane + (i - 4)); columns >= d do not exist
__global__ void __launch_bounds__(LN_WPB * 32)
layernorm_fwd_kernel(const float *__restrict__ x, int64_t n, int d, const float *__restrict__ gamma,
                     const float *__restrict__ beta, float eps, float *__restrict__ y, float *__restrict__ mean_out,
                     float *__restrict__ rstd_out)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c0 = 4 * lane, c1 = 128 + 4 * lane;
    const bool ok0 = c0 < d, ok1 = c1 < d;
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f), one = make_float4(1.f, 1.f, 1.f, 1.f);
    const float4 ga = ok0 ? (gamma ? __ldg(reinterpret_cast<const float4 *>(gamma + c0)) : one) : zero;
    const float4 gb = ok1 ? (gamma ? __ldg(reinterpret_cast<const float4 *>(gamma + c1)) : one) : zero;
    const float4 ba = ok0 && beta ? __ldg(reinterpret_cast<const float4 *>(beta + c0)) : zero;
    const float4 bb = ok1 && beta ? __ldg(reinterpret_cast<const float4 *>(beta + c1)) : zero;
    const float inv_d = 1.f / (float)d;
    for (int64_t row = (int64_t)blockIdx.x * LN_WPB + warp; row < n; row += (int64_t)gridDim.x * LN_WPB) {
        const float *xr = x + row * d;
        const float4 a = ok0 ? ld_stream_f4(reinterpret_cast<const float4 *>(xr + c0)) : zero;
        const float4 b = ok1 ? ld_stream_f4(reinterpret_cast<const float4 *>(xr + c1)) : zero;
        const float mean = warp_sum((a.x + a.y) + (a.z + a.w) + (b.x + b.y) + (b.z + b.w)) * inv_d;
        float4 da = make_float4(a.x - mean, a.y - mean, a.z - mean, a.w - mean);
        float4 db = make_float4(b.x - mean, b.y - mean, b.z - mean, b.w - mean);
        if (!ok0) da = zero;
        if (!ok1) db = zero;
        const float var = warp_sum((da.x * da.x + da.y * da.y) + (da.z * da.z + da.w * da.w) +
                                   (db.x * db.x + db.y * db.y) + (db.z * db.z + db.w * db.w)) * inv_d;
        const float rstd = rsqrtf(var + eps);
        float *yr = y + row * d;
        if (ok0)
            st_stream_f4(reinterpret_cast<float4 *>(yr + c0),
                         make_float4(fmaf(da.x * rstd, ga.x, ba.x), fmaf(da.y * rstd, ga.y, ba.y),
                                     fmaf(da.z * rstd, ga.z, ba.z), fmaf(da.w * rstd, ga.w, ba.w)));
        if (ok1)
            st_stream_f4(reinterpret_cast<float4 *>(yr + c1),
                         make_float4(fmaf(db.x * rstd, gb.x, bb.x), fmaf(db.y * rstd, gb.y, bb.y),
                                     fmaf(db.z * rstd, gb.z, bb.z), fmaf(db.w * rstd, gb.w, bb.w)));
        if (lane == 0) {
            mean_out[row] = mean;
            rstd_out[row] = rstd;
        }
    }
}

__global__ void __launch_bounds__(LN_WPB * 32)
layernorm_bwd_kernel(const float *__restrict__ g, const float *__restrict__ x, const float *__restrict__ mean_in,
                     const float *__restrict__ rstd_in, const float *__restrict__ gamma, int64_t n, int d,
                     float *__restrict__ dx, float *__restrict__ dgamma, float *__restrict__ dbeta)
{
    __shared__ float s_part[LN_WPB][2][MAXK_MAX_DIM];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c0 = 4 * lane, c1 = 128 + 4 * lane;
    const bool ok0 = c0 < d, ok1 = c1 < d;
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f), one = make_float4(1.f, 1.f, 1.f, 1.f);
    const float4 ga = ok0 ? (gamma ? __ldg(reinterpret_cast<const float4 *>(gamma + c0)) : one) : zero;
    const float4 gb = ok1 ? (gamma ? __ldg(reinterpret_cast<const float4 *>(gamma + c1)) : one) : zero;
    const float inv_d = 1.f / (float)d;
    float4 sga = zero, sgb = zero, sba = zero, sbb = zero;  // this warp's column sums of g * xhat and of g
    for (int64_t row = (int64_t)blockIdx.x * LN_WPB + warp; row < n; row += (int64_t)gridDim.x * LN_WPB) {
        const float mean = __ldg(mean_in + row), rstd = __ldg(rstd_in + row);
        const float *xr = x + row * d, *gr = g + row * d;
        const float4 xa = ok0 ? ld_stream_f4(reinterpret_cast<const float4 *>(xr + c0)) : zero;
        const float4 xb = ok1 ? ld_stream_f4(reinterpret_cast<const float4 *>(xr + c1)) : zero;
        const float4 va = ok0 ? ld_stream_f4(reinterpret_cast<const float4 *>(gr + c0)) : zero;
        const float4 vb = ok1 ? ld_stream_f4(reinterpret_cast<const float4 *>(gr + c1)) : zero;
        float4 ha = make_float4((xa.x - mean) * rstd, (xa.y - mean) * rstd, (xa.z - mean) * rstd, (xa.w - mean) * rstd);
        float4 hb = make_float4((xb.x - mean) * rstd, (xb.y - mean) * rstd, (xb.z - mean) * rstd, (xb.w - mean) * rstd);
        if (!ok0) ha = zero;
        if (!ok1) hb = zero;
        const float4 wa = make_float4(va.x * ga.x, va.y * ga.y, va.z * ga.z, va.w * ga.w);  // g * gamma
        const float4 wb = make_float4(vb.x * gb.x, vb.y * gb.y, vb.z * gb.z, vb.w * gb.w);
        const float m1 = warp_sum((wa.x + wa.y) + (wa.z + wa.w) + (wb.x + wb.y) + (wb.z + wb.w)) * inv_d;
        const float m2 = warp_sum((wa.x * ha.x + wa.y * ha.y) + (wa.z * ha.z + wa.w * ha.w) +
                                  (wb.x * hb.x + wb.y * hb.y) + (wb.z * hb.z + wb.w * hb.w)) * inv_d;
        float *dr = dx + row * d;
        if (ok0)
            st_stream_f4(reinterpret_cast<float4 *>(dr + c0),
                         make_float4(rstd * (wa.x - m1 - ha.x * m2), rstd * (wa.y - m1 - ha.y * m2),
                                     rstd * (wa.z - m1 - ha.z * m2), rstd * (wa.w - m1 - ha.w * m2)));
        if (ok1)
            st_stream_f4(reinterpret_cast<float4 *>(dr + c1),
                         make_float4(rstd * (wb.x - m1 - hb.x * m2), rstd * (wb.y - m1 - hb.y * m2),
                                     rstd * (wb.z - m1 - hb.z * m2), rstd * (wb.w - m1 - hb.w * m2)));
        sga.x = fmaf(va.x, ha.x, sga.x); sga.y = fmaf(va.y, ha.y, sga.y);
        sga.z = fmaf(va.z, ha.z, sga.z); sga.w = fmaf(va.w, ha.w, sga.w);
        sgb.x = fmaf(vb.x, hb.x, sgb.x); sgb.y = fmaf(vb.y, hb.y, sgb.y);
        sgb.z = fmaf(vb.z, hb.z, sgb.z); sgb.w = fmaf(vb.w, hb.w, sgb.w);
        sba.x += va.x; sba.y += va.y; sba.z += va.z; sba.w += va.w;
        sbb.x += vb.x; sbb.y += vb.y; sbb.z += vb.z; sbb.w += vb.w;
    }
    if (dgamma == nullptr && dbeta == nullptr) return;  // uniform
    // ---- the CTA's column sums: warps -> shared memory -> one reduction per column into global memory
    *reinterpret_cast<float4 *>(&s_part[warp][0][c0]) = sga;
    *reinterpret_cast<float4 *>(&s_part[warp][0][c1]) = sgb;
    *reinterpret_cast<float4 *>(&s_part[warp][1][c0]) = sba;
    *reinterpret_cast<float4 *>(&s_part[warp][1][c1]) = sbb;
    __syncthreads();
    for (int c = threadIdx.x; c < d; c += LN_WPB * 32) {
        float tg = 0.f, tb = 0.f;
#pragma unroll
        for (int w = 0; w < LN_WPB; ++w) {
            tg += s_part[w][0][c];
            tb += s_part[w][1][c];
        }
        if (dgamma != nullptr) red_add_f1(dgamma + c, tg);
        if (dbeta != nullptr) red_add_f1(dbeta + c, tb);
    }
}

// The backward of LayerNorm + TopK (topk_cbsr.cu: layernorm_topk_kernel): the gradient that reaches the normalised
// activation is k-sparse per row -- the CBSR gradient of the aggregation, at the columns the selection kept -- so it is
// read as it is (5k bytes per row instead of a scattered dense twin: no scatter kernel, no 4d-byte read) and spread
// over the lanes' columns through a shared-memory row; the arithmetic is layernorm_bwd_kernel's.
__global__ void __launch_bounds__(LN_WPB * 32)
layernorm_bwd_cbsr_kernel(const float *__restrict__ g_data, const unsigned char *__restrict__ sp_index, int k,
                          const float *__restrict__ x, const float *__restrict__ mean_in,
                          const float *__restrict__ rstd_in, const float *__restrict__ gamma, int64_t n, int d,
                          float *__restrict__ dx, float *__restrict__ dgamma, float *__restrict__ dbeta)
{
    __shared__ float s_part[LN_WPB][2][MAXK_MAX_DIM];   // first the warps' gradient rows, then their column sums
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c0 = 4 * lane, c1 = 128 + 4 * lane;
    const bool ok0 = c0 < d, ok1 = c1 < d;
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f), one = make_float4(1.f, 1.f, 1.f, 1.f);
    const float4 ga = ok0 ? (gamma ? __ldg(reinterpret_cast<const float4 *>(gamma + c0)) : one) : zero;
    const float4 gb = ok1 ? (gamma ? __ldg(reinterpret_cast<const float4 *>(gamma + c1)) : one) : zero;
    const float inv_d = 1.f / (float)d;
    float *grow = s_part[warp][0];
    float4 sga = zero, sgb = zero, sba = zero, sbb = zero;
    for (int64_t row = (int64_t)blockIdx.x * LN_WPB + warp; row < n; row += (int64_t)gridDim.x * LN_WPB) {
        // the row's gradient: zeros, then the k kept columns
        *reinterpret_cast<float4 *>(grow + c0) = zero;
        *reinterpret_cast<float4 *>(grow + c1) = zero;
        __syncwarp();
        for (int j = lane; j < k; j += 32) {
            const int c = sp_index[row * k + j];
            if (c < d) grow[c] = ld_stream_f32(g_data + row * k + j);
        }
        __syncwarp();
        const float4 va = *reinterpret_cast<const float4 *>(grow + c0);
        const float4 vb = *reinterpret_cast<const float4 *>(grow + c1);
        __syncwarp();
        const float mean = __ldg(mean_in + row), rstd = __ldg(rstd_in + row);
        const float *xr = x + row * d;
        const float4 xa = ok0 ? ld_stream_f4(reinterpret_cast<const float4 *>(xr + c0)) : zero;
        const float4 xb = ok1 ? ld_stream_f4(reinterpret_cast<const float4 *>(xr + c1)) : zero;
        float4 ha = make_float4((xa.x - mean) * rstd, (xa.y - mean) * rstd, (xa.z - mean) * rstd, (xa.w - mean) * rstd);
        float4 hb = make_float4((xb.x - mean) * rstd, (xb.y - mean) * rstd, (xb.z - mean) * rstd, (xb.w - mean) * rstd);
        if (!ok0) ha = zero;
        if (!ok1) hb = zero;
        const float4 wa = make_float4(va.x * ga.x, va.y * ga.y, va.z * ga.z, va.w * ga.w);
        const float4 wb = make_float4(vb.x * gb.x, vb.y * gb.y, vb.z * gb.z, vb.w * gb.w);
        const float m1 = warp_sum((wa.x + wa.y) + (wa.z + wa.w) + (wb.x + wb.y) + (wb.z + wb.w)) * inv_d;
        const float m2 = warp_sum((wa.x * ha.x + wa.y * ha.y) + (wa.z * ha.z + wa.w * ha.w) +
                                  (wb.x * hb.x + wb.y * hb.y) + (wb.z * hb.z + wb.w * hb.w)) * inv_d;
        float *dr = dx + row * d;
        if (ok0)
            st_stream_f4(reinterpret_cast<float4 *>(dr + c0),
                         make_float4(rstd * (wa.x - m1 - ha.x * m2), rstd * (wa.y - m1 - ha.y * m2),
                                     rstd * (wa.z - m1 - ha.z * m2), rstd * (wa.w - m1 - ha.w * m2)));
        if (ok1)
            st_stream_f4(reinterpret_cast<float4 *>(dr + c1),
                         make_float4(rstd * (wb.x - m1 - hb.x * m2), rstd * (wb.y - m1 - hb.y * m2),
                                     rstd * (wb.z - m1 - hb.z * m2), rstd * (wb.w - m1 - hb.w * m2)));
        sga.x = fmaf(va.x, ha.x, sga.x); sga.y = fmaf(va.y, ha.y, sga.y);
        sga.z = fmaf(va.z, ha.z, sga.z); sga.w = fmaf(va.w, ha.w, sga.w);
        sgb.x = fmaf(vb.x, hb.x, sgb.x); sgb.y = fmaf(vb.y, hb.y, sgb.y);
        sgb.z = fmaf(vb.z, hb.z, sgb.z); sgb.w = fmaf(vb.w, hb.w, sgb.w);
        sba.x += va.x; sba.y += va.y; sba.z += va.z; sba.w += va.w;
        sbb.x += vb.x; sbb.y += vb.y; sbb.z += vb.z; sbb.w += vb.w;
    }
    if (dgamma == nullptr && dbeta == nullptr) return;  // uniform
    __syncwarp();
    *reinterpret_cast<float4 *>(&s_part[warp][0][c0]) = sga;
    *reinterpret_cast<float4 *>(&s_part[warp][0][c1]) = sgb;
    *reinterpret_cast<float4 *>(&s_part[warp][1][c0]) = sba;
    *reinterpret_cast<float4 *>(&s_part[warp][1][c1]) = sbb;
    __syncthreads();
    for (int c = threadIdx.x; c < d; c += LN_WPB * 32) {
        float tg = 0.f, tb = 0.f;
#pragma unroll
        for (int w = 0; w < LN_WPB; ++w) {
            tg += s_part[w][0][c];
            tb += s_part[w][1][c];
        }
        if (dgamma != nullptr) red_add_f1(dgamma + c, tg);
        if (dbeta != nullptr) red_add_f1(dbeta + c, tb);
    }
}

}  // namespace maxk

using namespace maxk;

static int ln_grid(int64_t n)
{
    const int64_t want = ceil_div<int64_t>(n, LN_WPB);
    const int64_t cap = (int64_t)kNumSMs * 8;  // persistent: 8 CTAs of 8 warps per SM
    return (int)(want < cap ? want : cap);
}

extern "C" int maxk_layernorm_forward(const float *x, int64_t n, int32_t d, const float *gamma, const float *beta,
                                      float eps, float *y, float *mean, float *rstd, void *stream)
{
    MAXK_REQUIRE(n >= 0, "maxk_layernorm_forward: negative n");
    MAXK_REQUIRE(d > 0 && d <= MAXK_MAX_DIM && d % 4 == 0,
                 "maxk_layernorm_forward: d=%d must be a multiple of 4 in [4,256]", d);
    if (n == 0) return MAXK_OK;
    MAXK_REQUIRE(x != nullptr && y != nullptr && mean != nullptr && rstd != nullptr,
                 "maxk_layernorm_forward: NULL pointer");
    MAXK_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(gamma) |
                   reinterpret_cast<uintptr_t>(beta)) & 15) == 0,
                 "maxk_layernorm_forward: x, y, gamma and beta must be 16-byte aligned");
    layernorm_fwd_kernel<<<ln_grid(n), LN_WPB * 32, 0, static_cast<cudaStream_t>(stream)>>>(x, n, d, gamma, beta, eps, y,
                                                                                           mean, rstd);
    MAXK_LAUNCH_CHECK("layernorm_fwd_kernel");
    return MAXK_OK;
}

extern "C" int maxk_layernorm_backward(const float *grad_out, const float *x, const float *mean, const float *rstd,
                                       const float *gamma, int64_t n, int32_t d, float *dx, float *dgamma, float *dbeta,
                                       void *stream)
{
    MAXK_REQUIRE(n >= 0, "maxk_layernorm_backward: negative n");
    MAXK_REQUIRE(d > 0 && d <= MAXK_MAX_DIM && d % 4 == 0,
                 "maxk_layernorm_backward: d=%d must be a multiple of 4 in [4,256]", d);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (dgamma != nullptr) MAXK_CUDA(cudaMemsetAsync(dgamma, 0, sizeof(float) * (size_t)d, st));
    if (dbeta != nullptr) MAXK_CUDA(cudaMemsetAsync(dbeta, 0, sizeof(float) * (size_t)d, st));
    if (n == 0) return MAXK_OK;
    MAXK_REQUIRE(grad_out != nullptr && x != nullptr && mean != nullptr && rstd != nullptr && dx != nullptr,
                 "maxk_layernorm_backward: NULL pointer");
    MAXK_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(grad_out) |
                   reinterpret_cast<uintptr_t>(dx) | reinterpret_cast<uintptr_t>(gamma)) & 15) == 0,
                 "maxk_layernorm_backward: grad_out, x, dx and gamma must be 16-byte aligned");
    layernorm_bwd_kernel<<<ln_grid(n), LN_WPB * 32, 0, st>>>(grad_out, x, mean, rstd, gamma, n, d, dx, dgamma, dbeta);
    MAXK_LAUNCH_CHECK("layernorm_bwd_kernel");
    return MAXK_OK;
}

extern "C" int maxk_layernorm_backward_cbsr(const float *grad_data, const uint8_t *sp_index, int32_t k, const float *x,
                                            const float *mean, const float *rstd, const float *gamma, int64_t n,
                                            int32_t d, float *dx, float *dgamma, float *dbeta, void *stream)
{
    MAXK_REQUIRE(n >= 0, "maxk_layernorm_backward_cbsr: negative n");
    MAXK_REQUIRE(d > 0 && d <= MAXK_MAX_DIM && d % 4 == 0,
                 "maxk_layernorm_backward_cbsr: d=%d must be a multiple of 4 in [4,256]", d);
    MAXK_REQUIRE(k >= 1 && k <= d, "maxk_layernorm_backward_cbsr: k=%d must be in [1, d=%d]", k, d);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (dgamma != nullptr) MAXK_CUDA(cudaMemsetAsync(dgamma, 0, sizeof(float) * (size_t)d, st));
    if (dbeta != nullptr) MAXK_CUDA(cudaMemsetAsync(dbeta, 0, sizeof(float) * (size_t)d, st));
    if (n == 0) return MAXK_OK;
    MAXK_REQUIRE(grad_data != nullptr && sp_index != nullptr && x != nullptr && mean != nullptr && rstd != nullptr &&
                     dx != nullptr,
                 "maxk_layernorm_backward_cbsr: NULL pointer");
    MAXK_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dx) | reinterpret_cast<uintptr_t>(gamma)) &
                  15) == 0,
                 "maxk_layernorm_backward_cbsr: x, dx and gamma must be 16-byte aligned");
    layernorm_bwd_cbsr_kernel<<<ln_grid(n), LN_WPB * 32, 0, st>>>(grad_data, sp_index, k, x, mean, rstd, gamma, n, d, dx,
                                                                   dgamma, dbeta);
    MAXK_LAUNCH_CHECK("layernorm_bwd_cbsr_kernel");
    return MAXK_OK;
}
