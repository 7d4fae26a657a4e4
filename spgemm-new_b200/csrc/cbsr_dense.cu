// cbsr_dense.cu -- EXPERIMENTAL (written after round 1's GPU budget was spent; checked under the
// SIMT interpreter only, unmeasured): products of a CBSR operand with a small dense matrix, for
// the dense layers either side of the aggregation:
//
//   matmul      out[i, :]  (+)= bias + sum_j sp_data[i,j] * M[sp_index[i,j], :]
//   sampled_dot dv[i, j]   (+)= < G[i, :], M[sp_index[i,j], :] >
//
// With M = W^T these are `fc_self(h)` of the SAGE layer on the MaxK-sparsified h (upstream
// utils/models_old.py:55-60: SAGEConv feeds dropout(maxk(x)) to fc_self and fc_neigh; the fork's
// utils/models.py:242 applies fc_self to the dense x) and the gradient of that product with
// respect to the k kept values -- N*k*C multiply-adds each instead of the N*D*C of the dense
// fp32 GEMM on the scattered twin (8x fewer at k = 32, D = 256), and the dense [N, D] twin is
// neither written nor read.  The weight gradient stays a dense library GEMM.
//
// Tensor cores do not apply (every row of the left operand picks its own k rows of M).  The bound
// is shared-memory bandwidth: a 128-column tile of M (d x 512 bytes <= 128 KB) lives in shared
// memory, one warp owns one row at a time, lane l owns columns 4l..4l+3 of the tile, and every
// (row, entry) costs one conflict-free LDS.128 per lane (512 contiguous bytes per warp) plus one
// broadcast LDS.64 of the staged (value, row offset) pair: N*k*C*4 bytes through a 128 B/clk pipe
// -- 2.2 ms for the products shape (N = 2.45 M, k = 32, C = 256) against ~5 ms of fp32 SGEMM.
// Persistent grid: one 16-warp CTA per SM, rows dealt round-robin to warps, column tiles in turn.
#include "common.cuh"

namespace maxk {

constexpr int CD_WARPS = 16;
constexpr int CD_TILE = 128;  // columns of M per pass: 32 lanes x float4

// M[0:d, tile*128 .. +128) -> shared memory (zero beyond column c)
__device__ __forceinline__ void cd_load_tile(float *ms, const float *__restrict__ m, int d, int c, int col0)
{
    for (int u = threadIdx.x; u < d * (CD_TILE / 4); u += blockDim.x) {
        const int r = u >> 5, q = u & 31;
        const int col = col0 + 4 * q;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (col < c) v = __ldg(reinterpret_cast<const float4 *>(m + (int64_t)r * c + col));
        *reinterpret_cast<float4 *>(ms + r * CD_TILE + 4 * q) = v;
    }
}

// the k entries of one CBSR row -> this warp's staging area as (value, byte offset of the M row in
// the tile) pairs.  A selector outside [0, d) contributes nothing (same guard as the layout
// kernels): it is staged as value 0 on row 0.  Without values (sampled_dot) the first word is 1 / 0.
__device__ __forceinline__ void cd_stage_row(float2 *pr, const float *__restrict__ sp_data,
                                             const unsigned char *__restrict__ sp_index, int64_t row, int k,
                                             int d, int lane, bool with_values)
{
    for (int j = lane; j < k; j += 32) {
        const unsigned s = sp_index[row * k + j];
        const bool ok = (int)s < d;
        const float v = !ok ? 0.f : with_values ? __ldg(sp_data + row * k + j) : 1.f;
        pr[j] = make_float2(v, __uint_as_float(ok ? s * (unsigned)(CD_TILE * 4) : 0u));
    }
}

__global__ void __launch_bounds__(CD_WARPS * 32, 1)
cbsr_matmul_kernel(const float *__restrict__ sp_data, const unsigned char *__restrict__ sp_index, int64_t n,
                   int k, const float *__restrict__ m, int d, int c, const float *__restrict__ bias,
                   float *__restrict__ out, int accumulate)
{
    extern __shared__ __align__(16) float cd_smem[];
    float *ms = cd_smem;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float2 *pr = reinterpret_cast<float2 *>(cd_smem + d * CD_TILE) + warp * k;
    const int64_t gw = (int64_t)blockIdx.x * CD_WARPS + warp, stride = (int64_t)gridDim.x * CD_WARPS;

    for (int col0 = 0; col0 < c; col0 += CD_TILE) {
        __syncthreads();  // everyone is done with the previous tile
        cd_load_tile(ms, m, d, c, col0);
        __syncthreads();
        const int col = col0 + 4 * lane;
        const bool live = col < c;
        float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (bias != nullptr && live) b4 = __ldg(reinterpret_cast<const float4 *>(bias + col));
        const char *mt = reinterpret_cast<const char *>(ms + 4 * lane);
        for (int64_t row = gw; row < n; row += stride) {
            cd_stage_row(pr, sp_data, sp_index, row, k, d, lane, true);
            __syncwarp();
            float4 acc = b4;
#pragma unroll 4
            for (int j = 0; j < k; ++j) {
                const float2 p = pr[j];  // broadcast
                const float4 w = *reinterpret_cast<const float4 *>(mt + __float_as_uint(p.y));
                acc.x = fmaf(p.x, w.x, acc.x);
                acc.y = fmaf(p.x, w.y, acc.y);
                acc.z = fmaf(p.x, w.z, acc.z);
                acc.w = fmaf(p.x, w.w, acc.w);
            }
            if (live) {
                float4 *o = reinterpret_cast<float4 *>(out + row * c + col);
                if (accumulate) {
                    const float4 old = *o;
                    acc.x += old.x;
                    acc.y += old.y;
                    acc.z += old.z;
                    acc.w += old.w;
                }
                *o = acc;
            }
            __syncwarp();  // the staging area is rewritten for the next row
        }
    }
}

// NB per-lane partial sums p[0..NB) -> every lane t ends with the warp total of p[t mod NB]
// (halving exchange over the low lane bits: NB - 1 shuffles for NB sums instead of NB x 5; plain butterfly
// adds over the remaining lane bits)
template <int NB>
__device__ __forceinline__ float cd_transpose_reduce(float (&p)[NB], int lane)
{
#pragma unroll
    for (int half = NB / 2; half >= 1; half >>= 1) {
        const bool upper = (lane & half) != 0;
#pragma unroll
        for (int i = 0; i < half; ++i) {
            const float send = upper ? p[i] : p[i + half];
            const float keep = upper ? p[i + half] : p[i];
            p[i] = keep + __shfl_xor_sync(0xffffffffu, send, half);
        }
    }
    float v = p[0];
#pragma unroll
    for (int bit = NB; bit < 32; bit <<= 1) v += __shfl_xor_sync(0xffffffffu, v, bit);
    return v;
}

// NB = entries handled per exchange: 8 / 16 for k <= 8 / 16, else 32; the staged row is padded to a
// multiple of NB with entries that contribute nothing, so the inner loop has no bounds test.
template <int NB>
__global__ void __launch_bounds__(CD_WARPS * 32, 1)
cbsr_sampled_dot_kernel(const float *__restrict__ g, int64_t n, int c, const float *__restrict__ m, int d,
                        const unsigned char *__restrict__ sp_index, int k, float *__restrict__ out,
                        int accumulate)
{
    extern __shared__ __align__(16) float cd_smem[];
    float *ms = cd_smem;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int kpad = (k + NB - 1) / NB * NB;
    float2 *pr = reinterpret_cast<float2 *>(cd_smem + d * CD_TILE) + warp * kpad;
    const int64_t gw = (int64_t)blockIdx.x * CD_WARPS + warp, stride = (int64_t)gridDim.x * CD_WARPS;
    for (int j = k + lane; j < kpad; j += 32) pr[j] = make_float2(0.f, 0.f);  // padding, written once

    for (int col0 = 0; col0 < c; col0 += CD_TILE) {
        __syncthreads();
        cd_load_tile(ms, m, d, c, col0);
        __syncthreads();
        const int col = col0 + 4 * lane;
        const char *mt = reinterpret_cast<const char *>(ms + 4 * lane);
        const bool first = (col0 == 0) && !accumulate;  // later tiles add to what the same lane wrote
        for (int64_t row = gw; row < n; row += stride) {
            cd_stage_row(pr, nullptr, sp_index, row, k, d, lane, false);
            float4 g4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (col < c) g4 = __ldg(reinterpret_cast<const float4 *>(g + row * c + col));
            __syncwarp();
            for (int j0 = 0; j0 < kpad; j0 += NB) {
                float p[NB];
#pragma unroll
                for (int i = 0; i < NB; ++i) {
                    const float4 w = *reinterpret_cast<const float4 *>(mt + __float_as_uint(pr[j0 + i].y));
                    p[i] = fmaf(g4.x, w.x, fmaf(g4.y, w.y, fmaf(g4.z, w.z, g4.w * w.w)));
                }
                float tot = cd_transpose_reduce<NB>(p, lane);
                const int j = j0 + lane;
                if (lane < NB && j < k) {
                    if (pr[j].x == 0.f) tot = 0.f;  // selector outside [0, d)
                    float *o = out + row * k + j;
                    *o = first ? tot : *o + tot;
                }
            }
            __syncwarp();
        }
    }
}

// y[i, sp_index[i,j]] += alpha * sp_data[i,j]  (GIN's "(1 + eps) * h" term added to the aggregate in place,
// utils/models.py:361, without the dense twin h): one warp per row, k scattered read-modify-writes inside the
// row's own 4d bytes.  alpha = alpha_host + *alpha_dev (the learnable eps stays on the device: no host sync).
constexpr int AX_WPB = 8;
__global__ void __launch_bounds__(AX_WPB * 32)
cbsr_axpy_kernel(const float *__restrict__ sp_data, const unsigned char *__restrict__ sp_index, int64_t n, int k,
                 int d, const float *__restrict__ alpha_dev, float alpha_host, float *__restrict__ y)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t row = (int64_t)blockIdx.x * AX_WPB + warp;
    if (row >= n) return;
    const float alpha = alpha_host + (alpha_dev != nullptr ? __ldg(alpha_dev) : 0.f);
    float *yr = y + row * d;
    for (int j = lane; j < k; j += 32) {
        const int s = sp_index[row * k + j];
        if (s < d) yr[s] = fmaf(alpha, __ldg(sp_data + row * k + j), yr[s]);  // selectors of a row are distinct
    }
}

static int cd_check(const char *who, int64_t n, int k, int d, int c)
{
    MAXK_REQUIRE(n >= 0, "%s: negative n", who);
    MAXK_REQUIRE(d >= 1 && d <= MAXK_MAX_DIM, "%s: d=%d out of range [1,256]", who, d);
    MAXK_REQUIRE(k >= 1 && k <= 256, "%s: k=%d out of range", who, k);
    MAXK_REQUIRE(c >= 4 && c % 4 == 0 && c <= 65536, "%s: c=%d must be a multiple of 4 in [4,65536]", who, c);
    return MAXK_OK;
}

static size_t cd_smem_bytes(int d, int k, int nb = 1)
{
    return (size_t)d * CD_TILE * 4 + (size_t)CD_WARPS * ((k + nb - 1) / nb * nb) * 8;
}

static unsigned cd_grid(int64_t n)
{
    const int64_t g = ceil_div<int64_t>(n, CD_WARPS);
    return (unsigned)(g < kNumSMs ? g : kNumSMs);
}

template <int NB>
static int launch_sampled_dot(const float *g, int64_t n, int c, const float *m, int d, const uint8_t *sp_index, int k,
                              float *out, int accumulate, void *stream)
{
    if (int rc = opt_in_dynamic_smem(reinterpret_cast<const void *>(&cbsr_sampled_dot_kernel<NB>),
                                     "cbsr_sampled_dot_kernel"))
        return rc;
    cbsr_sampled_dot_kernel<NB><<<cd_grid(n), CD_WARPS * 32, cd_smem_bytes(d, k, NB), static_cast<cudaStream_t>(stream)>>>(
        g, n, c, m, d, sp_index, k, out, accumulate);
    MAXK_LAUNCH_CHECK("cbsr_sampled_dot_kernel");
    return MAXK_OK;
}

}  // namespace maxk

using namespace maxk;

extern "C" int maxk_cbsr_matmul(const float *sp_data, const uint8_t *sp_index, int64_t n, int32_t k,
                                const float *m, int32_t d, int32_t c, const float *bias, float *out,
                                int32_t accumulate, void *stream)
{
    if (int rc = cd_check("maxk_cbsr_matmul", n, k, d, c)) return rc;
    if (n == 0) return MAXK_OK;
    MAXK_REQUIRE(sp_data && sp_index && m && out, "maxk_cbsr_matmul: NULL pointer");
    MAXK_REQUIRE(((reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(out) |
                   reinterpret_cast<uintptr_t>(bias)) & 15) == 0,
                 "maxk_cbsr_matmul: m, bias and out must be 16-byte aligned");
    if (int rc = opt_in_dynamic_smem(reinterpret_cast<const void *>(&cbsr_matmul_kernel), "cbsr_matmul_kernel")) return rc;
    cbsr_matmul_kernel<<<cd_grid(n), CD_WARPS * 32, cd_smem_bytes(d, k), static_cast<cudaStream_t>(stream)>>>(
        sp_data, sp_index, n, k, m, d, c, bias, out, accumulate);
    MAXK_LAUNCH_CHECK("cbsr_matmul_kernel");
    return MAXK_OK;
}

extern "C" int maxk_cbsr_sampled_dot(const float *g, int64_t n, int32_t c, const float *m, int32_t d,
                                     const uint8_t *sp_index, int32_t k, float *out, int32_t accumulate,
                                     void *stream)
{
    if (int rc = cd_check("maxk_cbsr_sampled_dot", n, k, d, c)) return rc;
    if (n == 0) return MAXK_OK;
    MAXK_REQUIRE(g && sp_index && m && out, "maxk_cbsr_sampled_dot: NULL pointer");
    MAXK_REQUIRE(((reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(g)) & 15) == 0,
                 "maxk_cbsr_sampled_dot: m and g must be 16-byte aligned");
    return k <= 8 ? launch_sampled_dot<8>(g, n, c, m, d, sp_index, k, out, accumulate, stream)
           : k <= 16 ? launch_sampled_dot<16>(g, n, c, m, d, sp_index, k, out, accumulate, stream)
                     : launch_sampled_dot<32>(g, n, c, m, d, sp_index, k, out, accumulate, stream);
}

extern "C" int maxk_cbsr_axpy(const float *sp_data, const uint8_t *sp_index, int64_t n, int32_t k, int32_t d,
                              const float *alpha_dev, float alpha_host, float *y, void *stream)
{
    MAXK_REQUIRE(n >= 0, "maxk_cbsr_axpy: negative n");
    MAXK_REQUIRE(d >= 1 && d <= MAXK_MAX_DIM, "maxk_cbsr_axpy: d=%d out of range [1,256]", d);
    MAXK_REQUIRE(k >= 1 && k <= 256, "maxk_cbsr_axpy: k=%d out of range", k);
    if (n == 0) return MAXK_OK;
    MAXK_REQUIRE(sp_data && sp_index && y, "maxk_cbsr_axpy: NULL pointer");
    const int64_t grid = ceil_div<int64_t>(n, AX_WPB);
    MAXK_REQUIRE(grid <= 0x7fffffffLL, "maxk_cbsr_axpy: n too large");
    cbsr_axpy_kernel<<<(unsigned)grid, AX_WPB * 32, 0, static_cast<cudaStream_t>(stream)>>>(
        sp_data, sp_index, n, k, d, alpha_dev, alpha_host, y);
    MAXK_LAUNCH_CHECK("cbsr_axpy_kernel");
    return MAXK_OK;
}
