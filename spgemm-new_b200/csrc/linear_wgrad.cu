// linear_wgrad.cu -- the weight gradient of a Linear next to the aggregation path on the tensor cores:
//
//     dW[m][c] (+)= sum_r G[r][m] * X[r][c]        G [n, 256] = dL/d(output), X [n, in] = the layer's input
//
// i.e. `grad_output.t() @ input` of torch.nn.Linear's backward (what autograd runs for every Linear of
// utils/models.py:151-392; in fp32 it is a library SGEMM whose reduction dimension is the NODE count: 8.7 ms per
// layer on the products shape, the largest single cost of a training epoch after round 2's other changes).
// Same machinery as linear_topk.cu -- TMA ring, operand split x = hi + lo for fp32 accuracy from TF32 products,
// tcgen05.mma with fp32 accumulators in TMEM -- with the roles of the dimensions exchanged:
//   * M = 256 columns of G (two 128-row MMA tiles), N = `in` columns of X (<= 256, padded to 16 by TMA zero fill),
//     K = rows: both operands are MN-major in memory (the row index is the slow one), so the shared-memory tiles are
//     [16 rows x 32 columns] boxes (16 x 128 bytes, the 128-byte swizzle with a 32-byte base), one per 32 columns, and
//     the descriptors say MN-major (leading byte offset = distance between boxes, stride byte offset = four rows);
//   * the two 128 x 256 accumulators fill TMEM for the whole kernel: a CTA owns a contiguous range of row chunks,
//     accumulates all of them, and adds its 256 x `in` partial to dW once at the end (red.global.add.v4.f32);
//   * warps 2-11 split the boxes in place (hi) and into the lo copies; warps 4-11 then drain TMEM.
// Bound: 2 x 4nD bytes read once (HBM) against 12 MMAs x 128 cycles per 16 rows (tensor pipe): both ~0.8 ms on the
// products shape.
#include <mutex>
#include <set>

#include "common.cuh"
#include "tc_common.cuh"

namespace maxk {

constexpr int WG_M = 256;          // columns of G = rows of dW
constexpr int WG_KC = 16;          // rows per chunk
constexpr int WG_STAGES = 3;
constexpr int WG_BOX_BYTES = WG_KC * 128;            // one [16 x 32] fp32 box
constexpr int WG_OP_BYTES = 8 * WG_BOX_BYTES;        // 256 columns of one operand: 16 KB
constexpr int WG_STAGE_BYTES = 4 * WG_OP_BYTES;      // A_hi, A_lo, B_hi, B_lo
constexpr int WG_THREADS = 384;                      // warp 0 TMA, 1 MMA, 2-11 split, 4-11 also the epilogue
constexpr int WG_SPLIT_THREADS = 320;

struct WgShared {
    static constexpr int bars = WG_STAGES * WG_STAGE_BYTES;  // full, conv, empty per stage; done
    static constexpr int n_bars = 3 * WG_STAGES + 1;
    static constexpr int tmem_ptr = bars + n_bars * 8;
    static constexpr int total = tmem_ptr + 16;
};

// MN-major TF32 operand: the only shared-memory layout the tensor core takes is the 128-byte swizzle with a 32-byte
// base (cutlass/gemm/collective/builders/sm100_common.inl:92; layout type SWIZZLE_128B_BASE32B = 1, Swizzle<2,5,2>:
// 32-byte chunk index ^= row mod 4), which TMA writes as CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B.  Boxes of 32 columns
// (128 B) x 16 rows; atom = 128 B x 4 rows: LBO = next 32 columns = one box, SBO = next four rows = 512 B
// (cute/atom/mma_traits_sm100.hpp make_umma_desc<MN>)
__device__ __forceinline__ uint64_t smem_desc_mnmajor(uint32_t addr)
{
    constexpr uint32_t hi = (1u << 29) | (1u << 14) | (uint32_t)(512 / 16);
    return ((uint64_t)hi << 32) | ((uint32_t)(WG_BOX_BYTES / 16) << 16) | ((addr >> 4) & 0x3fffu);
}

__global__ void __launch_bounds__(WG_THREADS, 1)
linear_wgrad_kernel(const __grid_constant__ CUtensorMap map_g, const __grid_constant__ CUtensorMap map_x, int64_t n,
                    int in_dim, float *__restrict__ dw)
{
    extern __shared__ unsigned char wg_smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t sraw = smem_u32(wg_smem_raw);
    const uint32_t sbase = (sraw + 1023u) & ~1023u;
    unsigned char *smem = wg_smem_raw + (sbase - sraw);
    const uint32_t bar0 = sbase + WgShared::bars;
    auto bar_full = [&](int s) { return bar0 + 8u * (uint32_t)s; };
    auto bar_conv = [&](int s) { return bar0 + 8u * (uint32_t)(WG_STAGES + s); };
    auto bar_empty = [&](int s) { return bar0 + 8u * (uint32_t)(2 * WG_STAGES + s); };
    const uint32_t bar_done = bar0 + 8u * (uint32_t)(3 * WG_STAGES);
    uint32_t *tmem_ptr_smem = reinterpret_cast<uint32_t *>(smem + WgShared::tmem_ptr);

    const int n_mma = (in_dim + 15) & ~15;       // MMA N: multiple of 16; columns >= in_dim arrive as zeros
    const int x_boxes = (in_dim + 31) / 32;      // 32-column boxes of X
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) |
                           ((uint32_t)(n_mma >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);  // f32 += tf32 x tf32, MN-major
    // this CTA's row chunks
    const int64_t chunks = (n + WG_KC - 1) / WG_KC;
    const int64_t per = (chunks + gridDim.x - 1) / gridDim.x;
    const int64_t c_lo = (int64_t)blockIdx.x * per;
    const int64_t c_hi = c_lo + per < chunks ? c_lo + per : chunks;
    const int64_t my_chunks = c_hi > c_lo ? c_hi - c_lo : 0;

    if (threadIdx.x == 0) {
        for (int s = 0; s < WG_STAGES; ++s) {
            mbar_init(bar_full(s), 1);
            mbar_init(bar_conv(s), WG_SPLIT_THREADS);
            mbar_init(bar_empty(s), 1);
        }
        mbar_init(bar_done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)),
                     "r"(512)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    const uint32_t tx_bytes = (uint32_t)((8 + x_boxes) * WG_BOX_BYTES);

    if (warp == 0) {
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            for (int64_t c = 0; c < my_chunks; ++c) {
                mbar_wait<true>(bar_empty(s), ph ^ 1u);
                const uint32_t st = sbase + (uint32_t)(s * WG_STAGE_BYTES);
                const int r0 = (int)((c_lo + c) * WG_KC);
                mbar_expect_tx(bar_full(s), tx_bytes);
                for (int b = 0; b < 8; ++b) tma_load_2d(st + (uint32_t)(b * WG_BOX_BYTES), &map_g, 32 * b, r0, bar_full(s));
                for (int b = 0; b < x_boxes; ++b)
                    tma_load_2d(st + (uint32_t)(2 * WG_OP_BYTES + b * WG_BOX_BYTES), &map_x, 32 * b, r0, bar_full(s));
                if (++s == WG_STAGES) {
                    s = 0;
                    ph ^= 1u;
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            for (int64_t c = 0; c < my_chunks; ++c) {
                mbar_wait(bar_full(s), ph);
                mbar_wait(bar_conv(s), ph);
                tc_fence_after();
                const uint32_t st = sbase + (uint32_t)(s * WG_STAGE_BYTES);
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) {        // columns [128 mt, 128 mt + 128) of G: four boxes
#pragma unroll
                    for (int ks = 0; ks < WG_KC / 8; ++ks) {  // eight rows per MMA: 1024 B further into every box
                        const uint32_t off = (uint32_t)(mt * 4 * WG_BOX_BYTES + ks * 1024);
                        const uint64_t a_hi = smem_desc_mnmajor(st + off), a_lo = smem_desc_mnmajor(st + WG_OP_BYTES + off);
                        const uint64_t b_hi = smem_desc_mnmajor(st + 2 * WG_OP_BYTES + (uint32_t)(ks * 1024));
                        const uint64_t b_lo = smem_desc_mnmajor(st + 3 * WG_OP_BYTES + (uint32_t)(ks * 1024));
                        const uint32_t d = tmem_base + (uint32_t)(mt * 256);
                        tc_mma_tf32(d, a_hi, b_lo, idesc, (c | ks) != 0 ? 1u : 0u);
                        tc_mma_tf32(d, a_lo, b_hi, idesc, 1u);
                        tc_mma_tf32(d, a_hi, b_hi, idesc, 1u);
                    }
                }
                tc_commit(bar_empty(s));
                if (++s == WG_STAGES) {
                    s = 0;
                    ph ^= 1u;
                }
            }
            tc_commit(bar_done);
        }
    } else {
        // ===== split both operands x = hi + lo in place / into the lo copies ===========================================
        const int t = threadIdx.x - 64;
        int s = 0;
        uint32_t ph = 0;
        const int a4 = WG_OP_BYTES / 16, b4 = x_boxes * WG_BOX_BYTES / 16;  // float4 per operand
        for (int64_t c = 0; c < my_chunks; ++c) {
            mbar_wait(bar_full(s), ph);
            float4 *base = reinterpret_cast<float4 *>(smem + s * WG_STAGE_BYTES);
            for (int i = t; i < a4 + b4; i += WG_SPLIT_THREADS) {
                float4 *hi = i < a4 ? base + i : base + 2 * a4 + (i - a4);
                float4 *lo = hi + a4;
                const float4 v = *hi;
                float4 h, l;
                h.x = __uint_as_float(__float_as_uint(v.x) & 0xffffe000u);
                h.y = __uint_as_float(__float_as_uint(v.y) & 0xffffe000u);
                h.z = __uint_as_float(__float_as_uint(v.z) & 0xffffe000u);
                h.w = __uint_as_float(__float_as_uint(v.w) & 0xffffe000u);
                l.x = tf32_lo_part(v.x, h.x);
                l.y = tf32_lo_part(v.y, h.y);
                l.z = tf32_lo_part(v.z, h.z);
                l.w = tf32_lo_part(v.w, h.w);
                *hi = h;
                *lo = l;
            }
            fence_proxy_async();
            mbar_arrive(bar_conv(s));
            if (++s == WG_STAGES) {
                s = 0;
                ph ^= 1u;
            }
        }
        // ===== epilogue (warps 4-11): this CTA's partial -> dW ===========================================================
        if (warp >= 4 && my_chunks > 0) {
            const int q = warp & 3, mt = (warp - 4) >> 2;
            const int m = mt * 128 + 32 * q + lane;  // row of dW = column of G = TMEM lane of accumulator mt
            mbar_wait(bar_done, 0);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(mt * 256);
            float *drow = dw + (int64_t)m * in_dim;
            for (int c0 = 0; c0 < in_dim; c0 += 16) {  // in_dim % 4 == 0
                float v[16];
                tmem_ld16(taddr + (uint32_t)c0, v);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; i += 4)
                    if (c0 + i < in_dim) red_add_f4(drow + c0 + i, v[i], v[i + 1], v[i + 2], v[i + 3]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

}  // namespace maxk

using namespace maxk;

extern "C" int maxk_linear_wgrad(const float *grad_out, const float *x, int64_t n, int32_t d, int32_t in_dim, float *dw,
                                 int32_t accumulate, void *stream)
{
    MAXK_REQUIRE(n >= 0, "maxk_linear_wgrad: negative n");
    MAXK_REQUIRE(d == WG_M, "maxk_linear_wgrad: output width d=%d must be 256", d);
    MAXK_REQUIRE(in_dim >= 4 && in_dim <= 256 && in_dim % 4 == 0,
                 "maxk_linear_wgrad: in_dim=%d must be a multiple of 4 in [4, 256]", in_dim);
    MAXK_REQUIRE(dw != nullptr && (reinterpret_cast<uintptr_t>(dw) & 15) == 0,
                 "maxk_linear_wgrad: dw must be non-NULL and 16-byte aligned");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (!accumulate) MAXK_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)d * in_dim, st));
    if (n == 0) return MAXK_OK;
    MAXK_REQUIRE(grad_out != nullptr && x != nullptr, "maxk_linear_wgrad: NULL operand");
    MAXK_REQUIRE((reinterpret_cast<uintptr_t>(grad_out) & 15) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0,
                 "maxk_linear_wgrad: grad_out and x must be 16-byte aligned");
    MAXK_REQUIRE(n < (int64_t)1 << 31, "maxk_linear_wgrad: n too large for one tensor map");

    CUtensorMap map_g, map_x;
    if (int rc = make_map_f32(&map_g, grad_out, n, d, WG_KC, 32, "grad_out", true)) return rc;
    if (int rc = make_map_f32(&map_x, x, n, in_dim, WG_KC, 32, "x", true)) return rc;

    static std::mutex mu;
    static std::set<int> opted;
    int dev = 0, sms = 0;
    MAXK_CUDA(cudaGetDevice(&dev));
    {
        std::lock_guard<std::mutex> lock(mu);
        if (!opted.count(dev)) {
            MAXK_CUDA(cudaFuncSetAttribute(linear_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           WgShared::total + 1024));
            opted.insert(dev);
        }
    }
    MAXK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int64_t chunks = ceil_div<int64_t>(n, WG_KC);
    const unsigned grid = (unsigned)(chunks < sms ? chunks : sms);
    linear_wgrad_kernel<<<grid, WG_THREADS, WgShared::total + 1024, st>>>(map_g, map_x, n, in_dim, dw);
    MAXK_LAUNCH_CHECK("linear_wgrad_kernel");
    return MAXK_OK;
}
