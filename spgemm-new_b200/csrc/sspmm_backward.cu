// sspmm_backward.cu -- backward outer-product sampled SSpMM for sm_100a:
//     dXc[j, l] += val[e] * dY[r, sp_index[j, l]]      for every edge e = (r -> j)
// i.e. (A^T dY) evaluated only at the positions MaxK kept (CBSR layout).
//
// Replaces spmm_kernel_opt2_sparse_backward_v3 (ref: kernels/spmm_maxk_backward.cu:15-115).
// Same outer-product (scatter) form -- the gather form would read a full 1 KB dY row per
// edge instead of k selector bytes -- mapped differently (DESIGN.md "backward kernel"):
//   * a warp owns a contiguous run of tasks and stages dY[row,:] in shared memory once per
//     row, not once per task;
//   * work inside a task is flattened to (edge, 4-slot group) items so that all 32 lanes
//     are busy for every k (the reference idles lanes for k < 32 and warps for k < 32);
//   * each item reads 4 selector bytes with one 32-bit load and emits ONE 128-bit vector
//     reduction (red.global.add.v4.f32) instead of four scalar atomics: E*k/4 REDGs
//     instead of E*k (ref: spmm_maxk_backward.cu:80,101);
//   * (idx,val) pairs are fetched coalesced one task ahead, the next dY row one row ahead;
//     UNROLL items per lane are in flight so the selector gathers overlap;
//   * 64-bit offsets throughout.
// dXc (4k B/node) and the selectors (k B/node) stay L2-resident; HBM traffic is the CSR
// stream plus one read of dY.
#include "common.cuh"

namespace maxk {

constexpr int BWD_WPB = 8;
constexpr int BWD_UNROLL = 4;
constexpr int BWD_STAGE = 64;

struct BwdEdgeRegs {
    int i0, i1;
    float v0, v1;
};
__device__ __forceinline__ void bwd_fetch_edges(BwdEdgeRegs &r, const int *__restrict__ idx,
                                                const float *__restrict__ val, int64_t loc, int len,
                                                int lane)
{
    r.i0 = r.i1 = 0;
    r.v0 = r.v1 = 0.f;
    if (lane < len) {
        r.i0 = ld_stream_s32(idx + loc + lane);
        r.v0 = ld_stream_f32(val + loc + lane);
    }
    if (lane + 32 < len) {
        r.i1 = ld_stream_s32(idx + loc + lane + 32);
        r.v1 = ld_stream_f32(val + loc + lane + 32);
    }
}
__device__ __forceinline__ void bwd_stage_edges(const BwdEdgeRegs &r, int2 *stage, int lane)
{
    __syncwarp();
    stage[lane] = make_int2(r.i0, __float_as_int(r.v0));
    stage[lane + 32] = make_int2(r.i1, __float_as_int(r.v1));
    __syncwarp();
}

// GLOG2 >= 0: k == 4 << GLOG2 (vector path, shifts);  GLOG2 == -1: k % 4 == 0, runtime
// divide;  GLOG2 == -2: any k, scalar reductions.
// A16: the selectors arrive as the uint16 bank plan of the planned forward kernel (cbsr_plan.cu) -- on
// the multi-GPU peer path the plan travels over NVLink IN PLACE of the selectors -- and the column of
// an entry is decoded from its accumulator word (plan_word_column).
template <int GLOG2, bool A16>
__device__ __forceinline__ void bwd_process_batch(const int2 *stage, int len, const float *dys,
                                                  const unsigned char *__restrict__ sp_index,
                                                  float *__restrict__ dxc, int k, int lane)
{
    if constexpr (GLOG2 == -2) {
        const int items = len * k;
        for (int i = lane; i < items; i += 32) {
            const int e = i / k, l = i - e * k;
            const int2 ev = stage[e];
            const int64_t base = (int64_t)ev.x * k + l;
            unsigned c;
            if constexpr (A16)
                c = plan_word_column(__ldg(reinterpret_cast<const unsigned short *>(sp_index) + base));
            else
                c = ld_gather_u8(sp_index + base);
            red_add_f1(dxc + base, __int_as_float(ev.y) * dys[c]);
        }
    } else if constexpr (A16) {
        const int groups = (GLOG2 >= 0) ? (1 << GLOG2) : (k >> 2);
        const int items = len * groups;
        for (int i0 = lane; i0 < items; i0 += 32 * BWD_UNROLL) {
            int2 ev[BWD_UNROLL];
            int q[BWD_UNROLL];
            uint2 s4[BWD_UNROLL];
#pragma unroll
            for (int u = 0; u < BWD_UNROLL; ++u) {
                const int i = i0 + 32 * u;
                if (i < items) {
                    int e;
                    if constexpr (GLOG2 >= 0) {
                        e = i >> GLOG2;
                        q[u] = i & (groups - 1);
                    } else {
                        e = i / groups;
                        q[u] = i - e * groups;
                    }
                    ev[u] = stage[e];
                }
            }
#pragma unroll
            for (int u = 0; u < BWD_UNROLL; ++u) {
                if (i0 + 32 * u < items)
                    s4[u] = __ldg(reinterpret_cast<const uint2 *>(
                                      reinterpret_cast<const unsigned short *>(sp_index) + (int64_t)ev[u].x * k) + q[u]);
            }
#pragma unroll
            for (int u = 0; u < BWD_UNROLL; ++u) {
                if (i0 + 32 * u < items) {
                    const float w = __int_as_float(ev[u].y);
                    const float g0 = dys[plan_word_column(s4[u].x & 0xffffu)];
                    const float g1 = dys[plan_word_column(s4[u].x >> 16)];
                    const float g2 = dys[plan_word_column(s4[u].y & 0xffffu)];
                    const float g3 = dys[plan_word_column(s4[u].y >> 16)];
                    red_add_f4(dxc + (int64_t)ev[u].x * k + 4 * q[u], w * g0, w * g1, w * g2, w * g3);
                }
            }
        }
    } else {
        const int groups = (GLOG2 >= 0) ? (1 << GLOG2) : (k >> 2);
        const int items = len * groups;
        for (int i0 = lane; i0 < items; i0 += 32 * BWD_UNROLL) {
            int2 ev[BWD_UNROLL];
            int q[BWD_UNROLL];
            unsigned s4[BWD_UNROLL];
#pragma unroll
            for (int u = 0; u < BWD_UNROLL; ++u) {
                const int i = i0 + 32 * u;
                if (i < items) {
                    int e;
                    if constexpr (GLOG2 >= 0) {
                        e = i >> GLOG2;
                        q[u] = i & (groups - 1);
                    } else {
                        e = i / groups;
                        q[u] = i - e * groups;
                    }
                    ev[u] = stage[e];
                }
            }
#pragma unroll
            for (int u = 0; u < BWD_UNROLL; ++u) {
                if (i0 + 32 * u < items)
                    s4[u] = ld_gather_u32(
                        reinterpret_cast<const unsigned *>(sp_index + (int64_t)ev[u].x * k) + q[u]);
            }
#pragma unroll
            for (int u = 0; u < BWD_UNROLL; ++u) {
                if (i0 + 32 * u < items) {
                    const float w = __int_as_float(ev[u].y);
                    const float g0 = dys[s4[u] & 0xff];
                    const float g1 = dys[(s4[u] >> 8) & 0xff];
                    const float g2 = dys[(s4[u] >> 16) & 0xff];
                    const float g3 = dys[s4[u] >> 24];
                    red_add_f4(dxc + (int64_t)ev[u].x * k + 4 * q[u], w * g0, w * g1, w * g2, w * g3);
                }
            }
        }
    }
}

// dY[next row,:] is loaded into registers while the current task is processed: on graphs of short
// rows the row changes with every task and the kernel otherwise waits a full DRAM latency for it
// (ncu, products shape, round 1: 27 % of the stall samples sat on the staging stores that consume
// this load).  Measured in round 2 (profiles/r02_variants.md): 3.29 -> 3.17 ms on the products shape
// at k = 32, 0.064 -> 0.055 ms on the Flickr shape, 4.99 -> 4.74 ms on the Reddit shape at k = 64,
// unchanged at k = 32 / 16 there -- so it is the only form of the kernel.
template <int GLOG2, bool A16>
__global__ void __launch_bounds__(BWD_WPB * 32)
sspmm_bwd_kernel(const int4 *__restrict__ warp4, int64_t num_tasks, const int *__restrict__ idx,
                 const float *__restrict__ val, const float *__restrict__ dy,
                 const unsigned char *__restrict__ sp_index, int k, int d, float *__restrict__ dxc,
                 int tasks_per_warp)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float *dys = reinterpret_cast<float *>(smem_raw) + warp * d;
    int2 *stage = reinterpret_cast<int2 *>(reinterpret_cast<float *>(smem_raw) + BWD_WPB * d) +
                  warp * BWD_STAGE;

    const int64_t t0 = ((int64_t)blockIdx.x * BWD_WPB + warp) * tasks_per_warp;
    if (t0 >= num_tasks) return;
    const int64_t t1 = min(t0 + (int64_t)tasks_per_warp, num_tasks);

    int4 info = ld_stream_int4(warp4 + t0);
    BwdEdgeRegs pre;
    bwd_fetch_edges(pre, idx, val, info.y, min(info.z, BWD_STAGE), lane);

    int cur_row = -1;
    float4 pf0 = make_float4(0.f, 0.f, 0.f, 0.f), pf1 = pf0;  // the next row, columns 4*lane.. and 128+4*lane..
    {
        const float *src = dy + (int64_t)info.x * d;
        if (lane * 4 < d) pf0 = ld_stream_f4(reinterpret_cast<const float4 *>(src + lane * 4));
        if (lane * 4 + 128 < d) pf1 = ld_stream_f4(reinterpret_cast<const float4 *>(src + lane * 4 + 128));
    }
    for (int64_t t = t0; t < t1; ++t) {
        const bool has_next = (t + 1 < t1);
        int4 ninfo = info;
        if (has_next) ninfo = ld_stream_int4(warp4 + t + 1);

        if (info.x != cur_row) {  // stage dY[row,:] once per row: it is already here
            cur_row = info.x;
            __syncwarp();
            if (lane * 4 < d) *reinterpret_cast<float4 *>(dys + lane * 4) = pf0;
            if (lane * 4 + 128 < d) *reinterpret_cast<float4 *>(dys + lane * 4 + 128) = pf1;
            // visibility is ordered by the __syncwarp pair in bwd_stage_edges below
        }
        if (has_next && ninfo.x != cur_row) {  // rows ascend: the next different row is the next task's
            const float *src = dy + (int64_t)ninfo.x * d;
            if (lane * 4 < d) pf0 = ld_stream_f4(reinterpret_cast<const float4 *>(src + lane * 4));
            if (lane * 4 + 128 < d) pf1 = ld_stream_f4(reinterpret_cast<const float4 *>(src + lane * 4 + 128));
        }
        const int64_t loc = info.y;
        const int len = info.z;
        bwd_stage_edges(pre, stage, lane);
        if (has_next) bwd_fetch_edges(pre, idx, val, ninfo.y, min(ninfo.z, BWD_STAGE), lane);
        bwd_process_batch<GLOG2, A16>(stage, min(len, BWD_STAGE), dys, sp_index, dxc, k, lane);
        for (int off = BWD_STAGE; off < len; off += BWD_STAGE) {
            BwdEdgeRegs more;
            const int ln = min(len - off, BWD_STAGE);
            bwd_fetch_edges(more, idx, val, loc + off, ln, lane);
            bwd_stage_edges(more, stage, lane);
            bwd_process_batch<GLOG2, A16>(stage, ln, dys, sp_index, dxc, k, lane);
        }
        info = ninfo;
    }
}

template <int GLOG2, bool A16>
static int launch_bwd(const int32_t *warp4, int64_t W, const int32_t *idx, const float *val,
                      const float *dy, const uint8_t *sp_index, int k, int d, float *dxc,
                      int tasks_per_warp, cudaStream_t st)
{
    const size_t smem = (size_t)BWD_WPB * d * sizeof(float) + (size_t)BWD_WPB * BWD_STAGE * sizeof(int2);
    const int64_t chunks = ceil_div<int64_t>(W, tasks_per_warp);
    const int64_t grid = ceil_div<int64_t>(chunks, BWD_WPB);
    if (grid > 0x7fffffffLL) {
        set_error("maxk_sspmm_backward: grid too large");
        return MAXK_ERR_UNSUPPORTED;
    }
    sspmm_bwd_kernel<GLOG2, A16><<<(unsigned)grid, BWD_WPB * 32, smem, st>>>(
        reinterpret_cast<const int4 *>(warp4), W, idx, val, dy, sp_index, k, d, dxc, tasks_per_warp);
    MAXK_LAUNCH_CHECK("sspmm_bwd_kernel");
    return MAXK_OK;
}

int g_bwd_tasks_per_warp = 0;  // run length; 0 = automatic (as in the forward kernel)

}  // namespace maxk

using namespace maxk;

static int sspmm_backward_any(const char *who, const int32_t *warp4, int64_t num_tasks, const int32_t *indices,
                              const float *values, const float *dy, const void *sel, bool addr16, int32_t k,
                              int32_t d, float *dxc, int64_t num_cbsr_rows, int32_t accumulate, void *stream)
{
    MAXK_REQUIRE(num_tasks >= 0 && num_cbsr_rows >= 0, "%s: negative size", who);
    MAXK_REQUIRE(d > 0 && d <= MAXK_MAX_DIM && d % 4 == 0, "%s: d=%d must be a multiple of 4 in [4,256]", who, d);
    MAXK_REQUIRE(k >= 1 && k <= d, "%s: k=%d must be in [1,d=%d]", who, k, d);
    MAXK_REQUIRE(dxc != nullptr || num_cbsr_rows == 0, "%s: dxc is NULL", who);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (num_cbsr_rows == 0) return MAXK_OK;
    if (!accumulate) {
        MAXK_CUDA(cudaMemsetAsync(dxc, 0, (size_t)num_cbsr_rows * k * sizeof(float), st));
        count_launch();
    }
    if (num_tasks == 0) return MAXK_OK;
    MAXK_REQUIRE(warp4 && indices && values && dy && sel, "%s: NULL input pointer", who);
    MAXK_REQUIRE((reinterpret_cast<uintptr_t>(dy) & 15) == 0 && (reinterpret_cast<uintptr_t>(warp4) & 15) == 0,
                 "%s: dy and warp4 must be 16-byte aligned", who);
    int tpw = g_bwd_tasks_per_warp;
    if (tpw <= 0) tpw = auto_tasks_per_warp(num_tasks);
    const uint8_t *sp_index = static_cast<const uint8_t *>(sel);
    const bool vec = (k % 4 == 0) && ((reinterpret_cast<uintptr_t>(dxc) & 15) == 0) &&
                     ((reinterpret_cast<uintptr_t>(sel) & (addr16 ? 7 : 3)) == 0);
#define MAXK_BWD_GO(G)                                                                                      \
    return addr16 ? launch_bwd<G, true>(warp4, num_tasks, indices, values, dy, sp_index, k, d, dxc, tpw, st) \
                  : launch_bwd<G, false>(warp4, num_tasks, indices, values, dy, sp_index, k, d, dxc, tpw, st)
    if (!vec) MAXK_BWD_GO(-2);
    switch (k) {
        case 4: MAXK_BWD_GO(0);
        case 8: MAXK_BWD_GO(1);
        case 16: MAXK_BWD_GO(2);
        case 32: MAXK_BWD_GO(3);
        case 64: MAXK_BWD_GO(4);
        case 128: MAXK_BWD_GO(5);
        case 256: MAXK_BWD_GO(6);
        default: MAXK_BWD_GO(-1);
    }
#undef MAXK_BWD_GO
}

extern "C" int maxk_sspmm_backward(const int32_t *warp4, int64_t num_tasks, const int32_t *indices,
                                   const float *values, const float *dy, const uint8_t *sp_index,
                                   int32_t k, int32_t d, float *dxc, int64_t num_cbsr_rows,
                                   int32_t accumulate, void *stream)
{
    return sspmm_backward_any("maxk_sspmm_backward", warp4, num_tasks, indices, values, dy, sp_index, false, k, d,
                              dxc, num_cbsr_rows, accumulate, stream);
}

extern "C" int maxk_sspmm_backward_planned(const int32_t *warp4, int64_t num_tasks, const int32_t *indices,
                                           const float *values, const float *dy, const uint16_t *sp_addr,
                                           int32_t k, int32_t d, float *dxc, int64_t num_cbsr_rows,
                                           int32_t accumulate, void *stream)
{
    MAXK_REQUIRE(sp_addr == nullptr || (reinterpret_cast<uintptr_t>(sp_addr) & 1) == 0,
                 "maxk_sspmm_backward_planned: sp_addr must be 2-byte aligned");
    return sspmm_backward_any("maxk_sspmm_backward_planned", warp4, num_tasks, indices, values, dy, sp_addr, true, k,
                              d, dxc, num_cbsr_rows, accumulate, stream);
}
