// sspmm_backward_planned.cu -- backward sampled SSpMM through the bank plan of cbsr_plan.cu.
//
// EXPERIMENTAL (written when round 1's GPU budget was spent: compiled and checked under the SIMT
// interpreter of tests/simt, NOT yet run on a GPU; reachable only through
// maxk_sspmm_backward_planned, which nothing calls by default).
//
// Why.  ncu of sspmm_bwd_kernel<3> on the Reddit shape (gpurun_out/prof_bwd_k32.ncu-rep, round 1):
// the L1 data pipe is 90 % busy with 5.83 wavefronts per edge, and 2.23 of them are bank-conflict
// replays of the dY lookups (l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld = 255 M for
// 114.6 M edges): 32 lanes read 32 random columns of the staged dY row, the fullest bank serves
// 3.2 of them.  The forward's bank plan fixes the same problem for its scatter; it can serve the
// backward too if one lookup instruction covers the 32 entries of ONE edge (the grouping the plan
// was made for) instead of 8 entries of each of 4 edges:
//   * dY[row,:] is staged once per row in all H homes (word c of copy 0, plan_home_word(c, h) of
//     copy h), so entry l of CBSR row j finds its value at the planned word sp_addr[j,l] --
//     fullest bank 2.0 / 1.45 / 1.02 for H = 2 / 4 / 8 instead of 3.27;
//   * lane l handles entry 32*i + l of the edge: one 16-bit plan gather (64 B per edge at k = 32,
//     one wavefront), one shared-memory lookup, one scalar reduction -- the 32 lanes of an
//     instruction add to 32 consecutive floats of dXc (one 128-byte request, like the 128-bit
//     vector reductions of the default kernel, four edges at a time there, one here);
//   * (idx,val) pairs are staged per task and broadcast four edges at a time (two LDS.128).
// Inner loop: 35 instructions per 4 edges (default kernel: 12.2 per edge overall).
// Model (profiles/r01_static_forward_model.md): 0.5 (broadcast) + 1 (plan line) + fullest bank +
// 1 (reduction) + 0.2 (streams) = 3.7 wavefronts per edge at H = 8 against 5.83 measured:
// 2.58 -> ~1.7 ms on the Reddit shape if that pipe stays the bound -- the unknown is the L2, which
// now sees 32 scalar reductions per 128-byte line instead of 8 vector ones.
// k = 32, 64, 96, 128; k = 16 and 8 carry 2 / 4 edges per instruction with one staged copy per edge slot.
#include "common.cuh"

namespace maxk {

constexpr int BWP_WPB = 8;
constexpr int BWP_STAGE = 64;
constexpr int BWP_U = 4;  // edges in flight per warp

struct BwpEdgeRegs {
    int i0, i1;
    float v0, v1;
};
__device__ __forceinline__ void bwp_fetch_edges(BwpEdgeRegs &r, const int *__restrict__ idx,
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
__device__ __forceinline__ void bwp_stage_edges(const BwpEdgeRegs &r, int2 *stage, int lane)
{
    __syncwarp();  // everyone is done with the previous batch (and with the previous dY row)
    stage[lane] = make_int2(r.i0, __float_as_int(r.v0));
    stage[lane + 32] = make_int2(r.i1, __float_as_int(r.v1));
    __syncwarp();
}

// dY[row, 0..d) into every home of the warp's staging area (H * 256 words)
template <int H>
__device__ __forceinline__ void bwp_stage_row(float *dys, const float *__restrict__ src, int d, int lane)
{
    for (int c = lane * 4; c < d; c += 128) {
        const float4 v = ld_stream_f4(reinterpret_cast<const float4 *>(src + c));
        *reinterpret_cast<float4 *>(dys + c) = v;  // home 0
        // c % 4 == 0: columns c .. c+3 share c/8 and have c mod 8 = (c mod 8) + i
        const unsigned lo = c & 7, hi = c >> 3;
#pragma unroll
        for (int h = 1; h < H; ++h) {
            float *p = dys + h * 256;
            p[(lo + 0) * 32 + ((hi + plan_s(h) * (lo + 0) + plan_t(h)) & 31)] = v.x;
            p[(lo + 1) * 32 + ((hi + plan_s(h) * (lo + 1) + plan_t(h)) & 31)] = v.y;
            p[(lo + 2) * 32 + ((hi + plan_s(h) * (lo + 2) + plan_t(h)) & 31)] = v.z;
            p[(lo + 3) * 32 + ((hi + plan_s(h) * (lo + 3) + plan_t(h)) & 31)] = v.w;
        }
    }
}

template <int EPL>
__device__ __forceinline__ void bwp_process_batch(const int2 *stage, int len, const float *dys,
                                                  const unsigned short *__restrict__ sp_addr,
                                                  float *__restrict__ dxc, int lane)
{
    constexpr int K = 32 * EPL;
    static_assert(BWP_U == 4, "the full-group path reads four staged pairs as two 128-bit words");
    int j = 0;
    for (; j + BWP_U <= len; j += BWP_U) {  // four edges, no predicates
        const int4 e01 = *reinterpret_cast<const int4 *>(stage + j);      // warp-uniform address:
        const int4 e23 = *reinterpret_cast<const int4 *>(stage + j + 2);  // a broadcast
        const int col[BWP_U] = {e01.x, e01.z, e23.x, e23.z};
        const int wgt[BWP_U] = {e01.y, e01.w, e23.y, e23.w};
        unsigned short a[BWP_U][EPL];
#pragma unroll
        for (int u = 0; u < BWP_U; ++u) {
            const unsigned short *p = sp_addr + (uint64_t)(uint32_t)col[u] * (uint32_t)K + lane;
#pragma unroll
            for (int i = 0; i < EPL; ++i) a[u][i] = __ldg(p + 32 * i);
        }
#pragma unroll
        for (int u = 0; u < BWP_U; ++u) {
            const float w = __int_as_float(wgt[u]);
            float *q = dxc + (uint64_t)(uint32_t)col[u] * (uint32_t)K + lane;
#pragma unroll
            for (int i = 0; i < EPL; ++i) red_add_f1(q + 32 * i, w * dys[a[u][i]]);
        }
    }
    for (; j < len; ++j) {  // tail, one edge at a time
        const int2 ev = stage[j];
        const unsigned short *p = sp_addr + (uint64_t)(uint32_t)ev.x * (uint32_t)K + lane;
        float *q = dxc + (uint64_t)(uint32_t)ev.x * (uint32_t)K + lane;
        const float w = __int_as_float(ev.y);
#pragma unroll
        for (int i = 0; i < EPL; ++i) red_add_f1(q + 32 * i, w * dys[__ldg(p + 32 * i)]);
    }
}

// ---- variant that keeps the 128-bit vector reductions of the default kernel (knob: maxk_set_planned_
// backward_vector): work items are (edge, 4-entry group) as there, so one lookup instruction reads 8
// entries of each of 4 edges.  The plan still removes the conflicts INSIDE a row (its 8 entries sit in
// 8 different banks at H = 8); entries of different rows can still meet: fullest bank ~2.4 instead of
// 3.3.  The fallback if the L2 turns out to take 32 scalar reductions per line slower than 8 vector ones.
template <int EPL>
__device__ __forceinline__ void bwp_process_batch_vec(const int2 *stage, int len, const float *dys,
                                                      const unsigned short *__restrict__ sp_addr,
                                                      float *__restrict__ dxc, int lane)
{
    constexpr int K = 32 * EPL, GROUPS = K / 4;  // 4-entry groups per edge (a power of two times 1..4)
    const int items = len * GROUPS;
    for (int i0 = lane; i0 < items; i0 += 32 * BWP_U) {
        int2 ev[BWP_U];
        int q[BWP_U];
        uint2 a4[BWP_U];
#pragma unroll
        for (int u = 0; u < BWP_U; ++u) {
            const int i = i0 + 32 * u;
            if (i < items) {
                const int e = i / GROUPS;
                q[u] = i - e * GROUPS;
                ev[u] = stage[e];
            }
        }
#pragma unroll
        for (int u = 0; u < BWP_U; ++u)
            if (i0 + 32 * u < items)
                a4[u] = __ldg(reinterpret_cast<const uint2 *>(sp_addr + (uint64_t)(uint32_t)ev[u].x * (uint32_t)K) + q[u]);
#pragma unroll
        for (int u = 0; u < BWP_U; ++u) {
            if (i0 + 32 * u < items) {
                const float w = __int_as_float(ev[u].y);
                const float g0 = dys[a4[u].x & 0xffffu], g1 = dys[a4[u].x >> 16];
                const float g2 = dys[a4[u].y & 0xffffu], g3 = dys[a4[u].y >> 16];
                red_add_f4(dxc + (uint64_t)(uint32_t)ev[u].x * (uint32_t)K + 4 * q[u], w * g0, w * g1, w * g2, w * g3);
            }
        }
    }
}

// ---- k = 16, 8: E = 32/k edges per lookup instruction, one copy of the staged row per edge slot on
// its own 1/E of the banks (word * E + slot) -- the layout the plan of such a k was made for ----------
template <int H, int E>
__device__ __forceinline__ void bwp_stage_row_slots(float *dys, const float *__restrict__ src, int d, int lane)
{
    for (int c = lane; c < d; c += 32) {
        const float v = ld_stream_f32(src + c);
#pragma unroll
        for (int h = 0; h < H; ++h) {
            float *p = dys + plan_home_word((unsigned)c, h) * E;
#pragma unroll
            for (int s = 0; s < E; ++s) p[s] = v;
        }
    }
}

template <int KL>
__device__ __forceinline__ void bwp_process_batch_slots(const int2 *stage, int len, const float *dys,
                                                        const unsigned short *__restrict__ sp_addr,
                                                        float *__restrict__ dxc, int lane)
{
    constexpr int E = 32 / KL;
    const int slot = lane & (E - 1), entry = lane / E;
    for (int j = 0; j < len; j += BWP_U * E) {
        int2 ev[BWP_U];
        unsigned short a[BWP_U];
#pragma unroll
        for (int u = 0; u < BWP_U; ++u) {
            ev[u] = make_int2(0, 0);
            if (j + u * E + slot < len) ev[u] = stage[j + u * E + slot];
        }
#pragma unroll
        for (int u = 0; u < BWP_U; ++u)
            if (j + u * E + slot < len) a[u] = __ldg(sp_addr + (uint64_t)(uint32_t)ev[u].x * (uint32_t)KL + entry);
#pragma unroll
        for (int u = 0; u < BWP_U; ++u)
            if (j + u * E + slot < len)
                red_add_f1(dxc + (uint64_t)(uint32_t)ev[u].x * (uint32_t)KL + entry,
                           __int_as_float(ev[u].y) * dys[a[u] * E + slot]);
    }
}

// KL = 32: k = 32 * EPL;  KL = 16, 8: k = KL (EPL = 1);  VEC: the vector-reduction variant (KL = 32)
template <int KL, int EPL, int H, bool VEC = false>
__global__ void __launch_bounds__(BWP_WPB * 32)
sspmm_bwd_planned_kernel(const int4 *__restrict__ warp4, int64_t num_tasks, const int *__restrict__ idx,
                         const float *__restrict__ val, const float *__restrict__ dy,
                         const unsigned short *__restrict__ sp_addr, int d, float *__restrict__ dxc,
                         int tasks_per_warp)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int E = 32 / KL, ROW_WORDS = H * 256 * E;  // staged dY row: H homes x E edge slots
    float *dys = reinterpret_cast<float *>(smem_raw) + warp * ROW_WORDS;
    int2 *stage = reinterpret_cast<int2 *>(reinterpret_cast<float *>(smem_raw) + BWP_WPB * ROW_WORDS) +
                  warp * BWP_STAGE;

    const int64_t t0 = ((int64_t)blockIdx.x * BWP_WPB + warp) * tasks_per_warp;
    if (t0 >= num_tasks) return;  // no CTA-wide barrier below
    const int64_t t1 = min(t0 + (int64_t)tasks_per_warp, num_tasks);

    int4 info = ld_stream_int4(warp4 + t0);
    BwpEdgeRegs pre;
    bwp_fetch_edges(pre, idx, val, info.y, min(info.z, BWP_STAGE), lane);

    int cur_row = -1;
    for (int64_t t = t0; t < t1; ++t) {
        const bool has_next = (t + 1 < t1);
        int4 ninfo = info;
        if (has_next) ninfo = ld_stream_int4(warp4 + t + 1);

        if (info.x != cur_row) {  // stage dY[row,:] once per row, in every home
            cur_row = info.x;
            __syncwarp();  // lookups of the previous row are done
            if constexpr (KL == 32)
                bwp_stage_row<H>(dys, dy + (int64_t)cur_row * d, d, lane);
            else
                bwp_stage_row_slots<H, E>(dys, dy + (int64_t)cur_row * d, d, lane);
            // visibility is ordered by the __syncwarp pair in bwp_stage_edges below
        }
        const int64_t loc = info.y;
        const int len = info.z;
        bwp_stage_edges(pre, stage, lane);
        if (has_next) bwp_fetch_edges(pre, idx, val, ninfo.y, min(ninfo.z, BWP_STAGE), lane);
        if constexpr (VEC)
            bwp_process_batch_vec<EPL>(stage, min(len, BWP_STAGE), dys, sp_addr, dxc, lane);
        else if constexpr (KL == 32)
            bwp_process_batch<EPL>(stage, min(len, BWP_STAGE), dys, sp_addr, dxc, lane);
        else
            bwp_process_batch_slots<KL>(stage, min(len, BWP_STAGE), dys, sp_addr, dxc, lane);
        for (int off = BWP_STAGE; off < len; off += BWP_STAGE) {  // len > 64: foreign work lists
            BwpEdgeRegs more;
            const int ln = min(len - off, BWP_STAGE);
            bwp_fetch_edges(more, idx, val, loc + off, ln, lane);
            bwp_stage_edges(more, stage, lane);
            if constexpr (VEC)
                bwp_process_batch_vec<EPL>(stage, ln, dys, sp_addr, dxc, lane);
            else if constexpr (KL == 32)
                bwp_process_batch<EPL>(stage, ln, dys, sp_addr, dxc, lane);
            else
                bwp_process_batch_slots<KL>(stage, ln, dys, sp_addr, dxc, lane);
        }
        info = ninfo;
    }
}

template <int KL, int EPL, int H, bool VEC = false>
static int launch_bwp(const int32_t *warp4, int64_t W, const int32_t *idx, const float *val,
                      const float *dy, const uint16_t *sp_addr, int d, float *dxc, int tasks_per_warp,
                      cudaStream_t st)
{
    const size_t smem = (size_t)BWP_WPB * H * 256 * (32 / KL) * sizeof(float) +
                        (size_t)BWP_WPB * BWP_STAGE * sizeof(int2);
    if (smem > 48 * 1024)
        if (int rc = opt_in_dynamic_smem(reinterpret_cast<const void *>(&sspmm_bwd_planned_kernel<KL, EPL, H, VEC>),
                                         "sspmm_bwd_planned_kernel"))
            return rc;
    const int64_t chunks = ceil_div<int64_t>(W, tasks_per_warp);
    const int64_t grid = ceil_div<int64_t>(chunks, BWP_WPB);
    if (grid > 0x7fffffffLL) {
        set_error("maxk_sspmm_backward_planned: grid too large");
        return MAXK_ERR_UNSUPPORTED;
    }
    sspmm_bwd_planned_kernel<KL, EPL, H, VEC><<<(unsigned)grid, BWP_WPB * 32, smem, st>>>(
        reinterpret_cast<const int4 *>(warp4), W, idx, val, dy, sp_addr, d, dxc, tasks_per_warp);
    MAXK_LAUNCH_CHECK("sspmm_bwd_planned_kernel");
    return MAXK_OK;
}

extern int g_bwd_tasks_per_warp;
int g_bwp_vector = 0;  // 1: keep the 128-bit vector reductions (k = 32, 64)

}  // namespace maxk

using namespace maxk;

// Bench/test hook: 1 = the planned backward keeps 128-bit vector reductions (k = 32, 64), 0 = scalar
// reductions with one edge per lookup instruction (default).
extern "C" int maxk_set_planned_backward_vector(int on)
{
    g_bwp_vector = on ? 1 : 0;
    return MAXK_OK;
}

extern "C" int maxk_sspmm_backward_planned(const int32_t *warp4, int64_t num_tasks, const int32_t *indices,
                                           const float *values, const float *dy, const uint16_t *sp_addr,
                                           int32_t k, int32_t homes, int32_t d, float *dxc,
                                           int64_t num_cbsr_rows, int32_t accumulate, void *stream)
{
    MAXK_REQUIRE(num_tasks >= 0 && num_cbsr_rows >= 0, "maxk_sspmm_backward_planned: negative size");
    MAXK_REQUIRE(d > 0 && d <= MAXK_MAX_DIM && d % 4 == 0,
                 "maxk_sspmm_backward_planned: d=%d must be a multiple of 4 in [4,256]", d);
    MAXK_REQUIRE(k <= d && (k == 8 || k == 16 || (k >= 32 && k % 32 == 0 && k <= 128)),
                 "maxk_sspmm_backward_planned: k=%d must be 8, 16, 32, 64, 96 or 128, and <= d=%d", k, d);
    MAXK_REQUIRE(k >= 32 || homes == 2 || (k == 16 && homes == 4),
                 "maxk_sspmm_backward_planned: homes=%d is not built for k=%d (k = 16: 2, 4; k = 8: 2)", homes, k);
    MAXK_REQUIRE(homes == 2 || homes == 4 || homes == 8,
                 "maxk_sspmm_backward_planned: homes=%d must be 2, 4 or 8", homes);
    MAXK_REQUIRE(dxc != nullptr || num_cbsr_rows == 0, "maxk_sspmm_backward_planned: dxc is NULL");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (num_cbsr_rows == 0) return MAXK_OK;
    if (!accumulate) {
        MAXK_CUDA(cudaMemsetAsync(dxc, 0, (size_t)num_cbsr_rows * k * sizeof(float), st));
        count_launch();
    }
    if (num_tasks == 0) return MAXK_OK;
    MAXK_REQUIRE(warp4 && indices && values && dy && sp_addr, "maxk_sspmm_backward_planned: NULL input pointer");
    MAXK_REQUIRE((reinterpret_cast<uintptr_t>(dy) & 15) == 0 && (reinterpret_cast<uintptr_t>(warp4) & 15) == 0,
                 "maxk_sspmm_backward_planned: dy and warp4 must be 16-byte aligned");
    MAXK_REQUIRE((reinterpret_cast<uintptr_t>(sp_addr) & 1) == 0 && (reinterpret_cast<uintptr_t>(dxc) & 3) == 0,
                 "maxk_sspmm_backward_planned: sp_addr / dxc are misaligned");
    int tpw = g_bwd_tasks_per_warp;
    if (tpw <= 0) tpw = auto_tasks_per_warp(num_tasks);
    if (g_bwp_vector && (k == 32 || k == 64) && (reinterpret_cast<uintptr_t>(sp_addr) & 7) == 0 &&
        (reinterpret_cast<uintptr_t>(dxc) & 15) == 0) {
#define MAXK_BWP_VEC(E, HH) \
    return launch_bwp<32, E, HH, true>(warp4, num_tasks, indices, values, dy, sp_addr, d, dxc, tpw, st)
        switch (k / 32 * 16 + homes) {
            case 16 + 2: MAXK_BWP_VEC(1, 2);
            case 16 + 4: MAXK_BWP_VEC(1, 4);
            case 16 + 8: MAXK_BWP_VEC(1, 8);
            case 32 + 2: MAXK_BWP_VEC(2, 2);
            case 32 + 4: MAXK_BWP_VEC(2, 4);
            default: MAXK_BWP_VEC(2, 8);
        }
#undef MAXK_BWP_VEC
    }
#define MAXK_BWP_GO(E, HH) \
    return launch_bwp<32, E, HH>(warp4, num_tasks, indices, values, dy, sp_addr, d, dxc, tpw, st)
    if (k == 16 && homes == 2) return launch_bwp<16, 1, 2>(warp4, num_tasks, indices, values, dy, sp_addr, d, dxc, tpw, st);
    if (k == 16) return launch_bwp<16, 1, 4>(warp4, num_tasks, indices, values, dy, sp_addr, d, dxc, tpw, st);
    if (k == 8) return launch_bwp<8, 1, 2>(warp4, num_tasks, indices, values, dy, sp_addr, d, dxc, tpw, st);
#define MAXK_BWP_H(E)                \
    switch (homes) {                 \
        case 2: MAXK_BWP_GO(E, 2);   \
        case 4: MAXK_BWP_GO(E, 4);   \
        default: MAXK_BWP_GO(E, 8);  \
    }
    switch (k / 32) {
        case 1: MAXK_BWP_H(1)
        case 2: MAXK_BWP_H(2)
        case 3: MAXK_BWP_H(3)
        default: MAXK_BWP_H(4)
    }
#undef MAXK_BWP_H
#undef MAXK_BWP_GO
}
