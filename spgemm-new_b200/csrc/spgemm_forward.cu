// spgemm_forward.cu -- forward row-wise-product SpGEMM  Y = A * X_cbsr  for sm_100a.
//
// Replaces spmm_kernel_opt2_sparse_v3 (ref: kernels/spmm_maxk.cu:17-106).  Same math,
// different machine mapping (DESIGN.md "forward kernel"):
//   * a warp owns a CONTIGUOUS run of tasks from the edge-balanced work list and keeps
//     one output row (d floats) in shared memory across consecutive tasks of the same
//     row, so a row is written once per run instead of once per 64-nnz task;
//   * rows wholly inside a run are written with plain 128-bit streaming stores; only
//     rows that straddle two runs use 128-bit vector reductions (red.global.add.v4.f32);
//   * the task's (idx,val) pairs are fetched coalesced one task ahead and broadcast from
//     shared memory; neighbour CBSR rows are gathered UNROLL edges at a time (one 128 B
//     value line + one k-byte selector line per edge) before the shared-memory
//     accumulate, so each warp keeps 2*UNROLL independent gathers in flight;
//   * 64-bit row offsets (ref uses int: overflow beyond 8.39 M nodes).
// HBM-bound by design: CSR stream + one write of Y; the CBSR operand (5k B/node) lives
// in the 126 MB L2.  No tensor cores: 2 flops per >= 5 bytes gathered.
#include "common.cuh"

namespace maxk {

constexpr int FWD_WPB = 8;     // warps per CTA
// edges gathered per batch: gathers in flight per lane stay ~8, so fewer edges per batch
// when a lane owns more entries of each CBSR row
#ifndef MAXK_FWD_UNROLL1
#define MAXK_FWD_UNROLL1 4  // k <= 32
#endif
#ifndef MAXK_FWD_MINB1
#define MAXK_FWD_MINB1 4    // CTAs/SM the register budget is capped for (k <= 64)
#endif
__host__ __device__ constexpr int fwd_unroll(int epl)
{
    return epl == 1 ? MAXK_FWD_UNROLL1
                    : (epl == 2 ? (MAXK_FWD_UNROLL1 + 1) / 2 : (epl <= 4 ? 2 : 1));
}
__host__ __device__ constexpr int fwd_min_blocks(int epl) { return epl <= 2 ? MAXK_FWD_MINB1 : 2; }
constexpr int FWD_STAGE = 64;  // staged (idx,val) pairs per warp == MAXK_WARP_MAX_NZ

#ifndef MAXK_FWD_SYNCWARP
#define MAXK_FWD_SYNCWARP 1
#endif

__device__ __forceinline__ void fwd_flush_row(float *acc, float *__restrict__ out, int64_t row,
                                              int d, int lane, bool shared_row)
{
    __syncwarp();
    float *dst = out + row * (int64_t)d;
    for (int c = lane * 4; c < d; c += 128) {
        float4 v = *reinterpret_cast<float4 *>(acc + c);
        *reinterpret_cast<float4 *>(acc + c) = make_float4(0.f, 0.f, 0.f, 0.f);
        if (shared_row)
            red_add_f4(dst + c, v.x, v.y, v.z, v.w);
        else
            st_stream_f4(reinterpret_cast<float4 *>(dst + c), v);
    }
    __syncwarp();
}

// Per-lane slice of one neighbour's CBSR row: EPL entries.
template <int EPL, bool EXACT>
struct Slice {
    float v[EPL];
    unsigned char s[EPL];
};

template <int EPL, bool EXACT>
__device__ __forceinline__ void fwd_gather(Slice<EPL, EXACT> &sl, const float *__restrict__ sp_data,
                                           const unsigned char *__restrict__ sp_index, int col,
                                           int k, int lane)
{
    const int64_t base = (int64_t)col * k;
    if constexpr (EXACT && EPL == 1) {
        sl.v[0] = ld_gather_f32(sp_data + base + lane);
        sl.s[0] = ld_gather_u8(sp_index + base + lane);
    } else if constexpr (EXACT && EPL == 2) {
        // lane owns entries 2*lane, 2*lane+1: one 64-bit + one 16-bit load
        float2 f = __ldg(reinterpret_cast<const float2 *>(sp_data + base) + lane);
        unsigned short u = __ldg(reinterpret_cast<const unsigned short *>(sp_index + base) + lane);
        sl.v[0] = f.x; sl.v[1] = f.y;
        sl.s[0] = (unsigned char)(u & 0xff); sl.s[1] = (unsigned char)(u >> 8);
    } else if constexpr (EXACT && EPL == 4) {
        float4 f = ld_gather_f4(reinterpret_cast<const float4 *>(sp_data + base) + lane);
        unsigned u = ld_gather_u32(reinterpret_cast<const unsigned *>(sp_index + base) + lane);
        sl.v[0] = f.x; sl.v[1] = f.y; sl.v[2] = f.z; sl.v[3] = f.w;
        sl.s[0] = u & 0xff; sl.s[1] = (u >> 8) & 0xff; sl.s[2] = (u >> 16) & 0xff; sl.s[3] = u >> 24;
    } else {
#pragma unroll
        for (int i = 0; i < EPL; ++i) {
            const int l = lane + 32 * i;
            if (EXACT || l < k) {
                sl.v[i] = ld_gather_f32(sp_data + base + l);
                sl.s[i] = ld_gather_u8(sp_index + base + l);
            } else {
                sl.v[i] = 0.f;
                sl.s[i] = 0;
            }
        }
    }
}

template <int EPL, bool EXACT>
__device__ __forceinline__ void fwd_accumulate(const Slice<EPL, EXACT> &sl, float *acc, float w,
                                               int k, int lane)
{
#pragma unroll
    for (int i = 0; i < EPL; ++i) {
        // entries handled by one instruction belong to ONE CBSR row -> distinct columns
        // (precondition Q4), so the plain read-modify-write cannot collide inside a warp.
        bool on = true;
        if constexpr (!EXACT) on = (lane + 32 * i) < k;
        if (on) {
            float *a = acc + sl.s[i];
            *a = fmaf(w, sl.v[i], *a);  // FFMA, like the reference
        }
#if MAXK_FWD_SYNCWARP
        __syncwarp();
#endif
    }
}

template <int EPL, bool EXACT>
__device__ __forceinline__ void fwd_process_batch(const int2 *stage, int len, float *acc,
                                                  const float *__restrict__ sp_data,
                                                  const unsigned char *__restrict__ sp_index,
                                                  int k, int lane)
{
    constexpr int FWD_UNROLL = fwd_unroll(EPL);
    int j = 0;
    for (; j + FWD_UNROLL <= len; j += FWD_UNROLL) {
        int2 ev[FWD_UNROLL];
        Slice<EPL, EXACT> sl[FWD_UNROLL];
#pragma unroll
        for (int u = 0; u < FWD_UNROLL; ++u) ev[u] = stage[j + u];
#pragma unroll
        for (int u = 0; u < FWD_UNROLL; ++u) fwd_gather<EPL, EXACT>(sl[u], sp_data, sp_index, ev[u].x, k, lane);
#pragma unroll
        for (int u = 0; u < FWD_UNROLL; ++u)
            fwd_accumulate<EPL, EXACT>(sl[u], acc, __int_as_float(ev[u].y), k, lane);
    }
    if (j < len) {  // predicated tail group (len is warp-uniform)
        int2 ev[FWD_UNROLL];
        Slice<EPL, EXACT> sl[FWD_UNROLL];
#pragma unroll
        for (int u = 0; u < FWD_UNROLL; ++u)
            if (j + u < len) ev[u] = stage[j + u];
#pragma unroll
        for (int u = 0; u < FWD_UNROLL; ++u)
            if (j + u < len) fwd_gather<EPL, EXACT>(sl[u], sp_data, sp_index, ev[u].x, k, lane);
#pragma unroll
        for (int u = 0; u < FWD_UNROLL; ++u)
            if (j + u < len) fwd_accumulate<EPL, EXACT>(sl[u], acc, __int_as_float(ev[u].y), k, lane);
    }
}

// coalesced fetch of up to FWD_STAGE (idx,val) pairs of one task into registers
struct EdgeRegs {
    int i0, i1;
    float v0, v1;
};
__device__ __forceinline__ void fwd_fetch_edges(EdgeRegs &r, const int *__restrict__ idx,
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
__device__ __forceinline__ void fwd_stage_edges(const EdgeRegs &r, int2 *stage, int lane)
{
    __syncwarp();  // everyone is done reading the previous batch
    stage[lane] = make_int2(r.i0, __float_as_int(r.v0));
    stage[lane + 32] = make_int2(r.i1, __float_as_int(r.v1));
    __syncwarp();
}

template <int EPL, bool EXACT>
__global__ void __launch_bounds__(FWD_WPB * 32, fwd_min_blocks(EPL))
spgemm_fwd_kernel(const int4 *__restrict__ warp4, int64_t num_tasks, const int *__restrict__ idx,
                  const float *__restrict__ val, const float *__restrict__ sp_data,
                  const unsigned char *__restrict__ sp_index, int k, float *__restrict__ out, int d,
                  int tasks_per_warp, int accumulate)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float *acc = reinterpret_cast<float *>(smem_raw) + warp * d;
    int2 *stage = reinterpret_cast<int2 *>(reinterpret_cast<float *>(smem_raw) + FWD_WPB * d) +
                  warp * FWD_STAGE;

    const int64_t t0 = ((int64_t)blockIdx.x * FWD_WPB + warp) * tasks_per_warp;
    if (t0 >= num_tasks) return;  // no CTA-wide barrier below
    const int64_t t1 = min(t0 + (int64_t)tasks_per_warp, num_tasks);

    for (int c = lane * 4; c < d; c += 128)
        *reinterpret_cast<float4 *>(acc + c) = make_float4(0.f, 0.f, 0.f, 0.f);

    int4 info = ld_stream_int4(warp4 + t0);
    const int first_row = info.x;
    // accumulate mode (multi-GPU local/remote split): every row is added to what is there
    const bool first_shared = accumulate || ((t0 > 0) && (__ldg(&warp4[t0 - 1].x) == first_row));
    bool last_shared = accumulate;
    if (!accumulate && t1 < num_tasks)
        last_shared = (__ldg(&warp4[t1].x) == __ldg(&warp4[t1 - 1].x));

    EdgeRegs pre;
    fwd_fetch_edges(pre, idx, val, info.y, min(info.z, FWD_STAGE), lane);

    int cur_row = first_row;
    for (int64_t t = t0; t < t1; ++t) {
        const bool has_next = (t + 1 < t1);
        int4 ninfo = info;
        if (has_next) ninfo = ld_stream_int4(warp4 + t + 1);

        if (info.x != cur_row) {
            fwd_flush_row(acc, out, cur_row, d, lane,
                          accumulate || (first_shared && cur_row == first_row));
            cur_row = info.x;
        }
        const int64_t loc = info.y;
        const int len = info.z;
        const int len0 = min(len, FWD_STAGE);
        fwd_stage_edges(pre, stage, lane);
        if (has_next)  // next task's CSR slice streams in while this task's gathers run
            fwd_fetch_edges(pre, idx, val, ninfo.y, min(ninfo.z, FWD_STAGE), lane);
        fwd_process_batch<EPL, EXACT>(stage, len0, acc, sp_data, sp_index, k, lane);
        for (int off = FWD_STAGE; off < len; off += FWD_STAGE) {  // len > 64: foreign work lists
            EdgeRegs more;
            const int ln = min(len - off, FWD_STAGE);
            fwd_fetch_edges(more, idx, val, loc + off, ln, lane);
            fwd_stage_edges(more, stage, lane);
            fwd_process_batch<EPL, EXACT>(stage, ln, acc, sp_data, sp_index, k, lane);
        }
        info = ninfo;
    }
    fwd_flush_row(acc, out, cur_row, d, lane,
                  last_shared || (first_shared && cur_row == first_row));
}

template <int EPL, bool EXACT>
static int launch_fwd(const int32_t *warp4, int64_t W, const int32_t *idx, const float *val,
                      const float *sp_data, const uint8_t *sp_index, int k, float *out, int d,
                      int tasks_per_warp, int accumulate, cudaStream_t st)
{
    const size_t smem = (size_t)FWD_WPB * d * sizeof(float) + (size_t)FWD_WPB * FWD_STAGE * sizeof(int2);
    const int64_t chunks = ceil_div<int64_t>(W, tasks_per_warp);
    const int64_t grid = ceil_div<int64_t>(chunks, FWD_WPB);
    if (grid > 0x7fffffffLL) {
        set_error("maxk_spgemm_forward: grid too large");
        return MAXK_ERR_UNSUPPORTED;
    }
    spgemm_fwd_kernel<EPL, EXACT><<<(unsigned)grid, FWD_WPB * 32, smem, st>>>(
        reinterpret_cast<const int4 *>(warp4), W, idx, val, sp_data, sp_index, k, out, d,
        tasks_per_warp, accumulate);
    MAXK_LAUNCH_CHECK("spgemm_fwd_kernel");
    return MAXK_OK;
}

int g_fwd_tasks_per_warp = 16;  // tunable through maxk_set_tuning (bench sweeps)

}  // namespace maxk

using namespace maxk;

extern "C" int maxk_spgemm_forward(const int32_t *warp4, int64_t num_tasks, const int32_t *indices,
                                   const float *values, const float *sp_data,
                                   const uint8_t *sp_index, int32_t k, float *out,
                                   int64_t num_rows_out, int32_t d, void *stream)
{
    return maxk_spgemm_forward_acc(warp4, num_tasks, indices, values, sp_data, sp_index, k, out,
                                   num_rows_out, d, /*accumulate=*/0, stream);
}

extern "C" int maxk_spgemm_forward_acc(const int32_t *warp4, int64_t num_tasks,
                                       const int32_t *indices, const float *values,
                                       const float *sp_data, const uint8_t *sp_index, int32_t k,
                                       float *out, int64_t num_rows_out, int32_t d,
                                       int32_t accumulate, void *stream)
{
    MAXK_REQUIRE(num_tasks >= 0 && num_rows_out >= 0, "maxk_spgemm_forward: negative size");
    MAXK_REQUIRE(d > 0 && d <= MAXK_MAX_DIM && d % 4 == 0,
                 "maxk_spgemm_forward: d=%d must be a multiple of 4 in [4,256]", d);
    MAXK_REQUIRE(k >= 1 && k <= d, "maxk_spgemm_forward: k=%d must be in [1,d=%d]", k, d);
    MAXK_REQUIRE(out != nullptr || num_rows_out == 0, "maxk_spgemm_forward: out is NULL");
    MAXK_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0 &&
                     (reinterpret_cast<uintptr_t>(warp4) & 15) == 0,
                 "maxk_spgemm_forward: out and warp4 must be 16-byte aligned");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (num_rows_out == 0) return MAXK_OK;
    // Output is DEFINED by this call: empty rows and rows shared between
    // two warps' runs start from zero.
    if (!accumulate) {
        MAXK_CUDA(cudaMemsetAsync(out, 0, (size_t)num_rows_out * d * sizeof(float), st));
        count_launch();
    }
    if (num_tasks == 0) return MAXK_OK;
    MAXK_REQUIRE(warp4 && indices && values && sp_data && sp_index,
                 "maxk_spgemm_forward: NULL input pointer");
    const int epl = (k + 31) / 32;
    const bool exact = (k % 32 == 0);
    const bool vec_ok = ((reinterpret_cast<uintptr_t>(sp_data) & 15) == 0) &&
                        ((reinterpret_cast<uintptr_t>(sp_index) & 3) == 0);
    const int tpw = g_fwd_tasks_per_warp;
#define MAXK_FWD_CASE(E)                                                                          \
    case E:                                                                                       \
        return (exact && vec_ok)                                                                  \
                   ? launch_fwd<E, true>(warp4, num_tasks, indices, values, sp_data, sp_index, k, \
                                         out, d, tpw, accumulate, st)                             \
                   : launch_fwd<E, false>(warp4, num_tasks, indices, values, sp_data, sp_index,   \
                                          k, out, d, tpw, accumulate, st);
    switch (epl) {
        MAXK_FWD_CASE(1)
        MAXK_FWD_CASE(2)
        MAXK_FWD_CASE(3)
        MAXK_FWD_CASE(4)
        MAXK_FWD_CASE(5)
        MAXK_FWD_CASE(6)
        MAXK_FWD_CASE(7)
        MAXK_FWD_CASE(8)
    }
#undef MAXK_FWD_CASE
    set_error("maxk_spgemm_forward: unsupported k=%d", k);
    return MAXK_ERR_UNSUPPORTED;
}
