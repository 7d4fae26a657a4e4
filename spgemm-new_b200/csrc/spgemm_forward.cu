// spgemm_forward.cu -- forward row-wise-product SpGEMM  Y = A * X_cbsr  for sm_100a.
//
// Replaces spmm_kernel_opt2_sparse_v3 (ref: kernels/spmm_maxk.cu:17-106).  Same math,
// different machine mapping (DESIGN.md "forward kernel"):
//   * a warp owns a CONTIGUOUS run of tasks from the edge-balanced work list and keeps
//     one output row in shared memory across consecutive tasks of the same row, so a row
//     is written once per run instead of once per 64-nnz task;
//   * rows wholly inside a run are written with plain streaming stores; only rows that
//     straddle two runs use vector reductions (red.global.add.v4.f32);
//   * the task's (idx,val) pairs are fetched coalesced one task ahead and broadcast from
//     shared memory; neighbour CBSR rows are gathered U instructions ahead (one 128 B value
//     line + one k-byte selector line per edge) before the shared-memory accumulate;
//   * k < 32 (k = 16, 8, ...): 32/k edges share one instruction (lane l -> edge l % E,
//     entry l / E), each edge slot accumulating into its OWN copy of the row, so all 32
//     lanes work for every k (the reference idles 32-k lanes per edge and, for k < 32, the
//     warps >= 12k/32);
//   * the kernel is bound by L1TEX data-pipe wavefronts (ncu: 93 % of peak): per edge one
//     value-line gather, one selector-line gather, one staging broadcast and a read-modify-
//     write whose 32 random columns cost ~3.3 wavefronts for the load and again for the
//     store.  Interleaved accumulator copies (acc[c*R + lane % R]) were built to cut those
//     conflicts and measured: they do not pay (the worst bank still serialises ~3 lanes
//     unless R = 32), so R = E is the default and the other values stay as a tuning knob;
//   * 64-bit row offsets (ref uses int: overflow beyond 8.39 M nodes).
// HBM traffic is the CSR stream + one write of Y; the CBSR operand (5k B/node) lives in
// the 126 MB L2.  No tensor cores: 2 flops per >= 5 bytes gathered.
#include <type_traits>

#include "common.cuh"
#include "peer.cuh"

namespace maxk {

constexpr int FWD_STAGE = 64;  // staged (idx,val) pairs per warp == MAXK_WARP_MAX_NZ

#ifndef MAXK_FWD_SYNCWARP
#define MAXK_FWD_SYNCWARP 1
#endif

// One kernel variant.  KL lanes per edge (32 for k >= 32 or odd k; k itself for k = 16,8,4,2,1),
// EPL entries per lane (k > 32), EXACT: k == KL*EPL with aligned operands (vector gathers),
// R accumulator copies, WPB warps per CTA, U gather instructions in flight, MINB CTAs/SM
// the register budget is capped for.
template <int KL_, int EPL_, bool EXACT_, int R_, int WPB_, int U_, int MINB_, bool PEER_ = false,
          int RING_ = 0, bool PLAN_ = false>
struct FwdCfg {
    // PLAN: the second operand is the bank plan sp_addr (uint16 accumulator words, cbsr_plan.cu)
    // instead of sp_index; the output row lives in PLAN_HOMES 256-word copies (k = 32, 64: long rows
    // on an L2-resident operand, where the update's bank conflicts are the bound)
    static constexpr bool PLAN = PLAN_;
    static constexpr int HOMES = PLAN_ ? PLAN_HOMES : 0;
    static constexpr int PLAN_WORDS = PLAN_HOMES * 256;
    static_assert(!PLAN_ || (EXACT_ && KL_ == 32 && R_ == 1 && RING_ == 0), "the planned variant: k % 32 == 0");
    // RING > 0: gathers go through cp.async into a per-warp shared-memory ring of 2 x RING rows
    // (operands that miss the L2: launch shapes 8 and 12); K is the compile-time k of the EXACT variants
    static constexpr int RING = RING_;
    static constexpr int K = (KL_ == 32) ? 32 * EPL_ : KL_;
    static constexpr int ROW_BYTES = 5 * K;  // k fp32 values followed by k selector bytes
    static_assert(RING_ == 0 || (EXACT_ && K % 16 == 0 && !PEER_), "ring rows move in 16-byte chunks");
    static constexpr int KL = KL_, EPL = EPL_, R = R_, WPB = WPB_, U = U_, MINB = MINB_;
    static constexpr bool EXACT = EXACT_;
    // PEER: fused multi-GPU variant -- the first CTAs of the grid push this rank's CBSR block
    // to the other ranks over NVLink while the rest aggregate; gathers then use coherent loads
    // because the operand is being written by the peers during the kernel (peer.cuh)
    static constexpr bool PEER = PEER_;
    static constexpr int E = 32 / KL_;  // edges per instruction
    static_assert(KL_ * E == 32 && (R_ % E) == 0 && R_ >= E, "copies must separate edge slots");
    static_assert(KL_ == 32 || EPL_ == 1, "sub-warp edges have one entry per lane");
};

template <class C>
struct Slice {
    float v[C::EPL];
    // column id (uint8) -- or, PLAN, the accumulator word of the entry (uint16)
    typename std::conditional<C::PLAN, unsigned short, unsigned char>::type s[C::EPL];
};

// read-only path for an operand that is constant during the kernel; plain (coherent) loads
// when other CTAs of the same kernel are still filling it (PEER)
template <bool PEER, typename T>
__device__ __forceinline__ T fwd_ld(const T *p)
{
    if constexpr (PEER)
        return *p;
    else
        return __ldg(p);
}

template <class C>
__device__ __forceinline__ void fwd_gather(Slice<C> &sl, const float *sp_data,
                                           const unsigned char *sp_index, int col, int k, int lane)
{
    constexpr bool P = C::PEER;
    const int64_t base = (int64_t)col * k;
    if constexpr (C::PLAN) {
        // entry 32*i + lane: the grouping the plan was made for; sp_index points at sp_addr.
        // k is the compile-time C::K here and column ids are non-negative: one unsigned 32x32->64
        // multiply per edge instead of the sign-extended 64-bit product of the generic path
        const unsigned short *sp_addr = reinterpret_cast<const unsigned short *>(sp_index);
        const uint64_t ubase = (uint64_t)(uint32_t)col * (uint32_t)C::K;
#pragma unroll
        for (int i = 0; i < C::EPL; ++i) {
            sl.v[i] = fwd_ld<P>(sp_data + ubase + lane + 32 * i);
            sl.s[i] = fwd_ld<P>(sp_addr + ubase + lane + 32 * i);
        }
    } else if constexpr (C::KL < 32) {
        const int entry = lane / C::E;
        sl.v[0] = fwd_ld<P>(sp_data + base + entry);
        sl.s[0] = fwd_ld<P>(sp_index + base + entry);
    } else if constexpr (C::EXACT && C::EPL == 1) {
        sl.v[0] = fwd_ld<P>(sp_data + base + lane);
        sl.s[0] = fwd_ld<P>(sp_index + base + lane);
    } else if constexpr (C::EXACT && C::EPL == 2) {
        // lane owns entries 2*lane, 2*lane+1: one 64-bit + one 16-bit load
        float2 f = fwd_ld<P>(reinterpret_cast<const float2 *>(sp_data + base) + lane);
        unsigned short u = fwd_ld<P>(reinterpret_cast<const unsigned short *>(sp_index + base) + lane);
        sl.v[0] = f.x; sl.v[1] = f.y;
        sl.s[0] = (unsigned char)(u & 0xff); sl.s[1] = (unsigned char)(u >> 8);
    } else if constexpr (C::EXACT && C::EPL == 4) {
        float4 f = fwd_ld<P>(reinterpret_cast<const float4 *>(sp_data + base) + lane);
        unsigned u = fwd_ld<P>(reinterpret_cast<const unsigned *>(sp_index + base) + lane);
        sl.v[0] = f.x; sl.v[1] = f.y; sl.v[2] = f.z; sl.v[3] = f.w;
        sl.s[0] = u & 0xff; sl.s[1] = (u >> 8) & 0xff; sl.s[2] = (u >> 16) & 0xff; sl.s[3] = u >> 24;
    } else {
#pragma unroll
        for (int i = 0; i < C::EPL; ++i) {
            const int l = lane + 32 * i;
            if (C::EXACT || l < k) {
                sl.v[i] = fwd_ld<P>(sp_data + base + l);
                sl.s[i] = fwd_ld<P>(sp_index + base + l);
            } else {
                sl.v[i] = 0.f;
                sl.s[i] = 0;
            }
        }
    }
}

// acc_lane = acc + lane % R.  Lanes of one instruction that share a copy belong to ONE
// CBSR row -> distinct columns (precondition Q4); lanes of different edges use different
// copies.  So the plain read-modify-write cannot collide inside a warp.
template <class C>
__device__ __forceinline__ void fwd_accumulate(const Slice<C> &sl, float *acc_lane, float w, int k,
                                               int lane, bool valid)
{
#pragma unroll
    for (int i = 0; i < C::EPL; ++i) {
        bool on = valid;
        if constexpr (C::KL == 32 && !C::EXACT) on = on && ((lane + 32 * i) < k);
        if (on) {
            float *a = acc_lane + (int)sl.s[i] * C::R;
            *a = fmaf(w, sl.v[i], *a);  // FFMA, like the reference
        }
#if MAXK_FWD_SYNCWARP
        __syncwarp();
#endif
    }
}

// (A rolling pipeline -- re-issue a slot's gather right after its accumulate -- was built and
// measured: 1.5-2x SLOWER on every shape.  The staging read, the gather and the accumulate of
// one slot then form one dependent chain per edge, whereas the batch form below issues U
// staging reads, U gathers and U accumulates back to back.  profiles/r01_forward_pipeline.md)
template <class C>
__device__ __forceinline__ void fwd_process_batch(const int2 *stage, int len, float *acc_lane,
                                                  const float *sp_data,
                                                  const unsigned char *sp_index, int k, int lane)
{
    constexpr int U = C::U, E = C::E, STEP = U * E;
    const int e = lane & (E - 1);
    int j = 0;
    for (; j + STEP <= len; j += STEP) {  // every lane has an edge
        int2 ev[U];
        Slice<C> sl[U];
#pragma unroll
        for (int u = 0; u < U; ++u) ev[u] = stage[j + u * E + e];
#pragma unroll
        for (int u = 0; u < U; ++u) fwd_gather<C>(sl[u], sp_data, sp_index, ev[u].x, k, lane);
#pragma unroll
        for (int u = 0; u < U; ++u)
            fwd_accumulate<C>(sl[u], acc_lane, __int_as_float(ev[u].y), k, lane, true);
    }
    if (j < len) {  // predicated tail group
        int2 ev[U];
        Slice<C> sl[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            ev[u] = make_int2(0, 0);
            if (j + u * E + e < len) ev[u] = stage[j + u * E + e];
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (j + u * E + e < len) fwd_gather<C>(sl[u], sp_data, sp_index, ev[u].x, k, lane);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (j + u * E < len)  // warp-uniform: skip instructions no lane needs
                fwd_accumulate<C>(sl[u], acc_lane, __int_as_float(ev[u].y), k, lane,
                                  j + u * E + e < len);
        }
    }
}

// Launch shapes 8 and 12 (chosen automatically when the CBSR operand cannot live in the L2, see
// maxk_spgemm_forward_acc): gathers through LDGSTS (cp.async, 16-byte chunks) into a per-warp
// shared-memory ring of two groups of RING rows: loads in flight hold no registers, so a warp keeps
// up to 2 x RING CBSR rows in flight (the register path: U = 4) and a task of <= 32 edges issues all
// its gathers before it waits once.  On the products shape (392 MB operand, gathers pay DRAM latency)
// the register path keeps ~36 gathers in flight per SM where 64+ saturate the memory system:
// measured 4.01 -> 3.58 ms at k = 32, 3.39 -> 2.61 at k = 16, 6.35 -> 4.97 at k = 64
// (profiles/r02_variants.md).  Cost: the gathered row crosses the L1 data pipe twice more
// (shared-memory write + read) -- a loss where that pipe is the bound (L2-resident operands:
// 3.99 -> 5.27 ms on the Reddit shape), hence the choice by operand size.
// Lanes map to (row, 16-byte chunk) by shifts: K/4 lanes copy the values of one row (so one instruction
// covers 32 / (K/4) rows), K/16 lanes its selectors.  (Round 2, ncu of the products shape: the first form
// of this function -- one flat chunk index per lane, divided by the chunks per row -- and of the accumulate
// below made the kernel ISSUE-bound at 53 instructions per edge, 31 of them in the accumulate.)
template <class C>
__device__ __forceinline__ void fwd_ring_issue(const int2 *stage, int e0, int cnt, unsigned char *half,
                                               const float *sp_data, const unsigned char *sp_index,
                                               int lane)
{
    constexpr int K = C::K, LV = K / 4, LS = K / 16;  // lanes per row: values, selectors
    static_assert(LV <= 32 && (32 % LV) == 0 && LS >= 1 && (32 % LS) == 0, "ring: k = 16, 32, 64, 128");
    const int2 *st = stage + e0;
    {
        const int cv = lane % LV;
        const unsigned char *src = reinterpret_cast<const unsigned char *>(sp_data) + 16 * cv;
        unsigned char *dst = half + 16 * cv;
        for (int e = lane / LV; e < cnt; e += 32 / LV)
            cp_async16(dst + e * C::ROW_BYTES, src + (int64_t)st[e].x * (4 * K));
    }
    {
        const int cs = lane % LS;
        unsigned char *dst = half + 4 * K + 16 * cs;
        for (int e = lane / LS; e < cnt; e += 32 / LS)
            cp_async16(dst + e * C::ROW_BYTES, sp_index + (int64_t)st[e].x * K + 16 * cs);
    }
    cp_async_commit();  // every lane, also those without a chunk: group counts stay warp-uniform
}
template <class C>
__device__ __forceinline__ void fwd_ring_accumulate(const int2 *stage, int e0, int cnt,
                                                    const unsigned char *half, float *acc_lane, int k,
                                                    int lane)
{
    constexpr int U = C::U, E = C::E, K = C::K, RB = C::ROW_BYTES;
    const int es = lane & (E - 1);
    const int entry = (C::KL < 32) ? lane / E : lane;
    // this lane's view of the rows: value word `entry` (+ 32 i), selector byte 4K + entry (+ 32 i) of row es, es + E, ...
    const unsigned char *pv = half + es * RB + 4 * entry;
    const unsigned char *ps = half + es * RB + 4 * K + entry;
    const int2 *st = stage + e0 + es;
    int j = 0;
    for (; j + U * E <= cnt; j += U * E) {  // every lane has an edge in each of the U instructions: no predicates
        Slice<C> sl[U];
        float w[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            w[u] = __int_as_float(st[u * E].y);
#pragma unroll
            for (int i = 0; i < C::EPL; ++i) {  // conflict-free: consecutive words / bytes
                sl[u].v[i] = *reinterpret_cast<const float *>(pv + u * E * RB + 128 * i);
                sl[u].s[i] = ps[u * E * RB + 32 * i];
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) fwd_accumulate<C>(sl[u], acc_lane, w[u], k, lane, true);
        pv += U * E * RB;
        ps += U * E * RB;
        st += U * E;
    }
    for (; j < cnt; j += E) {  // tail: one instruction at a time, lanes without an edge idle
        const bool on = j + es < cnt;
        Slice<C> sl;
        float w = 0.f;
#pragma unroll
        for (int i = 0; i < C::EPL; ++i) {
            sl.v[i] = 0.f;
            sl.s[i] = 0;
        }
        if (on) {
            w = __int_as_float(st[0].y);
#pragma unroll
            for (int i = 0; i < C::EPL; ++i) {
                sl.v[i] = *reinterpret_cast<const float *>(pv + 128 * i);
                sl.s[i] = ps[32 * i];
            }
        }
        fwd_accumulate<C>(sl, acc_lane, w, k, lane, on);
        pv += E * RB;
        ps += E * RB;
        st += E;
    }
}
template <class C>
__device__ __forceinline__ void fwd_process_batch_ring(const int2 *stage, int len, float *acc_lane,
                                                       unsigned char *ring, const float *sp_data,
                                                       const unsigned char *sp_index, int k, int lane)
{
    constexpr int G = C::RING, HALF = G * C::ROW_BYTES;
    const int ng = (len + G - 1) / G;
    fwd_ring_issue<C>(stage, 0, min(G, len), ring, sp_data, sp_index, lane);
    if (ng > 1) fwd_ring_issue<C>(stage, G, min(G, len - G), ring + HALF, sp_data, sp_index, lane);
    for (int g = 0; g < ng; ++g) {
        if (g + 1 < ng)
            cp_async_wait<1>();  // this lane's chunks of group g have landed ...
        else
            cp_async_wait<0>();
        __syncwarp();            // ... and so have the other lanes'
        unsigned char *half = ring + (g & 1) * HALF;
        fwd_ring_accumulate<C>(stage, g * G, min(G, len - g * G), half, acc_lane, k, lane);
        __syncwarp();            // everyone is done reading this half
        if (g + 2 < ng)
            fwd_ring_issue<C>(stage, (g + 2) * G, min(G, len - (g + 2) * G), half, sp_data, sp_index, lane);
    }
}

// Sum the R copies of the row, write it out (plain store or vector reduction) and clear
// the accumulator.  Copies of one column are contiguous (acc[c*R + r]).
template <class C>
__device__ __forceinline__ void fwd_flush_row(float *acc, float *__restrict__ out, int64_t row,
                                              int d, int lane, bool shared_row)
{
    constexpr int R = C::R;
    __syncwarp();
    float *dst = out + row * (int64_t)d;
    const int total = d * R;
    for (int base = 0; base < total; base += 128) {  // warp-uniform trip count (shuffles below)
        const int i = base + lane * 4;
        const bool in = i < total;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (in) {
            v = *reinterpret_cast<float4 *>(acc + i);
            *reinterpret_cast<float4 *>(acc + i) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if constexpr (R == 1) {
            if (in) {
                if (shared_row)
                    red_add_f4(dst + i, v.x, v.y, v.z, v.w);
                else
                    st_stream_f4(reinterpret_cast<float4 *>(dst + i), v);
            }
        } else if constexpr (R == 2) {
            const float s0 = v.x + v.y, s1 = v.z + v.w;
            float *p = dst + (i >> 1);
            if (in) {
                if (shared_row) {
                    red_add_f1(p, s0);
                    red_add_f1(p + 1, s1);
                } else {
                    __stcs(reinterpret_cast<float2 *>(p), make_float2(s0, s1));
                }
            }
        } else {
            float s = (v.x + v.y) + (v.z + v.w);
#pragma unroll
            for (int o = 1; o < R / 4; o <<= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (in && (lane & (R / 4 - 1)) == 0) {
                float *p = dst + i / R;
                if (shared_row)
                    red_add_f1(p, s);
                else
                    __stcs(p, s);
            }
        }
    }
    __syncwarp();
}

// PLAN: out[row, c] = sum over the homes h of acc[plan_home_word(c, h)]; all homes are cleared.
template <class C>
__device__ __forceinline__ void fwd_flush_row_planned(float *acc, float *__restrict__ out, int64_t row,
                                                      int d, int lane, bool shared_row)
{
    constexpr int H = C::HOMES;
    __syncwarp();
    float *dst = out + row * (int64_t)d;
    for (int c = lane * 4; c < d; c += 128) {
        float4 v = *reinterpret_cast<float4 *>(acc + c);  // home 0
        // c % 4 == 0: columns c .. c+3 share c/8 and have c mod 8 = (c mod 8) + i
        const unsigned lo = c & 7, hi = c >> 3;
#pragma unroll
        for (int h = 1; h < H; ++h) {
            const float *p = acc + h * 256;
            v.x += p[(lo + 0) * 32 + ((hi + plan_s(h) * (lo + 0) + plan_t(h)) & 31)];
            v.y += p[(lo + 1) * 32 + ((hi + plan_s(h) * (lo + 1) + plan_t(h)) & 31)];
            v.z += p[(lo + 2) * 32 + ((hi + plan_s(h) * (lo + 2) + plan_t(h)) & 31)];
            v.w += p[(lo + 3) * 32 + ((hi + plan_s(h) * (lo + 3) + plan_t(h)) & 31)];
        }
        if (shared_row)
            red_add_f4(dst + c, v.x, v.y, v.z, v.w);
        else
            st_stream_f4(reinterpret_cast<float4 *>(dst + c), v);
    }
    __syncwarp();  // every lane has read all homes before anybody clears a word of them
    for (int i = lane * 4; i < C::PLAN_WORDS; i += 128)
        *reinterpret_cast<float4 *>(acc + i) = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncwarp();
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

// zero-fill rows [r0, r1) of the output (rows without tasks: isolated destination nodes)
__device__ __forceinline__ void fwd_zero_rows(float *__restrict__ out, int64_t r0, int64_t r1, int d,
                                              int lane)
{
    if (r0 >= r1) return;
    float4 *p = reinterpret_cast<float4 *>(out + r0 * d);
    const int64_t n4 = (r1 - r0) * (d >> 2);
    for (int64_t i = lane; i < n4; i += 32) st_stream_f4(p + i, make_float4(0.f, 0.f, 0.f, 0.f));
}

// Rows that straddle two warps' runs are combined with vector reductions and therefore have
// to start from zero: one warp per run boundary clears the row that crosses it.  Together
// with the gap fill inside the main kernel this replaces a memset of the whole output
// (4*N*D bytes of extra HBM writes: 2.5 GB, 0.4 ms, on the ogbn-products shape).
__global__ void __launch_bounds__(256)
spgemm_fwd_clear_shared_rows(const int4 *__restrict__ warp4, int64_t num_tasks, int tasks_per_warp,
                             float *__restrict__ out, int d)
{
    const int lane = threadIdx.x & 31;
    const int64_t b = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5) + 1;  // boundary index
    const int64_t t = b * tasks_per_warp;
    if (t >= num_tasks) return;
    const int row = __ldg(&warp4[t].x);
    if (__ldg(&warp4[t - 1].x) != row) return;
    fwd_zero_rows(out, row, row + 1, d, lane);
}

template <class C>
__global__ void __launch_bounds__(C::WPB * 32, C::MINB)
spgemm_fwd_kernel(const int4 *__restrict__ warp4, int64_t num_tasks, const int *__restrict__ idx,
                  const float *__restrict__ val, const float *sp_data,
                  const unsigned char *sp_index, int k, float *__restrict__ out, int d,
                  int tasks_per_warp, int accumulate, int64_t num_rows_out,
                  typename PeerParam<C::PEER>::type peer)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int64_t cta = blockIdx.x;
    bool gathered = true;  // PEER: the other ranks' CBSR blocks have arrived
    if constexpr (C::PEER) {
        if (cta < peer.copy_ctas) {  // lowest block ids: resident before any waiting CTA
            peer_push_block<(C::PLAN ? 2 : 1)>(peer, (int)cta, smem_raw);
            return;
        }
        cta -= peer.copy_ctas;
        gathered = false;
    }
    float *acc;
    int2 *stage;
    if constexpr (C::PLAN) {  // two 256-word copies of the row per warp
        acc = reinterpret_cast<float *>(smem_raw) + warp * C::PLAN_WORDS;
        stage = reinterpret_cast<int2 *>(reinterpret_cast<float *>(smem_raw) + C::WPB * C::PLAN_WORDS) +
                warp * FWD_STAGE;
    } else {
        acc = reinterpret_cast<float *>(smem_raw) + warp * d * C::R;
        stage = reinterpret_cast<int2 *>(reinterpret_cast<float *>(smem_raw) + C::WPB * d * C::R) +
                warp * FWD_STAGE;
    }
    float *acc_lane = acc + (lane & (C::R - 1));
    unsigned char *ring = nullptr;  // RING: 2 * RING gathered rows of this warp, behind the staging area
    if constexpr (C::RING > 0)
        ring = smem_raw + (size_t)C::WPB * d * C::R * sizeof(float) + (size_t)C::WPB * FWD_STAGE * sizeof(int2) +
               (size_t)warp * 2 * C::RING * C::ROW_BYTES;

    const int64_t t0 = (cta * C::WPB + warp) * tasks_per_warp;
    if (t0 >= num_tasks) return;  // no CTA-wide barrier below
    const int64_t t1 = min(t0 + (int64_t)tasks_per_warp, num_tasks);

    if constexpr (C::PLAN) {
        for (int i = lane * 4; i < C::PLAN_WORDS; i += 128)
            *reinterpret_cast<float4 *>(acc + i) = make_float4(0.f, 0.f, 0.f, 0.f);
    } else {
        for (int i = lane * 4; i < d * C::R; i += 128)
            *reinterpret_cast<float4 *>(acc + i) = make_float4(0.f, 0.f, 0.f, 0.f);
    }

    int4 info = ld_stream_int4(warp4 + t0);
    int4 ninfo = info;  // task t+1, loaded one task early so that its CSR slice can be
    if (t0 + 1 < t1) ninfo = ld_stream_int4(warp4 + t0 + 1);  // prefetched without a stall
    const int first_row = info.x;
    const int prev_row = (t0 > 0) ? __ldg(&warp4[t0 - 1].x) : -1;
    // accumulate mode (multi-GPU local/remote split): every row is added to what is there
    const bool first_shared = accumulate || (prev_row == first_row);
    bool last_shared = accumulate;
    if (!accumulate && t1 < num_tasks)
        last_shared = (__ldg(&warp4[t1].x) == __ldg(&warp4[t1 - 1].x));

    EdgeRegs pre;
    fwd_fetch_edges(pre, idx, val, info.y, min(info.z, FWD_STAGE), lane);
    // Output rows without tasks are DEFINED as zero: each run fills the gap in
    // front of its first row and the gaps between its rows; the last run fills the tail.
    if (!accumulate) fwd_zero_rows(out, (int64_t)prev_row + 1, first_row, d, lane);

    int cur_row = first_row;
    for (int64_t t = t0; t < t1; ++t) {
        const bool has_next = (t + 1 < t1);
        int4 nninfo = ninfo;
        if (t + 2 < t1) nninfo = ld_stream_int4(warp4 + t + 2);

        if (info.x != cur_row) {
            if constexpr (C::PLAN)
                fwd_flush_row_planned<C>(acc, out, cur_row, d, lane,
                                         accumulate || (first_shared && cur_row == first_row));
            else
                fwd_flush_row<C>(acc, out, cur_row, d, lane,
                                 accumulate || (first_shared && cur_row == first_row));
            if (!accumulate) fwd_zero_rows(out, (int64_t)cur_row + 1, info.x, d, lane);
            cur_row = info.x;
        }
        const int64_t loc = info.y;
        const int len = info.z;
        if constexpr (C::PEER) {
            // tasks from `wait_from_task` on read columns owned by other ranks
            if (!gathered && t >= peer.wait_from_task) {
                peer_wait_gathered(peer, lane);
                gathered = true;
            }
        }
        fwd_stage_edges(pre, stage, lane);
        if (has_next)  // next task's CSR slice streams in while this task's gathers run
            fwd_fetch_edges(pre, idx, val, ninfo.y, min(ninfo.z, FWD_STAGE), lane);
        if constexpr (C::RING > 0)
            fwd_process_batch_ring<C>(stage, min(len, FWD_STAGE), acc_lane, ring, sp_data, sp_index, k, lane);
        else
            fwd_process_batch<C>(stage, min(len, FWD_STAGE), acc_lane, sp_data, sp_index, k, lane);
        for (int off = FWD_STAGE; off < len; off += FWD_STAGE) {  // len > 64: foreign work lists
            EdgeRegs more;
            const int ln = min(len - off, FWD_STAGE);
            fwd_fetch_edges(more, idx, val, loc + off, ln, lane);
            fwd_stage_edges(more, stage, lane);
            if constexpr (C::RING > 0)
                fwd_process_batch_ring<C>(stage, ln, acc_lane, ring, sp_data, sp_index, k, lane);
            else
                fwd_process_batch<C>(stage, ln, acc_lane, sp_data, sp_index, k, lane);
        }
        info = ninfo;
        ninfo = nninfo;
    }
    if constexpr (C::PLAN)
        fwd_flush_row_planned<C>(acc, out, cur_row, d, lane,
                                 last_shared || (first_shared && cur_row == first_row));
    else
        fwd_flush_row<C>(acc, out, cur_row, d, lane,
                         last_shared || (first_shared && cur_row == first_row));
    if (!accumulate && t1 == num_tasks) fwd_zero_rows(out, (int64_t)cur_row + 1, num_rows_out, d, lane);
}

struct FwdArgs {
    const int32_t *warp4;
    int64_t W;
    const int32_t *idx;
    const float *val;
    const float *sp_data;
    const uint8_t *sp_index;
    int k;
    float *out;
    int d;
    int tasks_per_warp;
    int accumulate;
    cudaStream_t st;
    int depth;
    int64_t num_rows_out;
    const PeerArgs *peer;  // fused multi-GPU launch only
    bool ring_ok;          // operands 16-byte aligned: the cp.async ring variants may run
};

template <class C>
static int launch_fwd(const FwdArgs &a)
{
    size_t smem = (size_t)C::WPB * (C::PLAN ? C::PLAN_WORDS : a.d * C::R) * sizeof(float) +
                  (size_t)C::WPB * FWD_STAGE * sizeof(int2) +
                  (size_t)C::WPB * 2 * C::RING * C::ROW_BYTES;
    if constexpr (C::PEER) {  // the copy CTAs stage their TMA chunks in the same dynamic smem
        if (a.peer->use_tma && smem < PEER_TMA_SMEM + 128) smem = PEER_TMA_SMEM + 128;
    }
    if (smem > 48 * 1024)
        if (int rc = opt_in_dynamic_smem(reinterpret_cast<const void *>(&spgemm_fwd_kernel<C>), "spgemm_fwd_kernel")) return rc;
    const int64_t chunks = ceil_div<int64_t>(a.W, a.tasks_per_warp);
    int64_t grid = ceil_div<int64_t>(chunks, C::WPB);
    typename PeerParam<C::PEER>::type peer{};
    if constexpr (C::PEER) {
        peer = *a.peer;
        grid += peer.copy_ctas;
    }
    if (grid > 0x7fffffffLL) {
        set_error("maxk_spgemm_forward: grid too large");
        return MAXK_ERR_UNSUPPORTED;
    }
    spgemm_fwd_kernel<C><<<(unsigned)grid, C::WPB * 32, smem, a.st>>>(
        reinterpret_cast<const int4 *>(a.warp4), a.W, a.idx, a.val, a.sp_data, a.sp_index, a.k,
        a.out, a.d, a.tasks_per_warp, a.accumulate, a.num_rows_out, peer);
    MAXK_LAUNCH_CHECK("spgemm_fwd_kernel");
    return MAXK_OK;
}

int g_fwd_tasks_per_warp = 0;   // run length; 0 = automatic; tunable through maxk_set_tuning
int g_fwd_copies = 0;           // 0 = automatic choice of R
int g_fwd_depth = 0;            // launch shape, see maxk_set_forward_depth; 0 = automatic

// A CBSR operand of more than this many bytes cannot stay in the 126 MB L2 next to the CSR stream and
// the output: its gathers pay DRAM latency and the cp.async ring shapes win (measured on the products
// shape, 392 MB; on the 37 MB Reddit operand they lose 30 %).
constexpr int64_t FWD_L2_OPERAND_BYTES = 96LL << 20;

// launch shape per R:      R<=2: 8 warps, U=4, 64 regs   R=4: 8 warps, U=4
//                          R=8 : 8 warps (68 KB/CTA, 3 CTAs/SM), U=8
//                          R=16: 4 warps (66 KB/CTA, 3 CTAs/SM), U=8
template <int KL, int EPL, bool EXACT, int R>
static int launch_fwd_r(const FwdArgs &a)
{
    constexpr int U1 = (EPL == 1) ? 4 : (EPL == 2 ? 2 : (EPL <= 4 ? 2 : 1));
    if constexpr (R <= 2 && EPL <= 2) {
        // CTA shape (measured, profiles/r01_forward_pipeline.md, profiles/r02_variants.md): graphs of
        // short rows are latency-bound on the CBSR gathers and run lengths vary a lot, so 2-warp CTAs
        // (no warp idles until the slowest of 8 finishes) are 12 % faster on the products shape.
        // Register-only ways to keep more gathers in flight (2x / 4x U, double-buffered batches,
        // 48 / 40-register caps) were measured and lose to spills and lost occupancy; the cp.async
        // ring (shapes 8 and 12) is what pays when the operand misses the L2.
        if constexpr (EXACT && (KL == 32 ? 32 * EPL : KL) % 16 == 0) {
            if (a.ring_ok) {
                // 2-warp CTAs, ring of 2 x 16 rows (k <= 32) ...
                if (a.depth == 8) return launch_fwd<FwdCfg<KL, EPL, EXACT, R, 2, U1, 16, false, 16>>(a);
                // ... or 2 x 8 rows under a 48-register cap (loads in flight hold none): 40 warps per SM (k = 64)
                if (a.depth == 12) return launch_fwd<FwdCfg<KL, EPL, EXACT, R, 2, (U1 > 1 ? U1 / 2 : 1), 20, false, 8>>(a);
            }
        }
        if (a.depth == 2 || a.depth == 8 || a.depth == 12)  // (ring asked for but not applicable: 2-warp CTAs)
            return launch_fwd<FwdCfg<KL, EPL, EXACT, R, 2, U1, 16>>(a);
        if (a.depth == 3) return launch_fwd<FwdCfg<KL, EPL, EXACT, R, 8, 2 * U1, 3>>(a);  // deeper (measurement knob)
    }
    if constexpr (R <= 4)
        return launch_fwd<FwdCfg<KL, EPL, EXACT, R, 8, U1, (EPL <= 2 ? 4 : 2)>>(a);
    else if constexpr (R == 8)
        return launch_fwd<FwdCfg<KL, EPL, EXACT, R, 8, 2 * U1, 3>>(a);
    else if constexpr (R == 16)
        return launch_fwd<FwdCfg<KL, EPL, EXACT, R, 4, 2 * U1, 3>>(a);
    else
        return launch_fwd<FwdCfg<KL, EPL, EXACT, R, 2, U1, 2>>(a);
}

// pick R among the instantiated copies for this edge width (R >= E always)
template <int KL, int EPL, bool EXACT>
static int launch_fwd_pick(const FwdArgs &a, int want_r)
{
    constexpr int E = 32 / KL;
    if constexpr (KL == 32 && (EPL > 2 || !EXACT)) {
        return launch_fwd_r<KL, EPL, EXACT, 1>(a);
    } else if constexpr (E == 32) {
        return launch_fwd_r<KL, EPL, EXACT, 32>(a);
    } else {
        if constexpr (E < 16) {
            if (want_r >= 16) return launch_fwd_r<KL, EPL, EXACT, 16>(a);
        }
        if constexpr (E < 8) {
            if (want_r >= 8) return launch_fwd_r<KL, EPL, EXACT, 8>(a);
        }
        if constexpr (E < 4) {
            if (want_r >= 4) return launch_fwd_r<KL, EPL, EXACT, 4>(a);
        }
        if constexpr (E < 2) {
            if (want_r >= 2) return launch_fwd_r<KL, EPL, EXACT, 2>(a);
        }
        return launch_fwd_r<KL, EPL, EXACT, E>(a);
    }
}

// Rows shared between two runs are cleared by a pre-kernel, rows without tasks inside the main kernel:
// together they DEFINE the output without a memset of the whole of it.
static int clear_shared_rows(const FwdArgs &a)
{
    const int64_t boundaries = (a.W - 1) / a.tasks_per_warp;
    if (boundaries > 0) {
        spgemm_fwd_clear_shared_rows<<<(unsigned)ceil_div<int64_t>(boundaries, 8), 256, 0, a.st>>>(
            reinterpret_cast<const int4 *>(a.warp4), a.W, a.tasks_per_warp, a.out, a.d);
        MAXK_LAUNCH_CHECK("spgemm_fwd_clear_shared_rows");
    }
    return MAXK_OK;
}

// argument checks shared by every forward entry point; `who` names it in the message
static int check_fwd_common(const char *who, const int32_t *warp4, int64_t num_tasks, int k, const float *out,
                            int64_t num_rows_out, int d)
{
    MAXK_REQUIRE(num_tasks >= 0 && num_rows_out >= 0, "%s: negative size", who);
    MAXK_REQUIRE(d > 0 && d <= MAXK_MAX_DIM && d % 4 == 0, "%s: d=%d must be a multiple of 4 in [4,256]", who, d);
    MAXK_REQUIRE(k >= 1 && k <= d, "%s: k=%d must be in [1,d=%d]", who, k, d);
    MAXK_REQUIRE(out != nullptr || num_rows_out == 0, "%s: out is NULL", who);
    MAXK_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0 && (reinterpret_cast<uintptr_t>(warp4) & 15) == 0,
                 "%s: out and warp4 must be 16-byte aligned", who);
    return MAXK_OK;
}

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
    if (int rc = check_fwd_common("maxk_spgemm_forward", warp4, num_tasks, k, out, num_rows_out, d)) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (num_rows_out == 0) return MAXK_OK;
    // Output is DEFINED by this call: empty rows and rows shared between
    // two warps' runs start from zero.
    if (num_tasks == 0) {
        if (!accumulate) {
            MAXK_CUDA(cudaMemsetAsync(out, 0, (size_t)num_rows_out * d * sizeof(float), st));
            count_launch();
        }
        return MAXK_OK;
    }
    MAXK_REQUIRE(warp4 && indices && values && sp_data && sp_index,
                 "maxk_spgemm_forward: NULL input pointer");
    const bool vec_ok = ((reinterpret_cast<uintptr_t>(sp_data) & 15) == 0) &&
                        ((reinterpret_cast<uintptr_t>(sp_index) & 3) == 0);
    // Run length (tasks per warp): 16 amortises the row flush on big graphs; small graphs get
    // shorter runs so that the grid still fills 148 SMs x 32 warps several times over.
    int tpw = g_fwd_tasks_per_warp;
    if (tpw <= 0) tpw = auto_tasks_per_warp(num_tasks);
    // Launch shape.  Mostly single-task rows (W < 2N) -> 2-warp CTAs (see launch_fwd_r); if in
    // addition the CBSR operand cannot live in the L2 the gathers go through the cp.async ring.
    // The operand has at least as many rows as the graph has destination rows with edges
    // (num_rows_out is what the caller tells us about its size; 5k bytes per row).
    int shape = g_fwd_depth;
    if (shape <= 0) {
        shape = (num_tasks < 2 * num_rows_out) ? 2 : 1;
        if (shape == 2 && num_rows_out * 5LL * k > FWD_L2_OPERAND_BYTES) shape = (k > 32) ? 12 : 8;
    }
    const bool ring_ok = ((reinterpret_cast<uintptr_t>(sp_data) & 15) == 0) &&
                         ((reinterpret_cast<uintptr_t>(sp_index) & 15) == 0);
    FwdArgs a{warp4, num_tasks, indices, values, sp_data, sp_index, k, out, d,
              tpw, accumulate, st, shape, num_rows_out, nullptr, ring_ok};
    if (!accumulate)
        if (int rc = clear_shared_rows(a)) return rc;
    // R: accumulator copies.  Measured on B200 (profiles/r01_sweep_copies.md): R = E (one copy
    // per edge slot) is fastest on every BASELINE shape -- with 1 < R < 32 the worst bank
    // still serialises ~3 lanes, and R >= 16 costs more occupancy than it saves.
    const int want_r = g_fwd_copies > 0 ? g_fwd_copies : 1;
    switch (k) {
        case 16: return launch_fwd_pick<16, 1, true>(a, want_r);
        case 8: return launch_fwd_pick<8, 1, true>(a, want_r);
        case 4: return launch_fwd_pick<4, 1, true>(a, want_r);
        case 2: return launch_fwd_pick<2, 1, true>(a, want_r);
        case 1: return launch_fwd_pick<1, 1, true>(a, want_r);
        default: break;
    }
    const int epl = (k + 31) / 32;
    const bool exact = (k % 32 == 0) && vec_ok;
#define MAXK_FWD_CASE(EP)                                                         \
    case EP:                                                                      \
        return exact ? launch_fwd_pick<32, EP, true>(a, want_r)                   \
                     : launch_fwd_pick<32, EP, false>(a, want_r);
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

// ---- forward through a bank plan (cbsr_plan.cu): k = 32 and 64 ---------------------------------------
// Measured on the Reddit shape (profiles/r02_variants.md): the update's bank conflicts are the bound
// there (ncu: 4.5 of 9.9 L1 wavefronts per edge), two homes per column take the fullest bank from 3.3
// to ~2.1 entries and the kernel from 3.99 to ~3.1 ms (k = 32), 7.1 to ~5.9 ms (k = 64).  On short rows
// (products, Flickr) the doubled flush eats the gain: the caller chooses (maxk_b200.ops.use_bank_plan).
extern "C" int maxk_spgemm_forward_planned(const int32_t *warp4, int64_t num_tasks,
                                           const int32_t *indices, const float *values,
                                           const float *sp_data, const uint16_t *sp_addr, int32_t k,
                                           float *out, int64_t num_rows_out, int32_t d,
                                           int32_t accumulate, void *stream)
{
    if (int rc = check_fwd_common("maxk_spgemm_forward_planned", warp4, num_tasks, k, out, num_rows_out, d)) return rc;
    MAXK_REQUIRE(k == 32 || k == 64, "maxk_spgemm_forward_planned: k=%d is not built (32, 64)", k);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (num_rows_out == 0) return MAXK_OK;
    if (num_tasks == 0) {
        if (!accumulate) {
            MAXK_CUDA(cudaMemsetAsync(out, 0, (size_t)num_rows_out * d * sizeof(float), st));
            count_launch();
        }
        return MAXK_OK;
    }
    MAXK_REQUIRE(warp4 && indices && values && sp_data && sp_addr,
                 "maxk_spgemm_forward_planned: NULL input pointer");
    MAXK_REQUIRE((reinterpret_cast<uintptr_t>(sp_data) & 3) == 0 && (reinterpret_cast<uintptr_t>(sp_addr) & 1) == 0,
                 "maxk_spgemm_forward_planned: sp_data / sp_addr are misaligned");
    int tpw = g_fwd_tasks_per_warp;
    if (tpw <= 0) tpw = auto_tasks_per_warp(num_tasks);
    const bool two_warps = (g_fwd_depth == 2) || (g_fwd_depth <= 0 && num_tasks < 2 * num_rows_out);
    FwdArgs a{warp4, num_tasks, indices, values, sp_data, reinterpret_cast<const uint8_t *>(sp_addr), k, out, d,
              tpw, accumulate, st, two_warps ? 2 : 1, num_rows_out, nullptr, false};
    if (!accumulate)
        if (int rc = clear_shared_rows(a)) return rc;
    if (k == 32)
        return two_warps ? launch_fwd<FwdCfg<32, 1, true, 1, 2, 4, 16, false, 0, true>>(a)
                         : launch_fwd<FwdCfg<32, 1, true, 1, 8, 4, 4, false, 0, true>>(a);
    return two_warps ? launch_fwd<FwdCfg<32, 2, true, 1, 2, 2, 16, false, 0, true>>(a)
                     : launch_fwd<FwdCfg<32, 2, true, 1, 8, 2, 4, false, 0, true>>(a);
}

// ---- fused all-gather + forward over NVLink peer memory (peer.cuh) ---------------------------
// Checks and the PeerArgs block shared by the two fused entry points.
static int fused_setup(const char *who, const maxk_peer_ctx *ctx, uint32_t epoch, const int32_t *warp4,
                       int64_t num_tasks, int64_t wait_from_task, const float *out, int64_t num_rows_out, int32_t d,
                       int32_t copy_ctas, PeerArgs &pa)
{
    MAXK_REQUIRE(ctx != nullptr, "%s: ctx is NULL", who);
    const int world = ctx->world, rank = ctx->rank, k = ctx->k;
    MAXK_REQUIRE(world >= 1 && world <= MAXK_PEER_MAX && rank >= 0 && rank < world,
                 "%s: world=%d rank=%d out of range", who, world, rank);
    MAXK_REQUIRE(wait_from_task >= 0, "%s: negative size", who);
    if (int rc = check_fwd_common(who, warp4, num_tasks, k, out, num_rows_out, d)) return rc;
    MAXK_REQUIRE(copy_ctas != 0 && copy_ctas >= -1024 && copy_ctas <= 1024, "%s: copy_ctas out of range", who);
    pa = PeerArgs{};
    pa.copy_ctas = copy_ctas < 0 ? -copy_ctas : copy_ctas;
    pa.use_tma = copy_ctas > 0 ? 1 : 0;  // negative count: plain load/store copy (measurement knob)
    pa.world = world;
    pa.rank = rank;
    pa.k = k;
    pa.epoch = epoch;
    pa.wait_from_task = wait_from_task;
    for (int r = 0; r < world; ++r) {
        MAXK_REQUIRE(ctx->g_data[r] && ctx->g_index[r] && ctx->flags[r], "%s: buffers of rank %d are not mapped", who, r);
        MAXK_REQUIRE(ctx->rows_of[r] >= 0 && ctx->row_start[r] >= 0 &&
                         ctx->row_start[r] + ctx->rows_of[r] <= ctx->n_total,
                     "%s: row range of rank %d is outside [0, n_total)", who, r);
        MAXK_REQUIRE((ctx->row_start[r] * (int64_t)k) % 16 == 0,
                     "%s: row_start[%d] * k must be a multiple of 16", who, r);
        pa.rows_of[r] = ctx->rows_of[r];
        pa.row_start[r] = ctx->row_start[r];
        pa.send_rows[r] = (r == rank) ? nullptr : ctx->send_rows[r];
        pa.send_count[r] = ctx->send_count[r];
        pa.g_data[r] = static_cast<float *>(ctx->g_data[r]);
        pa.g_index[r] = static_cast<unsigned char *>(ctx->g_index[r]);
        pa.flags[r] = static_cast<unsigned *>(ctx->flags[r]);
    }
    // halo pairs are moved row by row with ordinary loads and stores: that needs many threads in
    // flight (a dependent list -> row -> store chain per 16 bytes), so the copy part of the grid
    // grows to four CTAs per SM -- the aggregating CTAs mostly wait for those rows anyway
    for (int r = 0; r < world; ++r)
        if (r != rank && pa.send_rows[r] != nullptr && pa.send_count[r] > 0 && pa.copy_ctas < 4 * kNumSMs)
            pa.copy_ctas = 4 * kNumSMs;
    pa.local_sync = static_cast<unsigned *>(ctx->local_sync);
    MAXK_REQUIRE(pa.local_sync != nullptr, "%s: local_sync is NULL", who);
    MAXK_REQUIRE(warp4 || num_tasks == 0, "%s: warp4 is NULL", who);
    pa.spin_limit = g_peer_spin_limit_cycles;
    return MAXK_OK;
}

extern "C" int maxk_spgemm_forward_fused(const maxk_peer_ctx *ctx, uint32_t epoch,
                                         const int32_t *warp4, int64_t num_tasks,
                                         int64_t wait_from_task, const int32_t *indices,
                                         const float *values, float *out, int64_t num_rows_out,
                                         int32_t d, int32_t copy_ctas, void *stream)
{
    PeerArgs pa;
    if (int rc = fused_setup("maxk_spgemm_forward_fused", ctx, epoch, warp4, num_tasks, wait_from_task, out,
                             num_rows_out, d, copy_ctas, pa))
        return rc;
    const int k = ctx->k;
    MAXK_REQUIRE(k == 8 || k == 16 || k == 32 || k == 64 || k == 96 || k == 128 || k == 256,
                 "maxk_spgemm_forward_fused: k=%d is not built for the fused path (8, 16, 32, 64, 96, 128, 256)", k);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // local-column and remote-column tasks of one row meet in `out` through reductions
    if (num_rows_out > 0) {
        MAXK_CUDA(cudaMemsetAsync(out, 0, (size_t)num_rows_out * d * sizeof(float), st));
        count_launch();
    }
    int tpw = g_fwd_tasks_per_warp;
    if (tpw <= 0) tpw = auto_tasks_per_warp(num_tasks);
    FwdArgs a{warp4, num_tasks, indices, values, pa.g_data[pa.rank], pa.g_index[pa.rank], k, out, d,
              tpw, /*accumulate=*/1, st, 1, num_rows_out, &pa, false};
    // the copy CTAs must run even when this rank has no task at all: launch_fwd adds them
    switch (k) {
        case 8: return launch_fwd<FwdCfg<8, 1, true, 4, 8, 4, 4, true>>(a);
        case 16: return launch_fwd<FwdCfg<16, 1, true, 2, 8, 4, 4, true>>(a);
        case 32: return launch_fwd<FwdCfg<32, 1, true, 1, 8, 4, 4, true>>(a);
        case 64: return launch_fwd<FwdCfg<32, 2, true, 1, 8, 2, 4, true>>(a);
        case 96: return launch_fwd<FwdCfg<32, 3, true, 1, 8, 2, 2, true>>(a);
        case 128: return launch_fwd<FwdCfg<32, 4, true, 1, 8, 2, 2, true>>(a);
        default: return launch_fwd<FwdCfg<32, 8, true, 1, 8, 1, 2, true>>(a);
    }
}

// ---- the same with the bank plan travelling in place of the selectors (k = 32, 64) ----
extern "C" int maxk_spgemm_forward_fused_planned(const maxk_peer_ctx *ctx, uint32_t epoch,
                                         const int32_t *warp4, int64_t num_tasks,
                                         int64_t wait_from_task, const int32_t *indices,
                                         const float *values, float *out, int64_t num_rows_out,
                                         int32_t d, int32_t copy_ctas, void *stream)
{
    PeerArgs pa;
    if (int rc = fused_setup("maxk_spgemm_forward_fused_planned", ctx, epoch, warp4, num_tasks, wait_from_task,
                             out, num_rows_out, d, copy_ctas, pa))
        return rc;
    const int k = ctx->k;
    MAXK_REQUIRE(k == 32 || k == 64, "maxk_spgemm_forward_fused_planned: k=%d is not built (32, 64)", k);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (num_rows_out > 0) {
        MAXK_CUDA(cudaMemsetAsync(out, 0, (size_t)num_rows_out * d * sizeof(float), st));
        count_launch();
    }
    int tpw = g_fwd_tasks_per_warp;
    if (tpw <= 0) tpw = auto_tasks_per_warp(num_tasks);
    // g_index[r] holds the uint16 bank plan (cbsr_plan.cu) of every row instead of its selectors
    FwdArgs a{warp4, num_tasks, indices, values, pa.g_data[pa.rank], pa.g_index[pa.rank], k, out, d,
              tpw, /*accumulate=*/1, st, 1, num_rows_out, &pa, false};
    if (k == 32) return launch_fwd<FwdCfg<32, 1, true, 1, 8, 4, 4, true, 0, true>>(a);
    return launch_fwd<FwdCfg<32, 2, true, 1, 8, 2, 4, true, 0, true>>(a);
}
