// linear_topk.cu -- SURVEY 8(f) rank 3 (i): the Linear in front of a MaxK with the selection in its epilogue.
//
//     (sp_data, sp_index) = TopK_k( X @ W^T + b )          X [n, in] fp32, W [256, in] fp32
//
// replaces `Linear -> MaxK` of the GCN / GIN layers (ref: utils/models.py:336-341, 365-369: torch.nn.Linear, then
// MaxK.apply) without the dense [n, 256] activation ever reaching HBM: the reference writes it (4nD bytes) and its
// TopK reads it back (4nD bytes).  One persistent warp-specialised kernel per SM:
//
//   warp 0      TMA producer: per 128-row tile and 16-column K chunk one X box [128 x 16] and two W boxes
//               [256 x 16] (high and low parts of W, split once per call by a prep kernel) into a four-stage
//               shared-memory ring (64-byte swizzle, 48 KB per stage), completion on an mbarrier;
//   warps 2-3   split the X box in place into its TF32 part and the remainder  x = hi + lo  (fp32 accuracy
//               from TF32 tensor-core products: hi*hi + hi*lo + lo*hi, the dropped lo*lo term is < 2^-21
//               relative), generic-proxy writes fenced for the tensor core;
//   warp 1      one thread issues tcgen05.mma.kind::tf32 (M = 128, N = 256, K = 8), three products per K step,
//               fp32 accumulation in TMEM; the 512 TMEM columns hold two 128 x 256 accumulators, so the next tile
//               is multiplied while the previous one is selected; tcgen05.commit frees ring slots / publishes
//               accumulators;
//   warps 4-11  epilogue: a thread owns HALF A ROW (128 contiguous columns of TMEM lane = row) in registers --
//               one tcgen05.ld pass (TMEM reads are 64 B / clock / SM: the search must not re-read it), the
//               accumulator is released at once; bias added; then the same value-space regula falsi as
//               topk_cbsr.cu with the two halves of a row exchanging their counts through shared memory
//               (named barrier per row quarter), emission by predicated stores into a shared-memory staging area
//               that has the layout of sp_data / sp_index (a thread's columns are contiguous: the lower half
//               starts at 0, the upper half at the lower half's count -- no scan), 128-bit copy-out; rows the fast
//               path cannot finish (ties at the k-th value) are passed through shared memory to the robust
//               key-space search of topk_select.cuh.
//
// Bound: tensor pipe (3 products x 32 K steps x 128 cycles = 12.3 k cycles per tile) against ~640 KB of L2 -> SM
// traffic per tile (W is re-streamed for every tile); measured 1.83 ms on 2 449 029 x 256 (tensor pipe 44 % busy, the
// epilogue's instruction issue is the limit: DESIGN.md 4.7).  d = 256 only (one MMA spans the row).
// The same kernel without the selection (sp_index == NULL) is the dense Linear of DESIGN.md 4.8: its epilogue drains
// the accumulator 16 columns at a time through shared memory into full 64-byte row segments (+ optional residual).
#include <mutex>
#include <set>

#include "common.cuh"
#include "tc_common.cuh"
#include "topk_select.cuh"

namespace maxk {

constexpr int LT_D = 256;          // output width = MMA N
constexpr int LT_TILE_M = 128;     // rows per tile = MMA M = TMEM lanes
#ifndef MAXK_LT_BK
#define MAXK_LT_BK 16
#endif
constexpr int LT_BK = MAXK_LT_BK;  // K chunk: 16 fp32 = one 64-byte swizzle row (32: 128-byte rows, two stages)
constexpr int LT_STAGES = 64 / LT_BK;  // 192 KB of operand ring either way
static_assert(LT_BK == 16 || LT_BK == 32, "the ring is laid out for 64- or 128-byte swizzle rows");
constexpr int LT_A_BYTES = LT_TILE_M * LT_BK * 4;   // 16 KB
constexpr int LT_B_BYTES = LT_D * LT_BK * 4;        // 32 KB
constexpr int LT_STAGE_BYTES = 2 * LT_A_BYTES + 2 * LT_B_BYTES;  // A_hi, A_lo, B_hi, B_lo = 96 KB
constexpr int LT_TX_BYTES = LT_A_BYTES + 2 * LT_B_BYTES;         // what the TMA unit delivers per stage
constexpr int LT_THREADS = 384;    // warp 0 TMA, 1 MMA, 2-3 split, 4-11 epilogue
constexpr int LT_SPLIT_THREADS = 64;
constexpr int LT_EPI_THREADS = 256;
constexpr int LT_HALF = LT_D / 2;  // columns per epilogue thread
constexpr int LT_STAGE_ENTRIES = 1216;  // CBSR entries staged per row quarter and pass (32 rows at k <= 38)

// tcgen05 instruction descriptor (cute/arch/mma_sm100_desc.hpp: InstrDescriptor): D = f32 [4,6) = 1,
// A = B = tf32 [7,10) / [10,13) = 2, both K-major (bits 15, 16 = 0), N >> 3 at [17,23), M >> 4 at [24,29)
constexpr uint32_t LT_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(LT_D >> 3) << 17) |
                              ((uint32_t)(LT_TILE_M >> 4) << 24);

struct LtShared {
    // offsets into dynamic shared memory (1024-byte aligned base)
    static constexpr int stage0 = 0;
    static constexpr int bias = LT_STAGES * LT_STAGE_BYTES;               // 256 floats
    static constexpr int xch_i = bias + LT_D * 4;                         // int [2 parities][128 rows][2 halves]
    static constexpr int xch_f = xch_i + 2 * LT_TILE_M * 2 * 4;           // float [128 rows][2 halves][2]
    // per row quarter: one row for the robust search (256 floats), then the staging area of the emission
    // (LT_STAGE_ENTRIES values, then as many column bytes; the robust search's 512 staging words overlay it)
    static constexpr int rowbuf = xch_f + LT_TILE_M * 2 * 2 * 4;
    static constexpr int quarter_bytes = 256 * 4 + LT_STAGE_ENTRIES * 5;
    static constexpr int bars = rowbuf + 4 * quarter_bytes;               // mbarriers (8 bytes each)
    static constexpr int n_bars = 3 * LT_STAGES + 4;                      // full, conv, empty per stage; tmem full/empty x2
    static constexpr int tmem_ptr = bars + n_bars * 8;
    static constexpr int total = tmem_ptr + 16;
};

// shared-memory matrix descriptor, K-major, rows of LT_BK fp32 = one swizzle span (cute/arch/mma_sm100_desc.hpp:
// SmemDescriptor): start address >> 4 at [0,14), leading byte offset (unused with a swizzle) = 1 at [16,30), stride
// byte offset = eight rows (1024 B or 512 B) >> 4 at [32,46), version 1 at [46,48), layout type at [61,64):
// SWIZZLE_128B = 2, SWIZZLE_64B = 4
__device__ __forceinline__ uint64_t smem_desc_kmajor(uint32_t addr)
{
    constexpr uint32_t hi = (LT_BK == 32 ? (2u << 29) : (4u << 29)) | (1u << 14) | (uint32_t)(8 * LT_BK * 4 / 16);
    return ((uint64_t)hi << 32) | (1u << 16) | ((addr >> 4) & 0x3fffu);
}
// if (!(x < t)) { value -> [w], column byte -> [wi], advance both }   (w, wi: 32-bit shared-memory addresses)
__device__ __forceinline__ void stage2_if_nlt(unsigned &w, unsigned &wi, float x, float t, unsigned col)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.geu.f32 p, %2, %3;\n\t"
        "@p st.shared.f32 [%0], %2;\n\t"
        "@p st.shared.u8 [%1], %4;\n\t"
        "@p add.u32 %0, %0, 4;\n\t"
        "@p add.u32 %1, %1, 1;\n\t}"
        : "+r"(w), "+r"(wi)
        : "f"(x), "f"(t), "r"(col)
        : "memory");
}

// ---- W -> (TF32 part, remainder), rows [0, 256) and [256, 512) of the workspace ------------------------------------
__global__ void lt_split_weight_kernel(const float *__restrict__ w, int64_t count, float *__restrict__ ws)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const float v = w[i];
    const float hi = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
    ws[i] = hi;
    ws[count + i] = tf32_lo_part(v, hi);
}

// ---- the fused kernel ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(LT_THREADS, 1)
linear_topk_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                   const float *__restrict__ bias, int64_t n, int in_dim, int k, float *__restrict__ out_val,
                   unsigned char *__restrict__ out_idx, float *__restrict__ linear_out, int vec_out,
                   const float *residual)
{
    extern __shared__ unsigned char lt_smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // the swizzled operand tiles need a 1024-byte aligned base (the launch asks for 1 KB of slack)
    const uint32_t sraw = smem_u32(lt_smem_raw);
    const uint32_t sbase = (sraw + 1023u) & ~1023u;
    unsigned char *lt_smem = lt_smem_raw + (sbase - sraw);
    const uint32_t bar0 = sbase + LtShared::bars;
    // barrier i of kind: full[s] = s, conv[s] = STAGES + s, empty[s] = 2 STAGES + s, tmem_full[b] = 3 STAGES + b,
    // tmem_empty[b] = 3 STAGES + 2 + b
    auto bar_full = [&](int s) { return bar0 + 8u * (uint32_t)s; };
    auto bar_conv = [&](int s) { return bar0 + 8u * (uint32_t)(LT_STAGES + s); };
    auto bar_empty = [&](int s) { return bar0 + 8u * (uint32_t)(2 * LT_STAGES + s); };
    auto bar_tfull = [&](int b) { return bar0 + 8u * (uint32_t)(3 * LT_STAGES + b); };
    auto bar_tempty = [&](int b) { return bar0 + 8u * (uint32_t)(3 * LT_STAGES + 2 + b); };
    uint32_t *tmem_ptr_smem = reinterpret_cast<uint32_t *>(lt_smem + LtShared::tmem_ptr);

    const int64_t num_tiles = (n + LT_TILE_M - 1) / LT_TILE_M;
    const int num_kc = (in_dim + LT_BK - 1) / LT_BK;

    // ---- one-time set-up -------------------------------------------------------------------------------------------
    if (threadIdx.x == 0) {
        for (int s = 0; s < LT_STAGES; ++s) {
            mbar_init(bar_full(s), 1);
            mbar_init(bar_conv(s), LT_SPLIT_THREADS);
            mbar_init(bar_empty(s), 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(bar_tfull(b), 1);
            mbar_init(bar_tempty(b), LT_EPI_THREADS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {  // TMEM: all 512 columns (two 128 x 256 fp32 accumulators); this warp frees them at the end
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)),
                     "r"(512)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    {
        float *sb = reinterpret_cast<float *>(lt_smem + LtShared::bias);
        for (int i = threadIdx.x; i < LT_D; i += LT_THREADS) sb[i] = bias != nullptr ? bias[i] : 0.f;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    // registers: the epilogue threads hold half a row (128 values); the first warpgroup gives its share away
    if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;" ::: "memory");
    if (warp == 0) {
        // ===== TMA producer ==========================================================================================
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                for (int kc = 0; kc < num_kc; ++kc) {
                    mbar_wait<true>(bar_empty(s), ph ^ 1u);
                    const uint32_t st = sbase + (uint32_t)(s * LT_STAGE_BYTES);
                    mbar_expect_tx(bar_full(s), LT_TX_BYTES);
                    tma_load_2d(st, &map_x, kc * LT_BK, (int)(tile * LT_TILE_M), bar_full(s));
                    tma_load_2d(st + 2 * LT_A_BYTES, &map_w, kc * LT_BK, 0, bar_full(s));
                    tma_load_2d(st + 2 * LT_A_BYTES + LT_B_BYTES, &map_w, kc * LT_BK, LT_D, bar_full(s));
                    if (++s == LT_STAGES) {
                        s = 0;
                        ph ^= 1u;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one thread) ================================================================================
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            int it = 0;
            for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
                const int b = it & 1;
                mbar_wait(bar_tempty(b), (((uint32_t)it >> 1) & 1u) ^ 1u);  // the epilogue has read accumulator b
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(b * LT_D);
                for (int kc = 0; kc < num_kc; ++kc) {
                    mbar_wait(bar_full(s), ph);
                    mbar_wait(bar_conv(s), ph);
                    tc_fence_after();
                    const uint32_t st = sbase + (uint32_t)(s * LT_STAGE_BYTES);
                    const uint64_t a_hi = smem_desc_kmajor(st), a_lo = smem_desc_kmajor(st + LT_A_BYTES);
                    const uint64_t b_hi = smem_desc_kmajor(st + 2 * LT_A_BYTES);
                    const uint64_t b_lo = smem_desc_kmajor(st + 2 * LT_A_BYTES + LT_B_BYTES);
#pragma unroll
                    for (int kk = 0; kk < LT_BK / 8; ++kk) {  // 8 tf32 = 32 bytes per K step: + 2 in 16-byte units
                        const uint64_t adv = (uint64_t)(2 * kk);
                        tc_mma_tf32(tmem_d, a_hi + adv, b_lo + adv, LT_IDESC, (kc | kk) != 0 ? 1u : 0u);
                        tc_mma_tf32(tmem_d, a_lo + adv, b_hi + adv, LT_IDESC, 1u);
                        tc_mma_tf32(tmem_d, a_hi + adv, b_hi + adv, LT_IDESC, 1u);
                    }
                    tc_commit(bar_empty(s));  // the ring slot is free once these MMAs have read it
                    if (kc == num_kc - 1) tc_commit(bar_tfull(b));
                    if (++s == LT_STAGES) {
                        s = 0;
                        ph ^= 1u;
                    }
                }
            }
        }
    } else if (warp < 4) {
        // ===== split X = hi + lo in shared memory =====================================================================
        const int t = threadIdx.x - 64;
        int s = 0;
        uint32_t ph = 0;
        for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            for (int kc = 0; kc < num_kc; ++kc) {
                mbar_wait(bar_full(s), ph);
                float4 *a_hi = reinterpret_cast<float4 *>(lt_smem + s * LT_STAGE_BYTES);
                float4 *a_lo = reinterpret_cast<float4 *>(lt_smem + s * LT_STAGE_BYTES + LT_A_BYTES);
#pragma unroll 4
                for (int i = t; i < LT_A_BYTES / 16; i += LT_SPLIT_THREADS) {
                    const float4 v = a_hi[i];
                    float4 h, l;
                    h.x = __uint_as_float(__float_as_uint(v.x) & 0xffffe000u);
                    h.y = __uint_as_float(__float_as_uint(v.y) & 0xffffe000u);
                    h.z = __uint_as_float(__float_as_uint(v.z) & 0xffffe000u);
                    h.w = __uint_as_float(__float_as_uint(v.w) & 0xffffe000u);
                    l.x = tf32_lo_part(v.x, h.x);
                    l.y = tf32_lo_part(v.y, h.y);
                    l.z = tf32_lo_part(v.z, h.z);
                    l.w = tf32_lo_part(v.w, h.w);
                    a_hi[i] = h;
                    a_lo[i] = l;
                }
                fence_proxy_async();  // generic-proxy writes -> visible to the tensor core's reads
                mbar_arrive(bar_conv(s));
                if (++s == LT_STAGES) {
                    s = 0;
                    ph ^= 1u;
                }
            }
        }
    }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 208;" ::: "memory");
        // ===== epilogue: bias, TopK, CBSR =============================================================================
        const int q = warp & 3;           // TMEM lane quarter this warp may read
        const int half = (warp - 4) >> 2; // columns [128 half, 128 half + 128)
        const int r = 32 * q + lane;      // row inside the tile = TMEM lane
        const int bar_id = 1 + q;         // named barrier of the two warps that share these 32 rows
        const float *sb = reinterpret_cast<const float *>(lt_smem + LtShared::bias) + half * LT_HALF;
        volatile int *xi = reinterpret_cast<volatile int *>(lt_smem + LtShared::xch_i);
        volatile float *xf = reinterpret_cast<volatile float *>(lt_smem + LtShared::xch_f);
        float *rowbuf = reinterpret_cast<float *>(lt_smem + LtShared::rowbuf + q * LtShared::quarter_bytes);
        float *sv = rowbuf + 256;                                                   // staged values
        unsigned char *si = reinterpret_cast<unsigned char *>(sv + LT_STAGE_ENTRIES);  // staged column bytes
        unsigned *rowstage = reinterpret_cast<unsigned *>(sv);                      // robust search: 512 words
        static_assert(LT_STAGE_ENTRIES * 5 >= 2 * MAXK_MAX_DIM * 4, "the robust search stages 512 words");
        const int rpp = LT_STAGE_ENTRIES / k < 32 ? LT_STAGE_ENTRIES / k : 32;      // rows per emission pass
        const unsigned full = 0xffffffffu;
        int it = 0;
        for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            const int b = it & 1;
            const int64_t row = tile * LT_TILE_M + r;
            float x[LT_HALF];
            mbar_wait(bar_tfull(b), ((uint32_t)it >> 1) & 1u);
            tc_fence_after();
            if (out_idx == nullptr) {
                // ---- the Linear alone (uniform): 16 columns at a time through shared memory, so that a store
                // instruction writes eight full 64-byte row segments instead of 32 half sectors 1 KB apart
                float *stg = reinterpret_cast<float *>(lt_smem + LtShared::rowbuf + q * LtShared::quarter_bytes) +
                             half * (32 * 20);  // 32 rows x 16 values, row stride 20 words: conflict-free STS.128
                const uint32_t taddr = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(b * LT_D + half * LT_HALF);
                const int64_t row0 = tile * LT_TILE_M + 32 * q;
#pragma unroll 1
                for (int c = 0; c < LT_HALF / 16; ++c) {
                    float v[16];
                    tmem_ld16(taddr + (uint32_t)(16 * c), v);
                    tmem_ld_wait();
                    if (c == LT_HALF / 16 - 1) {
                        tc_fence_before();
                        mbar_arrive(bar_tempty(b));
                    }
                    float4 *w4 = reinterpret_cast<float4 *>(stg + lane * 20);
#pragma unroll
                    for (int i = 0; i < 16; i += 4) {
                        const float4 bv = *reinterpret_cast<const float4 *>(sb + 16 * c + i);
                        w4[i >> 2] = make_float4(v[i] + bv.x, v[i + 1] + bv.y, v[i + 2] + bv.z, v[i + 3] + bv.w);
                    }
                    __syncwarp();
#pragma unroll
                    for (int i4 = 0; i4 < 4; ++i4) {
                        const int rr = i4 * 8 + (lane >> 2);
                        float4 val = *reinterpret_cast<const float4 *>(stg + rr * 20 + (lane & 3) * 4);
                        if (row0 + rr < n) {
                            const int64_t o = (row0 + rr) * LT_D + half * LT_HALF + 16 * c + (lane & 3) * 4;
                            if (residual != nullptr) {  // out = x W^T + b + residual (may be `out` itself)
                                const float4 rv = *reinterpret_cast<const float4 *>(residual + o);
                                val.x += rv.x;
                                val.y += rv.y;
                                val.z += rv.z;
                                val.w += rv.w;
                            }
                            st_stream_f4(reinterpret_cast<float4 *>(linear_out + o), val);
                        }
                    }
                    __syncwarp();
                }
                continue;
            }
            {
                const uint32_t taddr = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(b * LT_D + half * LT_HALF);
#pragma unroll
                for (int c = 0; c < LT_HALF / 32; ++c) tmem_ld32(taddr + (uint32_t)(32 * c), x + 32 * c);
                tmem_ld_wait();
            }
            tc_fence_before();
            mbar_arrive(bar_tempty(b));  // accumulator b may be overwritten: the search runs from registers
#pragma unroll
            for (int i = 0; i < LT_HALF; i += 4) {
                const float4 bv = *reinterpret_cast<const float4 *>(sb + i);
                x[i] += bv.x;
                x[i + 1] += bv.y;
                x[i + 2] += bv.z;
                x[i + 3] += bv.w;
            }
            if (linear_out != nullptr && row < n) {
                float4 *o = reinterpret_cast<float4 *>(linear_out + row * LT_D + half * LT_HALF);
#pragma unroll
                for (int i = 0; i < LT_HALF; i += 4) o[i >> 2] = make_float4(x[i], x[i + 1], x[i + 2], x[i + 3]);
            }

            // ---- bracket and first threshold (topk_select.cuh: TopkSeed).  k <= 64: ORDER STATISTICS instead of the row
            // minimum -- the minima over the maxima of 8 / 16 / 32 consecutive columns (32 / 16 / 8 groups per row) have
            // expected ranks 100.6 / 48.3 / 20.7 of 256 whatever the distribution of the values, the first threshold is
            // interpolated by rank between the two that enclose k, and the lower end of the bracket is the minimum over
            // the 32 (k <= 32) or 64 (groups of 4) maxima, which has that many elements not below it.  The maxima
            // fall out of the tree that computes the row maximum; the minimum pass of the old bracket is gone.  The
            // products of a Linear are close to normally distributed: the interpolation between minimum and maximum
            // put the first probe ~26 elements off there (4 on uniform values).  Both halves combine the partner's
            // four numbers in the same operand order: bit-identical brackets, hence identical control flow in the two
            // warps that meet at the named barrier.  The numbers travel through the row buffer of the robust search,
            // which is free until this tile's emission has passed its barriers.
            const TopkSeed seed = topk_seed(LT_D, k);
            const float nanf_ = __uint_as_float(0x7fc00000u);
            float fl, fh, tn = nanf_;
            if (seed.mode != 0) {
                const float pinf = __uint_as_float(0x7f800000u), ninf = __uint_as_float(0xff800000u);
                float sa = pinf, sb2 = pinf, sc = pinf, sg = pinf;
                fh = ninf;
#pragma unroll
                for (int j32 = 0; j32 < LT_HALF; j32 += 32) {
                    float m32 = ninf;
#pragma unroll
                    for (int j16 = 0; j16 < 32; j16 += 16) {
                        float m16 = ninf;
#pragma unroll
                        for (int j8 = 0; j8 < 16; j8 += 8) {
                            const int o = j32 + j16 + j8;
                            const float q0 = fmaxf(fmaxf(x[o], x[o + 1]), fmaxf(x[o + 2], x[o + 3]));
                            const float q1 = fmaxf(fmaxf(x[o + 4], x[o + 5]), fmaxf(x[o + 6], x[o + 7]));
                            if (seed.low == 2) sg = fminf(sg, fminf(q0, q1));
                            const float m8 = fmaxf(q0, q1);
                            sa = fminf(sa, m8);
                            m16 = fmaxf(m16, m8);
                        }
                        sb2 = fminf(sb2, m16);
                        m32 = fmaxf(m32, m16);
                    }
                    sc = fminf(sc, m32);
                    fh = fmaxf(fh, m32);
                }
                volatile float *xb = reinterpret_cast<volatile float *>(rowbuf);
                const int me = (lane * 2 + half) * 4, p0 = (lane * 2) * 4, p1 = (lane * 2 + 1) * 4;
                xb[me] = fh;
                xb[me + 1] = seed.low == 2 ? sg : sa;
                xb[me + 2] = seed.mode == 1 ? sb2 : (seed.mode == 2 ? sc : sa);
                xb[me + 3] = seed.mode == 1 ? sc : sb2;
                named_bar_sync(bar_id, 64);
                fh = fmaxf(xb[p0], xb[p1]);
                fl = fminf(xb[p0 + 1], xb[p1 + 1]);
                const float lo = fminf(xb[p0 + 2], xb[p1 + 2]);
                const float hi = seed.mode == 2 ? fh : fminf(xb[p0 + 3], xb[p1 + 3]);
                tn = fmaf(hi - lo, seed.w, lo);
            } else {
                fl = x[0];
                fh = x[0];
#pragma unroll
                for (int i = 1; i < LT_HALF; ++i) {
                    fl = fminf(fl, x[i]);
                    fh = fmaxf(fh, x[i]);
                }
                xf[(r * 2 + half) * 2] = fl;
                xf[(r * 2 + half) * 2 + 1] = fh;
                named_bar_sync(bar_id, 64);
                fl = fminf(xf[(r * 2) * 2], xf[(r * 2 + 1) * 2]);
                fh = fmaxf(xf[(r * 2) * 2 + 1], xf[(r * 2 + 1) * 2 + 1]);
            }

            // ---- regula falsi (Illinois weights) on g(t) = count(!(x < t)) - k; both halves of a row run the same
            // state machine on the same totals (identical control flow in the two warps that meet at the named
            // barrier).  state: 0 searching, 1 t separates exactly k elements, 2 robust search needed.
            // The 32 rows of a warp probe in lock step: 11.8 probes per warp against 3.9 per row (ncu, round 2, before
            // the order-statistics seed above: 1.83 -> 1.70 ms on the products-sized input with it).  An
            // exact finish -- bracket down to <= 16 candidates per half, write them to shared memory, select the
            // j-th largest, prove it with one counting pass -- was built and measured: SLOWER (3.3 vs 2.0 ms on the
            // products shape), its own extraction / selection passes cost more than the probes they replace.
            int state = 0, c = 0, c_other = 0, run = 0, par = 0;
            float t = fl;
            float gl = seed.gl, gh = seed.gh;
            if (k == LT_D) {
                t = __uint_as_float(0xff800000u);
                state = 1;
                c = LT_HALF;
                c_other = LT_HALF;
            }
#pragma unroll 1
            for (int probe = 0; probe < TOPK_FAST_MAX_PROBES; ++probe) {
                if (__all_sync(full, state != 0)) break;  // the partner warp holds the same rows: same decision
                if (state == 0) {
                    float fm = tn;  // the seeded first threshold; NaN afterwards: interpolate
                    tn = nanf_;
                    if (!(fm > fl && fm < fh)) fm = fmaf(fh - fl, gl * rcp_approx(gl - gh), fl);
                    if (!(fm > fl && fm < fh)) fm = 0.5f * fl + 0.5f * fh;
                    if (!(fm > fl && fm < fh)) state = 2;  // adjacent floats, infinite or NaN end points
                    t = fm;
                }
                int ca = 0, cb = 0, cc = 0, cd = 0;
#pragma unroll
                for (int i = 0; i < LT_HALF; i += 4) {
                    count_nlt(ca, x[i], t);
                    count_nlt(cb, x[i + 1], t);
                    count_nlt(cc, x[i + 2], t);
                    count_nlt(cd, x[i + 3], t);
                }
                c = (ca + cb) + (cc + cd);
                xi[(par * LT_TILE_M + r) * 2 + half] = c;
                named_bar_sync(bar_id, 64);
                c_other = xi[(par * LT_TILE_M + r) * 2 + (half ^ 1)];
                par ^= 1;
                const int g = (half ? c_other + c : c + c_other) - k;
                if (state == 0) {
                    if (g == 0) {
                        state = 1;
                    } else if (g > 0) {
                        fl = t;
                        gl = (float)g;
                        if (run > 0) gh *= 0.5f;
                        run = run > 0 ? run + 1 : 1;
                    } else {
                        fh = t;
                        gh = (float)g;
                        if (run < 0) gl *= 0.5f;
                        run = run < 0 ? run - 1 : -1;
                    }
                    if (run >= TOPK_FAST_MAX_RUN || run <= -TOPK_FAST_MAX_RUN) state = 2;
                }
            }
            if (state == 0) state = 2;

            // ---- emission: ascending columns, the upper half behind the lower half's entries; staged in shared
            // memory (rpp rows per pass), then copied out coalesced -- the 32 rows of a quarter are contiguous
            const unsigned okmask = __ballot_sync(full, state == 1 && row < n);
#pragma unroll 1
            for (int first = 0; first < 32; first += rpp) {
                const int last = first + rpp < 32 ? first + rpp : 32;
                if (lane >= first && lane < last && ((okmask >> lane) & 1u)) {
                    const unsigned pos = (unsigned)((lane - first) * k + (half ? c_other : 0));
                    unsigned w = smem_u32(sv + pos), wi = smem_u32(si + pos);
                    const unsigned col0 = (unsigned)(half * LT_HALF);
#pragma unroll
                    for (int i = 0; i < LT_HALF; ++i) stage2_if_nlt(w, wi, x[i], t, col0 + i);
                }
                named_bar_sync(bar_id, 64);
                if (rpp == 32 && okmask == full && vec_out) {
                    // the common case: all 32 rows staged, contiguous in sp_data / sp_index -> 128-bit copies
                    const int64_t o = (tile * LT_TILE_M + 32 * q) * k;
                    const int tid = half * 32 + lane;
                    float4 *gv = reinterpret_cast<float4 *>(out_val + o);
                    uint4 *gi = reinterpret_cast<uint4 *>(out_idx + o);
                    const float4 *s4 = reinterpret_cast<const float4 *>(sv);
                    const uint4 *i4 = reinterpret_cast<const uint4 *>(si);
                    for (int e = tid; e < 8 * k; e += 64) gv[e] = s4[e];
                    for (int e = tid; e < 2 * k; e += 64) gi[e] = i4[e];
                } else
#pragma unroll 1
                for (int rr = first + half; rr < last; rr += 2) {  // the two warps take alternate rows
                    if (!((okmask >> rr) & 1u)) continue;
                    const int64_t o = (tile * LT_TILE_M + 32 * q + rr) * k;
                    const float *rv = sv + (rr - first) * k;
                    const unsigned char *ri = si + (rr - first) * k;
                    for (int e = lane; e < k; e += 32) {
                        out_val[o + e] = rv[e];
                        out_idx[o + e] = ri[e];
                    }
                }
                named_bar_sync(bar_id, 64);  // the staging area may be refilled (next pass, robust search)
            }

            // ---- rows for the robust search: through shared memory into the one-row-per-warp layout
            unsigned todo = __ballot_sync(full, state == 2);
            while (todo) {  // identical in both partner warps
                const int src = __ffs(todo) - 1;
                todo &= todo - 1;
                const int64_t frow = tile * LT_TILE_M + 32 * q + src;
                if (frow >= n) continue;
                if (lane == src) {
#pragma unroll
                    for (int i = 0; i < LT_HALF; ++i) rowbuf[half * LT_HALF + i] = x[i];
                }
                named_bar_sync(bar_id, 64);
                if (half == 0) {
                    const int c0 = 4 * lane, c1 = 128 + 4 * lane;
                    const float4 a = *reinterpret_cast<const float4 *>(rowbuf + c0);
                    const float4 bb = *reinterpret_cast<const float4 *>(rowbuf + c1);
                    const unsigned raw[8] = {__float_as_uint(a.x),  __float_as_uint(a.y),  __float_as_uint(a.z),
                                             __float_as_uint(a.w),  __float_as_uint(bb.x), __float_as_uint(bb.y),
                                             __float_as_uint(bb.z), __float_as_uint(bb.w)};
                    const unsigned sel = topk_select_slow_f32(raw[0], raw[1], raw[2], raw[3], raw[4], raw[5], raw[6],
                                                              raw[7], true, true, LT_D, k, lane);
                    topk_emit<false>(sel, raw, rowstage, lane, k, frow, c0, c1, out_val, out_idx);
                }
                named_bar_sync(bar_id, 64);  // the row buffer may be refilled
            }
        }
    }

    // ---- teardown ------------------------------------------------------------------------------------------------
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

}  // namespace maxk

using namespace maxk;

extern "C" int64_t maxk_linear_topk_workspace_bytes(int32_t d, int32_t in_dim)
{
    if (d <= 0 || in_dim <= 0) return 0;
    return (int64_t)2 * d * in_dim * (int64_t)sizeof(float);
}

static int linear_topk_launch(const float *x, int64_t n, int32_t in_dim, const float *weight, const float *bias,
                              int32_t d, int32_t k, float *sp_data, uint8_t *sp_index, float *linear_out,
                              const float *residual, void *workspace, void *stream)
{
    MAXK_REQUIRE(n >= 0, "maxk_linear_topk_cbsr / maxk_linear_f32: negative n");
    MAXK_REQUIRE(d == LT_D, "maxk_linear_topk_cbsr / maxk_linear_f32: output width d=%d must be 256 (one MMA spans the row)", d);
    MAXK_REQUIRE(in_dim >= 4 && in_dim % 4 == 0, "maxk_linear_topk_cbsr / maxk_linear_f32: in_dim=%d must be a positive multiple of 4",
                 in_dim);
    MAXK_REQUIRE(k >= 1 && k <= d, "maxk_linear_topk_cbsr / maxk_linear_f32: Invalid k value (k=%d, d=%d)", k, d);
    if (n == 0) return MAXK_OK;
    MAXK_REQUIRE((sp_data == nullptr) == (sp_index == nullptr),
                 "maxk_linear_topk_cbsr / maxk_linear_f32: sp_data and sp_index must both be given or both NULL");
    MAXK_REQUIRE(sp_index != nullptr || linear_out != nullptr, "maxk_linear_topk_cbsr / maxk_linear_f32: no output");
    MAXK_REQUIRE(x != nullptr && weight != nullptr && workspace != nullptr,
                 "maxk_linear_topk_cbsr / maxk_linear_f32: x, weight and workspace must not be NULL");
    MAXK_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(workspace) & 15) == 0 &&
                     (reinterpret_cast<uintptr_t>(linear_out) & 15) == 0,
                 "maxk_linear_topk_cbsr / maxk_linear_f32: x, workspace and linear_out must be 16-byte aligned");
    MAXK_REQUIRE(n < (int64_t)1 << 31, "maxk_linear_topk_cbsr / maxk_linear_f32: n too large for one tensor map");
    cudaStream_t st = static_cast<cudaStream_t>(stream);

    float *ws = static_cast<float *>(workspace);
    const int64_t wcount = (int64_t)d * in_dim;
    lt_split_weight_kernel<<<(unsigned)ceil_div<int64_t>(wcount, 256), 256, 0, st>>>(weight, wcount, ws);
    MAXK_LAUNCH_CHECK("lt_split_weight_kernel");

    CUtensorMap map_x, map_w;
    if (int rc = make_map_f32(&map_x, x, n, in_dim, LT_TILE_M, LT_BK, "x")) return rc;
    if (int rc = make_map_f32(&map_w, ws, 2 * (int64_t)d, in_dim, LT_D, LT_BK, "weight")) return rc;

    static std::mutex mu;
    static std::set<int> opted;
    int dev = 0, sms = 0;
    MAXK_CUDA(cudaGetDevice(&dev));
    {
        std::lock_guard<std::mutex> lock(mu);
        if (!opted.count(dev)) {
            MAXK_CUDA(cudaFuncSetAttribute(linear_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           LtShared::total + 1024));
            opted.insert(dev);
        }
    }
    MAXK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int64_t tiles = ceil_div<int64_t>(n, LT_TILE_M);
    const unsigned grid = (unsigned)(tiles < sms ? tiles : sms);
    const int vec_out = (k % 4 == 0) && (reinterpret_cast<uintptr_t>(sp_data) & 15) == 0 &&
                        (reinterpret_cast<uintptr_t>(sp_index) & 15) == 0;
    linear_topk_kernel<<<grid, LT_THREADS, LtShared::total + 1024, st>>>(map_x, map_w, bias, n, in_dim, k, sp_data,
                                                                         sp_index, linear_out, vec_out, residual);
    MAXK_LAUNCH_CHECK("linear_topk_kernel");
    return MAXK_OK;
}

extern "C" int maxk_linear_topk_cbsr(const float *x, int64_t n, int32_t in_dim, const float *weight, const float *bias,
                                     int32_t d, int32_t k, float *sp_data, uint8_t *sp_index, float *linear_out,
                                     void *workspace, void *stream)
{
    return linear_topk_launch(x, n, in_dim, weight, bias, d, k, sp_data, sp_index, linear_out, nullptr, workspace, stream);
}

extern "C" int maxk_linear_f32(const float *x, int64_t n, int32_t in_dim, const float *weight, const float *bias,
                               const float *residual, int32_t d, float *out, void *workspace, void *stream)
{
    MAXK_REQUIRE(out != nullptr || n == 0, "maxk_linear_f32: out is NULL");
    MAXK_REQUIRE((reinterpret_cast<uintptr_t>(residual) & 15) == 0, "maxk_linear_f32: residual must be 16-byte aligned");
    return linear_topk_launch(x, n, in_dim, weight, bias, d, 1, nullptr, nullptr, out, residual, workspace, stream);
}
