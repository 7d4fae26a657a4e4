// topk_cbsr.cu -- MaxK nonlinearity: exact per-row top-k of d <= 256 features, emitted as
// CBSR (k values + k uint8 column ids, ascending column order, ties -> lowest column).
//
// Replaces `topk` (ref: kernels/maxk_kernel.cu:23-96) and its float facade
// (cuda_kernel_bindings.cpp:203-238).  The reference idea -- search for a threshold with
// warp-wide counts, then compact with a scan -- is kept, but done on the fp32 values
// themselves (no uint8 quantisation), with a warp-uniform threshold (Q5/Q6), for
// any k and N (Q7) and with an exact tie rule (Q8):
//   * one warp per row, 8 elements per lane in registers (two 128-bit loads per lane);
//   * fp32 fast path: the threshold is searched in VALUE space with plain float compares
//     (`!(x < t)`, so NaN ranks largest and -0 == +0 for free); one REDUX-counted probe (8 FSETP + 8 SEL +
//     4 IADD3 + POPC + REDUX) per step, ending on the first probe that separates exactly k elements.
//     The first probe comes from ORDER STATISTICS of the lane maxima (CREDUX.MIN/MAX.F32: no keys, no
//     conversions; topk_select.cuh: TopkSeed) and lands at the right rank whatever the distribution; once a
//     probe is within 8 elements of k the next threshold is NAMED by extraction from the lanes' nearest values
//     (topk_extract_threshold) instead of interpolated: 2.4 probes + 1.1 extractions per row on uniform data,
//     2.7 + 1.1 on normal data (round 1 / early round 2: Illinois regula falsi between the row minimum and
//     maximum, 4.0 / 6.4 probes);
//   * rows the fast path cannot finish (ties at the k-th value, +-inf end points, a bracket
//     that collapsed to adjacent floats) fall to the robust key-space search, which snaps its probes to
//     elements (whole value levels per step) and resolves ties by column order exactly;
//   * emission positions from one packed warp scan, staged in shared memory so the global
//     stores are coalesced; optional dense x*mask output (MaxK.forward,
//     utils/models.py:48-52) from the same registers.
// Pure streaming kernel: 4*d bytes read + 5*k bytes written per row -> HBM-bound by design;
// the instruction count per row decides how close it gets: ncu (rounds 1 and 2) shows it ISSUE-bound
// (smsp__issue_active 78 %, DRAM 27 %).  ~350 warp instructions per row now (set-up ~66, 2.45 probes x 55,
// 1.1 extractions x 45, emission ~97); 368 before the order-statistics seed and the extraction (510 on
// normal data).  232 965 x 256, k = 32: 79 us uniform / 83 us normal (99.4 / 131.2 before).
// Measured and dropped in round 2 (profiles/r02_variants.md, r02_topk.md): several rows per warp with
// the next row prefetched (8-17 % slower: the loop costs more than the latency it hides) and two rows
// per warp in half-warps (the warp-uniform overhead is shared, but 16 elements per lane compile to a
// longer compare / staging sequence: 128 vs 100 us).
#include "common.cuh"
#include "topk_select.cuh"

namespace maxk {

// ---- uint8 rows (the reference kernel's own data type): key-space search only ------------
__global__ void __launch_bounds__(TOPK_WPB * 32)
topk_cbsr_u8_kernel(const unsigned char *__restrict__ x_, int64_t n, int d, int k,
                    unsigned char *__restrict__ out_val, unsigned char *__restrict__ out_idx)
{
    __shared__ unsigned s_stage[TOPK_WPB][2 * MAXK_MAX_DIM];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t row = (int64_t)blockIdx.x * TOPK_WPB + warp;
    if (row >= n) return;  // warp-uniform; no CTA barrier below
    const int c0 = 4 * lane, c1 = 128 + 4 * lane;
    const bool ok0 = c0 < d, ok1 = c1 < d;
    const unsigned char *x = x_ + row * d;
    const unsigned wa = ok0 ? __ldcs(reinterpret_cast<const unsigned *>(x + c0)) : 0u;
    const unsigned wb = ok1 ? __ldcs(reinterpret_cast<const unsigned *>(x + c1)) : 0u;
    unsigned key[8], raw[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        raw[i] = (wa >> (8 * i)) & 0xff;
        raw[4 + i] = (wb >> (8 * i)) & 0xff;
        key[i] = ok0 ? raw[i] + 1u : 0u;  // 0 is reserved for "no element"
        key[4 + i] = ok1 ? raw[4 + i] + 1u : 0u;
    }
    const unsigned sel = topk_select_by_keys<true>(key, d, k, lane);
    topk_emit<true>(sel, raw, s_stage[warp], lane, k, row, c0, c1, out_val, out_idx);
}

// ---- fp32 rows ----------------------------------------------------------------------------
// One row, held by the warp as two float4 per lane (columns 4*lane.. and 128 + 4*lane..; -inf where
// the column does not exist): threshold search, emission, optional dense output.
// FULL: d == 256 at compile time -- every column exists, no -inf fill, no column predicates.
// OUT: what the launch writes, known at compile time -- bit 0 the CBSR pair, bit 1 the dense x*mask, bit 2: k <= 32 --
// or 0: decided from the pointers and k at run time (every instruction of this function is paid per row)
// CFG: the seed's (mode, low) pair at compile time -- 1: (1, 1) k = 21..32, 2: (2, 1) k <= 20, 3: (1, 2) k = 33..48,
// 4: (3, 2) k = 49..64 -- or 0: read from `seed` (TopkSeed: what the numbers mean)
template <bool FULL, int OUT = 0, int CFG = 0>
__device__ __forceinline__ void topk_row_f32(float4 a, float4 b, bool ok0, bool ok1, int c0, int c1, int lane,
                                             int64_t row, int d, int k, TopkSeed seed, unsigned *stage,
                                             float *__restrict__ out_val, unsigned char *__restrict__ out_idx,
                                             float *__restrict__ dense_out)
{
    const float pinf = __uint_as_float(0x7f800000u);
    const float x[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    if constexpr (FULL) {
        ok0 = ok1 = true;
        d = MAXK_MAX_DIM;
    }

    // ---- bracket and first threshold ------------------------------------------------------------------------------
    // fmaxf / CREDUX skip NaN (NaN elements rank largest and pass every `!(x < t)` on their own).
    const float hi0 = fmaxf(fmaxf(a.x, a.y), fmaxf(a.z, a.w)), hi1 = fmaxf(fmaxf(b.x, b.y), fmaxf(b.z, b.w));
    const float lmax = fmaxf(hi0, hi1);
    float fh = warp_max_f32(lmax), fl, gl = seed.gl, gh = seed.gh, t;
    const int mode = CFG == 0 ? seed.mode : (CFG == 2 ? 2 : (CFG == 4 ? 3 : 1));
    const int low = CFG == 0 ? seed.low : (CFG <= 2 ? 1 : 2);
    if (mode != 0) {
        // Order statistics of the lane maxima instead of the row minimum (full-width rows, k <= 64; topk_seed() has
        // the ranks): a shuffle + FMNMX + CREDUX each, and the first probe lands at the right RANK whatever the shape
        // of the distribution (the linear interpolation between minimum and maximum it replaces needed 6.4 probes per
        // row on normal data against 4.0 on uniform data).
        const float A = warp_min_f32(lmax);
        const float pm = fmaxf(lmax, __shfl_xor_sync(0xffffffffu, lmax, 1));
        const float B = warp_min_f32(pm);
        float lo = A, hi = B;
        if (mode != 3) {
            const float C = warp_min_f32(fmaxf(pm, __shfl_xor_sync(0xffffffffu, pm, 2)));
            lo = mode == 1 ? B : C;
            hi = mode == 1 ? C : fh;
        }
        fl = low == 1 ? A : warp_min_f32(fminf(hi0, hi1));
        t = fmaf(hi - lo, seed.w, lo);
    } else {
        const float lo0 = fminf(fminf(a.x, a.y), fminf(a.z, a.w)), lo1 = fminf(fminf(b.x, b.y), fminf(b.z, b.w));
        fl = warp_min_f32(FULL ? fminf(lo0, lo1) : fminf(ok0 ? lo0 : pinf, ok1 ? lo1 : pinf));
        t = fmaf(fh - fl, seed.w, fl);  // gl = d - k is exact: every element, NaN included, is not below the minimum
    }

    // ---- fast path: regula falsi (Illinois weights) on g(t) = count(!(x < t)) - k, finished by extraction -------
    // gl / gh only steer the probes.  A probe with g == 0 is a proof -- exactly k elements are not below t -- so the
    // result does not depend on any estimate; once a probe is within TOPK_EXTRACT_MAX elements of k the next
    // threshold is NAMED (topk_extract_threshold) instead of interpolated, which usually ends the search with the
    // following probe: 2.4 probes + 1.1 extractions per row on uniform data, 2.7 + 1.1 on normal data (was 4.0 / 6.4
    // probes).  Anything the loop cannot finish leaves to the key search.
    unsigned sel = 0u;
    bool done = false;
    {   // (k == d needs no special case: no probe can prove it, the key search keeps every element)
        int run = 0;  // +n: fl moved n times in a row, -n: fh did
#pragma unroll 1
        for (int it = 0; it < TOPK_FAST_MAX_PROBES; ++it) {
            if (!(t > fl && t < fh)) {
                t = fmaf(fh - fl, gl * rcp_approx(gl - gh), fl);
                if (!(t > fl && t < fh)) {
                    t = 0.5f * fl + 0.5f * fh;
                    if (!(t > fl && t < fh)) break;  // adjacent floats, infinite or NaN end points
                }
            }
            sel = lane_mask_nlt(x, t);
            const int g = __reduce_add_sync(0xffffffffu, __popc(sel)) - k;
            if (g == 0) {
                done = true;
                break;
            }
            if (g > 0) {
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
            if (run >= TOPK_FAST_MAX_RUN || run <= -TOPK_FAST_MAX_RUN) break;
            if (g >= -TOPK_EXTRACT_MAX && g <= TOPK_EXTRACT_MAX)
                t = topk_extract_threshold(x, t, g);
            else
                t = fl;  // outside the bracket: interpolate
        }
    }
    unsigned raw[8];  // fp32 bit patterns (bit-copied to the output)
#pragma unroll
    for (int i = 0; i < 8; ++i) raw[i] = __float_as_uint(x[i]);
    if (!done)
        sel = topk_select_slow_f32(raw[0], raw[1], raw[2], raw[3], raw[4], raw[5], raw[6], raw[7],
                                   ok0, ok1, d, k, lane);

    if (OUT == 0 ? out_idx != nullptr : (OUT & 1) != 0)
        topk_emit<false, (OUT & 4) != 0>(sel, raw, stage, lane, k, row, c0, c1, out_val, out_idx);
    if (OUT == 0 ? dense_out != nullptr : (OUT & 2) != 0) {
        float *o = dense_out + row * d;
        float4 va, vb;
        va.x = (sel & 1u) ? x[0] : 0.f;
        va.y = (sel & 2u) ? x[1] : 0.f;
        va.z = (sel & 4u) ? x[2] : 0.f;
        va.w = (sel & 8u) ? x[3] : 0.f;
        vb.x = (sel & 16u) ? x[4] : 0.f;
        vb.y = (sel & 32u) ? x[5] : 0.f;
        vb.z = (sel & 64u) ? x[6] : 0.f;
        vb.w = (sel & 128u) ? x[7] : 0.f;
        if (ok0) st_stream_f4(reinterpret_cast<float4 *>(o + c0), va);
        if (ok1) st_stream_f4(reinterpret_cast<float4 *>(o + c1), vb);
    }
}

// 48 warps per SM by launch bounds (40 registers): the out-of-line robust search would otherwise set the kernel's
// register count (45: 40 warps, measured 6 % slower); 64 warps (32 registers, two spills) measure the same as 48
// (profiles/r02_topk_occupancy.jsonl) -- the kernel is bound by issue slots, not by latency.  CTAs of TWO warps: a CTA
// keeps its slot until its slowest row is done (ncu: 59 % of the warp slots active with eight warps per CTA against
// 75 % allocated); 8 / 4 / 2 warps measure 853 / 841 / 837 us on the products-sized activation, and a persistent
// grid-stride loop over rows 1051-1083 us (`profiles/r02_topk_cta_shape.jsonl`; the loop form lost in round 2's first
// attempt as well).
template <bool FULL, int OUT = 0, int CFG = 0>
__global__ void __launch_bounds__(TOPK_F32_WPB * 32, 48 / TOPK_F32_WPB)
topk_cbsr_f32_kernel(const float *__restrict__ x_, int64_t n, int d, int k, TopkSeed seed,
                     float *__restrict__ out_val, unsigned char *__restrict__ out_idx,
                     float *__restrict__ dense_out)
{
    __shared__ unsigned s_stage[TOPK_F32_WPB][2 * MAXK_MAX_DIM];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t row = (int64_t)blockIdx.x * TOPK_F32_WPB + warp;
    if (row >= n) return;  // warp-uniform; no CTA barrier below

    const int c0 = 4 * lane, c1 = 128 + 4 * lane;
    if constexpr (FULL) {
        const float4 *xr = reinterpret_cast<const float4 *>(x_ + row * MAXK_MAX_DIM);
        const float4 a = ld_stream_f4(xr + lane), b = ld_stream_f4(xr + 32 + lane);
        topk_row_f32<true, OUT, CFG>(a, b, true, true, c0, c1, lane, row, MAXK_MAX_DIM, k, seed, s_stage[warp], out_val,
                                     out_idx, dense_out);
    } else {
        const bool ok0 = c0 < d, ok1 = c1 < d;
        const float *xr = x_ + row * d;
        const float ninf = __uint_as_float(0xff800000u);
        float4 a = make_float4(ninf, ninf, ninf, ninf), b = a;  // columns >= d never qualify
        if (ok0) a = ld_stream_f4(reinterpret_cast<const float4 *>(xr + c0));
        if (ok1) b = ld_stream_f4(reinterpret_cast<const float4 *>(xr + c1));
        topk_row_f32<false>(a, b, ok0, ok1, c0, c1, lane, row, d, k, seed, s_stage[warp], out_val, out_idx, dense_out);
    }
}


// ---- LayerNorm + TopK: the normalisation that closes layer i and the selection that opens layer i + 1 -----------
// (SAGE: `h = self.norm(h)` then `MaxK(h)` of the next layer, utils/models.py:252 / 230; the normalised activation
// feeds nothing else, so it need not reach HBM.)  One warp per row: mean / variance / affine in registers exactly as
// layernorm.cu does, then the row goes into the search above; mean and rstd are kept for the backward
// (layernorm.cu: layernorm_bwd_cbsr_kernel takes the k-sparse gradient as it comes back from the aggregation).
// FULL: d == 256 and k <= 32 at compile time (what the 256-wide layers launch): no column predicates, no -inf fill, the
// emission without its k test.
template <bool FULL, int CFG = 0>
__global__ void __launch_bounds__(TOPK_F32_WPB * 32, 48 / TOPK_F32_WPB)
layernorm_topk_kernel(const float *__restrict__ x_, int64_t n, int d, const float *__restrict__ gamma,
                      const float *__restrict__ beta, float eps, int k, TopkSeed seed, float *__restrict__ out_val,
                      unsigned char *__restrict__ out_idx, float *__restrict__ mean_out, float *__restrict__ rstd_out,
                      float *__restrict__ y_out)
{
    __shared__ unsigned s_stage[TOPK_F32_WPB][2 * MAXK_MAX_DIM];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t row = (int64_t)blockIdx.x * TOPK_F32_WPB + warp;
    if (row >= n) return;  // warp-uniform; no CTA barrier below
    const int c0 = 4 * lane, c1 = 128 + 4 * lane;
    if constexpr (FULL) d = MAXK_MAX_DIM;
    const bool ok0 = FULL || c0 < d, ok1 = FULL || c1 < d;
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f), one = make_float4(1.f, 1.f, 1.f, 1.f);
    const float *xr = x_ + row * d;
    const float4 a = ok0 ? ld_stream_f4(reinterpret_cast<const float4 *>(xr + c0)) : zero;
    const float4 b = ok1 ? ld_stream_f4(reinterpret_cast<const float4 *>(xr + c1)) : zero;
    const float4 ga = ok0 ? (gamma ? __ldg(reinterpret_cast<const float4 *>(gamma + c0)) : one) : zero;
    const float4 gb = ok1 ? (gamma ? __ldg(reinterpret_cast<const float4 *>(gamma + c1)) : one) : zero;
    const float4 ba = ok0 && beta ? __ldg(reinterpret_cast<const float4 *>(beta + c0)) : zero;
    const float4 bb = ok1 && beta ? __ldg(reinterpret_cast<const float4 *>(beta + c1)) : zero;
    const float inv_d = 1.f / (float)d;
    float sum = (a.x + a.y) + (a.z + a.w) + (b.x + b.y) + (b.z + b.w);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum * inv_d;
    float4 da = make_float4(a.x - mean, a.y - mean, a.z - mean, a.w - mean);
    float4 db = make_float4(b.x - mean, b.y - mean, b.z - mean, b.w - mean);
    if (!ok0) da = zero;
    if (!ok1) db = zero;
    float sq = (da.x * da.x + da.y * da.y) + (da.z * da.z + da.w * da.w) + (db.x * db.x + db.y * db.y) +
               (db.z * db.z + db.w * db.w);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    const float rstd = rsqrtf(sq * inv_d + eps);
    if (lane == 0) {
        mean_out[row] = mean;
        rstd_out[row] = rstd;
    }
    const float ninf = __uint_as_float(0xff800000u);
    float4 ya = make_float4(ninf, ninf, ninf, ninf), yb = ya;  // columns >= d never qualify
    if (ok0)
        ya = make_float4(fmaf(da.x * rstd, ga.x, ba.x), fmaf(da.y * rstd, ga.y, ba.y), fmaf(da.z * rstd, ga.z, ba.z),
                         fmaf(da.w * rstd, ga.w, ba.w));
    if (ok1)
        yb = make_float4(fmaf(db.x * rstd, gb.x, bb.x), fmaf(db.y * rstd, gb.y, bb.y), fmaf(db.z * rstd, gb.z, bb.z),
                         fmaf(db.w * rstd, gb.w, bb.w));
    if (y_out != nullptr) {
        if (ok0) st_stream_f4(reinterpret_cast<float4 *>(y_out + row * d + c0), ya);
        if (ok1) st_stream_f4(reinterpret_cast<float4 *>(y_out + row * d + c1), yb);
    }
    if constexpr (FULL)
        topk_row_f32<true, 5, CFG>(ya, yb, true, true, c0, c1, lane, row, MAXK_MAX_DIM, k, seed, s_stage[warp], out_val,
                                   out_idx, nullptr);
    else
        topk_row_f32<false>(ya, yb, ok0, ok1, c0, c1, lane, row, d, k, seed, s_stage[warp], out_val, out_idx, nullptr);
}

// ---- fp32 rows, several rows per warp (round 2, second form) -------------------------------------
// The one-row-per-warp kernel above is issue-bound (ncu: ~365 warp instructions per row, issue slots 80 %
// busy, DRAM 27 %) and most of what it issues is warp-UNIFORM: the bracket update, the reductions over 32
// lanes and the 32-lane scans cost the same whether a lane holds 8 elements or 64.  Here L = 4 or 8 lanes
// own a row (E = D / L elements per lane IN CONTIGUOUS COLUMNS, so that one short scan over L lanes places
// a lane's output), a warp works on R = 32 / L rows at once, and every warp-wide instruction serves R rows:
//   * the rows arrive through cp.async (coalesced 16-byte chunks, no registers held while in flight) in a
//     padded shared-memory layout -- every lane segment shifted by 4 words -- from which each lane reads its
//     E contiguous values with conflict-free LDS.128 (a quarter-warp's eight segments start in eight
//     different bank quads);
//   * minimum / maximum: E FMNMX each + log2(L) shuffles (no keys, no REDUX);
//   * the same value-space regula falsi as above, every row group with its own bracket; a probe costs two
//     instructions per element (compare + predicated add) and log2(L) shuffle steps, the bracket update is
//     paid once per R rows; groups that are finished keep their threshold and idle along until the slowest
//     group of the warp is done (measured probe counts: 3.7 on average per row, ~5 for the slowest of 8);
//   * emission: exclusive scan of the lanes' counts within the group, then per element compare +
//     predicated 64-bit store of (value bits, column) into a staging area that has the layout of the
//     OUTPUT (the R rows of a warp are contiguous in sp_data and in sp_index), copied out coalesced;
//   * rows the fast path cannot finish (ties at the k-th value, infinite end points, NaN-only groups) are
//     re-read in the one-row layout and go through the robust key search above, one at a time.
// Used for D = 128 (L = 4) and D = 64 (L = 2) with k <= D / 2 (the staging area of the output overlays the input
// rows).  Measured in round 2 with L = 1, 2, 4, 8 at every width (profiles/r02_topk_grouped.md): at D = 256 the rows
// of a warp probing in lock step cost more than the shared overhead saves, and the one-row kernel stays.
template <int L>
__device__ __forceinline__ int group_sum(int v)
{
#pragma unroll
    for (int o = L / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// staging of one robustly selected row (one-row layout of `sel`): values at sv[], columns at si[]
__device__ __forceinline__ void topk_stage_row(unsigned sel, const unsigned (&raw)[8], unsigned *sv, unsigned *si,
                                               int lane, int c0, int c1)
{
    const unsigned n0 = __popc(sel & 0xfu), n1 = __popc(sel >> 4);
    const unsigned inc = warp_incl_scan(n0 | (n1 << 16), lane);
    const unsigned tot0 = __shfl_sync(0xffffffffu, inc, 31) & 0xffffu;
    unsigned p0 = (inc & 0xffffu) - n0, p1 = tot0 + (inc >> 16) - n1;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (sel & (1u << i)) {
            sv[p0] = raw[i];
            si[p0] = (unsigned)(c0 + i);
            ++p0;
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (sel & (16u << i)) {
            sv[p1] = raw[4 + i];
            si[p1] = (unsigned)(c1 + i);
            ++p1;
        }
    }
}

// if (!(x < t)) { value -> [w], column -> [w + IDX_OFF], w += 4 }   (w: 32-bit shared-memory address)
template <int IDX_OFF>
__device__ __forceinline__ void stage_if_nlt(unsigned &w, float x, float t, unsigned col)
{
    asm volatile(
        "{.reg .pred p; setp.geu.f32 p, %1, %2; @p st.shared.f32 [%0], %1; @p st.shared.u32 [%0+%4], %3; "
        "@p add.u32 %0, %0, 4;}"
        : "+r"(w)
        : "f"(x), "f"(t), "r"(col), "n"(IDX_OFF)
        : "memory");
}

constexpr int TOPK_GROUP_WPB = 8;
template <int D, int L>
constexpr int topk_group_smem_words_per_warp() { return (32 / L) * (D + 4 * L); }

template <int D, int L>
__global__ void __launch_bounds__(TOPK_GROUP_WPB * 32)
topk_cbsr_f32_group_kernel(const float *__restrict__ x_, int64_t n, int k, float *__restrict__ out_val,
                           unsigned char *__restrict__ out_idx, float *__restrict__ dense_out, int vec_out)
{
    constexpr int E = D / L;        // elements per lane
    constexpr int R = 32 / L;       // rows per warp
    constexpr int J = E / 4;        // float4 per lane
    constexpr int C4 = D / 4;       // float4 per row
    constexpr int SEG = E + 4;      // words per lane segment in shared memory (4 words of padding)
    constexpr int RS = L * SEG;     // words per row in shared memory
    constexpr int IDXW = R * D / 2; // staging area of the output: R * k <= IDXW values, then as many columns
    static_assert(E % 4 == 0 && (R * C4) % 32 == 0, "row groups must tile the warp");
    extern __shared__ __align__(16) unsigned s_topk_group[];

    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t row0 = ((int64_t)blockIdx.x * TOPK_GROUP_WPB + warp) * R;
    if (row0 >= n) return;  // warp-uniform; no CTA barrier below
    unsigned *sm = s_topk_group + warp * (R * RS);
    const int g = lane / L, l = lane % L;

    // ---- the warp's R rows -> shared memory (rows past the end: the last row again, never written out)
    const int rows_here = (n - row0 < R) ? (int)(n - row0) : R;
    {   // rows past the end are not loaded: their groups work on stale shared memory and are never written out
        const float *xw = x_ + row0 * D;
#pragma unroll
        for (int it = 0; it < R * C4 / 32; ++it) {
            const int q = it * 32 + lane;
            const int gg = q / C4, c4 = q % C4;
            if (gg < rows_here) cp_async16(sm + gg * RS + (c4 / J) * SEG + (c4 % J) * 4, xw + gg * D + 4 * c4);
        }
    }
    cp_async_commit();
    cp_async_wait<0>();
    __syncwarp();
    float x[E];
    {
        const float4 *seg = reinterpret_cast<const float4 *>(sm + g * RS + l * SEG);
#pragma unroll
        for (int j = 0; j < J; ++j) {
            const float4 v = seg[j];
            x[4 * j] = v.x;
            x[4 * j + 1] = v.y;
            x[4 * j + 2] = v.z;
            x[4 * j + 3] = v.w;
        }
    }
    __syncwarp();  // the input rows are in registers: the area becomes the staging area of the output

    // ---- bracket: the group's minimum and maximum (fminf / fmaxf skip NaN; NaN-only groups fail below)
    float fl = x[0], fh = x[0];
#pragma unroll
    for (int i = 1; i < E; ++i) {
        fl = fminf(fl, x[i]);
        fh = fmaxf(fh, x[i]);
    }
#pragma unroll
    for (int o = L / 2; o > 0; o >>= 1) {
        fl = fminf(fl, __shfl_xor_sync(full, fl, o));
        fh = fmaxf(fh, __shfl_xor_sync(full, fh, o));
    }

    // ---- regula falsi (Illinois weights) on g(t) = count(!(x < t)) - k, one bracket per row group.
    // state (uniform within a group): 0 searching, 1 t separates exactly k elements, 2 robust search needed
    int state = 0, c = 0, c_lo = 0, run = 0;  // c / c_lo: the lane's count at t (all / first half of its columns)
    float t = fl;
    float gl = (float)(D - k), gh = fminf((float)(1 - k), -0.5f);
    if (k == D) {  // everything is kept: nothing is below -inf, NaN included
        t = __uint_as_float(0xff800000u);
        state = 1;
        c = E;
        c_lo = E / 2;
    }
#pragma unroll 1
    for (int it = 0; it < TOPK_FAST_MAX_PROBES; ++it) {
        if (__all_sync(full, state != 0)) break;
        if (state == 0) {
            float fm = fmaf(fh - fl, gl * rcp_approx(gl - gh), fl);
            if (!(fm > fl && fm < fh)) fm = 0.5f * fl + 0.5f * fh;
            if (!(fm > fl && fm < fh)) state = 2;  // adjacent floats, infinite or NaN end points
            t = fm;
        }
        // four independent chains of predicated adds; ca + cb counts the first half of the lane's columns
        int ca = 0, cb = 0, cc = 0, cd = 0;
#pragma unroll
        for (int i = 0; i < E / 2; i += 2) {
            count_nlt(ca, x[i], t);
            count_nlt(cb, x[i + 1], t);
            count_nlt(cc, x[E / 2 + i], t);
            count_nlt(cd, x[E / 2 + i + 1], t);
        }
        c_lo = ca + cb;
        c = c_lo + cc + cd;
        const int gcount = group_sum<L>(c) - k;
        if (state == 0) {
            if (gcount == 0) {
                state = 1;
            } else if (gcount > 0) {
                fl = t;
                gl = (float)gcount;
                if (run > 0) gh *= 0.5f;
                run = run > 0 ? run + 1 : 1;
            } else {
                fh = t;
                gh = (float)gcount;
                if (run < 0) gl *= 0.5f;
                run = run < 0 ? run - 1 : -1;
            }
            if (run >= TOPK_FAST_MAX_RUN || run <= -TOPK_FAST_MAX_RUN) state = 2;
        }
    }
    if (state == 0) state = 2;
    const unsigned okmask = __ballot_sync(full, state == 1);  // bit g * L: row group g finished on the fast path

    // ---- emission of the finished groups into the staging area (pairs in output order)
    if (out_idx != nullptr) {
        int inc = c;
#pragma unroll
        for (int o = 1; o < L; o <<= 1) {
            const int v = __shfl_up_sync(full, inc, o, L);
            if (l >= o) inc += v;
        }
        if (state == 1) {
            // two independent store chains: the lane's first and second half of columns
            unsigned w = (unsigned)__cvta_generic_to_shared(sm + (g * k + inc - c));
            unsigned w2 = w + 4u * (unsigned)c_lo;
            const unsigned col0 = (unsigned)(l * E);
#pragma unroll
            for (int i = 0; i < E / 2; ++i) {
                stage_if_nlt<4 * IDXW>(w, x[i], t, col0 + i);
                stage_if_nlt<4 * IDXW>(w2, x[E / 2 + i], t, col0 + E / 2 + i);
            }
        }
    }

    // ---- the rows left to the robust search, one at a time in the one-row layout
    unsigned todo = ~okmask;
#pragma unroll 1
    for (int gg = 0; gg < R; ++gg) {
        if (!((todo >> (gg * L)) & 1u) || row0 + gg >= n) continue;  // warp-uniform
        const int64_t row = row0 + gg;
        const int c0 = 4 * lane, c1 = 128 + 4 * lane;
        const bool ok0 = c0 < D, ok1 = c1 < D;
        const float *xr = x_ + row * D;
        const float ninf = __uint_as_float(0xff800000u);
        float4 a = make_float4(ninf, ninf, ninf, ninf), b = a;
        if (ok0) a = ld_stream_f4(reinterpret_cast<const float4 *>(xr + c0));
        if (ok1) b = ld_stream_f4(reinterpret_cast<const float4 *>(xr + c1));
        const unsigned raw[8] = {__float_as_uint(a.x), __float_as_uint(a.y), __float_as_uint(a.z), __float_as_uint(a.w),
                                 __float_as_uint(b.x), __float_as_uint(b.y), __float_as_uint(b.z), __float_as_uint(b.w)};
        const unsigned sel = topk_select_slow_f32(raw[0], raw[1], raw[2], raw[3], raw[4], raw[5], raw[6], raw[7],
                                                  ok0, ok1, D, k, lane);
        if (out_idx != nullptr) topk_stage_row(sel, raw, sm + gg * k, sm + IDXW + gg * k, lane, c0, c1);
        if (dense_out != nullptr) {
            float *o = dense_out + row * D;
            float4 va, vb;
            va.x = (sel & 1u) ? a.x : 0.f;
            va.y = (sel & 2u) ? a.y : 0.f;
            va.z = (sel & 4u) ? a.z : 0.f;
            va.w = (sel & 8u) ? a.w : 0.f;
            vb.x = (sel & 16u) ? b.x : 0.f;
            vb.y = (sel & 32u) ? b.y : 0.f;
            vb.z = (sel & 64u) ? b.z : 0.f;
            vb.w = (sel & 128u) ? b.w : 0.f;
            if (ok0) st_stream_f4(reinterpret_cast<float4 *>(o + c0), va);
            if (ok1) st_stream_f4(reinterpret_cast<float4 *>(o + c1), vb);
        }
    }

    // ---- staging area -> sp_data / sp_index (the warp's rows are contiguous in both)
    if (out_idx != nullptr) {
        __syncwarp();
        const int total = rows_here * k;
        unsigned *ov = reinterpret_cast<unsigned *>(out_val) + row0 * k;
        unsigned char *oi = out_idx + row0 * k;
        if (vec_out) {  // 16-byte aligned sp_data, 4-byte aligned sp_index, k % 4 == 0
            const uint4 *sv4 = reinterpret_cast<const uint4 *>(sm), *si4 = reinterpret_cast<const uint4 *>(sm + IDXW);
#pragma unroll 1
            for (int e = lane; 4 * e < total; e += 32) {
                const uint4 v = sv4[e], ci = si4[e];
                reinterpret_cast<uint4 *>(ov)[e] = v;
                reinterpret_cast<unsigned *>(oi)[e] = ci.x | (ci.y << 8) | (ci.z << 16) | (ci.w << 24);
            }
        } else {
#pragma unroll 1
            for (int e = lane; e < total; e += 32) {
                ov[e] = sm[e];
                oi[e] = (unsigned char)sm[IDXW + e];
            }
        }
    }

    // ---- dense x * mask of the finished groups (MaxK.forward): back through the padded rows
    if (dense_out != nullptr) {
        __syncwarp();
        if (state == 1) {
            float4 *seg = reinterpret_cast<float4 *>(sm + g * RS + l * SEG);
#pragma unroll
            for (int j = 0; j < J; ++j) {
                float4 v;
                v.x = !(x[4 * j] < t) ? x[4 * j] : 0.f;
                v.y = !(x[4 * j + 1] < t) ? x[4 * j + 1] : 0.f;
                v.z = !(x[4 * j + 2] < t) ? x[4 * j + 2] : 0.f;
                v.w = !(x[4 * j + 3] < t) ? x[4 * j + 3] : 0.f;
                seg[j] = v;
            }
        }
        __syncwarp();
#pragma unroll
        for (int it = 0; it < R * C4 / 32; ++it) {
            const int q = it * 32 + lane;
            const int gg = q / C4, c4 = q % C4;
            if (gg < rows_here && ((okmask >> (gg * L)) & 1u))
                st_stream_f4(reinterpret_cast<float4 *>(dense_out + (row0 + gg) * D + 4 * c4),
                             *reinterpret_cast<const float4 *>(sm + gg * RS + (c4 / J) * SEG + (c4 % J) * 4));
        }
    }
}

template <int D, int L>
int launch_topk_group(const float *x, int64_t n, int k, float *sp_data, uint8_t *sp_index, float *dense_out,
                      cudaStream_t stream)
{
    constexpr int R = 32 / L;
    constexpr size_t smem = (size_t)TOPK_GROUP_WPB * topk_group_smem_words_per_warp<D, L>() * sizeof(unsigned);
    const int64_t grid = ceil_div<int64_t>(n, (int64_t)TOPK_GROUP_WPB * R);
    MAXK_REQUIRE(grid <= 0x7fffffffLL, "maxk_topk_cbsr_f32: n too large");
    if (smem > 48 * 1024) {
        const int rc = opt_in_dynamic_smem(reinterpret_cast<const void *>(&topk_cbsr_f32_group_kernel<D, L>),
                                           "topk_cbsr_f32_group_kernel");
        if (rc != MAXK_OK) return rc;
    }
    const int vec_out = (k % 4 == 0) && (reinterpret_cast<uintptr_t>(sp_data) & 15) == 0 &&
                        (reinterpret_cast<uintptr_t>(sp_index) & 3) == 0;
    topk_cbsr_f32_group_kernel<D, L><<<(unsigned)grid, TOPK_GROUP_WPB * 32, smem, stream>>>(
        x, n, k, sp_data, sp_index, dense_out, vec_out);
    MAXK_LAUNCH_CHECK("topk_cbsr_f32_group_kernel");
    return MAXK_OK;
}

// grouped kernel where it is the faster one (d = 128, 64): -1 / 1 = yes, 0 = one row per warp everywhere
int g_topk_group = -1;

}  // namespace maxk

using namespace maxk;

extern "C" int maxk_set_topk_group(int on)
{
    MAXK_REQUIRE(on == -1 || on == 0 || on == 1, "maxk_set_topk_group: -1 (automatic), 0 (one row per warp) or 1");
    g_topk_group = on;
    return MAXK_OK;
}

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
    const int64_t grid = ceil_div<int64_t>(n, TOPK_F32_WPB);
    MAXK_REQUIRE(grid <= 0x7fffffffLL, "maxk_topk_cbsr_f32: n too large");
    {
        cudaStream_t st = static_cast<cudaStream_t>(stream);
        // round 2 measured 1, 2, 4 and 8 lanes per row (profiles/r02_topk_grouped.md): at d = 128 four lanes
        // (134 vs 246 us one row per warp), at d = 64 two lanes (172 vs 432 us) win; at d = 256 the one-row kernel
        // stays ahead (101 vs 108 us: the rows of a warp probe in lock step and the slowest of 8 sets the pace)
        if (2 * k <= d && g_topk_group != 0) {
            if (d == 128) return launch_topk_group<128, 4>(x, n, k, sp_data, sp_index, dense_out, st);
            if (d == 64) return launch_topk_group<64, 2>(x, n, k, sp_data, sp_index, dense_out, st);
        }
    }
    if (d == MAXK_MAX_DIM) {
        cudaStream_t st = static_cast<cudaStream_t>(stream);
        const TopkSeed seed = topk_seed(d, k);
        const int what = (sp_index != nullptr ? 1 : 0) | (dense_out != nullptr ? 2 : 0) |
                         ((sp_index != nullptr && k <= 32) ? 4 : 0);
        const int cfg = seed.mode == 0 ? 0 : (seed.low == 1 ? (seed.mode == 1 ? 1 : 2) : (seed.mode == 1 ? 3 : 4));
#define MAXK_TOPK_FULL(OUT, CFG)                                                                                 \
    else if (what == OUT && cfg == CFG)                                                                          \
        topk_cbsr_f32_kernel<true, OUT, CFG><<<(unsigned)grid, TOPK_F32_WPB * 32, 0, st>>>(x, n, d, k, seed,     \
                                                                                           sp_data, sp_index,    \
                                                                                           dense_out);
        if (false) {}
        MAXK_TOPK_FULL(5, 1) MAXK_TOPK_FULL(5, 2) MAXK_TOPK_FULL(7, 1) MAXK_TOPK_FULL(7, 2)        // k <= 32
        MAXK_TOPK_FULL(1, 3) MAXK_TOPK_FULL(1, 4) MAXK_TOPK_FULL(1, 0)                             // k > 32
        MAXK_TOPK_FULL(3, 3) MAXK_TOPK_FULL(3, 4) MAXK_TOPK_FULL(3, 0)
        MAXK_TOPK_FULL(2, 0) MAXK_TOPK_FULL(2, 1) MAXK_TOPK_FULL(2, 2) MAXK_TOPK_FULL(2, 3) MAXK_TOPK_FULL(2, 4)
        else {
            set_error("maxk_topk_cbsr_f32: no kernel for outputs %d / seed %d", what, cfg);
            return MAXK_ERR_UNSUPPORTED;
        }
#undef MAXK_TOPK_FULL
    } else
        topk_cbsr_f32_kernel<false><<<(unsigned)grid, TOPK_F32_WPB * 32, 0, static_cast<cudaStream_t>(stream)>>>(
            x, n, d, k, topk_seed(d, k), sp_data, sp_index, dense_out);
    MAXK_LAUNCH_CHECK("topk_cbsr_f32_kernel");
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
    topk_cbsr_u8_kernel<<<(unsigned)grid, TOPK_WPB * 32, 0, static_cast<cudaStream_t>(stream)>>>(
        x, n, d, k, value, index);
    MAXK_LAUNCH_CHECK("topk_cbsr_u8_kernel");
    return MAXK_OK;
}

extern "C" int maxk_layernorm_topk_cbsr(const float *x, int64_t n, int32_t d, const float *gamma, const float *beta,
                                        float eps, int32_t k, float *sp_data, uint8_t *sp_index, float *mean,
                                        float *rstd, float *y_out, void *stream)
{
    MAXK_REQUIRE(n >= 0, "maxk_layernorm_topk_cbsr: negative n");
    MAXK_REQUIRE(d > 0 && d <= MAXK_MAX_DIM && d % 4 == 0,
                 "maxk_layernorm_topk_cbsr: d=%d must be a multiple of 4 in [4,256]", d);
    MAXK_REQUIRE(k >= 1 && k <= d, "maxk_layernorm_topk_cbsr: Invalid k value (k=%d, d=%d)", k, d);
    if (n == 0) return MAXK_OK;
    MAXK_REQUIRE(x != nullptr && sp_data != nullptr && sp_index != nullptr && mean != nullptr && rstd != nullptr,
                 "maxk_layernorm_topk_cbsr: NULL pointer");
    MAXK_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y_out) |
                   reinterpret_cast<uintptr_t>(gamma) | reinterpret_cast<uintptr_t>(beta)) & 15) == 0,
                 "maxk_layernorm_topk_cbsr: x, y_out, gamma and beta must be 16-byte aligned");
    const int64_t grid = ceil_div<int64_t>(n, TOPK_F32_WPB);
    MAXK_REQUIRE(grid <= 0x7fffffffLL, "maxk_layernorm_topk_cbsr: n too large");
    if (d == MAXK_MAX_DIM && k <= 32 && k > 20)
        layernorm_topk_kernel<true, 1><<<(unsigned)grid, TOPK_F32_WPB * 32, 0, static_cast<cudaStream_t>(stream)>>>(
            x, n, d, gamma, beta, eps, k, topk_seed(d, k), sp_data, sp_index, mean, rstd, y_out);
    else if (d == MAXK_MAX_DIM && k <= 20)
        layernorm_topk_kernel<true, 2><<<(unsigned)grid, TOPK_F32_WPB * 32, 0, static_cast<cudaStream_t>(stream)>>>(
            x, n, d, gamma, beta, eps, k, topk_seed(d, k), sp_data, sp_index, mean, rstd, y_out);
    else
        layernorm_topk_kernel<false><<<(unsigned)grid, TOPK_F32_WPB * 32, 0, static_cast<cudaStream_t>(stream)>>>(
            x, n, d, gamma, beta, eps, k, topk_seed(d, k), sp_data, sp_index, mean, rstd, y_out);
    MAXK_LAUNCH_CHECK("layernorm_topk_kernel");
    return MAXK_OK;
}
