// topk_select.cuh -- device functions shared by the TopK kernels (topk_cbsr.cu) and the TopK epilogue of the
// fused Linear (linear_topk.cu): order-preserving keys, the robust key-space selection with its exact tie rule,
// the staged emission of one row held by a warp, and the compare-and-count primitives of the value-space search.
#pragma once
#include "common.cuh"

namespace maxk {

constexpr int TOPK_WPB = 8;      // warps per CTA of the uint8 kernel
constexpr int TOPK_F32_WPB = 2;  // ... of the fp32 one-row-per-warp kernels (topk_cbsr.cu: why two)
constexpr int TOPK_FAST_MAX_PROBES = 14;  // then the key-space search takes over
constexpr int TOPK_FAST_MAX_RUN = 5;      // same end point moved this often in a row -> ties likely
constexpr int TOPK_EXTRACT_MAX = 8;       // a probe within this many elements of k is finished by extraction

// Where the fp32 search of a one-row-per-warp kernel starts; made once per launch on the host (k and d are launch
// constants, the kernel is issue-bound and every instruction of its set-up is paid per row).
//   mode 0: between the row minimum and maximum, gl = d - k (exact), w = gl / (gl - gh);
//   mode 1..3 (d == 256, k <= 64): from ORDER STATISTICS of the lanes' maxima.  With 8 elements per lane, the minimum
//   over the 32 lane maxima (A), over the 16 maxima of lane pairs (B) and over the 8 maxima of lane quads (C) have
//   expected ranks 100.6, 48.3 and 20.7 of 256 for ANY continuous distribution (measured on 20 000 rows; standard
//   deviations 22 / 14 / 8), the row maximum has rank 1: the first threshold is interpolated BY RANK between the two
//   statistics that enclose k -- (C, max) for k <= 20, (B, C) for k <= 48, (A, B) above.  The lower end of the
//   bracket must have at least k elements not below it: A has >= 32 (one per lane), and for 32 < k <= 64 the minimum G
//   over the 64 maxima of the lanes' four-element groups has >= 64 (expected rank 175).  gl = rank - k is an estimate
//   that only steers the interpolation; gh = min(1 - k, -0.5) as before.
struct TopkSeed {
    float w, gl, gh;
    int mode;  // 0: minimum / maximum; 1: (B, C); 2: (C, max); 3: (A, B)
    int low;   // lower end of the bracket: 0 = row minimum, 1 = A, 2 = G
};
__host__ __device__ inline TopkSeed topk_seed(int d, int k)
{
    TopkSeed s;
    s.gh = (float)(1 - k) < -0.5f ? (float)(1 - k) : -0.5f;
    if (d == MAXK_MAX_DIM && k <= 64) {
        const float ra = 100.6f, rb = 48.3f, rc = 20.7f, rg = 175.f;
        if (k <= 20) {
            s.mode = 2;
            s.w = (rc - (float)k) / (rc - 1.f);
        } else if (k <= 48) {
            s.mode = 1;
            s.w = (rb - (float)k) / (rb - rc);
        } else {
            s.mode = 3;
            s.w = (ra - (float)k) / (ra - rb);
        }
        s.low = k <= 32 ? 1 : 2;
        s.gl = (k <= 32 ? ra : rg) - (float)k;
    } else {
        s.mode = 0;
        s.low = 0;
        s.gl = (float)(d - k);
        s.w = s.gl / (s.gl - s.gh);
    }
    return s;
}

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

// The same order without the canonical forms (NaN, -0): two instructions each way.  Enough for the END
// POINTS of the fast path's bracket, which only have to lie at or outside the row's extreme values: -0 / +0
// differ by one key, and a NaN end point (a lane whose eight values are all NaN) makes the first probe fail
// its range test, which sends the row to the robust search.
__device__ __forceinline__ unsigned f32_rough_key(float f)
{
    const unsigned b = __float_as_uint(f);
    return b ^ ((unsigned)((int)b >> 31) | 0x80000000u);
}
__device__ __forceinline__ float rough_key_to_f32(unsigned key)
{
    return __uint_as_float(key ^ (~(unsigned)((int)key >> 31) | 0x80000000u));
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

// lane mask (bit i) of elements NOT below t: NaN elements always qualify (they rank largest).
// Sums of distinct bits compile to FSETP + SEL per element and one IADD3 per two elements;
// the count is the POPC of the mask and the final probe's mask IS the selection.
__device__ __forceinline__ unsigned lane_mask_nlt(const float (&x)[8], float t)
{
    unsigned m = 0u;
#pragma unroll
    for (int i = 0; i < 8; ++i) m += !(x[i] < t) ? (1u << i) : 0u;
    return m;
}

__device__ __forceinline__ float rcp_approx(float v)
{
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
}

// warp-wide minimum / maximum of fp32 values in ONE instruction (sm_100a: CREDUX.MIN/MAX.F32, result in a uniform
// register): NaN operands are skipped, an all-NaN warp yields NaN.  No order-preserving keys, no conversions.
__device__ __forceinline__ float warp_max_f32(float v)
{
    float r;
    asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v));
    return r;
}
__device__ __forceinline__ float warp_min_f32(float v)
{
    float r;
    asm volatile("redux.sync.min.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v));
    return r;
}

// one step of an inclusive warp scan: the shuffle's own "source lane in range" predicate guards the add
// (SHFL.UP + predicated IADD; the C form costs SHFL + ISETP + SEL + IADD)
__device__ __forceinline__ void scan_step_up(unsigned &v, int o)
{
    asm volatile("{.reg .pred p; .reg .b32 r; shfl.sync.up.b32 r|p, %0, %1, 0, 0xffffffff; @p add.u32 %0, %0, r;}"
                 : "+r"(v)
                 : "r"(o));
}

// The float just above v in the order of `<` (v finite; -0 counts as +0).
__device__ __forceinline__ float f32_next_up(float v)
{
    const int b = __float_as_int(v);
    return __int_as_float(v > 0.f ? b + 1 : (v < 0.f ? b - 1 : 1));
}

// ---- finishing a search by extraction ----------------------------------------------------------------------------
// A probe at t selected the elements with !(x[i] < t) and left g = count - k, g != 0 and |g| small.  Instead of probing
// on, name the threshold that WOULD separate exactly k elements if the |g| elements to move sat in different lanes:
//   g < 0 (too few):  the |g|-th largest of the lanes' largest values below t  (every lane offers its best unselected)
//   g > 0 (too many): just above the g-th smallest of the lanes' smallest selected values
// |g| rounds of one CREDUX each; the caller probes the result like any other threshold, so a lane that held two of
// the wanted elements, or a tie, only costs another step -- never a wrong answer.
__device__ __forceinline__ float topk_extract_threshold(const float (&x)[8], float t, int g)
{
    const float pinf = __uint_as_float(0x7f800000u), ninf = __uint_as_float(0xff800000u);
    float m;
    if (g < 0) {
        float u = (x[0] < t) ? x[0] : ninf;
#pragma unroll
        for (int i = 1; i < 8; ++i) u = fmaxf(u, (x[i] < t) ? x[i] : ninf);
#pragma unroll 1
        while (true) {
            m = warp_max_f32(u);
            if (++g == 0) break;
            u = (u == m) ? ninf : u;
        }
        return m;
    }
    float u = (x[0] < t) ? pinf : x[0];  // (a NaN here is skipped by the CREDUX like by fminf)
#pragma unroll
    for (int i = 1; i < 8; ++i) u = fminf(u, (x[i] < t) ? pinf : x[i]);
#pragma unroll 1
    while (true) {
        m = warp_min_f32(u);
        if (--g == 0) break;
        u = (u == m) ? pinf : u;
    }
    return f32_next_up(m);
}

// Element i of a lane sits at column  (i < 4 ? 4*lane + i : 128 + 4*lane + (i-4)):
// group 0 of every lane precedes group 1 of every lane in column order.

// ---- robust selection on order-preserving keys (u8 rows; fp32 rows the fast path left) ----
// key 0 marks "no element" (columns >= d).  Returns the lane's selection mask (bit i).
// Invariant of the search: count(key >= L) = cL >= k  and  count(key >= H) = cH < k,  L < H.
// Probes interpolate on the counts -- in VALUE space for fp32 -- and fall back to key-space
// bisection whenever the same end point moved twice in a row; the search ends on a probe
// that separates exactly k elements or with H == L + 1, i.e. L is the k-th largest key and
// the ties at L are taken in ascending column order.
template <bool U8>
__device__ __forceinline__ unsigned topk_select_by_keys(const unsigned (&key)[8], int d, int k, int lane)
{
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
    } else if (H - L == 1u) {
        thr = L;
    } else {
        // ties at the row minimum (e.g. post-ReLU zeros filling the tail of the top-k):
        // one probe settles it instead of a walk down the whole key range
        const int c1 = warp_count_ge(key, L + 1u);
        if (c1 < k) {
            thr = L;
        } else if (c1 == k) {
            thr = L + 1u;
            exact = true;
        } else {
            L = L + 1u;
            cL = c1;
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
                // Snap the probe to an ELEMENT before it becomes an end point: the bracket then shrinks by whole value
                // levels, and a probe that fell into the gap next to the k-th largest value ends the search at once.
                // Without this, ties at the k-th value (values on a grid, zeros after a ReLU) cost one bisection
                // step per key bit between two levels -- ~23 probes per row (round 2: 8-level rows 520-715 us per
                // 233 k rows against 85 us for tie-free rows).
                if (c > k) {
                    unsigned lm = 0xffffffffu;  // smallest element not below M: same count
#pragma unroll
                    for (int i = 0; i < 8; ++i) lm = min(lm, key[i] >= M ? key[i] : 0xffffffffu);
                    const unsigned m = __reduce_min_sync(0xffffffffu, lm);
                    if (m == 0xffffffffu) {  // more than k NaNs
                        thr = m;
                        break;
                    }
                    const int c1 = warp_count_ge(key, m + 1u);
                    if (c1 < k) {  // m is the k-th largest value, tied
                        thr = m;
                        break;
                    }
                    if (c1 == k) {
                        thr = m + 1u;
                        exact = true;
                        break;
                    }
                    L = m + 1u;
                    cL = c1;
                    stuck = stuck > 0 ? stuck + 1 : 1;
                } else {
                    unsigned lm = 0u;  // largest element below M (there is one: cL >= k > c)
#pragma unroll
                    for (int i = 0; i < 8; ++i) lm = max(lm, key[i] < M ? key[i] : 0u);
                    const unsigned m = __reduce_max_sync(0xffffffffu, lm);
                    const int cm = warp_count_ge(key, m);
                    if (cm >= k) {  // the elements above m are the c < k not below M: m is the k-th largest value
                        thr = m;
                        exact = (cm == k);
                        break;
                    }
                    H = m;
                    cH = cm;
                    stuck = stuck < 0 ? stuck - 1 : -1;
                }
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
    return sel;
}

// Out-of-line copy for the fp32 kernel: the fast path must not pay registers for it.  The
// bit patterns travel by value (registers), the keys are built here.
static __device__ __noinline__ unsigned topk_select_slow_f32(unsigned r0, unsigned r1, unsigned r2,
                                                      unsigned r3, unsigned r4, unsigned r5,
                                                      unsigned r6, unsigned r7, bool ok0, bool ok1,
                                                      int d, int k, int lane)
{
    const unsigned raw[8] = {r0, r1, r2, r3, r4, r5, r6, r7};
    unsigned key[8];
#pragma unroll
    for (int i = 0; i < 8; ++i)  // valid keys are >= 0x007fffff (-inf); 0 = "no element"
        key[i] = ((i < 4) ? ok0 : ok1) ? f32_order_key(raw[i]) : 0u;
    return topk_select_by_keys<false>(key, d, k, lane);
}

// ---- emission: ascending column order, staged in shared memory so that the global stores
// are coalesced (one 4k-byte value line + one k-byte selector line per row) -----------------
// stage: 2 * MAXK_MAX_DIM words of this warp (values, then column ids as words so that one
// address register serves both stores).
// K32: k <= 32 known at compile time (one store pair per lane, no test)
template <bool U8, bool K32 = false>
__device__ __forceinline__ void topk_emit(unsigned sel, const unsigned (&raw)[8], unsigned *stage,
                                          int lane, int k, int64_t row, int c0, int c1,
                                          void *__restrict__ out_val_,
                                          unsigned char *__restrict__ out_idx)
{
    const unsigned n0 = __popc(sel & 0xfu), n1 = __popc(sel >> 4);
    unsigned inc = n0 | (n1 << 16);
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) scan_step_up(inc, o);
    const unsigned tot0 = __shfl_sync(0xffffffffu, inc, 31) & 0xffffu;
    unsigned *w0 = stage + ((inc & 0xffffu) - n0);
    unsigned *w1 = stage + (tot0 + (inc >> 16) - n1);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (sel & (1u << i)) {
            w0[0] = raw[i];
            w0[MAXK_MAX_DIM] = (unsigned)(c0 + i);
            ++w0;
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (sel & (16u << i)) {
            w1[0] = raw[4 + i];
            w1[MAXK_MAX_DIM] = (unsigned)(c1 + i);
            ++w1;
        }
    }
    __syncwarp();
    unsigned char *oi = out_idx + row * k + lane;
    const unsigned *sv = stage + lane;
    if constexpr (!U8) {
        if (K32 || k <= 32) {  // the common widths: one store pair per lane, no loop
            if (lane < k) {
                static_cast<unsigned *>(out_val_)[row * k + lane] = sv[0];
                *oi = (unsigned char)sv[MAXK_MAX_DIM];
            }
            return;
        }
    }
    if constexpr (U8) {
        unsigned char *ov = static_cast<unsigned char *>(out_val_) + row * k + lane;
#pragma unroll 1
        for (int j = lane; j < k; j += 32, sv += 32, ov += 32, oi += 32) {
            *ov = (unsigned char)sv[0];
            *oi = (unsigned char)sv[MAXK_MAX_DIM];
        }
    } else {
        unsigned *ov = static_cast<unsigned *>(out_val_) + row * k + lane;
#pragma unroll 1
        for (int j = lane; j < k; j += 32, sv += 32, ov += 32, oi += 32) {
            *ov = sv[0];
            *oi = (unsigned char)sv[MAXK_MAX_DIM];
        }
    }
}


// c += !(x < t) as one compare and one predicated add (the C form compiles to three instructions per
// element in a single dependency chain)
__device__ __forceinline__ void count_nlt(int &c, float x, float t)
{
    asm("{.reg .pred p; setp.geu.f32 p, %1, %2; @p add.s32 %0, %0, 1;}" : "+r"(c) : "f"(x), "f"(t));
}

}  // namespace maxk
