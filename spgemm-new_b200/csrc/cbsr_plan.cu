// cbsr_plan.cu -- bank plan of a CBSR operand for the forward SpGEMM's shared-memory scatter.
//
// EXPERIMENTAL (written when round 1's GPU budget was spent: compiled and checked under the SIMT
// interpreter of tests/simt, NOT yet run on a GPU; reachable only through the two entry points
// below, which nothing calls by default).
//
// Why.  The forward kernel (spgemm_forward.cu) is bound by L1/shared-memory wavefronts on
// operands that live in the L2 (ncu: 93 % of peak on the Reddit shape): per edge one value-line
// gather, one selector-line gather, one staging broadcast -- and a read-modify-write of the
// output row in shared memory whose 32 lanes hit 32 random columns.  32 random columns over 32
// banks collide: the fullest bank holds 3.3 of them on average, so the load and the store of that
// update cost ~3.3 wavefronts each, more than everything else together.  No fixed column->bank
// mapping helps (any mapping sees the same random sets).  What helps is a CHOICE: give every
// column two homes in different banks,
//       copy 0:  word  c                          bank  c mod 32
//       copy 1:  word  256 + (c mod 8)*32 + c/8   bank  c / 8
// and decide once per CBSR row -- not once per edge -- which home each of its entries uses so
// that the 32 entries of one instruction spread over the banks.  Greedy two-choice placement plus
// two repair passes brings the fullest bank to 2 in 98 % of random rows (mean 2.02 vs 3.27), i.e.
// ~2.5 of ~10 wavefronts per edge go away; the row is summed over its two copies when it is
// flushed.  A row is planned once per TopK and used once per incident edge (492 times on the Reddit
// shape, 25 on the products shape).
//
// The plan is a uint16 array sp_addr[N,k] holding the accumulator WORD of every entry (0..511): it
// replaces sp_index as the forward kernel's second gather (64 instead of 32 bytes per row at
// k = 32 -- still one wavefront, 6k instead of 5k bytes per node in the L2).  Entries are planned
// in groups of 32 consecutive entries, which is how the planned forward kernel maps entries to
// lanes (entry 32*i + lane in its i-th update instruction).  For k = 16 and k = 8 one instruction
// carries 32/k edges, each with its own accumulator set on its own 1/E of the banks (word*E + slot),
// so a row's k entries are planned over k banks (b mod k).
#include "common.cuh"

namespace maxk {

constexpr int PLAN_GROUP = 32;

__device__ __forceinline__ unsigned plan_word(unsigned c, unsigned copy)
{
    return copy ? (256u + ((c & 7u) << 5) + (c >> 3)) : c;
}

struct BankLoad {  // 32 four-bit counters (an instruction has at most 32 entries)
    unsigned long long lo = 0ull, hi = 0ull;
    __device__ __forceinline__ unsigned get(unsigned b) const
    {
        return (unsigned)(((b < 16u) ? lo : hi) >> (4u * (b & 15u))) & 15u;
    }
    __device__ __forceinline__ void add(unsigned b, int delta)
    {
        const unsigned long long one = 1ull << (4u * (b & 15u));
        if (b < 16u)
            lo = delta > 0 ? lo + one : lo - one;
        else
            hi = delta > 0 ? hi + one : hi - one;
    }
    __device__ __forceinline__ unsigned fullest() const
    {
        unsigned m = 0u;
#pragma unroll
        for (int b = 0; b < 16; ++b) {
            const unsigned a = (unsigned)(lo >> (4 * b)) & 15u, c = (unsigned)(hi >> (4 * b)) & 15u;
            m = a > m ? a : m;
            m = c > m ? c : m;
        }
        return m;
    }
};

// one thread per (row, group of 32 entries)
__global__ void __launch_bounds__(128)
cbsr_bank_plan_kernel(const unsigned char *__restrict__ sp_index, int64_t n, int k, int groups,
                      int group_len, unsigned bank_mask, unsigned short *__restrict__ sp_addr)
{
    const int64_t item = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (item >= n * groups) return;
    const int64_t row = item / groups;
    const int g = (int)(item - row * groups);
    const int j0 = g * group_len;
    const int cnt = min(group_len, k - j0);
    const unsigned char *src = sp_index + row * k + j0;

    unsigned char c[PLAN_GROUP];
#pragma unroll
    for (int i = 0; i < PLAN_GROUP; ++i) c[i] = (i < cnt) ? __ldg(src + i) : (unsigned char)0;

    BankLoad load;
    unsigned copy = 0u;  // bit i: entry i lives in copy 1
#pragma unroll
    for (int i = 0; i < PLAN_GROUP; ++i) {
        if (i < cnt) {
            const unsigned b0 = c[i] & bank_mask, b1 = (c[i] >> 3) & bank_mask;
            const bool second = load.get(b1) < load.get(b0);
            copy |= second ? (1u << i) : 0u;
            load.add(second ? b1 : b0, +1);
        }
    }
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {  // repair: move entries out of the fullest banks
        const unsigned m = load.fullest();
        if (m <= 1u) break;
        bool moved = false;
#pragma unroll
        for (int i = 0; i < PLAN_GROUP; ++i) {
            if (i < cnt) {
                const unsigned b0 = c[i] & bank_mask, b1 = (c[i] >> 3) & bank_mask;
                const bool second = (copy >> i) & 1u;
                const unsigned cur = second ? b1 : b0, alt = second ? b0 : b1;
                if (load.get(cur) == m && load.get(alt) + 1u < m) {
                    load.add(cur, -1);
                    load.add(alt, +1);
                    copy ^= 1u << i;
                    moved = true;
                }
            }
        }
        if (!moved) break;
    }
    unsigned short *dst = sp_addr + row * k + j0;
#pragma unroll
    for (int i = 0; i < PLAN_GROUP; ++i)
        if (i < cnt) dst[i] = (unsigned short)plan_word(c[i], (copy >> i) & 1u);
}

}  // namespace maxk

using namespace maxk;

extern "C" int maxk_cbsr_bank_plan(const uint8_t *sp_index, int64_t n, int32_t k, uint16_t *sp_addr,
                                   void *stream)
{
    MAXK_REQUIRE(n >= 0, "maxk_cbsr_bank_plan: negative n");
    MAXK_REQUIRE(k >= 1 && k <= MAXK_MAX_DIM, "maxk_cbsr_bank_plan: k=%d out of range", k);
    if (n == 0) return MAXK_OK;
    MAXK_REQUIRE(sp_index != nullptr && sp_addr != nullptr, "maxk_cbsr_bank_plan: NULL pointer");
    MAXK_REQUIRE((reinterpret_cast<uintptr_t>(sp_addr) & 1) == 0, "maxk_cbsr_bank_plan: sp_addr must be 2-byte aligned");
    // entries of one update instruction and the banks they share: k >= 32 -> 32 entries over 32
    // banks; k = 16 / 8 -> the row's k entries over k banks (the other edges of the instruction
    // use the other banks)
    const int group_len = (k == 16 || k == 8) ? k : PLAN_GROUP;
    const unsigned bank_mask = (unsigned)group_len - 1u;
    const int groups = ceil_div<int>(k, group_len);
    const int64_t grid = ceil_div<int64_t>(n * groups, 128);
    MAXK_REQUIRE(grid <= 0x7fffffffLL, "maxk_cbsr_bank_plan: n too large");
    cbsr_bank_plan_kernel<<<(unsigned)grid, 128, 0, static_cast<cudaStream_t>(stream)>>>(
        sp_index, n, k, groups, group_len, bank_mask, sp_addr);
    MAXK_LAUNCH_CHECK("cbsr_bank_plan_kernel");
    return MAXK_OK;
}
