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
// column H homes in different banks (H copies of the output row, common.cuh: plan_home_word),
//       home 0:  word  c                                            bank  c mod 32
//       home h:  word  h*256 + (c mod 8)*32 + (c/8 + S_h*(c mod 8) + T_h) mod 32
// and decide once per CBSR row -- not once per edge -- which home each of its entries uses so
// that the 32 entries of one update instruction spread over the banks.  The plan is a maximum
// bipartite matching entries -> banks (augmenting paths, one thread per row); entries it leaves
// over go to their least loaded bank.  Fullest bank of 32 random columns (simulation, 500 rows):
//       one home 3.27   H = 2: 2.0   H = 4: 1.46 (54 % of rows conflict-free)   H = 8: 1.01
// i.e. of ~10.9 wavefronts per edge on the Reddit shape ~2.9 / ~4 / ~4.9 go away; the row is
// summed over its homes when it is flushed (H x the flush work and shared memory, so H is a
// trade: 8 for long rows, 2 for short ones).  A row is planned once per TopK and used once per
// incident edge (492 times on the Reddit shape, 25 on the products shape).
//
// The plan is a uint16 array sp_addr[N,k] holding the accumulator WORD of every entry (< H*256): it
// replaces sp_index as the forward kernel's second gather (64 instead of 32 bytes per row at
// k = 32 -- still one wavefront, 6k instead of 5k bytes per node in the L2).  Entries are planned
// in groups of 32 consecutive entries, which is how the planned forward kernel maps entries to
// lanes (entry 32*i + lane in its i-th update instruction).  For k = 16 and k = 8 one instruction
// carries 32/k edges, each with its own accumulator set on its own 1/E of the banks (word*E + slot),
// so a row's k entries are planned over k banks (b mod k).
#include "common.cuh"

namespace maxk {

constexpr int PLAN_GROUP = 32;

// one thread per (row, group of entries that share an update instruction)
__global__ void __launch_bounds__(128)
cbsr_bank_plan_kernel(const unsigned char *__restrict__ sp_index, int64_t n, int k, int groups,
                      int group_len, unsigned bank_mask, int homes, unsigned short *__restrict__ sp_addr)
{
    const int64_t item = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (item >= n * groups) return;
    const int64_t row = item / groups;
    const int g = (int)(item - row * groups);
    const int j0 = g * group_len;
    const int cnt = min(group_len, k - j0);
    const unsigned char *src = sp_index + row * k + j0;

    unsigned char c[PLAN_GROUP];       // columns of the entries
    unsigned char pos[PLAN_GROUP][PLAN_MAX_HOMES];  // entry, home -> word of the home modulo 32 (its bank)
    signed char owner[PLAN_GROUP];     // bank -> entry that holds it (-1: free)
    signed char home[PLAN_GROUP];      // entry -> home it uses (-1: not placed yet)
    unsigned char done[PLAN_GROUP];    // entry -> placed for good
    for (int i = 0; i < PLAN_GROUP; ++i) {
        const unsigned ci = (i < cnt) ? __ldg(src + i) : 0u;
        c[i] = (unsigned char)ci;
#pragma unroll
        for (int h = 0; h < PLAN_MAX_HOMES; ++h) pos[i][h] = (unsigned char)(plan_home_word(ci, h) & 31u);
        owner[i] = -1;
        home[i] = -1;
        done[i] = 0;
    }
    auto bank = [&](int e, int h) -> unsigned { return pos[e][h] & bank_mask; };

    // Layer 0: a maximum matching entries -> banks (Kuhn's augmenting paths, depth-first, explicit
    // stack): every matched entry has a bank to itself.  Entries it leaves over are matched among
    // themselves onto a second slot of every bank (layer 1), and so on: the fullest bank holds as
    // many entries as layers were needed (1 for most rows with 8 homes, 2 with 2 homes).
    int placed = 0;
    for (int layer = 0; layer < PLAN_GROUP && placed < cnt; ++layer) {
        for (int b = 0; b < PLAN_GROUP; ++b) owner[b] = -1;
        for (int i = 0; i < cnt; ++i) {
            if (done[i]) continue;  // placed by an earlier layer
            unsigned visited = 0u;
            signed char st_entry[PLAN_GROUP], st_next[PLAN_GROUP], st_bank[PLAN_GROUP], st_home[PLAN_GROUP];
            int sp = 0;
            st_entry[0] = (signed char)i;
            st_next[0] = 0;
            bool found = false;
            while (sp >= 0) {
                const int e = st_entry[sp], h = st_next[sp];
                if (h == homes) {  // this entry cannot be re-homed: back up
                    --sp;
                    continue;
                }
                st_next[sp] = (signed char)(h + 1);
                const unsigned b = bank(e, h);
                if ((visited >> b) & 1u) continue;
                visited |= 1u << b;
                st_bank[sp] = (signed char)b;
                st_home[sp] = (signed char)h;
                if (owner[b] < 0) {
                    found = true;
                    break;
                }
                ++sp;  // ask the holder of bank b (an entry of this layer) to move
                st_entry[sp] = owner[b];
                st_next[sp] = 0;
            }
            if (found) {
                for (int l = sp; l >= 0; --l) {  // every entry on the path takes the bank it asked for
                    owner[st_bank[l]] = st_entry[l];
                    home[st_entry[l]] = st_home[l];
                }
            }
        }
        for (int b = 0; b < PLAN_GROUP; ++b) {  // what this layer matched is final
            if (owner[b] >= 0 && !done[owner[b]]) {
                done[owner[b]] = 1;
                ++placed;
            }
        }
    }
    unsigned short *dst = sp_addr + row * k + j0;
    for (int i = 0; i < cnt; ++i) {
        const unsigned ci = c[i], h = (unsigned)home[i];
        dst[i] = (unsigned short)(h == 0u ? ci : h * 256u + (ci & 7u) * 32u + pos[i][h]);
    }
}

}  // namespace maxk

using namespace maxk;

extern "C" int maxk_cbsr_bank_plan(const uint8_t *sp_index, int64_t n, int32_t k, int32_t homes,
                                   uint16_t *sp_addr, void *stream)
{
    MAXK_REQUIRE(n >= 0, "maxk_cbsr_bank_plan: negative n");
    MAXK_REQUIRE(k >= 1 && k <= MAXK_MAX_DIM, "maxk_cbsr_bank_plan: k=%d out of range", k);
    MAXK_REQUIRE(homes == 2 || homes == 4 || homes == 8, "maxk_cbsr_bank_plan: homes=%d must be 2, 4 or 8", homes);
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
        sp_index, n, k, groups, group_len, bank_mask, homes, sp_addr);
    MAXK_LAUNCH_CHECK("cbsr_bank_plan_kernel");
    return MAXK_OK;
}
