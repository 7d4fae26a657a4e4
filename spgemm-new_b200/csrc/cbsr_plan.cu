// cbsr_plan.cu -- bank plan of a CBSR operand for the forward SpGEMM's shared-memory scatter.
//
// Why.  On operands that live in the L2 the forward kernel (spgemm_forward.cu) is bound by L1 /
// shared-memory wavefronts (ncu, Reddit shape: 98 % of peak): per edge one value-line gather, one
// selector-line gather, one staging broadcast -- and a read-modify-write of the output row in shared
// memory whose 32 lanes hit 32 random columns.  32 random columns over 32 banks collide: the fullest
// bank holds 3.3 of them on average, so the load and the store of that update cost ~3.3 wavefronts
// each, more than everything else together.  No fixed column -> bank mapping helps (any mapping sees
// the same random sets).  What helps is a CHOICE: every column gets two homes in different banks
// (common.cuh: plan_home_word) and a plan decides ONCE PER CBSR ROW -- not once per edge -- which home
// each of its entries uses, so that the 32 entries of one update instruction spread over the banks.
// A row is planned once per TopK and used once per incident edge (492 times on the Reddit shape).
//
// The plan is a uint16 array sp_addr[N,k] holding the accumulator WORD of every entry (< 512): it
// replaces sp_index as the forward kernel's second gather (64 instead of 32 bytes per row at k = 32:
// still one wavefront).  Entries are planned in groups of 32 consecutive entries, which is how the
// planned forward kernel maps entries to lanes (entry 32*i + lane in its i-th update instruction).
//
// Round 1 made the plan with an exact maximum matching, one THREAD per group: 1.4 ms for the Reddit
// shape's 233 k rows -- as much as the plan saves.  This kernel gives a group to a WARP (lane = entry)
// and places the entries greedily in a fixed number of warp-wide rounds, conflicts being settled
// through a 32-entry owner table in shared memory (whoever's store lands last owns the bank):
//   1. every entry asks for its home-0 bank;              2. the losers ask for their home-1 bank if free;
//   3. one augmenting step per home: a loser u whose bank b is held by o takes b if o can move to
//      its own other home (o's move and u's claim are settled through the same table);
//   4. what is left (about 5 of 32 entries) shares a bank with one placed entry, spread by the same
//      two rounds over a second layer (and a third ... until everything is placed).
// Fullest bank on random 32-of-256 column sets: 3.27 without a plan, 2.00 with the exact two-home
// matching, ~2.1 with this one (tests/test_simt_kernels.py measures it on the kernel's own output).
#include "common.cuh"

namespace maxk {

constexpr int PLAN_WPB = 8;  // warps (groups) per CTA

__global__ void __launch_bounds__(PLAN_WPB * 32)
cbsr_bank_plan_kernel(const unsigned char *__restrict__ sp_index, int64_t n_groups, int groups_per_row, int k,
                      unsigned short *__restrict__ sp_addr)
{
    __shared__ unsigned char s_owner[PLAN_WPB][32];  // bank -> lane that holds it (layer 1)
    __shared__ unsigned char s_req[PLAN_WPB][32];    // bank -> lane that claims it in the current round
    __shared__ unsigned char s_move[PLAN_WPB][32];   // lane -> asked to move to its other home
    __shared__ unsigned short s_banks[PLAN_WPB][32]; // lane -> b0 | b1 << 8
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t item = (int64_t)blockIdx.x * PLAN_WPB + warp;
    if (item >= n_groups) return;  // warp-uniform; no CTA barrier below
    const int64_t row = item / groups_per_row;
    const int j = (int)(item - row * groups_per_row) * 32 + lane;  // entry of the row
    const bool valid = j < k;
    const unsigned c = valid ? (unsigned)__ldg(sp_index + row * k + j) : 0u;
    const unsigned b0 = c & 31u, b1 = c >> 3;  // banks of the two homes
    unsigned char *owner = s_owner[warp], *req = s_req[warp], *move = s_move[warp];
    int home = -1;  // -1: not placed

    s_banks[warp][lane] = (unsigned short)(b0 | (b1 << 8));
    move[lane] = 0;
    // 1. home 0
    if (valid) owner[b0] = (unsigned char)lane;
    __syncwarp();
    if (valid && owner[b0] == lane) home = 0;
    unsigned taken = __reduce_or_sync(full, home == 0 ? (1u << b0) : 0u);
    // 2. losers: home 1, if that bank is free
    const bool ask1 = valid && home < 0 && !((taken >> b1) & 1u);
    __syncwarp();
    if (ask1) owner[b1] = (unsigned char)lane;
    __syncwarp();
    if (ask1 && owner[b1] == lane) home = 1;
    taken |= __reduce_or_sync(full, (ask1 && home == 1) ? (1u << b1) : 0u);
    // 3. one augmenting step through the holder of the bank of home hm (hm = 0, then 1)
#pragma unroll
    for (int hm = 0; hm < 2; ++hm) {
        const unsigned mine = hm ? b1 : b0;
        bool ask = valid && home < 0;
        unsigned alt = 0u;
        int o = 0;
        if (ask) {
            o = owner[mine];  // placed in step 1 or 2 (the bank is taken, or this lane would hold it)
            const unsigned ob = s_banks[warp][o];
            const unsigned o0 = ob & 0xffu, o1 = ob >> 8;
            alt = (o0 == mine) ? o1 : o0;  // the holder's other home (o0 == o1: nowhere to go)
            ask = alt != mine && !((taken >> alt) & 1u);
        }
        __syncwarp();
        if (ask) req[alt] = (unsigned char)lane;
        __syncwarp();
        const bool won = ask && req[alt] == lane;
        if (won) {
            move[o] = 1;
            home = hm;
        }
        __syncwarp();
        if (won) {
            owner[alt] = (unsigned char)o;
            owner[mine] = (unsigned char)lane;
        }
        if (move[lane]) {  // this lane was asked to leave its bank for its other home
            home ^= 1;
            move[lane] = 0;
        }
        taken |= __reduce_or_sync(full, won ? (1u << alt) : 0u);
        __syncwarp();
    }
    // 4. further layers for what is left (one in almost every row): the same two rounds over a table
    //    of second, third ... slots; every pass places at least one entry per bank asked for
    while (__ballot_sync(full, valid && home < 0)) {
        const bool left = valid && home < 0;
        if (left) req[b0] = (unsigned char)lane;
        __syncwarp();
        const bool l0 = left && req[b0] == lane;
        if (l0) home = 0;
        const unsigned taken2 = __reduce_or_sync(full, l0 ? (1u << b0) : 0u);
        const bool ask2 = left && home < 0 && !((taken2 >> b1) & 1u);
        __syncwarp();
        if (ask2) req[b1] = (unsigned char)lane;
        __syncwarp();
        if (ask2 && req[b1] == lane) home = 1;
        __syncwarp();
    }
    if (valid) sp_addr[row * k + j] = (unsigned short)plan_home_word(c, home);
}

}  // namespace maxk

using namespace maxk;

extern "C" int maxk_cbsr_bank_plan(const uint8_t *sp_index, int64_t n, int32_t k, uint16_t *sp_addr, void *stream)
{
    MAXK_REQUIRE(n >= 0, "maxk_cbsr_bank_plan: negative n");
    MAXK_REQUIRE(k >= 1 && k <= MAXK_MAX_DIM, "maxk_cbsr_bank_plan: k=%d out of range", k);
    if (n == 0) return MAXK_OK;
    MAXK_REQUIRE(sp_index != nullptr && sp_addr != nullptr, "maxk_cbsr_bank_plan: NULL pointer");
    MAXK_REQUIRE((reinterpret_cast<uintptr_t>(sp_addr) & 1) == 0, "maxk_cbsr_bank_plan: sp_addr must be 2-byte aligned");
    const int groups = ceil_div<int>(k, 32);
    const int64_t n_groups = n * groups;
    const int64_t grid = ceil_div<int64_t>(n_groups, PLAN_WPB);
    MAXK_REQUIRE(grid <= 0x7fffffffLL, "maxk_cbsr_bank_plan: n too large");
    cbsr_bank_plan_kernel<<<(unsigned)grid, PLAN_WPB * 32, 0, static_cast<cudaStream_t>(stream)>>>(
        sp_index, n_groups, groups, k, sp_addr);
    MAXK_LAUNCH_CHECK("cbsr_bank_plan_kernel");
    return MAXK_OK;
}
