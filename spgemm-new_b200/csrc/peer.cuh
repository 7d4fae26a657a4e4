// peer.cuh -- device side of the multi-GPU path over NVLink peer memory (one process per GPU,
// buffers shared with CUDA IPC; the reference has no multi-GPU code at all).
//
// Protocol (all flags are monotonically increasing step numbers, "epochs"):
//   * every rank owns a gathered CBSR buffer [n_total, k] (values + selectors, global node
//     numbering) that its peers have mapped.  TopK writes the rank's own rows in place (stream
//     order makes them visible to the fused forward kernel that follows);
//   * the fused forward kernel's first `copy_ctas` CTAs PUSH that block into every peer's
//     buffer with 128-bit stores over NVLink -- posted writes stream at link speed from a few
//     CTAs, whereas reads are capped by the outstanding requests one SM may have (measured:
//     a pull delivered 1.7 GB/s per CTA whatever the unroll) -- and the last of them to finish
//     raises flag word [16 + rank] = epoch in every peer's flag array (system-scope release
//     after a system fence);
//   * per peer either the whole block travels (dense pair, TMA path below) or only the rows
//     that peer's edges reference (halo pair: a row list, 128-bit load/store copies) -- on
//     sparse graphs an all-gather of everything costs more than the aggregation itself;
//   * the other CTAs aggregate the tasks on this rank's own columns straight away and, before
//     the first task that reads other ranks' columns, wait until the flag words of all peers
//     in their OWN memory (local polling, nothing crosses NVLink) show the epoch;
//   * the backward publishes flag word [1] = epoch after the SSpMM; the reduce-scatter kernel
//     waits for it on every peer and pulls its block of each partial gradient;
//   * buffers are double-buffered by epoch parity on the host side, so no "done reading"
//     handshake is needed: a rank only PUSHES step e after every peer's block of step e-1 has
//     arrived in its own buffer (the copy CTAs wait for that, unconditionally -- also on a rank
//     whose own tasks never read a remote column), i.e. after every peer has entered the forward
//     of step e-1 and is therefore done with the buffers of step e-2, which step e overwrites.
// Every spin is bounded (maxk_set_peer_timeout_ms, default 10 minutes -- ordinary rank skew such as
// rank-0-only evaluation, checkpointing or a lazy module load must not kill the job): a peer that
// never arrives makes ONE lane report and the kernel trap, instead of hanging the GPU.
#pragma once

#include "common.cuh"

namespace maxk {

struct PeerArgs {
    int copy_ctas, world, rank, k;
    int use_tma;  // push through cp.async.bulk (needs PEER_TMA_SMEM + 128 bytes of dynamic smem)
    unsigned epoch;
    long long wait_from_task;
    long long row_start[MAXK_PEER_MAX];
    long long rows_of[MAXK_PEER_MAX];
    const int *send_rows[MAXK_PEER_MAX];  // NULL: dense pair (whole block); else rows peer p needs
    long long send_count[MAXK_PEER_MAX];
    float *g_data[MAXK_PEER_MAX];           // rank r's gathered value buffer ([rank] = own)
    unsigned char *g_index[MAXK_PEER_MAX];  // rank r's gathered selector buffer
    unsigned *flags[MAXK_PEER_MAX];         // rank r's flag words ([rank] = own)
    unsigned *local_sync;                   // own memory: [1] arrival counter of the copy CTAs
    long long spin_limit;                   // clock64 ticks a device-side wait may take (0 = no limit)
};
struct NoPeer {};
template <bool P>
struct PeerParam {
    using type = NoPeer;
};
template <>
struct PeerParam<true> {
    using type = PeerArgs;
};

__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned *p, unsigned v)
{
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_release_gpu(unsigned *p, unsigned v)
{
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// Peer data is immutable between its publication and the epoch after next, and this kernel
// has not touched these lines before (L1 starts clean at launch), so a weak streaming load is
// safe -- and unlike a volatile (system-scope) load it moves full 128-byte requests over NVLink.
__device__ __forceinline__ uint4 ld_peer_u4(const uint4 *p)
{
    uint4 v;
    asm volatile("ld.global.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(p)
                 : "memory");
    return v;
}

// host-side default of PeerArgs::spin_limit / RsArgs::spin_limit (peer.cu: maxk_set_peer_timeout_ms)
extern long long g_peer_spin_limit_cycles;
static __device__ unsigned g_peer_timeout_reported;  // first lane to time out reports, the others only trap

template <bool SYS>
__device__ __forceinline__ void peer_spin_until(const unsigned *flag, unsigned epoch, int who, long long limit)
{
    const long long t0 = clock64();
    while (true) {
        const unsigned v = SYS ? ld_acquire_sys(flag) : ld_acquire_gpu(flag);
        if ((int)(v - epoch) >= 0) return;
        __nanosleep(200);  // back off: thousands of waiting warps must not starve the copy CTAs
        if (limit > 0 && clock64() - t0 > limit) {
            if (atomicExch(&g_peer_timeout_reported, 1u) == 0u)
                printf("libmaxk_b200: timed out waiting for %s %d (flag %u, want %u); raise the limit with "
                       "maxk_set_peer_timeout_ms\n",
                       SYS ? "peer rank" : "the gather of block", who, v, epoch);
            __trap();
        }
    }
}

// this CTA's share (cta of nctas) of a `bytes`-long copy; 8 x 128 bits in flight per thread
__device__ __forceinline__ void peer_copy_slice(unsigned char *dst, const unsigned char *src,
                                                long long bytes, int cta, int nctas)
{
    constexpr int U = 8;  // 16 would spill under the aggregation kernel's 64-register cap
    const long long n16 = bytes >> 4;
    const long long lo = n16 * cta / nctas, hi = n16 * (cta + 1) / nctas;
    const uint4 *s = reinterpret_cast<const uint4 *>(src);
    uint4 *d = reinterpret_cast<uint4 *>(dst);
    const int nthr = blockDim.x, tid = threadIdx.x;
    for (long long base = lo; base < hi; base += (long long)nthr * U) {
        uint4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long i = base + (long long)u * nthr + tid;
            if (i < hi) v[u] = ld_peer_u4(s + i);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long i = base + (long long)u * nthr + tid;
            if (i < hi) d[i] = v[u];
        }
    }
    if (cta == 0) {  // tail bytes of a length that is not a multiple of 16
        for (long long i = (n16 << 4) + tid; i < bytes; i += nthr)
            dst[i] = *reinterpret_cast<const volatile unsigned char *>(src + i);
    }
}

// ---- the same push through the TMA unit -----------------------------------------------------
// The aggregating CTAs keep the SM's load/store pipe ~93 % busy, so copy warps that share an SM
// with them crawl (measured: 10 GB/s per copy CTA with 128-bit LDG/STG).  Bulk asynchronous
// copies do not go through that pipe: one thread stages 16 KB chunks of the block in shared
// memory (cp.async.bulk global -> shared, mbarrier completion) and fans each chunk out to
// every peer (cp.async.bulk shared -> global over NVLink, bulk-group completion); the block
// is read once, not world-1 times.
constexpr int PEER_TMA_CHUNK = 16384;
constexpr int PEER_TMA_STAGES = 2;
constexpr int PEER_TMA_SMEM = PEER_TMA_CHUNK * PEER_TMA_STAGES;

__device__ __forceinline__ unsigned smem_u32(const void *p)
{
    return (unsigned)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void tma_load_chunk(unsigned dst_smem, const void *src, unsigned bytes,
                                               unsigned mbar)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes)
                 : "memory");
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            dst_smem),
        "l"(src), "r"(bytes), "r"(mbar)
        : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned mbar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "PEER_MBAR_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra PEER_MBAR_DONE;\n"
        "bra PEER_MBAR_WAIT;\n"
        "PEER_MBAR_DONE:\n"
        "}" ::"r"(mbar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_store_chunk(void *dst, unsigned src_smem, unsigned bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst),
                 "r"(src_smem), "r"(bytes)
                 : "memory");
}

// thread 0 of a copy CTA: this CTA's share of `bytes` (a multiple of 16) from own memory to the
// same offset of every peer buffer.  `stage_smem`: PEER_TMA_SMEM bytes, 128-byte aligned;
// `mbar`: PEER_TMA_STAGES initialised mbarriers; `phase`: their parity bits, carried between calls.
__device__ __forceinline__ void peer_tma_fanout(const PeerArgs &a, unsigned char *const *dst_base,
                                                const unsigned char *src, long long byte_off,
                                                long long bytes, int cta, unsigned stage_smem,
                                                unsigned mbar0, unsigned &phase, int &chunks_issued)
{
    const long long n16 = bytes >> 4;
    const long long lo = (n16 * cta / a.copy_ctas) << 4, hi = (n16 * (cta + 1) / a.copy_ctas) << 4;
    const long long nchunks = (hi - lo + PEER_TMA_CHUNK - 1) / PEER_TMA_CHUNK;
    if (nchunks <= 0) return;
    auto chunk_bytes = [&](long long c) -> unsigned {
        const long long b = hi - (lo + c * PEER_TMA_CHUNK);
        return (unsigned)(b < PEER_TMA_CHUNK ? b : PEER_TMA_CHUNK);
    };
    auto load = [&](long long c) {
        const int st = chunks_issued % PEER_TMA_STAGES;
        // the stores that last read this stage (two chunks ago) must be done with it
        if (chunks_issued >= PEER_TMA_STAGES) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        tma_load_chunk(stage_smem + st * PEER_TMA_CHUNK, src + byte_off + lo + c * PEER_TMA_CHUNK,
                       chunk_bytes(c), mbar0 + 8 * st);
        ++chunks_issued;
    };
    int first = chunks_issued;
    load(0);
    for (long long c = 0; c < nchunks; ++c) {
        if (c + 1 < nchunks) load(c + 1);
        const int st = (first + (int)c) % PEER_TMA_STAGES;
        mbar_wait(mbar0 + 8 * st, (phase >> st) & 1u);
        phase ^= (1u << st);
        const unsigned nb = chunk_bytes(c);
        const long long off = byte_off + lo + c * PEER_TMA_CHUNK;
        for (int s = 1; s < a.world; ++s) {
            const int p = (a.rank + s) % a.world;
            if (a.send_rows[p] == nullptr)  // dense pair
                tma_store_chunk(dst_base[p] + off, stage_smem + st * PEER_TMA_CHUNK, nb);
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
}

constexpr int PEER_FLAG_BWD = 1;       // flag word: partial gradients published
constexpr int PEER_FLAG_ARRIVED = 16;  // flag words [16 + r]: CBSR block of rank r has arrived

// halo pair: copy the listed rows (global ids) of the own block into the same rows of peer p's
// buffers; a row is k*4 bytes of values + k bytes of selectors, moved in 16-byte units when
// k % 16 == 0 (4-byte units otherwise).  All copy CTAs share every list.
// IB: bytes per entry of the second buffer -- 1 = uint8 selectors, 2 = the uint16 bank plan of the
// planned kernels (cbsr_plan.cu), which then travels in place of the selectors.
template <int IB = 1>
__device__ __forceinline__ void peer_push_rows(const PeerArgs &a, int p, int cta)
{
    const int *rows = a.send_rows[p];
    const long long n = a.send_count[p];
    const long long tid = (long long)cta * blockDim.x + threadIdx.x;
    const long long nthr = (long long)a.copy_ctas * blockDim.x;
    const int k = a.k;
    if ((k & 15) == 0) {
        const int uv = k >> 2, upr = uv + ((k * IB) >> 4);  // 16-byte units: values, then selectors
        for (long long i = tid; i < n * upr; i += nthr) {
            const long long r = __ldg(rows + i / upr);
            const int u = (int)(i % upr);
            if (u < uv) {
                const uint4 v = ld_peer_u4(reinterpret_cast<const uint4 *>(a.g_data[a.rank] + r * k) + u);
                reinterpret_cast<uint4 *>(a.g_data[p] + r * k)[u] = v;
            } else {
                const uint4 v = ld_peer_u4(reinterpret_cast<const uint4 *>(a.g_index[a.rank] + r * k * IB) + (u - uv));
                reinterpret_cast<uint4 *>(a.g_index[p] + r * k * IB)[u - uv] = v;
            }
        }
    } else {
        const int uv = k, upr = k + ((k * IB) >> 2);  // 4-byte units (k % 4 == 0)
        for (long long i = tid; i < n * upr; i += nthr) {
            const long long r = __ldg(rows + i / upr);
            const int u = (int)(i % upr);
            if (u < uv)
                reinterpret_cast<unsigned *>(a.g_data[p] + r * k)[u] =
                    reinterpret_cast<const unsigned *>(a.g_data[a.rank] + r * k)[u];
            else
                reinterpret_cast<unsigned *>(a.g_index[p] + r * k * IB)[u - uv] =
                    reinterpret_cast<const unsigned *>(a.g_index[a.rank] + r * k * IB)[u - uv];
        }
    }
}

// copy CTAs of the fused forward kernel: push this rank's CBSR rows to every other rank
template <int IB = 1>
__device__ __forceinline__ void peer_push_block(const PeerArgs &a, int cta, unsigned char *smem)
{
    const long long row0 = a.row_start[a.rank];
    const long long nel = a.rows_of[a.rank] * a.k;
    const long long nib = nel * IB;  // bytes of the second buffer (selectors or plan)
    // The buffers of this step's parity were last used by step epoch-2.  A peer is done reading them
    // once it has entered the forward of step epoch-1, i.e. once its block of epoch-1 has arrived here:
    // wait for that before overwriting anything over there -- in EVERY rank, also one whose own tasks
    // never wait for a remote column (the aggregating warps' wait alone would let such a rank run ahead).
    // No deadlock: the peers' pushes of epoch-1 wait for arrivals of epoch-2 only, and kernels of one
    // stream run in launch order.
    if (a.epoch >= 2u && (int)threadIdx.x < a.world && (int)threadIdx.x != a.rank)
        peer_spin_until<true>(a.flags[a.rank] + PEER_FLAG_ARRIVED + threadIdx.x, a.epoch - 1u, (int)threadIdx.x,
                              a.spin_limit);
    __syncthreads();
    bool any_dense = false;
    for (int r = 0; r < a.world; ++r) any_dense |= (r != a.rank && a.send_rows[r] == nullptr);
    if (any_dense && a.use_tma) {
        __shared__ __align__(8) unsigned long long mbar[PEER_TMA_STAGES];
        if (threadIdx.x == 0) {
            const unsigned mbar0 = smem_u32(mbar);
            for (int i = 0; i < PEER_TMA_STAGES; ++i)
                asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar0 + 8 * i) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("fence.proxy.async;" ::: "memory");
            unsigned phase = 0u;
            int issued = 0;
            unsigned char *dd[MAXK_PEER_MAX], *di[MAXK_PEER_MAX];
            for (int r = 0; r < a.world; ++r) {
                dd[r] = reinterpret_cast<unsigned char *>(a.g_data[r]);
                di[r] = a.g_index[r];
            }
            const unsigned stage = (smem_u32(smem) + 127u) & ~127u;
            peer_tma_fanout(a, dd, dd[a.rank], row0 * a.k * 4, (nel * 4) & ~15LL, cta, stage, mbar0, phase, issued);
            peer_tma_fanout(a, di, di[a.rank], row0 * a.k * IB, nib & ~15LL, cta, stage, mbar0, phase, issued);
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // all remote writes performed
            asm volatile("fence.proxy.async;" ::: "memory");
        }
        if (cta == 0 && (nib & 15)) {  // selector tail of a block whose byte count is not 16-aligned
            const long long t0 = nib & ~15LL;
            for (int s = 1; s < a.world; ++s) {
                const int p = (a.rank + s) % a.world;
                if (a.send_rows[p] != nullptr) continue;
                for (long long i = t0 + threadIdx.x; i < nib; i += blockDim.x)
                    a.g_index[p][row0 * a.k * IB + i] = a.g_index[a.rank][row0 * a.k * IB + i];
            }
        }
    } else if (any_dense) {
        const unsigned char *own_d = reinterpret_cast<const unsigned char *>(a.g_data[a.rank] + row0 * a.k);
        const unsigned char *own_i = a.g_index[a.rank] + row0 * a.k * IB;
        for (int s = 1; s < a.world; ++s) {
            const int p = (a.rank + s) % a.world;  // staggered: rank r writes to r+1 first
            if (a.send_rows[p] != nullptr) continue;
            peer_copy_slice(reinterpret_cast<unsigned char *>(a.g_data[p] + row0 * a.k), own_d, nel * 4,
                            cta, a.copy_ctas);
            peer_copy_slice(a.g_index[p] + row0 * a.k * IB, own_i, nib, cta, a.copy_ctas);
        }
    }
    for (int s = 1; s < a.world; ++s) {  // halo pairs
        const int p = (a.rank + s) % a.world;
        if (a.send_rows[p] != nullptr && a.send_count[p] > 0) peer_push_rows<IB>(a, p, cta);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();  // this CTA's remote stores before the arrival count
        unsigned arrived;
        asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;"
                     : "=r"(arrived)
                     : "l"(a.local_sync + 1)
                     : "memory");
        if (arrived == (unsigned)a.copy_ctas - 1u) {  // last copy CTA: the whole block is out
            a.local_sync[1] = 0u;
            __threadfence_system();
            for (int s = 1; s < a.world; ++s) {
                const int p = (a.rank + s) % a.world;
                st_release_sys(a.flags[p] + PEER_FLAG_ARRIVED + a.rank, a.epoch);
            }
        }
    }
}

// aggregating warps: block until every peer's block has arrived (lane r watches rank r's word
// in this rank's own flag array)
__device__ __forceinline__ void peer_wait_gathered(const PeerArgs &a, int lane)
{
    if (lane < a.world && lane != a.rank)
        peer_spin_until<true>(a.flags[a.rank] + PEER_FLAG_ARRIVED + lane, a.epoch, lane, a.spin_limit);
    __syncwarp();
}

}  // namespace maxk
