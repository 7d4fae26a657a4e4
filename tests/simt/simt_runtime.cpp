// tests/simt/simt_runtime.cpp -- TEST INFRASTRUCTURE (see include/cuda_runtime.h).
//
// SIMT interpreter: every CUDA thread of a CTA is a fibre with its own stack; CTAs run one
// after another on the calling host thread.  All interpreter state is per host thread, so several
// host threads can each run a kernel at the same time: that is how the tests run the ranks of the
// NVLink peer-memory path, whose kernels wait for each other through flag words (one host thread
// per rank, buffers shared; a kernel that only spins yields the CPU).  A fibre runs until it reaches a rendezvous
// (__syncthreads, __syncwarp, a warp collective, __nanosleep) or returns; the scheduler then
// runs the next one.  This is the weakest schedule CUDA allows inside a warp -- lanes are
// NOT in lock step between rendezvous points -- so code that relies on implicit warp
// synchrony shows up as a wrong result here instead of passing by accident.
//
// Checks made on the way (each aborts the launch with a message and makes cudaGetLastError
// return cudaErrorLaunchFailure):
//   * every lane named in the mask of a collective must arrive at the SAME collective with the
//     SAME mask (lanes that have already returned from the kernel are skipped, as the hardware
//     does for exited threads);
//   * a lane must be in the mask it passes;
//   * deadlock: no fibre can make progress (a barrier some threads never reach);
//   * an access outside mapped memory (the tests put guard pages around every buffer): SIGSEGV is
//     caught while a kernel runs and reported with the faulting address and thread;
//   * launch configuration: block size <= 1024, dynamic shared memory <= 48 KB unless the kernel
//     opted in with cudaFuncSetAttribute, <= 227 KB in any case.
#include <cuda_runtime.h>
#include <sched.h>
#include <signal.h>
#include <sys/mman.h>
#include <time.h>

#include <map>
#include <mutex>
#include <unordered_map>
#include <string>
#include <vector>

namespace simt {

thread_local Thread *self_ = nullptr;

namespace {

constexpr size_t kStack = 256 * 1024;  // per fibre; pages are touched lazily
constexpr int kMaxThreads = 1024;

struct Warp {
    unsigned arrived = 0, released = 0, exited = 0;
    unsigned mask_of[32];
    int op_of[32];
    uint64_t slot[32];
    uint64_t result[32];
};

enum Wait { RUN, WAIT_WARP, WAIT_BLOCK, WAIT_SPIN, DONE };

struct AsyncCopy {
    void *dst;
    const void *src;
    int bytes;
    long group;
};

struct BulkCopy {  // cp.async.bulk issued, not yet performed
    void *dst;
    const void *src;
    unsigned bytes;
    uint64_t *mbar;  // loads: the barrier that counts the bytes; stores: nullptr
};

struct Fibre {
    Thread t;
    std::vector<BulkCopy> bulk;
    std::vector<AsyncCopy> copies;  // cp.async issued, not yet performed
    long cp_group = 0;              // the group new copies join (open until commit_group)
    void *sp = nullptr;
    Wait wait = RUN;
    unsigned wait_mask = 0;
    unsigned long long wait_gen = 0;
    Warp *w = nullptr;
};

struct Block {
    int nthreads = 0, alive = 0, arrived = 0;
    unsigned long long gen = 0;
    std::vector<Fibre> f;
    std::vector<Warp> warps;
};

thread_local Block g_blk;
thread_local Fibre *g_cur = nullptr;
thread_local void *g_sched_sp = nullptr;
thread_local char *g_stacks = nullptr;  // fibre stacks of this host thread (released when it exits)
thread_local const std::function<void()> *g_body = nullptr;
thread_local unsigned long long g_progress = 0;
thread_local bool g_failed = false;
thread_local std::string g_fail_msg;
thread_local cudaError_t g_last_error = cudaSuccess;
std::map<const void *, int> g_max_dyn_smem;
std::mutex g_max_dyn_smem_lock;
constexpr size_t kDynSmem = 227 * 1024 + 128;
thread_local unsigned char *g_dyn_smem = nullptr;
struct ThreadResources {  // host threads come and go (one per rank in the peer tests)
    ~ThreadResources()
    {
        if (g_stacks) munmap(g_stacks, kStack * kMaxThreads);
        free(g_dyn_smem);
        g_stacks = nullptr;
        g_dyn_smem = nullptr;
    }
};
thread_local ThreadResources g_thread_resources;

extern "C" void simt_ctx_switch(void **save_sp, void *load_sp);
#if defined(__x86_64__)
asm(R"(
.text
.globl simt_ctx_switch
.type simt_ctx_switch,@function
simt_ctx_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
.size simt_ctx_switch,.-simt_ctx_switch
)");
#else
#error "the SIMT test interpreter has a context switch for x86-64 only"
#endif

void to_scheduler() { simt_ctx_switch(&g_cur->sp, g_sched_sp); }

void drain_copies(Fibre *f, long below_group)
{
    size_t keep = 0;
    for (size_t i = 0; i < f->copies.size(); ++i) {
        const AsyncCopy &c = f->copies[i];
        if (c.group < below_group)
            memcpy(c.dst, c.src, (size_t)c.bytes);
        else
            f->copies[keep++] = c;
    }
    f->copies.resize(keep);
}

void drain_bulk_stores(Fibre *f)
{
    size_t keep = 0;
    for (size_t i = 0; i < f->bulk.size(); ++i) {
        const BulkCopy &c = f->bulk[i];
        if (c.mbar == nullptr)
            memcpy(c.dst, c.src, c.bytes);
        else
            f->bulk[keep++] = c;
    }
    f->bulk.resize(keep);
}

[[noreturn]] void fibre_exit()
{
    Fibre *f = g_cur;
    drain_copies(f, f->cp_group + 1);  // outstanding copies still land after the thread returns
    drain_bulk_stores(f);
    f->wait = DONE;
    f->w->exited |= 1u << f->t.lane;
    --g_blk.alive;
    ++g_progress;
    to_scheduler();
    abort();  // never resumed
}

void fibre_main()
{
    (*g_body)();
    fibre_exit();
}

[[noreturn]] void fail(const std::string &msg)
{
    if (!g_failed) {
        g_failed = true;
        g_fail_msg = msg;
        fprintf(stderr, "simt: %s\n", msg.c_str());
    }
    // the launch is abandoned: park this fibre for good
    g_cur->wait = DONE;
    ++g_progress;
    to_scheduler();
    abort();
}

std::string where()
{
    const Thread &t = g_cur->t;
    char b[128];
    snprintf(b, sizeof b, "block %u thread %u (warp %d lane %d)", t.bid.x, t.tid.x, t.warp, t.lane);
    return b;
}

void prepare(Fibre &f)
{
    const size_t i = &f - g_blk.f.data();
    char *top = g_stacks + (i + 1) * kStack;
    uintptr_t a = (reinterpret_cast<uintptr_t>(top) & ~uintptr_t(15)) - 64;  // return-address slot, 16-aligned
    void **slot = reinterpret_cast<void **>(a);
    slot[1] = nullptr;                                   // a fake return address for fibre_main
    slot[0] = reinterpret_cast<void *>(&fibre_main);     // popped by `ret`
    for (int r = 1; r <= 6; ++r) slot[-r] = nullptr;     // rbp rbx r12 r13 r14 r15
    f.sp = slot - 6;
    f.wait = RUN;
}

// can this fibre make progress if resumed?
bool ready(const Fibre &f)
{
    switch (f.wait) {
        case RUN:
        case WAIT_SPIN: return true;
        case WAIT_WARP: {
            const unsigned bit = 1u << f.t.lane;
            if (f.w->released & bit) return true;
            const unsigned need = f.wait_mask & ~f.w->exited;
            return (f.w->arrived & need) == need;
        }
        case WAIT_BLOCK: return g_blk.gen != f.wait_gen || g_blk.arrived >= g_blk.alive;
        case DONE: return false;
    }
    return false;
}

const char *op_name(int op)
{
    static const char *n[] = {"shfl", "shfl_up", "shfl_down", "shfl_xor", "reduce_add", "reduce_min", "reduce_max",
                              "reduce_umin", "reduce_umax", "reduce_and", "reduce_or", "ballot", "syncwarp",
                              "collective-end"};
    return (op >= 0 && op < 14) ? n[op] : "?";
}
constexpr int OP_SYNCWARP = 12, OP_END = 13;

// rendezvous of the lanes in `mask` (minus lanes that have returned)
void warp_barrier(unsigned mask, int op)
{
    Fibre *f = g_cur;
    Warp &w = *f->w;
    const int lane = f->t.lane;
    const unsigned bit = 1u << lane;
    if (!(mask & bit)) fail(where() + ": " + op_name(op) + " with mask " + std::to_string(mask) + " that does not name the calling lane");
    w.mask_of[lane] = mask;
    w.op_of[lane] = op;
    w.arrived |= bit;
    for (;;) {
        if (w.released & bit) {
            w.released &= ~bit;
            return;
        }
        const unsigned need = mask & ~w.exited;
        if ((w.arrived & need) == need) {
            for (int l = 0; l < 32; ++l) {
                if (!(need & (1u << l))) continue;
                if (w.mask_of[l] != mask || w.op_of[l] != op) {
                    char b[256];
                    snprintf(b, sizeof b, ": divergent collective -- lane %d is in %s(mask %#x) while lane %d is in %s(mask %#x)",
                             lane, op_name(op), mask, l, op_name(w.op_of[l]), w.mask_of[l]);
                    fail(where() + b);
                }
            }
            w.arrived &= ~need;
            w.released |= need & ~bit;
            ++g_progress;
            return;
        }
        f->wait = WAIT_WARP;
        f->wait_mask = mask;
        to_scheduler();
        f->wait = RUN;
    }
}

// A fibre touched unmapped memory (the tests surround every buffer with guard pages).  The handler
// runs on the faulting fibre's stack; SA_NODEFER keeps the signal unblocked, so leaving through
// the scheduler instead of returning is harmless.
struct sigaction g_old_segv, g_old_bus;

void on_segv(int sig, siginfo_t *si, void *)
{
    if (g_cur == nullptr) {  // not inside a kernel of this thread: not ours
        sigaction(sig, sig == SIGBUS ? &g_old_bus : &g_old_segv, nullptr);
        raise(sig);
        return;
    }
    char b[96];
    snprintf(b, sizeof b, ": invalid memory access at address %p", si->si_addr);
    fail(where() + b);
}

bool run_block(const LaunchCfg &cfg, unsigned bx, unsigned by, unsigned bz)
{
    Block &b = g_blk;
    const int n = (int)(cfg.block.x * cfg.block.y * cfg.block.z);
    b.nthreads = b.alive = n;
    b.arrived = 0;
    b.gen = 0;
    b.f.assign(n, Fibre());
    b.warps.assign((n + 31) / 32, Warp());
    for (int i = 0; i < n; ++i) {
        Fibre &f = b.f[i];
        f.t.tid = uint3{(unsigned)i % cfg.block.x, ((unsigned)i / cfg.block.x) % cfg.block.y,
                        (unsigned)i / (cfg.block.x * cfg.block.y)};
        f.t.bid = uint3{bx, by, bz};
        f.t.bdim = cfg.block;
        f.t.gdim = cfg.grid;
        f.t.lane = i & 31;
        f.t.warp = i >> 5;
        f.w = &b.warps[i >> 5];
        prepare(f);
    }
    // lanes beyond the block size in the last warp do not exist
    if (n & 31) b.warps.back().exited = ~0u << (n & 31);

    while (b.alive > 0 && !g_failed) {
        const unsigned long long before = g_progress;
        bool ran = false, spun = false;
        for (int i = 0; i < n && !g_failed; ++i) {
            Fibre &f = b.f[i];
            if (!ready(f)) continue;
            if (f.wait == WAIT_SPIN) spun = true; else ran = true;
            g_cur = &f;
            self_ = &f.t;
            simt_ctx_switch(&g_sched_sp, f.sp);
        }
        if (g_failed) break;
        if (!ran && g_progress == before) {
            if (spun) {  // only spin-waits are left: another host thread (rank) has to move first;
                sched_yield();  // the waits are bounded and trap by themselves
                continue;
            }
            int stuck = -1;
            for (int i = 0; i < n; ++i)
                if (b.f[i].wait != DONE) { stuck = i; break; }
            char m[256];
            snprintf(m, sizeof m, "deadlock in block %u: thread %d waits at %s and nobody can release it (%d of %d threads still alive)",
                     bx, stuck, stuck >= 0 && b.f[stuck].wait == WAIT_BLOCK ? "__syncthreads" : "a warp rendezvous",
                     b.alive, n);
            g_failed = true;
            g_fail_msg = m;
            fprintf(stderr, "simt: %s\n", m);
        }
    }
    g_cur = nullptr;
    self_ = nullptr;
    return !g_failed;
}

}  // namespace

void launch(const void *kernel, const LaunchCfg &cfg, const std::function<void()> &body)
{
    const unsigned long long n = (unsigned long long)cfg.block.x * cfg.block.y * cfg.block.z;
    const unsigned long long nblocks = (unsigned long long)cfg.grid.x * cfg.grid.y * cfg.grid.z;
    if (n == 0 || n > kMaxThreads || nblocks == 0 || cfg.grid.x > 0x7fffffffu || cfg.grid.y > 65535 || cfg.grid.z > 65535) {
        g_last_error = cudaErrorInvalidConfiguration;
        return;
    }
    size_t limit = 48 * 1024;
    {
        std::lock_guard<std::mutex> hold(g_max_dyn_smem_lock);
        auto it = g_max_dyn_smem.find(kernel);
        if (it != g_max_dyn_smem.end()) limit = std::max<size_t>(limit, (size_t)it->second);
    }
    if (cfg.smem > limit) {
        g_last_error = cudaErrorInvalidValue;
        return;
    }
    if (g_cur != nullptr) {
        fprintf(stderr, "simt: nested launch\n");
        abort();
    }
    if (!g_stacks) {
        (void)&g_thread_resources;  // instantiate the releaser for this thread
        g_stacks = static_cast<char *>(mmap(nullptr, kStack * kMaxThreads, PROT_READ | PROT_WRITE,
                                            MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0));
        if (g_stacks == MAP_FAILED) abort();
    }
    g_body = &body;
    g_failed = false;
    static std::once_flag handlers;  // process-wide and for good: kernels of several threads overlap
    std::call_once(handlers, [] {
        struct sigaction sa;
        memset(&sa, 0, sizeof sa);
        sa.sa_sigaction = on_segv;
        sa.sa_flags = SA_SIGINFO | SA_NODEFER;
        sigaction(SIGSEGV, &sa, &g_old_segv);
        sigaction(SIGBUS, &sa, &g_old_bus);
    });
    // poison dynamic shared memory: kernels must not rely on zeroed shared memory
    if (!g_dyn_smem) g_dyn_smem = static_cast<unsigned char *>(aligned_alloc(128, kDynSmem));
    memset(g_dyn_smem, 0xcb, kDynSmem);
    for (unsigned z = 0; z < cfg.grid.z && !g_failed; ++z)
        for (unsigned y = 0; y < cfg.grid.y && !g_failed; ++y)
            for (unsigned x = 0; x < cfg.grid.x && !g_failed; ++x) run_block(cfg, x, y, z);
    g_body = nullptr;
    if (g_failed) g_last_error = cudaErrorLaunchFailure;
}

void *dynamic_smem()
{
    return reinterpret_cast<void *>((reinterpret_cast<uintptr_t>(g_dyn_smem) + 127) & ~uintptr_t(127));
}

void set_func_max_dynamic_smem(const void *kernel, int bytes)
{
    std::lock_guard<std::mutex> hold(g_max_dyn_smem_lock);
    g_max_dyn_smem[kernel] = bytes;
}

void syncthreads()
{
    Fibre *f = g_cur;
    Block &b = g_blk;
    const unsigned long long gen = b.gen;
    ++b.arrived;
    for (;;) {
        if (b.gen != gen) return;
        if (b.arrived >= b.alive) {
            b.arrived = 0;
            ++b.gen;
            ++g_progress;
            return;
        }
        f->wait = WAIT_BLOCK;
        f->wait_gen = gen;
        to_scheduler();
        f->wait = RUN;
    }
}

void syncwarp(unsigned mask) { warp_barrier(mask, OP_SYNCWARP); }

uint64_t collective(unsigned mask, Op op, uint64_t value, int arg, int width)
{
    Fibre *f = g_cur;
    Warp &w = *f->w;
    const int lane = f->t.lane;
    w.slot[lane] = value;
    warp_barrier(mask, (int)op);
    const unsigned live = mask & ~w.exited;
    uint64_t r = value;
    auto from = [&](int src) -> uint64_t {
        // reading a lane that is not part of the collective is undefined in CUDA: own value
        return (src >= 0 && src < 32 && (live & (1u << src))) ? w.slot[src] : value;
    };
    const int seg = lane & ~(width - 1);
    switch (op) {
        case OP_SHFL: r = from(seg + (arg & (width - 1))); break;
        case OP_SHFL_UP: r = (lane - arg >= seg) ? from(lane - arg) : value; break;
        case OP_SHFL_DOWN: r = (lane + arg < seg + width) ? from(lane + arg) : value; break;
        case OP_SHFL_XOR: r = ((lane ^ arg) < seg + width) ? from(lane ^ arg) : value; break;
        case OP_BALLOT: {
            uint32_t m = 0;
            for (int l = 0; l < 32; ++l)
                if ((live & (1u << l)) && w.slot[l]) m |= 1u << l;
            r = m;
            break;
        }
        default: {
            bool first = true;
            uint32_t acc = 0;
            for (int l = 0; l < 32; ++l) {
                if (!(live & (1u << l))) continue;
                const uint32_t v = (uint32_t)w.slot[l];
                if (first) { acc = v; first = false; continue; }
                switch (op) {
                    case OP_ADD: acc += v; break;
                    case OP_MIN: acc = (uint32_t)std::min((int32_t)acc, (int32_t)v); break;
                    case OP_MAX: acc = (uint32_t)std::max((int32_t)acc, (int32_t)v); break;
                    case OP_UMIN: acc = std::min(acc, v); break;
                    case OP_UMAX: acc = std::max(acc, v); break;
                    case OP_AND: acc &= v; break;
                    case OP_OR: acc |= v; break;
                    default: break;
                }
            }
            r = acc;
        }
    }
    warp_barrier(mask, OP_END);  // nobody overwrites a slot before everyone has read
    return r;
}

bool g_cp_async_eager = false;  // tests run both: data lands at once / as late as PTX allows

void cp_async(void *smem_dst, const void *gmem_src, int bytes)
{
    Fibre *f = g_cur;
    if ((reinterpret_cast<uintptr_t>(smem_dst) | reinterpret_cast<uintptr_t>(gmem_src)) & (uintptr_t)(bytes - 1))
        fail(where() + ": cp.async with a misaligned address");
    volatile char probe = *static_cast<const volatile char *>(gmem_src);  // fault now, in the issuing thread
    (void)probe;
    probe = static_cast<const volatile char *>(gmem_src)[bytes - 1];
    if (g_cp_async_eager)
        memcpy(smem_dst, gmem_src, (size_t)bytes);
    else
        f->copies.push_back(AsyncCopy{smem_dst, gmem_src, bytes, f->cp_group});
}

void cp_async_commit() { ++g_cur->cp_group; }

void cp_async_wait(int all_but_newest) { drain_copies(g_cur, g_cur->cp_group - all_but_newest); }

// ---- shared-memory address tokens ------------------------------------------------------------------
thread_local std::unordered_map<uint32_t, uintptr_t> g_smem_pages;  // low 20 bits of the page number -> page number

size_t smem_token(const void *p)
{
    const uintptr_t page = reinterpret_cast<uintptr_t>(p) >> 12;
    for (uintptr_t i = 0; i < 64; ++i) {  // the object and what follows it (dynamic shared memory: 227 KB)
        const uint32_t key = (uint32_t)((page + i) & 0xfffffu);
        auto it = g_smem_pages.find(key);
        if (it == g_smem_pages.end())
            g_smem_pages.emplace(key, page + i);
        else if (it->second != page + i) {
            fprintf(stderr, "simt: two shared-memory objects share the low 32 address bits; rerun\n");
            abort();
        }
    }
    return (size_t)(uint32_t)reinterpret_cast<uintptr_t>(p);
}

void *smem_resolve(size_t token)
{
    const uint32_t t = (uint32_t)token;
    auto it = g_smem_pages.find(t >> 12);
    if (it == g_smem_pages.end()) fail(where() + ": a 32-bit shared-memory address that no cvta produced");
    return reinterpret_cast<void *>((it->second << 12) | (t & 4095u));
}

// ---- mbarrier: bit 63 phase, bits 32..62 pending arrivals, bits 0..31 pending bytes (count 1 only) -----
void mbar_init(size_t mbar, int count)
{
    *static_cast<uint64_t *>(smem_resolve(mbar)) = (uint64_t)count << 32;
}

static void mbar_settle(uint64_t *m)
{
    if ((*m & 0x7fffffffffffffffull) == 0) *m = ((*m ^ (1ull << 63)) & (1ull << 63)) | (1ull << 32);
}

void mbar_expect_tx(size_t mbar, unsigned bytes)
{
    uint64_t *m = static_cast<uint64_t *>(smem_resolve(mbar));
    if (((*m >> 32) & 0x7fffffffu) == 0) fail(where() + ": mbarrier arrival without a pending arrival count");
    *m = *m - (1ull << 32) + bytes;  // one arrival, `bytes` more expected
    mbar_settle(m);
}

void bulk_load(size_t smem_dst, const void *gmem_src, unsigned bytes, size_t mbar)
{
    void *dst = smem_resolve(smem_dst);
    if ((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(gmem_src) | bytes) & 15u)
        fail(where() + ": cp.async.bulk needs 16-byte aligned addresses and sizes");
    volatile char probe = static_cast<const volatile char *>(gmem_src)[bytes - 1];  // fault in the issuing thread
    (void)probe;
    g_cur->bulk.push_back(BulkCopy{dst, gmem_src, bytes, static_cast<uint64_t *>(smem_resolve(mbar))});
}

void mbar_wait(size_t mbar, unsigned parity)
{
    Fibre *f = g_cur;
    uint64_t *m = static_cast<uint64_t *>(smem_resolve(mbar));
    size_t keep = 0;  // the loads counted by this barrier land now, the latest moment PTX allows
    for (size_t i = 0; i < f->bulk.size(); ++i) {
        const BulkCopy c = f->bulk[i];
        if (c.mbar == m) {
            memcpy(c.dst, c.src, c.bytes);
            if ((uint32_t)*m < c.bytes) fail(where() + ": more bytes arrived at an mbarrier than it expected");
            *m -= c.bytes;
            mbar_settle(m);
        } else {
            f->bulk[keep++] = c;
        }
    }
    f->bulk.resize(keep);
    if ((unsigned)(*m >> 63) == (parity & 1u))
        fail(where() + ": mbarrier wait that nothing in flight can complete (deadlock on the GPU)");
}

void bulk_store(void *gmem_dst, size_t smem_src, unsigned bytes)
{
    const void *src = smem_resolve(smem_src);
    if ((reinterpret_cast<uintptr_t>(gmem_dst) | reinterpret_cast<uintptr_t>(src) | bytes) & 15u)
        fail(where() + ": cp.async.bulk needs 16-byte aligned addresses and sizes");
    volatile char *q = static_cast<volatile char *>(gmem_dst);  // fault in the issuing thread
    q[bytes - 1] = q[bytes - 1];
    g_cur->bulk.push_back(BulkCopy{gmem_dst, src, bytes, nullptr});
}

void bulk_wait() { drain_bulk_stores(g_cur); }

void yield_spin()
{
    Fibre *f = g_cur;
    f->wait = WAIT_SPIN;
    to_scheduler();
    f->wait = RUN;
}

long long clock_ticks()  // a "cycle" every 4 ns of real time: the kernels' 10^10-cycle spin limits expire
{                        // after 40 s, generous enough for a loaded test machine
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ((long long)ts.tv_sec * 1000000000LL + ts.tv_nsec) / 4;
}

void trap(const char *why) { fail(where() + ": " + why); }

void unsupported_asm(const char *ptx)
{
    fail(where() + ": inline PTX without a SIMT stand-in was executed: " + ptx);
}

}  // namespace simt

// ---- runtime API --------------------------------------------------------------------------------
cudaError_t cudaMalloc(void **p, size_t bytes)
{
    void *q = nullptr;
    if (posix_memalign(&q, 256, bytes ? bytes : 1) != 0) return cudaErrorMemoryAllocation;
    memset(q, 0xcd, bytes);
    *p = q;
    return cudaSuccess;
}
cudaError_t cudaFree(void *p)
{
    free(p);
    return cudaSuccess;
}
cudaError_t cudaMemset(void *p, int v, size_t bytes)
{
    memset(p, v, bytes);
    return cudaSuccess;
}
cudaError_t cudaMemsetAsync(void *p, int v, size_t bytes, cudaStream_t) { return cudaMemset(p, v, bytes); }
cudaError_t cudaMemcpy(void *dst, const void *src, size_t bytes, cudaMemcpyKind)
{
    memmove(dst, src, bytes);
    return cudaSuccess;
}
cudaError_t cudaMemcpyAsync(void *dst, const void *src, size_t bytes, cudaMemcpyKind k, cudaStream_t)
{
    return cudaMemcpy(dst, src, bytes, k);
}
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaGetLastError(); }
cudaError_t cudaDeviceSynchronize() { return cudaGetLastError(); }
cudaError_t cudaGetLastError()
{
    const cudaError_t e = simt::g_last_error;
    simt::g_last_error = cudaSuccess;
    return e;
}
const char *cudaGetErrorName(cudaError_t e)
{
    switch (e) {
        case cudaSuccess: return "cudaSuccess";
        case cudaErrorInvalidValue: return "cudaErrorInvalidValue";
        case cudaErrorMemoryAllocation: return "cudaErrorMemoryAllocation";
        case cudaErrorInvalidConfiguration: return "cudaErrorInvalidConfiguration";
        case cudaErrorLaunchFailure: return "cudaErrorLaunchFailure";
        case cudaErrorNotSupported: return "cudaErrorNotSupported";
        default: return "cudaErrorUnknown";
    }
}
const char *cudaGetErrorString(cudaError_t e)
{
    if (e == cudaErrorLaunchFailure && !simt::g_fail_msg.empty()) return simt::g_fail_msg.c_str();
    return cudaGetErrorName(e);
}
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t *, void *) { return cudaErrorNotSupported; }
cudaError_t cudaIpcOpenMemHandle(void **, cudaIpcMemHandle_t, unsigned) { return cudaErrorNotSupported; }
cudaError_t cudaIpcCloseMemHandle(void *) { return cudaErrorNotSupported; }

// what the tests read back after a failed launch
extern "C" __attribute__((visibility("default"))) void simt_set_cp_async_eager(int eager) { simt::g_cp_async_eager = eager != 0; }
extern "C" __attribute__((visibility("default"))) const char *simt_last_failure(void) { return simt::g_fail_msg.c_str(); }
