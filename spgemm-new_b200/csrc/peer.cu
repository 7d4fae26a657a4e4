// peer.cu -- host side of the NVLink peer-memory path: IPC allocation, flag publication and
// the pull reduce-scatter of the partial gradients.  (The pull all-gather lives inside the
// fused forward kernel, spgemm_forward.cu + peer.cuh.)  New code: the reference is single-GPU.
#include "peer.cuh"

namespace maxk {

// clock64 ticks a device-side wait for a peer may take before it reports and traps; 0 = wait for ever.
// Default: 10 minutes at 1.965 GHz (NCCL's default watchdog is of the same order).
long long g_peer_spin_limit_cycles = 600LL * 1965000000LL;

__global__ void peer_publish_kernel(unsigned *flag_word, unsigned epoch)
{
    __threadfence_system();  // everything this stream wrote before is visible to the peers first
    st_release_sys(flag_word, epoch);
}

struct RsArgs {
    int world, rank, k;
    unsigned epoch;
    long long n4;   // float4 elements of this rank's block
    long long off;  // element offset of the block inside every partial buffer
    const float *partial[MAXK_PEER_MAX];
    const unsigned *flags[MAXK_PEER_MAX];
    const int *rows[MAXK_PEER_MAX];  // NULL: dense pair (every row of the block), else halo list
    long long count[MAXK_PEER_MAX];
    long long row0;
    long long spin_limit;
};

// dense pairs: out = sum_r partial_r[block of this rank]; r in rank order (deterministic)
__global__ void __launch_bounds__(256) reduce_scatter_pull_kernel(RsArgs a, float *__restrict__ out)
{
    if (threadIdx.x < a.world && threadIdx.x != a.rank)
        peer_spin_until<true>(a.flags[threadIdx.x] + PEER_FLAG_BWD, a.epoch, threadIdx.x, a.spin_limit);
    __syncthreads();
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.n4; i += stride) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        uint4 v[MAXK_PEER_MAX];
#pragma unroll
        for (int r = 0; r < MAXK_PEER_MAX; ++r)
            if (r < a.world && (r == a.rank || a.rows[r] == nullptr))
                v[r] = ld_peer_u4(reinterpret_cast<const uint4 *>(a.partial[r] + a.off) + i);
#pragma unroll
        for (int r = 0; r < MAXK_PEER_MAX; ++r) {
            if (r < a.world && (r == a.rank || a.rows[r] == nullptr)) {
                acc.x += __uint_as_float(v[r].x);
                acc.y += __uint_as_float(v[r].y);
                acc.z += __uint_as_float(v[r].z);
                acc.w += __uint_as_float(v[r].w);
            }
        }
        reinterpret_cast<float4 *>(out)[i] = acc;
    }
}

// halo pairs (runs after the kernel above): out[row - row0, :] += partial_p[row, :] for the rows
// peer p touched.  Different peers can list the same row -> vector reductions.
__global__ void __launch_bounds__(256) reduce_scatter_rows_kernel(RsArgs a, float *__restrict__ out)
{
    const int g4 = a.k >> 2;  // float4 groups per row
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (int p = 0; p < a.world; ++p) {
        if (p == a.rank || a.rows[p] == nullptr) continue;
        const long long items = a.count[p] * g4;
        for (long long i = tid; i < items; i += stride) {
            const long long row = __ldg(a.rows[p] + i / g4);
            const int q = (int)(i % g4);
            const uint4 v = ld_peer_u4(reinterpret_cast<const uint4 *>(a.partial[p] + row * a.k) + q);
            red_add_f4(out + (row - a.row0) * a.k + 4 * q, __uint_as_float(v.x), __uint_as_float(v.y),
                       __uint_as_float(v.z), __uint_as_float(v.w));
        }
    }
}

struct ZeroArgs {
    int world, rank, k;
    float *partial;
    const int *rows[MAXK_PEER_MAX];
    long long count[MAXK_PEER_MAX];
};

__global__ void __launch_bounds__(256) zero_rows_kernel(ZeroArgs a)
{
    const int g4 = a.k >> 2;
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (int p = 0; p < a.world; ++p) {
        if (p == a.rank || a.rows[p] == nullptr) continue;
        const long long items = a.count[p] * g4;
        for (long long i = tid; i < items; i += stride) {
            const long long row = __ldg(a.rows[p] + i / g4);
            reinterpret_cast<float4 *>(a.partial + row * a.k)[i % g4] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
}

static int check_ctx(const maxk_peer_ctx *ctx, const char *who)
{
    MAXK_REQUIRE(ctx != nullptr, "%s: ctx is NULL", who);
    MAXK_REQUIRE(ctx->world >= 1 && ctx->world <= MAXK_PEER_MAX && ctx->rank >= 0 &&
                     ctx->rank < ctx->world,
                 "%s: world=%d rank=%d out of range", who, ctx->world, ctx->rank);
    MAXK_REQUIRE(ctx->k >= 4 && ctx->k % 4 == 0, "%s: k=%d must be a positive multiple of 4", who, ctx->k);
    for (int r = 0; r < ctx->world; ++r) {
        MAXK_REQUIRE(ctx->rows_of[r] >= 0 && ctx->row_start[r] >= 0 &&
                         ctx->row_start[r] + ctx->rows_of[r] <= ctx->n_total,
                     "%s: row range of rank %d is outside [0, n_total)", who, r);
        MAXK_REQUIRE((ctx->row_start[r] * ctx->k) % 16 == 0,
                     "%s: row_start[%d] * k must be a multiple of 16 (align the partition)", who, r);
    }
    return MAXK_OK;
}

}  // namespace maxk

using namespace maxk;

extern "C" int maxk_set_peer_timeout_ms(int64_t ms)
{
    MAXK_REQUIRE(ms >= 0, "maxk_set_peer_timeout_ms: negative limit");
    g_peer_spin_limit_cycles = ms * 1965000LL;  // clock64 ticks at the B200's 1.965 GHz maximum; 0 = no limit
    return MAXK_OK;
}

extern "C" int maxk_peer_alloc(int64_t bytes, void **dev_ptr, void *handle_out)
{
    MAXK_REQUIRE(bytes > 0 && dev_ptr && handle_out, "maxk_peer_alloc: bad argument");
    void *p = nullptr;
    MAXK_CUDA(cudaMalloc(&p, (size_t)bytes));
    cudaError_t e = cudaMemset(p, 0, (size_t)bytes);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        return cuda_fail(e, "maxk_peer_alloc");
    }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    memcpy(handle_out, &h, sizeof(h));
    *dev_ptr = p;
    return MAXK_OK;
}

extern "C" int maxk_peer_free(void *dev_ptr)
{
    if (dev_ptr) MAXK_CUDA(cudaFree(dev_ptr));
    return MAXK_OK;
}

extern "C" int maxk_peer_open(const void *handle, void **dev_ptr)
{
    MAXK_REQUIRE(handle && dev_ptr, "maxk_peer_open: bad argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    void *p = nullptr;
    MAXK_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    *dev_ptr = p;
    return MAXK_OK;
}

extern "C" int maxk_peer_close(void *dev_ptr)
{
    if (dev_ptr) MAXK_CUDA(cudaIpcCloseMemHandle(dev_ptr));
    return MAXK_OK;
}

extern "C" int maxk_peer_publish(void *flag_word, uint32_t epoch, void *stream)
{
    MAXK_REQUIRE(flag_word != nullptr, "maxk_peer_publish: flag is NULL");
    peer_publish_kernel<<<1, 1, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<unsigned *>(flag_word), epoch);
    MAXK_LAUNCH_CHECK("peer_publish_kernel");
    return MAXK_OK;
}

extern "C" int maxk_peer_zero_partial(const maxk_peer_ctx *ctx, void *stream)
{
    if (int rc = check_ctx(ctx, "maxk_peer_zero_partial")) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    float *partial = static_cast<float *>(ctx->partial[ctx->rank]);
    MAXK_REQUIRE(partial != nullptr, "maxk_peer_zero_partial: own partial buffer is NULL");
    ZeroArgs z{};
    z.world = ctx->world;
    z.rank = ctx->rank;
    z.k = ctx->k;
    z.partial = partial;
    long long items = 0;
    for (int r = 0; r < ctx->world; ++r) {
        const bool whole = (r == ctx->rank) || ctx->recv_rows[r] == nullptr;
        if (whole) {
            if (ctx->rows_of[r] > 0) {
                MAXK_CUDA(cudaMemsetAsync(partial + ctx->row_start[r] * ctx->k, 0,
                                          (size_t)ctx->rows_of[r] * ctx->k * sizeof(float), st));
                count_launch();
            }
        } else {
            z.rows[r] = ctx->recv_rows[r];
            z.count[r] = ctx->recv_count[r];
            items += ctx->recv_count[r] * (ctx->k / 4);
        }
    }
    if (items > 0) {
        const long long blocks = ceil_div<long long>(items, 256);
        const unsigned grid = (unsigned)(blocks < 8LL * kNumSMs ? blocks : 8LL * kNumSMs);
        zero_rows_kernel<<<grid, 256, 0, st>>>(z);
        MAXK_LAUNCH_CHECK("zero_rows_kernel");
    }
    return MAXK_OK;
}

extern "C" int maxk_reduce_scatter_pull(const maxk_peer_ctx *ctx, uint32_t epoch, float *out,
                                        void *stream)
{
    if (int rc = check_ctx(ctx, "maxk_reduce_scatter_pull")) return rc;
    const long long n = ctx->rows_of[ctx->rank] * ctx->k;
    if (n == 0) return MAXK_OK;
    MAXK_REQUIRE(out != nullptr && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
                 "maxk_reduce_scatter_pull: out must be non-NULL and 16-byte aligned");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    RsArgs a{};
    a.world = ctx->world;
    a.rank = ctx->rank;
    a.k = ctx->k;
    a.epoch = epoch;
    a.spin_limit = g_peer_spin_limit_cycles;
    a.n4 = n / 4;
    a.row0 = ctx->row_start[ctx->rank];
    a.off = a.row0 * ctx->k;
    long long halo_items = 0;
    for (int r = 0; r < ctx->world; ++r) {
        MAXK_REQUIRE(ctx->partial[r] && ctx->flags[r], "maxk_reduce_scatter_pull: rank %d not mapped", r);
        a.partial[r] = static_cast<const float *>(ctx->partial[r]);
        a.flags[r] = static_cast<const unsigned *>(ctx->flags[r]);
        a.rows[r] = (r == ctx->rank) ? nullptr : ctx->send_rows[r];
        a.count[r] = ctx->send_count[r];
        if (a.rows[r]) halo_items += a.count[r] * (ctx->k / 4);
    }
    const long long blocks = ceil_div<long long>(a.n4, 256);
    const unsigned grid = (unsigned)(blocks < 4LL * kNumSMs ? blocks : 4LL * kNumSMs);
    reduce_scatter_pull_kernel<<<grid, 256, 0, st>>>(a, out);
    MAXK_LAUNCH_CHECK("reduce_scatter_pull_kernel");
    if (halo_items > 0) {
        const long long hb = ceil_div<long long>(halo_items, 256);
        const unsigned hgrid = (unsigned)(hb < 8LL * kNumSMs ? hb : 8LL * kNumSMs);
        reduce_scatter_rows_kernel<<<hgrid, 256, 0, st>>>(a, out);
        MAXK_LAUNCH_CHECK("reduce_scatter_rows_kernel");
    }
    return MAXK_OK;
}
