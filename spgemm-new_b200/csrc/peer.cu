// peer.cu -- host side of the NVLink peer-memory path: IPC allocation, flag publication and
// the pull reduce-scatter of the partial gradients.  (The pull all-gather lives inside the
// fused forward kernel, spgemm_forward.cu + peer.cuh.)  New code: the reference is single-GPU.
#include "peer.cuh"

namespace maxk {

__global__ void peer_publish_kernel(unsigned *flag_word, unsigned epoch)
{
    __threadfence_system();  // everything this stream wrote before is visible to the peers first
    st_release_sys(flag_word, epoch);
}

struct RsArgs {
    int world, rank;
    unsigned epoch;
    long long n4;   // float4 elements of this rank's block
    long long off;  // element offset of the block inside every partial buffer
    const float *partial[MAXK_PEER_MAX];
    const unsigned *flags[MAXK_PEER_MAX];
};

// out = sum_r partial_r[block of this rank]; r in rank order, so the sum is deterministic
__global__ void __launch_bounds__(256) reduce_scatter_pull_kernel(RsArgs a, float *__restrict__ out)
{
    if (threadIdx.x < a.world && threadIdx.x != a.rank)
        peer_spin_until<true>(a.flags[threadIdx.x] + PEER_FLAG_BWD, a.epoch, threadIdx.x);
    __syncthreads();
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.n4; i += stride) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        uint4 v[MAXK_PEER_MAX];
#pragma unroll
        for (int r = 0; r < MAXK_PEER_MAX; ++r)
            if (r < a.world) v[r] = ld_peer_u4(reinterpret_cast<const uint4 *>(a.partial[r] + a.off) + i);
#pragma unroll
        for (int r = 0; r < MAXK_PEER_MAX; ++r) {
            if (r < a.world) {
                acc.x += __uint_as_float(v[r].x);
                acc.y += __uint_as_float(v[r].y);
                acc.z += __uint_as_float(v[r].z);
                acc.w += __uint_as_float(v[r].w);
            }
        }
        reinterpret_cast<float4 *>(out)[i] = acc;
    }
}

}  // namespace maxk

using namespace maxk;

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

extern "C" int maxk_reduce_scatter_pull(const maxk_peer_ctx *ctx, uint32_t epoch, float *out,
                                        void *stream)
{
    MAXK_REQUIRE(ctx != nullptr, "maxk_reduce_scatter_pull: ctx is NULL");
    MAXK_REQUIRE(ctx->world >= 1 && ctx->world <= MAXK_PEER_MAX && ctx->rank >= 0 &&
                     ctx->rank < ctx->world,
                 "maxk_reduce_scatter_pull: world=%d rank=%d out of range", ctx->world, ctx->rank);
    MAXK_REQUIRE(ctx->k % 4 == 0 && ctx->block_rows % 16 == 0,
                 "maxk_reduce_scatter_pull: k %% 4 and block_rows %% 16 must be 0");
    const long long n = (long long)ctx->rows_of[ctx->rank] * ctx->k;
    if (n == 0) return MAXK_OK;
    MAXK_REQUIRE(out != nullptr && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
                 "maxk_reduce_scatter_pull: out must be non-NULL and 16-byte aligned");
    RsArgs a{};
    a.world = ctx->world;
    a.rank = ctx->rank;
    a.epoch = epoch;
    a.n4 = n / 4;
    a.off = (long long)ctx->rank * ctx->block_rows * ctx->k;
    for (int r = 0; r < ctx->world; ++r) {
        MAXK_REQUIRE(ctx->partial[r] && ctx->flags[r], "maxk_reduce_scatter_pull: rank %d not mapped", r);
        a.partial[r] = static_cast<const float *>(ctx->partial[r]);
        a.flags[r] = static_cast<const unsigned *>(ctx->flags[r]);
    }
    const long long blocks = ceil_div<long long>(a.n4, 256);
    const unsigned grid = (unsigned)(blocks < 4LL * kNumSMs ? blocks : 4LL * kNumSMs);
    reduce_scatter_pull_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(a, out);
    MAXK_LAUNCH_CHECK("reduce_scatter_pull_kernel");
    return MAXK_OK;
}
