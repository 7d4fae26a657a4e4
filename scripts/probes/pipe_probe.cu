// pipe_probe.cu -- microbenchmarks behind the round-2 decisions (not product code).  Each question
// is one the planned kernels' models (profiles/r01_static_forward_model.md) had to guess:
//
//   1. reductions: how fast does the L2 take 32 scalar `red.global.add.f32` to 32 consecutive floats
//      of a random row (planned backward) compared with 8 `red.global.add.v4.f32` to the same 128
//      bytes (default backward)?  Rows from an L2-resident (37 MB) and a DRAM-sized (313 MB) array.
//   2. shared-memory updates: cycles per read-modify-write instruction as a function of the
//      conflict degree (1, 2, 3, 4 lanes per bank) -- the thing the bank plan removes.
//   3. broadcasts: LDS.64 of one address by 32 lanes vs SHFL of a register (staging the (col, w) pairs).
//   4. gathers into shared memory: 16-byte cp.async (LDGSTS) and cp.async.bulk (TMA unit, 128 B + 32 B
//      per row, one lane per row) against plain LDG, rows from a DRAM-sized operand (cp.async ring).
//
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a pipe_probe.cu -o pipe_probe && ./pipe_probe
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t err_ = (x); if (err_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(err_)); exit(1);} } while (0)

template <typename F>
static float time_ms(F launch, int iters = 5)
{
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    launch(); launch();
    CK(cudaEventRecord(a));
    for (int i = 0; i < iters; ++i) launch();
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    CK(cudaGetLastError());
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    return ms / iters;
}

// ---- 1. reductions ---------------------------------------------------------------------------
// one warp per run of rows; row r gets 32 floats added.  VEC: 8 lanes x v4 (4 rows per instruction),
// else 32 lanes x scalar (1 row per instruction).
template <bool VEC>
__global__ void __launch_bounds__(256) red_kernel(const int *__restrict__ rows, int64_t e, float *dst, int per_warp)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int64_t j0 = warp * per_warp, j1 = min(j0 + per_warp, e);
    for (int64_t j = j0; j + 4 <= j1; j += 4) {
        if (VEC) {
            const int r = __ldg(rows + j + (lane >> 3));
            float *p = dst + (int64_t)r * 32 + 4 * (lane & 7);
            asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %1, %1, %1};" ::"l"(p), "f"(1.0f) : "memory");
        } else {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int r = __ldg(rows + j + u);
                asm volatile("red.relaxed.gpu.global.add.f32 [%0], %1;" ::"l"(dst + (int64_t)r * 32 + lane), "f"(1.0f) : "memory");
            }
        }
    }
}

// ---- 2. shared-memory read-modify-write under a chosen conflict degree -----------------------------
template <int DEG>
__global__ void __launch_bounds__(256) rmw_kernel(float *out, int iters)
{
    __shared__ float acc[8][512];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 8 * 512; i += 256) (&acc[0][0])[i] = 0.f;
    __syncthreads();
    // DEG lanes share a bank (different words): lane l -> word (l / DEG) + 32 * (l % DEG)
    float *a = &acc[warp][(lane / DEG) + 32 * (lane % DEG)];
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            *a = fmaf(1.0f, 0.5f, *a);
            __syncwarp();
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) out[DEG] = (float)(t1 - t0) / (8.f * iters);
    if (*a == -1.f) out[0] = 1.f;
}

// ---- 3. broadcast: LDS.64 vs SHFL -------------------------------------------------------------------
template <bool SHFL>
__global__ void __launch_bounds__(256) bcast_kernel(float *out, int iters)
{
    __shared__ int2 stage[8][64];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    stage[warp][lane] = make_int2(lane, lane);
    stage[warp][lane + 32] = make_int2(lane, lane);
    __syncthreads();
    int acc = 0;
    int2 mine = stage[warp][lane];
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 32; ++u) {
            if (SHFL) {
                acc += __shfl_sync(0xffffffffu, mine.x, u) + __shfl_sync(0xffffffffu, mine.y, u);
            } else {
                int2 v;  // ld.shared.v2 of one address by all lanes (what the staging broadcast is)
                asm volatile("ld.shared.v2.s32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y)
                             : "r"((unsigned)__cvta_generic_to_shared(&stage[warp][(u + i) & 63])) : "memory");
                acc += v.x + v.y;
            }
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) out[SHFL ? 9 : 8] = (float)(t1 - t0) / (32.f * iters);
    if (acc == -1) out[0] = 1.f;
}

// ---- 4. gathers of 160-byte CBSR rows (k = 32) into shared memory ---------------------------------------
constexpr int RING = 16;  // rows per group, two groups per warp
template <int MODE>       // 0: LDG to registers (baseline), 1: cp.async 16 B, 2: cp.async.bulk per row
__global__ void __launch_bounds__(64) gather_kernel(const int *__restrict__ idx, int64_t e, const float *__restrict__ data,
                                                    const unsigned char *__restrict__ sel, float *__restrict__ out,
                                                    int per_warp)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char *ring = smem + warp * (2 * RING * 160);
    __shared__ __align__(8) unsigned long long mbar[2][2];
    const int64_t w = (int64_t)blockIdx.x * 2 + warp;
    const int64_t j0 = w * per_warp, j1 = min(j0 + per_warp, e);
    float acc = 0.f;
    if (MODE == 2 && lane == 0) {
        for (int s = 0; s < 2; ++s)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((unsigned)__cvta_generic_to_shared(&mbar[warp][s])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    unsigned phase = 0;
    auto issue = [&](int64_t j, int g) {
        unsigned char *half = ring + g * RING * 160;
        if (MODE == 1) {
            for (int c = lane; c < RING * 10; c += 32) {
                const int r = c / 10, p = c - r * 10;
                const int64_t col = __ldg(idx + j + r);
                const unsigned char *src = p < 8 ? (const unsigned char *)data + col * 128 + 16 * p : sel + col * 32 + 16 * (p - 8);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"l"(__cvta_generic_to_shared(half + r * 160 + 16 * p)), "l"(src) : "memory");
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        } else if (MODE == 2) {
            const unsigned mb = (unsigned)__cvta_generic_to_shared(&mbar[warp][g]);
            if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(RING * 160) : "memory");
            __syncwarp();
            if (lane < RING) {
                const int64_t col = __ldg(idx + j + lane);
                const unsigned d0 = (unsigned)__cvta_generic_to_shared(half + lane * 160);
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 128, [%2];" ::"r"(d0), "l"(data + col * 32), "r"(mb) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 32, [%2];" ::"r"(d0 + 128), "l"(sel + col * 32), "r"(mb) : "memory");
            }
        }
    };
    auto wait = [&](int g, bool last) {
        if (MODE == 1) {
            if (last) asm volatile("cp.async.wait_group 0;" ::: "memory"); else asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else if (MODE == 2) {
            const unsigned mb = (unsigned)__cvta_generic_to_shared(&mbar[warp][g]);
            unsigned done = 0;
            while (!done)
                asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(mb), "r"((phase >> g) & 1u) : "memory");
            phase ^= 1u << g;
        }
        __syncwarp();
    };
    if (MODE == 0) {
        for (int64_t j = j0; j + 4 <= j1; j += 4) {
            float v[4]; unsigned char s[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) { const int64_t col = __ldg(idx + j + u); v[u] = __ldg(data + col * 32 + lane); s[u] = __ldg(sel + col * 32 + lane); }
#pragma unroll
            for (int u = 0; u < 4; ++u) acc = fmaf(v[u], (float)s[u], acc);
        }
    } else {
        const int64_t groups = (j1 - j0) / RING;
        if (groups > 0) issue(j0, 0);
        if (groups > 1) issue(j0 + RING, 1);
        for (int64_t g = 0; g < groups; ++g) {
            wait((int)(g & 1), g + 1 >= groups);
            const unsigned char *half = ring + (g & 1) * RING * 160;
            for (int r = 0; r < RING; ++r)
                acc = fmaf(((const float *)(half + r * 160))[lane], (float)half[r * 160 + 128 + lane], acc);
            __syncwarp();
            if (g + 2 < groups) issue(j0 + (g + 2) * RING, (int)(g & 1));
        }
    }
    out[w * 32 + lane] = acc;
}

int main()
{
    float *out; CK(cudaMalloc(&out, 1 << 22)); CK(cudaMemset(out, 0, 1 << 22));
    float h[16] = {0};
    // 2. + 3. (one CTA per SM is enough: these are per-SM pipe costs in cycles)
    rmw_kernel<1><<<148, 256>>>(out, 2000); rmw_kernel<2><<<148, 256>>>(out, 2000);
    rmw_kernel<3><<<148, 256>>>(out, 2000); rmw_kernel<4><<<148, 256>>>(out, 2000);
    bcast_kernel<false><<<148, 256>>>(out, 500); bcast_kernel<true><<<148, 256>>>(out, 500);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h, out, sizeof h, cudaMemcpyDeviceToHost));
    printf("shared RMW (LDS+FFMA+STS+syncwarp), 8 warps/SM, cycles per instruction triple by conflict degree: 1:%.1f 2:%.1f 3:%.1f 4:%.1f\n", h[1], h[2], h[3], h[4]);
    printf("broadcast of a staged pair: LDS.64 %.1f cycles, 2 x SHFL %.1f cycles (8 warps/SM)\n", h[8], h[9]);

    // 1. + 4. need index streams
    const int64_t e = 60000000;
    for (int64_t n : {233000LL, 2449029LL}) {
        std::vector<int> hi(e);
        uint64_t st = 88172645463325252ull;
        for (int64_t i = 0; i < e; ++i) { st ^= st << 13; st ^= st >> 7; st ^= st << 17; hi[i] = (int)(st % (uint64_t)n); }
        int *idx; float *data; unsigned char *sel; float *dst;
        CK(cudaMalloc(&idx, e * 4)); CK(cudaMalloc(&data, n * 128)); CK(cudaMalloc(&sel, n * 32)); CK(cudaMalloc(&dst, n * 128));
        CK(cudaMemcpy(idx, hi.data(), e * 4, cudaMemcpyHostToDevice));
        CK(cudaMemset(data, 0, n * 128)); CK(cudaMemset(sel, 1, n * 32)); CK(cudaMemset(dst, 0, n * 128));
        const int per_warp = 1024;
        const unsigned grid8 = (unsigned)((e / per_warp + 7) / 8);
        float tv = time_ms([&] { red_kernel<true><<<grid8, 256>>>(idx, e, dst, per_warp); });
        float ts = time_ms([&] { red_kernel<false><<<grid8, 256>>>(idx, e, dst, per_warp); });
        printf("n=%lld (%.0f MB of rows): 128-byte row reductions  v4 x 8 lanes %.3f ms (%.1f G rows/s)   scalar x 32 lanes %.3f ms (%.1f G rows/s)\n",
               (long long)n, n * 128 / 1e6, tv, e / tv / 1e6, ts, e / ts / 1e6);
        const unsigned grid2 = (unsigned)((e / per_warp + 1) / 2);
        const size_t smem = 2 * 2 * RING * 160;
        float t0 = time_ms([&] { gather_kernel<0><<<grid2, 64, smem>>>(idx, e, data, sel, out, per_warp); });
        float t1 = time_ms([&] { gather_kernel<1><<<grid2, 64, smem>>>(idx, e, data, sel, out, per_warp); });
        float t2 = time_ms([&] { gather_kernel<2><<<grid2, 64, smem>>>(idx, e, data, sel, out, per_warp); });
        printf("n=%lld: gather of 160-byte rows, 2-warp CTAs  LDG U=4 %.3f ms (%.1f G rows/s)  cp.async ring %.3f ms (%.1f)  cp.async.bulk ring %.3f ms (%.1f)\n",
               (long long)n, t0, e / t0 / 1e6, t1, e / t1 / 1e6, t2, e / t2 / 1e6);
        CK(cudaFree(idx)); CK(cudaFree(data)); CK(cudaFree(sel)); CK(cudaFree(dst));
    }
    return 0;
}
