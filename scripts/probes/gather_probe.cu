// gather_probe.cu -- microbenchmark behind the forward kernel's gather design (not product code):
// how fast can a B200 gather random CBSR rows (k=32: 128 B of values + 32 B of selectors)
// that do NOT fit the L2, as a function of layout (two arrays vs one 160 B record) and of
// the number of rows in flight per warp?   nvcc -O3 -arch=sm_100a gather_probe.cu -o gather_probe
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t err_ = (x); if (err_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(err_)); exit(1);} } while (0)

template <int U, bool PACKED>
__global__ void __launch_bounds__(256) gather_kernel(const int *__restrict__ idx, int64_t e,
                                                     const float *__restrict__ data,
                                                     const unsigned char *__restrict__ sel,
                                                     const unsigned char *__restrict__ rec,
                                                     float *__restrict__ out, int per_warp)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    int64_t j0 = warp * per_warp;
    if (j0 >= e) return;
    int64_t j1 = j0 + per_warp < e ? j0 + per_warp : e;
    float acc = 0.f;
    for (int64_t j = j0; j + U <= j1; j += U) {
        int c[U];
        float v[U];
        unsigned char s[U];
#pragma unroll
        for (int u = 0; u < U; ++u) c[u] = __ldcs(idx + j + u);   // warp-uniform broadcast load
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (PACKED) {
                const unsigned char *r = rec + (int64_t)c[u] * 160;
                v[u] = __ldg(reinterpret_cast<const float *>(r) + lane);
                s[u] = __ldg(r + 128 + lane);
            } else {
                v[u] = __ldg(data + (int64_t)c[u] * 32 + lane);
                s[u] = __ldg(sel + (int64_t)c[u] * 32 + lane);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) acc = fmaf(v[u], (float)s[u], acc);
    }
    out[warp * 32 + lane] = acc;
}

template <int U, bool PACKED>
float run(const int *idx, int64_t e, const float *data, const unsigned char *sel,
          const unsigned char *rec, float *out, int per_warp, int iters)
{
    const int64_t warps = (e + per_warp - 1) / per_warp;
    const unsigned grid = (unsigned)((warps + 7) / 8);
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    for (int i = 0; i < 2; ++i) gather_kernel<U, PACKED><<<grid, 256>>>(idx, e, data, sel, rec, out, per_warp);
    CK(cudaEventRecord(a));
    for (int i = 0; i < iters; ++i) gather_kernel<U, PACKED><<<grid, 256>>>(idx, e, data, sel, rec, out, per_warp);
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    return ms / iters;
}

int main(int argc, char **argv)
{
    const int64_t n = argc > 1 ? atoll(argv[1]) : 2449029;
    const int64_t e = argc > 2 ? atoll(argv[2]) : 61859140;
    const int per_warp = argc > 3 ? atoi(argv[3]) : 256;
    std::vector<int> h(e);
    uint64_t st = 88172645463325252ull;
    for (int64_t i = 0; i < e; ++i) { st ^= st << 13; st ^= st >> 7; st ^= st << 17; h[i] = (int)(st % (uint64_t)n); }
    int *idx; float *data; unsigned char *sel, *rec; float *out;
    CK(cudaMalloc(&idx, e * 4)); CK(cudaMalloc(&data, n * 128)); CK(cudaMalloc(&sel, n * 32));
    CK(cudaMalloc(&rec, n * 160)); CK(cudaMalloc(&out, ((e + per_warp - 1) / per_warp + 8) * 32 * 4));
    CK(cudaMemcpy(idx, h.data(), e * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(data, 0, n * 128)); CK(cudaMemset(sel, 1, n * 32)); CK(cudaMemset(rec, 0, n * 160));
    printf("n=%lld e=%lld per_warp=%d  operand %.0f MB\n", (long long)n, (long long)e, per_warp, n * 160 / 1e6);
#define RUN(U, P) { float ms = run<U, P>(idx, e, data, sel, rec, out, per_warp, 5); \
    printf("U=%2d %s: %.3f ms  %.2f G rows/s  %.0f GB/s\n", U, P ? "packed  " : "separate", ms, e / ms / 1e6, e * 164.0 / ms / 1e6); }
    RUN(1, false) RUN(2, false) RUN(4, false) RUN(8, false) RUN(16, false)
    RUN(1, true) RUN(2, true) RUN(4, true) RUN(8, true) RUN(16, true)
    return 0;
}
