// worklist.cu -- on-device generation of the warp-level edge-balanced work list.
//
// Replaces kernels/generate_meta.py:8-48 (an interpreted Python loop over all rows plus a
// disk round trip through w12_nz64_warp_4/<graph>.warp4) with three small kernels whose
// output is bit-identical to that file:
//     for every row with deg > 0, in row order, chunks of <= warp_max_nz nnz:
//         (row, indptr[row] + nz*c, min(nz, deg - nz*c), 0)
//   1. wl_block_count : per 1024-row block, number of tasks           (reads indptr once)
//   2. wl_scan_blocks : exclusive scan of the block counts (one CTA)  -> total W
//   3. wl_fill        : per block, re-derive the in-block offsets in shared memory, then one
//                       thread per TASK (binary search over the block's row offsets) writes a
//                       coalesced int4 record -- hub rows with hundreds of tasks do not
//                       serialise on one thread.
// Pure integer, HBM-bound: 4(N+1) bytes read twice + 16 W bytes written.
#include "common.cuh"

namespace maxk {

constexpr int WL_ROWS = 1024;  // rows per CTA == threads per CTA

__device__ __forceinline__ int wl_row_tasks(const int *__restrict__ indptr, int64_t row, int64_t n,
                                            int nz)
{
    if (row >= n) return 0;
    const int deg = __ldg(indptr + row + 1) - __ldg(indptr + row);
    return (deg + nz - 1) / nz;
}

// block-wide exclusive scan of one int per thread (1024 threads); returns the exclusive
// prefix, *total receives the block sum.
__device__ __forceinline__ int wl_block_excl_scan(int v, int *total)
{
    __shared__ int warp_sums[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int n = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += n;
    }
    if (lane == 31) warp_sums[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int s = warp_sums[lane];
        int si = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int n = __shfl_up_sync(0xffffffffu, si, o);
            if (lane >= o) si += n;
        }
        warp_sums[lane] = si - s;  // exclusive
        if (lane == 31) *total = si;
    }
    __syncthreads();
    return inc - v + warp_sums[warp];
}

__global__ void __launch_bounds__(WL_ROWS)
wl_block_count(const int *__restrict__ indptr, int64_t n, int nz, int *__restrict__ block_counts)
{
    __shared__ int total;
    const int64_t row = (int64_t)blockIdx.x * WL_ROWS + threadIdx.x;
    const int c = wl_row_tasks(indptr, row, n, nz);
    (void)wl_block_excl_scan(c, &total);
    if (threadIdx.x == 0) block_counts[blockIdx.x] = total;
}

// one CTA; block_counts[nb] -> exclusive offsets in place, block_counts[nb] = total.
// Totals are < 2^31 because E < 2^31 (int32 CSR offsets).
__global__ void __launch_bounds__(WL_ROWS)
wl_scan_blocks(int *__restrict__ block_counts, int64_t nb)
{
    __shared__ int total;
    __shared__ int carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int64_t base = 0; base < nb; base += WL_ROWS) {
        const int64_t i = base + threadIdx.x;
        const int v = (i < nb) ? block_counts[i] : 0;
        const int ex = wl_block_excl_scan(v, &total);
        const int carry = carry_s;
        if (i < nb) block_counts[i] = carry + ex;
        __syncthreads();
        if (threadIdx.x == 0) carry_s = carry + total;
        __syncthreads();
    }
    if (threadIdx.x == 0) block_counts[nb] = carry_s;
}

__global__ void __launch_bounds__(WL_ROWS)
wl_fill(const int *__restrict__ indptr, int64_t n, int nz, const int *__restrict__ block_offsets,
        int4 *__restrict__ warp4)
{
    __shared__ int total;
    __shared__ int row_first[WL_ROWS + 1];  // first in-block task index of each row
    const int64_t row0 = (int64_t)blockIdx.x * WL_ROWS;
    const int64_t row = row0 + threadIdx.x;
    const int c = wl_row_tasks(indptr, row, n, nz);
    const int ex = wl_block_excl_scan(c, &total);
    row_first[threadIdx.x] = ex;
    if (threadIdx.x == 0) row_first[WL_ROWS] = 0x7fffffff;
    __syncthreads();
    const int ntasks = total;
    const int64_t out0 = block_offsets[blockIdx.x];
    for (int t = threadIdx.x; t < ntasks; t += WL_ROWS) {
        // last row r with row_first[r] <= t (rows without tasks share their successor's
        // offset and are skipped by taking the LAST such r)
        int lo = 0, hi = WL_ROWS;  // invariant: row_first[lo] <= t < row_first[hi]
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (row_first[mid] <= t) lo = mid; else hi = mid;
        }
        const int64_t r = row0 + lo;
        const int c_in_row = t - row_first[lo];
        const int start = __ldg(indptr + r);
        const int deg = __ldg(indptr + r + 1) - start;
        int4 rec;
        rec.x = (int)r;
        rec.y = start + nz * c_in_row;
        rec.z = min(nz, deg - nz * c_in_row);
        rec.w = 0;
        warp4[out0 + t] = rec;
    }
}

}  // namespace maxk

using namespace maxk;

extern "C" int64_t maxk_worklist_workspace_bytes(int64_t num_rows)
{
    if (num_rows < 0) return 0;
    const int64_t nb = ceil_div<int64_t>(num_rows, WL_ROWS);
    return (nb + 1) * (int64_t)sizeof(int) + 16;
}

extern "C" int maxk_worklist_count(const int32_t *indptr, int64_t num_rows, int32_t warp_max_nz,
                                   void *workspace, int64_t *num_tasks_host, void *stream)
{
    MAXK_REQUIRE(num_rows >= 0, "maxk_worklist_count: negative num_rows");
    MAXK_REQUIRE(warp_max_nz >= 1, "maxk_worklist_count: warp_max_nz must be >= 1");
    MAXK_REQUIRE(num_tasks_host != nullptr, "maxk_worklist_count: num_tasks_host is NULL");
    *num_tasks_host = 0;
    if (num_rows == 0) return MAXK_OK;
    MAXK_REQUIRE(indptr != nullptr && workspace != nullptr, "maxk_worklist_count: NULL pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int64_t nb = ceil_div<int64_t>(num_rows, WL_ROWS);
    MAXK_REQUIRE(nb <= 0x7fffffffLL, "maxk_worklist_count: too many rows");
    int *bc = static_cast<int *>(workspace);
    wl_block_count<<<(unsigned)nb, WL_ROWS, 0, st>>>(indptr, num_rows, warp_max_nz, bc);
    MAXK_LAUNCH_CHECK("wl_block_count");
    wl_scan_blocks<<<1, WL_ROWS, 0, st>>>(bc, nb);
    MAXK_LAUNCH_CHECK("wl_scan_blocks");
    int total = 0;
    MAXK_CUDA(cudaMemcpyAsync(&total, bc + nb, sizeof(int), cudaMemcpyDeviceToHost, st));
    MAXK_CUDA(cudaStreamSynchronize(st));
    *num_tasks_host = total;
    return MAXK_OK;
}

extern "C" int maxk_worklist_fill(const int32_t *indptr, int64_t num_rows, int32_t warp_max_nz,
                                  const void *workspace, int32_t *warp4, int64_t num_tasks,
                                  void *stream)
{
    MAXK_REQUIRE(num_rows >= 0 && num_tasks >= 0, "maxk_worklist_fill: negative size");
    MAXK_REQUIRE(warp_max_nz >= 1, "maxk_worklist_fill: warp_max_nz must be >= 1");
    if (num_rows == 0 || num_tasks == 0) return MAXK_OK;
    MAXK_REQUIRE(indptr && workspace && warp4, "maxk_worklist_fill: NULL pointer");
    MAXK_REQUIRE((reinterpret_cast<uintptr_t>(warp4) & 15) == 0,
                 "maxk_worklist_fill: warp4 must be 16-byte aligned");
    const int64_t nb = ceil_div<int64_t>(num_rows, WL_ROWS);
    wl_fill<<<(unsigned)nb, WL_ROWS, 0, static_cast<cudaStream_t>(stream)>>>(
        indptr, num_rows, warp_max_nz, static_cast<const int *>(workspace),
        reinterpret_cast<int4 *>(warp4));
    MAXK_LAUNCH_CHECK("wl_fill");
    return MAXK_OK;
}
