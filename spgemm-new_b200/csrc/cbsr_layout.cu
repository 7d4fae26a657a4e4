// cbsr_layout.cu -- CBSR <-> dense conversions and the multi-GPU wire format.
//
// scatter / gather / mask replace the host-side loops around the reference kernels:
//   * utils/models.py:136-141  (N*k .item() calls to rebuild the dense gradient),
//   * utils/models.py:55-59    (MaxK.backward: grad * mask),
//   * main.cu:135-146 / direct_kernel_interface.py:241-242 (dense twin for cuSPARSE).
// All three are one-warp-per-row streaming kernels: each dense row is staged in shared
// memory so global traffic is 128-bit and coalesced; HBM-bound (4d + 5k bytes per row).
// pack / unpack build the per-node record  k*fp32 || k*u8 (padded to 16 B) that lets one
// all-gather move a row partition's CBSR features.
#include "common.cuh"

namespace maxk {

constexpr int LAY_WPB = 8;

enum class LayOp { Scatter, Mask };

template <LayOp OP>
__global__ void __launch_bounds__(LAY_WPB * 32)
cbsr_to_dense_kernel(const float *__restrict__ src, const unsigned char *__restrict__ sp_index,
                     int64_t n, int k, int d, float *__restrict__ dense)
{
    extern __shared__ __align__(16) float rowbuf_all[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float *rowbuf = rowbuf_all + warp * d;
    const int64_t row = (int64_t)blockIdx.x * LAY_WPB + warp;
    if (row >= n) return;
    for (int c = lane * 4; c < d; c += 128)
        *reinterpret_cast<float4 *>(rowbuf + c) = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncwarp();
    for (int j = lane; j < k; j += 32) {
        const int c = sp_index[row * k + j];
        if (c < d) {  // utils/models.py:140 guard
            if constexpr (OP == LayOp::Scatter)
                rowbuf[c] = __ldcs(src + row * k + j);
            else
                rowbuf[c] = __ldcs(src + row * d + c);
        }
    }
    __syncwarp();
    float *o = dense + row * d;
    for (int c = lane * 4; c < d; c += 128)
        st_stream_f4(reinterpret_cast<float4 *>(o + c), *reinterpret_cast<float4 *>(rowbuf + c));
}

__global__ void __launch_bounds__(256)
cbsr_gather_kernel(const float *__restrict__ dense, const unsigned char *__restrict__ sp_index,
                   int64_t n, int k, int d, float *__restrict__ sp_data)
{
    const int64_t total = n * (int64_t)k;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = i / k;
        const int c = sp_index[i];  // selectors >= d (possible when d < 256) read nothing, like the other layout kernels
        sp_data[i] = c < d ? __ldg(dense + row * d + c) : 0.f;
    }
}

__global__ void __launch_bounds__(256)
cbsr_pack_kernel(const float *__restrict__ sp_data, const unsigned char *__restrict__ sp_index,
                 int64_t n, int k, int rec_bytes, unsigned char *__restrict__ records)
{
    // one thread per 4-byte word of the record
    const int words = rec_bytes >> 2;
    const int data_words = k;               // k floats
    const int idx_words = (k + 3) >> 2;     // k bytes, rounded up
    const int64_t total = n * (int64_t)words;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = i / words;
        const int w = (int)(i - row * words);
        unsigned v = 0u;
        if (w < data_words) {
            v = __float_as_uint(sp_data[row * k + w]);
        } else if (w < data_words + idx_words) {
            const int b0 = (w - data_words) * 4;
#pragma unroll
            for (int b = 0; b < 4; ++b)
                if (b0 + b < k) v |= (unsigned)sp_index[row * k + b0 + b] << (8 * b);
        }
        reinterpret_cast<unsigned *>(records)[i] = v;
    }
}

__global__ void __launch_bounds__(256)
cbsr_unpack_kernel(const unsigned char *__restrict__ records, int64_t n, int k, int rec_bytes,
                   float *__restrict__ sp_data, unsigned char *__restrict__ sp_index)
{
    const int words = rec_bytes >> 2;
    const int data_words = k;
    const int idx_words = (k + 3) >> 2;
    const int64_t total = n * (int64_t)words;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = i / words;
        const int w = (int)(i - row * words);
        const unsigned v = reinterpret_cast<const unsigned *>(records)[i];
        if (w < data_words) {
            sp_data[row * k + w] = __uint_as_float(v);
        } else if (w < data_words + idx_words) {
            const int b0 = (w - data_words) * 4;
#pragma unroll
            for (int b = 0; b < 4; ++b)
                if (b0 + b < k) sp_index[row * k + b0 + b] = (unsigned char)(v >> (8 * b));
        }
    }
}

static int check_layout_args(const char *who, int64_t n, int k, int d)
{
    MAXK_REQUIRE(n >= 0, "%s: negative n", who);
    MAXK_REQUIRE(d > 0 && d <= MAXK_MAX_DIM && d % 4 == 0,
                 "%s: d=%d must be a multiple of 4 in [4,256]", who, d);
    MAXK_REQUIRE(k >= 1 && k <= 256, "%s: k=%d out of range", who, k);
    return MAXK_OK;
}

static unsigned stream_grid(int64_t total)
{
    int64_t g = ceil_div<int64_t>(total, 256);
    const int64_t cap = (int64_t)kNumSMs * 16;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (unsigned)g;
}

}  // namespace maxk

using namespace maxk;

extern "C" int maxk_cbsr_scatter(const float *sp_data, const uint8_t *sp_index, int64_t n, int32_t k,
                                 int32_t d, float *dense, void *stream)
{
    if (int rc = check_layout_args("maxk_cbsr_scatter", n, k, d)) return rc;
    if (n == 0) return MAXK_OK;
    MAXK_REQUIRE(sp_data && sp_index && dense, "maxk_cbsr_scatter: NULL pointer");
    MAXK_REQUIRE((reinterpret_cast<uintptr_t>(dense) & 15) == 0,
                 "maxk_cbsr_scatter: dense must be 16-byte aligned");
    const int64_t grid = ceil_div<int64_t>(n, LAY_WPB);
    MAXK_REQUIRE(grid <= 0x7fffffffLL, "maxk_cbsr_scatter: n too large");
    cbsr_to_dense_kernel<LayOp::Scatter>
        <<<(unsigned)grid, LAY_WPB * 32, LAY_WPB * d * sizeof(float), static_cast<cudaStream_t>(stream)>>>(
            sp_data, sp_index, n, k, d, dense);
    MAXK_LAUNCH_CHECK("cbsr_to_dense_kernel<scatter>");
    return MAXK_OK;
}

extern "C" int maxk_cbsr_mask(const float *in, const uint8_t *sp_index, int64_t n, int32_t k,
                              int32_t d, float *out, void *stream)
{
    if (int rc = check_layout_args("maxk_cbsr_mask", n, k, d)) return rc;
    if (n == 0) return MAXK_OK;
    MAXK_REQUIRE(in && sp_index && out, "maxk_cbsr_mask: NULL pointer");
    MAXK_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0,
                 "maxk_cbsr_mask: out must be 16-byte aligned");
    const int64_t grid = ceil_div<int64_t>(n, LAY_WPB);
    MAXK_REQUIRE(grid <= 0x7fffffffLL, "maxk_cbsr_mask: n too large");
    cbsr_to_dense_kernel<LayOp::Mask>
        <<<(unsigned)grid, LAY_WPB * 32, LAY_WPB * d * sizeof(float), static_cast<cudaStream_t>(stream)>>>(
            in, sp_index, n, k, d, out);
    MAXK_LAUNCH_CHECK("cbsr_to_dense_kernel<mask>");
    return MAXK_OK;
}

extern "C" int maxk_cbsr_gather(const float *dense, const uint8_t *sp_index, int64_t n, int32_t k,
                                int32_t d, float *sp_data, void *stream)
{
    if (int rc = check_layout_args("maxk_cbsr_gather", n, k, d)) return rc;
    if (n == 0) return MAXK_OK;
    MAXK_REQUIRE(dense && sp_index && sp_data, "maxk_cbsr_gather: NULL pointer");
    cbsr_gather_kernel<<<stream_grid(n * (int64_t)k), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        dense, sp_index, n, k, d, sp_data);
    MAXK_LAUNCH_CHECK("cbsr_gather_kernel");
    return MAXK_OK;
}

extern "C" int32_t maxk_cbsr_record_bytes(int32_t k)
{
    if (k < 1) return 0;
    return ((5 * k + 15) / 16) * 16;
}

extern "C" int maxk_cbsr_pack(const float *sp_data, const uint8_t *sp_index, int64_t n, int32_t k,
                              void *records, void *stream)
{
    MAXK_REQUIRE(n >= 0 && k >= 1 && k <= 256, "maxk_cbsr_pack: bad n/k");
    if (n == 0) return MAXK_OK;
    MAXK_REQUIRE(sp_data && sp_index && records, "maxk_cbsr_pack: NULL pointer");
    const int rb = maxk_cbsr_record_bytes(k);
    cbsr_pack_kernel<<<stream_grid(n * (int64_t)(rb / 4)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        sp_data, sp_index, n, k, rb, static_cast<unsigned char *>(records));
    MAXK_LAUNCH_CHECK("cbsr_pack_kernel");
    return MAXK_OK;
}

extern "C" int maxk_cbsr_unpack(const void *records, int64_t n, int32_t k, float *sp_data,
                                uint8_t *sp_index, void *stream)
{
    MAXK_REQUIRE(n >= 0 && k >= 1 && k <= 256, "maxk_cbsr_unpack: bad n/k");
    if (n == 0) return MAXK_OK;
    MAXK_REQUIRE(sp_data && sp_index && records, "maxk_cbsr_unpack: NULL pointer");
    const int rb = maxk_cbsr_record_bytes(k);
    cbsr_unpack_kernel<<<stream_grid(n * (int64_t)(rb / 4)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const unsigned char *>(records), n, k, rb, sp_data, sp_index);
    MAXK_LAUNCH_CHECK("cbsr_unpack_kernel");
    return MAXK_OK;
}
