// TEST INFRASTRUCTURE: entry points of the tensor-core kernels (spgemm-new_b200/csrc/linear_topk.cu: tcgen05, TMA
// tensor maps, TMEM) for the SIMT interpreter's library.  The interpreter has no stand-in for that hardware, so
// these only report that; the kernels themselves are tested on a B200 (tests/test_gpu_linear_topk.py).
#include <stdint.h>

#include "maxk_b200.h"

namespace maxk {
void set_error(const char *fmt, ...);
}

extern "C" MAXK_API int64_t maxk_linear_topk_workspace_bytes(int32_t d, int32_t in_dim)
{
    return d > 0 && in_dim > 0 ? (int64_t)2 * d * in_dim * 4 : 0;
}

extern "C" MAXK_API int maxk_linear_topk_cbsr(const float *, int64_t, int32_t, const float *, const float *, int32_t,
                                              int32_t, float *, uint8_t *, float *, void *, void *)
{
    maxk::set_error("maxk_linear_topk_cbsr: tensor-core kernel, not available under the SIMT interpreter");
    return MAXK_ERR_UNSUPPORTED;
}

extern "C" MAXK_API int maxk_linear_f32(const float *, int64_t, int32_t, const float *, const float *, const float *,
                                        int32_t, float *, void *, void *)
{
    maxk::set_error("maxk_linear_f32: tensor-core kernel, not available under the SIMT interpreter");
    return MAXK_ERR_UNSUPPORTED;
}

extern "C" MAXK_API int maxk_linear_wgrad(const float *, const float *, int64_t, int32_t, int32_t, float *, int32_t, void *)
{
    maxk::set_error("maxk_linear_wgrad: tensor-core kernel, not available under the SIMT interpreter");
    return MAXK_ERR_UNSUPPORTED;
}
