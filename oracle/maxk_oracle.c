/*
 * maxk_oracle.c -- CPU restatement of the MaxK-GNN sparse-aggregation hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may build, load or call it, and only as the checker / the timed CPU
 * baseline.  The product path (spgemm-new_b200/) never links or imports it.
 *
 * Every function cites the reference file:line (paths relative to the reference
 * repository root) whose algorithm it restates.  No reference source is copied;
 * the math is restated from SURVEY.md Appendix A.
 *
 * Pinning status (see DESIGN.md "Oracle"):
 *   - work list        : pinned bit-exact against kernels/generate_meta.py run in the
 *                        build container (tests/golden/worklist_*.npz).
 *   - TopK -> CBSR     : pinned against torch.topk (the arithmetic the reference models
 *                        actually call, utils/models.py:48, kernels/spmm_bindings.cpp:179)
 *                        on tie-free rows; tie order is the oracle's own rule
 *                        (lowest column wins) -- the reference pins none ("parity unpinned"
 *                        for ties).
 *   - forward/backward : pinned against the reference CUDA kernels compiled unmodified
 *                        (oracle/_ref) and run on the B200 box; vectors committed under
 *                        tests/golden/ref_kernels_*.npz once generated there.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define ORACLE_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------- */
/* a4: warp-level edge-balanced work list                                     */
/* restates kernels/generate_meta.py:30-48 (row loop, <=warp_max_nz chunks,   */
/* record (row, loc, len, 0)); consumers spmm_maxk.cu:37-40.                  */
/* Returns W.  If out == NULL only counts.                                    */
/* ------------------------------------------------------------------------- */
ORACLE_API int64_t oracle_worklist(const int32_t *indptr, int64_t n, int32_t max_nz,
                                   int32_t *out)
{
    int64_t w = 0;
    int64_t cur_loc = 0; /* generate_meta.py:30 running counter `cur_loc` */
    for (int64_t i = 0; i < n; ++i) {
        int64_t deg = (int64_t)indptr[i + 1] - (int64_t)indptr[i];
        if (deg == 0)
            continue;
        int64_t done = 0;
        while (done < deg) {
            int64_t len = deg - done;
            if (len > max_nz)
                len = max_nz;
            if (out) {
                out[4 * w + 0] = (int32_t)i;
                out[4 * w + 1] = (int32_t)cur_loc;
                out[4 * w + 2] = (int32_t)len;
                out[4 * w + 3] = 0;
            }
            cur_loc += len;
            done += len;
            ++w;
        }
    }
    return w;
}

/* ------------------------------------------------------------------------- */
/* a1: MaxK / TopK -> CBSR                                                    */
/* Selection rule of kernels/maxk_kernel.cu:35-88 (threshold with >=k above,  */
/* first k in ascending column order) restated as exact top-k with lowest-    */
/* column tie-break and ascending-column emission; the set equals */
/* torch.topk's (kernels/spmm_bindings.cpp:179, utils/models.py:48) whenever   */
/* the k-th and (k+1)-th values differ.  NaN ranks largest (torch rule);       */
/* -0.0 == +0.0.  Values are bit-copied.                                       */
/* ------------------------------------------------------------------------- */
static inline uint32_t f32_key(float f)
{
    uint32_t b;
    memcpy(&b, &f, 4);
    if ((b & 0x7fffffffu) > 0x7f800000u)
        return 0xffffffffu; /* NaN: largest */
    if (b == 0x80000000u)
        b = 0; /* -0 == +0 */
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

typedef struct {
    uint32_t key;
    int32_t col;
} keycol_t;

static int cmp_desc_key_asc_col(const void *a, const void *b)
{
    const keycol_t *x = (const keycol_t *)a, *y = (const keycol_t *)b;
    if (x->key != y->key)
        return (x->key > y->key) ? -1 : 1;
    return (x->col < y->col) ? -1 : (x->col > y->col);
}

ORACLE_API void oracle_topk_cbsr_f32(const float *x, int64_t n, int32_t d, int32_t k,
                                     float *sp_data, uint8_t *sp_index)
{
#pragma omp parallel
    {
        keycol_t *kc = (keycol_t *)malloc(sizeof(keycol_t) * (size_t)d);
        uint8_t *chosen = (uint8_t *)malloc((size_t)d);
#pragma omp for schedule(static)
        for (int64_t i = 0; i < n; ++i) {
            const float *row = x + i * d;
            for (int32_t c = 0; c < d; ++c) {
                kc[c].key = f32_key(row[c]);
                kc[c].col = c;
            }
            qsort(kc, (size_t)d, sizeof(keycol_t), cmp_desc_key_asc_col);
            memset(chosen, 0, (size_t)d);
            for (int32_t j = 0; j < k; ++j)
                chosen[kc[j].col] = 1;
            int32_t p = 0;
            for (int32_t c = 0; c < d; ++c) {
                if (chosen[c]) {
                    sp_data[i * k + p] = row[c]; /* bit copy */
                    sp_index[i * k + p] = (uint8_t)c;
                    ++p;
                }
            }
        }
        free(kc);
        free(chosen);
    }
}

/* uint8 variant: the data type the reference kernel itself works on
 * (kernels/maxk_kernel.cu:23, cuda_kernel_bindings.cpp:164-201), with the same
 * exact-top-k / lowest-column / ascending-emission rule. */
ORACLE_API void oracle_topk_cbsr_u8(const uint8_t *x, int64_t n, int32_t d, int32_t k,
                                    uint8_t *value, uint8_t *index)
{
#pragma omp parallel
    {
        keycol_t *kc = (keycol_t *)malloc(sizeof(keycol_t) * (size_t)d);
        uint8_t *chosen = (uint8_t *)malloc((size_t)d);
#pragma omp for schedule(static)
        for (int64_t i = 0; i < n; ++i) {
            const uint8_t *row = x + i * d;
            for (int32_t c = 0; c < d; ++c) {
                kc[c].key = row[c];
                kc[c].col = c;
            }
            qsort(kc, (size_t)d, sizeof(keycol_t), cmp_desc_key_asc_col);
            memset(chosen, 0, (size_t)d);
            for (int32_t j = 0; j < k; ++j)
                chosen[kc[j].col] = 1;
            int32_t p = 0;
            for (int32_t c = 0; c < d; ++c) {
                if (chosen[c]) {
                    value[i * k + p] = row[c];
                    index[i * k + p] = (uint8_t)c;
                    ++p;
                }
            }
        }
        free(kc);
        free(chosen);
    }
}

/* Dense MaxK nonlinearity: out = x * mask (utils/models.py:48-52). */
ORACLE_API void oracle_maxk_dense_f32(const float *x, int64_t n, int32_t d, int32_t k,
                                      float *out)
{
    float *sd = (float *)malloc(sizeof(float) * (size_t)n * (size_t)k);
    uint8_t *si = (uint8_t *)malloc((size_t)n * (size_t)k);
    oracle_topk_cbsr_f32(x, n, d, k, sd, si);
    memset(out, 0, sizeof(float) * (size_t)n * (size_t)d);
    for (int64_t i = 0; i < n; ++i)
        for (int32_t j = 0; j < k; ++j)
            out[i * d + si[i * k + j]] = sd[i * k + j];
    free(sd);
    free(si);
}

/* ------------------------------------------------------------------------- */
/* a5: CBSR -> dense scatter (utils/models.py:136-141; main.cu:135-146)       */
/* ------------------------------------------------------------------------- */
ORACLE_API void oracle_cbsr_scatter_f32(const float *sp_data, const uint8_t *sp_index,
                                        int64_t n, int32_t k, int32_t d, float *dense)
{
    memset(dense, 0, sizeof(float) * (size_t)n * (size_t)d);
    for (int64_t i = 0; i < n; ++i)
        for (int32_t j = 0; j < k; ++j) {
            int32_t c = sp_index[i * k + j];
            if (c < d) /* models.py:140 guard */
                dense[i * d + c] = sp_data[i * k + j];
        }
}

/* gather the dense tensor at the selector (MaxK.backward's non-zero part,
 * utils/models.py:55-59 in CBSR layout) */
ORACLE_API void oracle_cbsr_gather_f32(const float *dense, const uint8_t *sp_index,
                                       int64_t n, int32_t k, int32_t d, float *sp_data)
{
    for (int64_t i = 0; i < n; ++i)
        for (int32_t j = 0; j < k; ++j)
            sp_data[i * k + j] = dense[i * d + sp_index[i * k + j]];
}

/* ------------------------------------------------------------------------- */
/* a2: forward row-wise-product SpGEMM                                        */
/* restates kernels/spmm_maxk.cu:62-105: per task, acc[sel[idx[e],l]] +=      */
/* val[e]*data[idx[e],l] over the task's edges in order (fused multiply-add,   */
/* SURVEY Q14), then vout[row,:] += acc.                                       */
/* f32 flavour: per-task fp32 FMA accumulation, partials added to the row in   */
/* task order (one of the orders the reference's atomics can produce).         */
/* f64 flavour: everything in double -- the tolerance anchor.                  */
/* ------------------------------------------------------------------------- */
ORACLE_API void oracle_spgemm_forward_f32(const int32_t *warp4, int64_t W, const int32_t *idx,
                                          const float *val, const float *sp_data,
                                          const uint8_t *sp_index, int32_t k, int32_t d,
                                          int64_t n_out, float *y)
{
    memset(y, 0, sizeof(float) * (size_t)n_out * (size_t)d);
    float *acc = (float *)malloc(sizeof(float) * (size_t)d);
    for (int64_t t = 0; t < W; ++t) {
        int64_t row = warp4[4 * t + 0], loc = warp4[4 * t + 1], len = warp4[4 * t + 2];
        for (int32_t c = 0; c < d; ++c)
            acc[c] = 0.0f; /* spmm_maxk.cu:52-55 */
        for (int64_t e = loc; e < loc + len; ++e) {
            float lv = val[e];
            int64_t j = idx[e];
            for (int32_t l = 0; l < k; ++l) { /* spmm_maxk.cu:85-96 */
                int32_t c = sp_index[j * k + l];
                acc[c] = fmaf(lv, sp_data[j * k + l], acc[c]);
            }
        }
        for (int32_t c = 0; c < d; ++c) /* spmm_maxk.cu:101-105 */
            y[row * d + c] += acc[c];
    }
    free(acc);
}

ORACLE_API void oracle_spgemm_forward_f64(const int32_t *warp4, int64_t W, const int32_t *idx,
                                          const float *val, const float *sp_data,
                                          const uint8_t *sp_index, int32_t k, int32_t d,
                                          int64_t n_out, double *y)
{
    memset(y, 0, sizeof(double) * (size_t)n_out * (size_t)d);
    /* tasks of one row are contiguous and rows ascend (generate_meta.py:30-45), so
     * distinct threads touching the same row is possible only at chunk seams; use
     * atomics there to stay generic for arbitrary task lists. */
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t t = 0; t < W; ++t) {
        int64_t row = warp4[4 * t + 0], loc = warp4[4 * t + 1], len = warp4[4 * t + 2];
        for (int64_t e = loc; e < loc + len; ++e) {
            double lv = val[e];
            int64_t j = idx[e];
            for (int32_t l = 0; l < k; ++l) {
                int32_t c = sp_index[j * k + l];
                double p = lv * (double)sp_data[j * k + l];
#pragma omp atomic
                y[row * d + c] += p;
            }
        }
    }
}

/* ------------------------------------------------------------------------- */
/* a3: backward outer-product sampled SSpMM                                   */
/* restates kernels/spmm_maxk_backward.cu:52-113: g = dY[row,:]; for every    */
/* edge (row -> j) of the task: dXc[j,l] += val[e] * g[sel[j,l]].              */
/* ------------------------------------------------------------------------- */
ORACLE_API void oracle_sspmm_backward_f32(const int32_t *warp4, int64_t W, const int32_t *idx,
                                          const float *val, const float *dy,
                                          const uint8_t *sp_index, int32_t k, int32_t d,
                                          int64_t n_cbsr, float *dxc)
{
    memset(dxc, 0, sizeof(float) * (size_t)n_cbsr * (size_t)k);
    for (int64_t t = 0; t < W; ++t) {
        int64_t row = warp4[4 * t + 0], loc = warp4[4 * t + 1], len = warp4[4 * t + 2];
        const float *g = dy + row * d;
        for (int64_t e = loc; e < loc + len; ++e) {
            float lv = val[e];
            int64_t j = idx[e];
            for (int32_t l = 0; l < k; ++l) /* spmm_maxk_backward.cu:92-111 */
                dxc[j * k + l] += lv * g[sp_index[j * k + l]];
        }
    }
}

ORACLE_API void oracle_sspmm_backward_f64(const int32_t *warp4, int64_t W, const int32_t *idx,
                                          const float *val, const float *dy,
                                          const uint8_t *sp_index, int32_t k, int32_t d,
                                          int64_t n_cbsr, double *dxc)
{
    memset(dxc, 0, sizeof(double) * (size_t)n_cbsr * (size_t)k);
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t t = 0; t < W; ++t) {
        int64_t row = warp4[4 * t + 0], loc = warp4[4 * t + 1], len = warp4[4 * t + 2];
        const float *g = dy + row * d;
        for (int64_t e = loc; e < loc + len; ++e) {
            double lv = val[e];
            int64_t j = idx[e];
            for (int32_t l = 0; l < k; ++l) {
                double p = lv * (double)g[sp_index[j * k + l]];
#pragma omp atomic
                dxc[j * k + l] += p;
            }
        }
    }
}

/* ------------------------------------------------------------------------- */
/* Independent cross-check: plain CSR SpMM on the dense twin (the cuSPARSE     */
/* comparator of kernels/spmm_cusparse.cu:6-62 / direct_kernel_interface.py:   */
/* 215-298 restated in double).  y = A * x_dense,  dx = A^T * dy.              */
/* ------------------------------------------------------------------------- */
ORACLE_API void oracle_csr_spmm_f64(const int32_t *indptr, const int32_t *idx, const float *val,
                                    int64_t n_rows, const float *xd, int32_t d, double *y)
{
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t r = 0; r < n_rows; ++r) {
        double *yr = y + r * d;
        for (int32_t c = 0; c < d; ++c)
            yr[c] = 0.0;
        for (int64_t e = indptr[r]; e < indptr[r + 1]; ++e) {
            double lv = val[e];
            const float *xr = xd + (int64_t)idx[e] * d;
            for (int32_t c = 0; c < d; ++c)
                yr[c] += lv * (double)xr[c];
        }
    }
}

ORACLE_API void oracle_csr_spmm_t_f64(const int32_t *indptr, const int32_t *idx,
                                      const float *val, int64_t n_rows, int64_t n_cols,
                                      const float *dy, int32_t d, double *dx)
{
    memset(dx, 0, sizeof(double) * (size_t)n_cols * (size_t)d);
    for (int64_t r = 0; r < n_rows; ++r) {
        const float *g = dy + r * d;
        for (int64_t e = indptr[r]; e < indptr[r + 1]; ++e) {
            double lv = val[e];
            double *o = dx + (int64_t)idx[e] * d;
            for (int32_t c = 0; c < d; ++c)
                o[c] += lv * (double)g[c];
        }
    }
}

/* ------------------------------------------------------------------------- */
/* Q1: closed-form description of the reference forward kernel's tail defect   */
/* (kernels/spmm_maxk.cu:27,57-58,101-105): for dim_sparse < 32, column c of    */
/* task T = 12*b + w is dropped iff 12*b + w*floor(32/k) + floor((c%32)/k) >= W */
/* dropped[t*d + c] = 1 where the reference loses the contribution.            */
/* Known answer: debug.txt:106-128 (k=18, bad columns all have c%32 >= 18).    */
/* ------------------------------------------------------------------------- */
ORACLE_API void oracle_q1_forward_dropped(int64_t W, int32_t k, int32_t d, uint8_t *dropped)
{
    memset(dropped, 0, (size_t)W * (size_t)d);
    if (k >= 32)
        return;
    int32_t per = 32 / k;
    for (int64_t t = 0; t < W; ++t) {
        int64_t b = t / 12, w = t % 12;
        for (int32_t c = 0; c < d; ++c) {
            int64_t sparse_wid = w * per + (c % 32) / k;
            if (12 * b + sparse_wid >= W)
                dropped[t * d + c] = 1;
        }
    }
}

ORACLE_API int oracle_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
