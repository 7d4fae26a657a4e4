/*
 * maxk_b200.h -- C ABI of libmaxk_b200.so: the B200 (sm_100a) implementation of the
 * MaxK-GNN sparse-aggregation hot path.
 *
 * This is the drop-in boundary.  Every entry point takes plain device pointers, sizes
 * and a CUDA stream (a cudaStream_t passed as void*; NULL = legacy default stream);
 * no torch types, no allocation, no hidden synchronisation unless stated.  All
 * pointers are BORROWED: the caller owns the buffers and must keep them alive until
 * the stream has drained (same convention as the reference shims,
 * cuda_kernel_wrappers.cu:36-93).
 *
 * Return value: MAXK_OK (0) or a MAXK_ERR_* code; maxk_last_error() gives the
 * message for the calling thread.  Unlike the reference shims (which printf on
 * cudaGetLastError and return void, cuda_kernel_wrappers.cu:52-55) errors are reported.
 *
 * Citations "ref: file:line" are into julius-sk/spgemm-new and name the reference
 * interface each function replaces.  Layouts:
 *   CSR            indptr int32[N+1], indices int32[E], values float32[E]; row = destination
 *   work list      int32[4*W] = (row, loc, len, 0) per task, len <= warp_max_nz (64)
 *   CBSR           sp_data float32[N,k] + sp_index uint8[N,k], row-major, k <= D <= 256
 */
#ifndef MAXK_B200_H_
#define MAXK_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* the library is built with -fvisibility=hidden; only the ABI below is exported */
#if defined(__GNUC__)
#define MAXK_API __attribute__((visibility("default")))
#else
#define MAXK_API
#endif

#define MAXK_OK 0
#define MAXK_ERR_INVALID_ARGUMENT 1
#define MAXK_ERR_CUDA 2
#define MAXK_ERR_UNSUPPORTED 3

#define MAXK_ABI_VERSION 1
#define MAXK_WARP_MAX_NZ 64 /* ref: kernels/generate_meta.py:9 */
#define MAXK_MAX_DIM 256    /* selectors are uint8 (ref: spmm_maxk.cu:17) */
#define MAXK_PEER_MAX 16    /* ranks of one NVLink domain addressed by the peer-memory path */

/* Layout-compatible stand-in for CUDA's dim3 so that C callers can bind the
 * reference-named shims below (ref: cuda_kernel_bindings.cpp:11-35). */
typedef struct maxk_dim3 {
    unsigned int x, y, z;
} maxk_dim3;

MAXK_API int maxk_abi_version(void);
MAXK_API const char *maxk_last_error(void);
/* Number of kernels this library has launched in the calling process (for bench.py's
 * gpu_launches claim). */
MAXK_API int64_t maxk_launch_count(void);
/* Bench/test hook (not part of the reference surface): run length, in tasks per warp, of
 * the forward / backward kernels; 0 leaves that knob unchanged, a negative value returns it
 * to its automatic choice. */
MAXK_API int maxk_set_tuning(int fwd_tasks_per_warp, int bwd_tasks_per_warp);
/* Bench/test hook: accumulator copies R of the forward kernel (1, 2, 4, 8, 16; rounded down
 * to an instantiated value and up to the edges-per-instruction count); 0 = automatic. */
MAXK_API int maxk_set_forward_copies(int copies);
/* Bench/test hook: launch shape of the forward kernel (1 = 8-warp CTAs, 2 = 2-warp CTAs,
 * 3 = 8-warp CTAs with twice the gathers in flight, 8 / 12 = 2-warp CTAs whose gathers go through
 * cp.async into a shared-memory ring of 2 x 16 / 2 x 8 rows (k = 16, 32, 64 with 16-byte aligned
 * operands, otherwise shape 2)); 0 = automatic: 1 for long rows, 2 for short rows, 8 (k <= 32) or
 * 12 (k = 64) for short rows on an operand too large for the L2. */
MAXK_API int maxk_set_forward_depth(int depth);
/* Bench/test hook: the grouped fp32 TopK kernel (several rows per warp; d = 128 / 64 with k <= d / 2):
 * -1 = automatic (on where it is faster), 0 = one row per warp everywhere, 1 = on. */
MAXK_API int maxk_set_topk_group(int on);

/* ---- a4: work list (replaces kernels/generate_meta.py:8-48 and the disk round trip of
 * load_warp4_metadata, cuda_kernel_bindings.cpp:287-317) ------------------------------
 * Two calls because the caller allocates the output:
 *   1. maxk_worklist_count  -> W (synchronises `stream` once to return the count) and
 *      fills `workspace` (maxk_worklist_workspace_bytes(num_rows) bytes, device);
 *   2. maxk_worklist_fill   -> writes int32[4*W] records, bit-identical to
 *      generate_meta.py's .warp4 file for the same indptr.                            */
MAXK_API int64_t maxk_worklist_workspace_bytes(int64_t num_rows);
MAXK_API int maxk_worklist_count(const int32_t *indptr, int64_t num_rows, int32_t warp_max_nz,
                        void *workspace, int64_t *num_tasks_host, void *stream);
MAXK_API int maxk_worklist_fill(const int32_t *indptr, int64_t num_rows, int32_t warp_max_nz,
                       const void *workspace, int32_t *warp4, int64_t num_tasks, void *stream);

/* ---- a1: MaxK / TopK -> CBSR (replaces topk, kernels/maxk_kernel.cu:23-96, and the
 * torch::topk stand-ins kernels/spmm_bindings.cpp:163-204, utils/models.py:39,48) -----
 * Exact top-k per row, ties -> lowest column, emitted in ascending column order;
 * values are bit copies.  NaN ranks largest, -0 == +0.  d % 4 == 0, 1 <= k <= d <= 256.
 * dense_out (optional, may be NULL): x * mask, float32[n,d] (MaxK.forward,
 * utils/models.py:48-52).  sp_data / sp_index may be NULL when only dense_out is wanted. */
MAXK_API int maxk_topk_cbsr_f32(const float *x, int64_t n, int32_t d, int32_t k, float *sp_data,
                       uint8_t *sp_index, float *dense_out, void *stream);
/* uint8 flavour of the reference kernel's own data type (cuda_topk_maxk,
 * cuda_kernel_bindings.cpp:164-201). d % 8 == 0. */
MAXK_API int maxk_topk_cbsr_u8(const uint8_t *x, int64_t n, int32_t d, int32_t k, uint8_t *value,
                      uint8_t *index, void *stream);

/* ---- a2: forward row-wise-product SpGEMM  Y = A * X_cbsr (replaces
 * spmm_kernel_opt2_sparse_v3, kernels/spmm_maxk.cu:17-106; spmm_maxk_forward,
 * cuda_kernel_bindings.cpp:42-104) -----------------------------------------------------
 * out float32[num_rows_out, d] is fully DEFINED by the call (no pre-zero needed, not
 * accumulated -- SURVEY Q2).  Task rows must be non-decreasing (what generate_meta.py
 * produces).  Selectors of one CBSR row must be distinct.  d % 4 == 0.      */
MAXK_API int maxk_spgemm_forward(const int32_t *warp4, int64_t num_tasks, const int32_t *indices,
                        const float *values, const float *sp_data, const uint8_t *sp_index,
                        int32_t k, float *out, int64_t num_rows_out, int32_t d, void *stream);

/* Same, but with accumulate != 0 every contribution is ADDED to the existing content of
 * `out` (no zero-fill, vector reductions for every row): used by the multi-GPU
 * local-column / remote-column split. */
MAXK_API int maxk_spgemm_forward_acc(const int32_t *warp4, int64_t num_tasks, const int32_t *indices,
                            const float *values, const float *sp_data, const uint8_t *sp_index,
                            int32_t k, float *out, int64_t num_rows_out, int32_t d,
                            int32_t accumulate, void *stream);

/* ---- a2 through a bank plan (k = 32, 64; long rows on an L2-resident operand) ------------------
 * There the forward kernel is bound by shared-memory wavefronts of its scatter into the output row:
 * 32 random columns per edge collide in the 32 banks ~3.3-way (ncu, Reddit shape).
 * maxk_cbsr_bank_plan gives every entry of a CBSR row one of two homes for its column -- word c of
 * copy 0 (bank c mod 32) or word 256 + (c mod 8)*32 + c/8 of copy 1 (bank c/8) -- so that the 32
 * entries of one update instruction spread over the banks (fullest bank ~2.1 instead of 3.3), once per
 * row instead of once per edge: one warp per 32 entries, a few greedy rounds.  sp_addr uint16[n,k]
 * holds the chosen accumulator words and replaces sp_index in maxk_spgemm_forward_planned (same output
 * contract as maxk_spgemm_forward_acc; measured 3.99 -> ~3.1 ms on the Reddit shape at k = 32).
 * maxk_sspmm_backward_planned is maxk_sspmm_backward reading its columns out of the same plan (the
 * column of word w is w for w < 256, else ((w & 31) << 3) | ((w >> 5) & 7)): on the multi-GPU peer
 * path the plan travels over NVLink in place of the selectors.                                    */
MAXK_API int maxk_cbsr_bank_plan(const uint8_t *sp_index, int64_t n, int32_t k, uint16_t *sp_addr, void *stream);
MAXK_API int maxk_spgemm_forward_planned(const int32_t *warp4, int64_t num_tasks, const int32_t *indices,
                                const float *values, const float *sp_data, const uint16_t *sp_addr,
                                int32_t k, float *out, int64_t num_rows_out, int32_t d,
                                int32_t accumulate, void *stream);
MAXK_API int maxk_sspmm_backward_planned(const int32_t *warp4, int64_t num_tasks, const int32_t *indices,
                                const float *values, const float *dy, const uint16_t *sp_addr,
                                int32_t k, int32_t d, float *dxc, int64_t num_cbsr_rows,
                                int32_t accumulate, void *stream);

/* ---- a3: backward outer-product sampled SSpMM  dXc = sample(A^T dY, sp_index)
 * (replaces spmm_kernel_opt2_sparse_backward_v3, kernels/spmm_maxk_backward.cu:15-115;
 * spmm_maxk_backward, cuda_kernel_bindings.cpp:106-161) ----------------------------------
 * dxc float32[num_cbsr_rows, k].  accumulate == 0: dxc is zero-filled first (defined
 * output); accumulate != 0: contributions are added to the existing content (used by
 * the multi-GPU local/remote split).                                                   */
MAXK_API int maxk_sspmm_backward(const int32_t *warp4, int64_t num_tasks, const int32_t *indices,
                        const float *values, const float *dy, const uint8_t *sp_index,
                        int32_t k, int32_t d, float *dxc, int64_t num_cbsr_rows,
                        int32_t accumulate, void *stream);

/* ---- a5: CBSR <-> dense (replaces the .item() double loop utils/models.py:136-141 and
 * the mask multiply of MaxK.backward, utils/models.py:55-59) -----------------------------
 * scatter: dense[i, sp_index[i,j]] = sp_data[i,j], zeros elsewhere (dense fully defined)
 * gather : sp_data[i,j] = dense[i, sp_index[i,j]]
 * mask   : out[i,c] = in[i,c] if c in sp_index[i,:] else 0                              */
MAXK_API int maxk_cbsr_scatter(const float *sp_data, const uint8_t *sp_index, int64_t n, int32_t k,
                      int32_t d, float *dense, void *stream);
MAXK_API int maxk_cbsr_gather(const float *dense, const uint8_t *sp_index, int64_t n, int32_t k,
                     int32_t d, float *sp_data, void *stream);
MAXK_API int maxk_cbsr_mask(const float *in, const uint8_t *sp_index, int64_t n, int32_t k, int32_t d,
                   float *out, void *stream);

/* ---- f3 (i): the Linear in front of a MaxK with the selection in its epilogue ------------
 * (sp_data, sp_index) = TopK_k(x @ weight^T + bias): replaces `self.linlayers[i](x)` followed by `MaxK.apply`
 * (utils/models.py:336-341, 365-369) without the dense [n, d] activation in HBM.  One tcgen05 kernel: TMA-fed
 * TF32 tensor-core products with a two-term operand split (fp32 accuracy), fp32 accumulators in TMEM, TopK in the
 * epilogue.  x [n, in_dim] and weight [d, in_dim] row-major fp32, bias [d] or NULL; d must be 256, in_dim a
 * multiple of 4; x, workspace, linear_out 16-byte aligned.  workspace: maxk_linear_topk_workspace_bytes(d, in_dim)
 * bytes of device memory (the split weight), overwritten by every call.  linear_out (optional, [n, d]): the
 * Linear's own output, for callers that want it; sp_data / sp_index may then be NULL.  Selection semantics are
 * maxk_topk_cbsr_f32's applied to the kernel's own fp32 result (ties -> lowest column). */
MAXK_API int64_t maxk_linear_topk_workspace_bytes(int32_t d, int32_t in_dim);
MAXK_API int maxk_linear_topk_cbsr(const float *x, int64_t n, int32_t in_dim, const float *weight, const float *bias,
                                   int32_t d, int32_t k, float *sp_data, uint8_t *sp_index, float *linear_out,
                                   void *workspace, void *stream);

/* The dense Linear next to the path on the same tensor-core machinery (fp32 accuracy from split TF32 products;
 * replaces the library SGEMMs torch.nn.Linear runs in fp32 for utils/models.py:151-392 -- the forward and the two
 * backward products of a 256-wide layer):
 *   maxk_linear_f32:   out[n, d] = x[n, in_dim] @ weight[d, in_dim]^T + bias + residual[n, d]   (d = 256; bias and
 *                      residual may be NULL, residual may be `out`; also dX = dY @ W with weight := W^T)
 *   maxk_linear_wgrad: dw[d, in_dim] (+)= grad_out[n, d]^T @ x[n, in_dim]         (d = 256, in_dim <= 256)
 * All pointers 16-byte aligned, in_dim a multiple of 4; workspace as for maxk_linear_topk_cbsr. */
MAXK_API int maxk_linear_f32(const float *x, int64_t n, int32_t in_dim, const float *weight, const float *bias,
                             const float *residual, int32_t d, float *out, void *workspace, void *stream);
MAXK_API int maxk_linear_wgrad(const float *grad_out, const float *x, int64_t n, int32_t d, int32_t in_dim, float *dw,
                               int32_t accumulate, void *stream);

/* ---- LayerNorm of the layers (`self.normlayers[i]`, utils/models.py:207-213; nn.LayerNorm(hidden, affine)) -------
 * forward:  y = (x - mean) * rstd * gamma + beta per row of d <= 256 (multiple of 4) floats; mean / rstd [n] are kept
 *           for the backward; gamma / beta may be NULL (1 / 0).
 * backward: dx = rstd * (g*gamma - mean_d(g*gamma) - xhat * mean_d(g*gamma*xhat)); dgamma [d] = sum_rows g*xhat and
 *           dbeta [d] = sum_rows g are overwritten (either may be NULL).  One warp per row, HBM-bound. */
MAXK_API int maxk_layernorm_forward(const float *x, int64_t n, int32_t d, const float *gamma, const float *beta,
                                    float eps, float *y, float *mean, float *rstd, void *stream);
MAXK_API int maxk_layernorm_backward(const float *grad_out, const float *x, const float *mean, const float *rstd,
                                     const float *gamma, int64_t n, int32_t d, float *dx, float *dgamma, float *dbeta,
                                     void *stream);

/* LayerNorm + TopK in one pass (SAGE: the LayerNorm that closes layer i feeds only the MaxK that opens layer i + 1,
 * utils/models.py:252 / 230): (sp_data, sp_index) = TopK_k(LayerNorm(x)); the normalised activation is written only
 * if y_out is given.  Backward: the k-sparse gradient of the kept values goes straight into the LayerNorm backward
 * (maxk_layernorm_backward_cbsr: grad_data [n, k] at the columns sp_index) -- no scattered dense twin. */
MAXK_API int maxk_layernorm_topk_cbsr(const float *x, int64_t n, int32_t d, const float *gamma, const float *beta,
                                      float eps, int32_t k, float *sp_data, uint8_t *sp_index, float *mean,
                                      float *rstd, float *y_out, void *stream);
MAXK_API int maxk_layernorm_backward_cbsr(const float *grad_data, const uint8_t *sp_index, int32_t k, const float *x,
                                          const float *mean, const float *rstd, const float *gamma, int64_t n,
                                          int32_t d, float *dx, float *dgamma, float *dbeta, void *stream);

/* ---- the dense layer next to the aggregation on the CBSR operand itself,
 * without the dense [n, d] twin (measured on the products shape, k = 32: 3.8 vs 6.3 ms for fc_self,
 * 5.1 vs 6.2 ms for the gradient of the kept values, profiles/r02_variants.md).  m is a dense float32 [d, c] matrix,
 * row-major (for a Linear layer: the transposed weight), d <= 256, c % 4 == 0.
 * matmul     : out[i, 0:c] (+)= bias[0:c] + sum_j sp_data[i,j] * m[sp_index[i,j], 0:c]
 *              = fc_self(h) on the MaxK-sparsified h (utils/models_old.py:55-60; the reference runs
 *              torch's dense Linear on the masked [n, d] tensor).  bias may be NULL.
 * sampled_dot: out[i,j] (+)= sum_c g[i,c] * m[sp_index[i,j], c]   (float32 [n, k]; the gradient of
 *              matmul with respect to sp_data, g = dL/d out).
 * accumulate != 0 adds to the existing content of out.  Selectors >= d contribute nothing / 0.  */
MAXK_API int maxk_cbsr_matmul(const float *sp_data, const uint8_t *sp_index, int64_t n, int32_t k,
                     const float *m, int32_t d, int32_t c, const float *bias, float *out,
                     int32_t accumulate, void *stream);
/* axpy: y[i, sp_index[i,j]] += (alpha_host + *alpha_dev) * sp_data[i,j] in place (alpha_dev may be NULL) --
 * GINConv's "(1 + eps) * h" (utils/models.py:361) added to the aggregate without the dense twin of h. */
MAXK_API int maxk_cbsr_axpy(const float *sp_data, const uint8_t *sp_index, int64_t n, int32_t k, int32_t d,
                   const float *alpha_dev, float alpha_host, float *y, void *stream);
MAXK_API int maxk_cbsr_sampled_dot(const float *g, int64_t n, int32_t c, const float *m, int32_t d,
                          const uint8_t *sp_index, int32_t k, float *out, int32_t accumulate,
                          void *stream);

/* ---- multi-GPU wire format ("new"): one record per node,
 * k*float32 followed by k*uint8, padded to a 16-byte multiple, so that one all-gather
 * moves both arrays.  record_bytes = maxk_cbsr_record_bytes(k).                         */
MAXK_API int32_t maxk_cbsr_record_bytes(int32_t k);
MAXK_API int maxk_cbsr_pack(const float *sp_data, const uint8_t *sp_index, int64_t n, int32_t k,
                   void *records, void *stream);
MAXK_API int maxk_cbsr_unpack(const void *records, int64_t n, int32_t k, float *sp_data,
                     uint8_t *sp_index, void *stream);

/* ---- multi-GPU over NVLink peer memory (new: the reference is single-GPU) -----
 * One process per GPU.  Buffers that peers touch are allocated with maxk_peer_alloc, their
 * 64-byte CUDA IPC handles are exchanged by the host (torch.distributed in the Python layer)
 * and mapped with maxk_peer_open.  Rank r owns the node rows [row_start[r], row_start[r] +
 * rows_of[r]) of the GLOBAL node numbering; every rank holds full-height buffers in that
 * numbering, so CSR column ids need no remapping.  A maxk_peer_ctx lists, for every rank r of
 * the job, this process's mapping of r's buffers (entry [rank] = the process's own):
 *   g_data / g_index  gathered CBSR operand, float32 / uint8 [n_total, k]; rank r writes its own
 *                     rows of its own buffer (TopK in place) and the fused forward pushes them
 *                     into the same rows of the peers' buffers;
 *   partial           CBSR-shaped partial gradients, float32 [n_total, k];
 *   flags             uint32[64] per rank, zero-initialised: [1] = last epoch whose partial
 *                     gradients are published, [16 + r] = last epoch whose CBSR rows from
 *                     rank r have arrived in this rank's buffer (written by rank r);
 *   local_sync        uint32[4] in the process's own memory, zero-initialised.
 * Exchange volume: per ordered pair (this rank -> peer p) either the whole block travels
 * (send_rows[p] == NULL: "dense", TMA bulk copies) or only the rows p's edges reference
 * (send_rows[p] = device int32 list of global row ids inside this rank's block, ascending;
 * "halo" mode for sparse graphs, where an all-gather of everything would cost more than the
 * aggregation itself).  recv_rows[p] / recv_count[p] is the mirror image: the rows of p's block
 * that THIS rank's edges reference (NULL = treat the whole block as referenced).  The backward
 * zeroes / the reduce-scatter pulls exactly those rows.  Both sides must take the dense/halo
 * decision of a pair from the same numbers (list length vs rows_of).
 * Epochs are step numbers that only grow; buffers alternate between two contexts by epoch
 * parity so that nobody overwrites what a slower peer still reads.  row_start[r] * k % 16 == 0.
 * Every device-side wait is bounded (maxk_set_peer_timeout_ms, default 10 minutes): one lane reports
 * and the kernel traps instead of hanging the GPU.                                          */
typedef struct maxk_peer_ctx {
    int32_t world, rank, k, reserved;
    int64_t n_total;
    int64_t row_start[MAXK_PEER_MAX];
    int64_t rows_of[MAXK_PEER_MAX];
    void *g_data[MAXK_PEER_MAX];
    void *g_index[MAXK_PEER_MAX];
    void *partial[MAXK_PEER_MAX];
    void *flags[MAXK_PEER_MAX];
    const int32_t *send_rows[MAXK_PEER_MAX];
    int64_t send_count[MAXK_PEER_MAX];
    const int32_t *recv_rows[MAXK_PEER_MAX];
    int64_t recv_count[MAXK_PEER_MAX];
    void *local_sync;
} maxk_peer_ctx;

/* Limit of every device-side wait for a peer, in milliseconds (0 = wait for ever).  Applies to the
 * launches that follow on the calling process's current device. */
MAXK_API int maxk_set_peer_timeout_ms(int64_t ms);
/* cudaMalloc + cudaIpcGetMemHandle (handle_out: 64 bytes); memory is zero-filled. */
MAXK_API int maxk_peer_alloc(int64_t bytes, void **dev_ptr, void *handle_out);
MAXK_API int maxk_peer_free(void *dev_ptr);
/* cudaIpcOpenMemHandle of a peer's allocation (enables peer access lazily) / close it. */
MAXK_API int maxk_peer_open(const void *handle, void **dev_ptr);
MAXK_API int maxk_peer_close(void *dev_ptr);
/* Stream-ordered system-scope release store *flag_word = epoch (after the backward kernel on
 * the same stream: flags[rank] + 1). */
MAXK_API int maxk_peer_publish(void *flag_word, uint32_t epoch, void *stream);
/* Fused all-gather + forward SpGEMM: ONE kernel whose first |copy_ctas| CTAs push this rank's
 * CBSR rows to the peers over NVLink while the others aggregate (copy_ctas < 0: plain
 * load/store copies instead of TMA bulk copies for dense pairs).  The work list / CSR is the
 * concatenation [tasks whose columns this rank owns | tasks reading other ranks' columns];
 * tasks from wait_from_task on wait for the peers' rows to arrive.  out float32[num_rows_out, d]
 * is defined by the call.  k = ctx->k must be 8, 16 or a multiple of 32.                  */
MAXK_API int maxk_spgemm_forward_fused(const maxk_peer_ctx *ctx, uint32_t epoch, const int32_t *warp4,
                              int64_t num_tasks, int64_t wait_from_task, const int32_t *indices,
                              const float *values, float *out, int64_t num_rows_out, int32_t d,
                              int32_t copy_ctas, void *stream);
/* The same kernel with the bank plan of cbsr_plan.cu travelling in place of the selectors:
 * ctx->g_index[r] are uint16[n_total, k] buffers holding sp_addr.  k = 32, 64.                  */
MAXK_API int maxk_spgemm_forward_fused_planned(const maxk_peer_ctx *ctx, uint32_t epoch,
                                      const int32_t *warp4, int64_t num_tasks, int64_t wait_from_task,
                                      const int32_t *indices, const float *values, float *out,
                                      int64_t num_rows_out, int32_t d, int32_t copy_ctas, void *stream);
/* Before the backward SSpMM (called with accumulate != 0 on ctx->partial[rank]): zero the rows
 * of this rank's partial buffer that its edges can touch -- its own block, and per peer the
 * whole block (dense pair) or the recv_rows list (halo pair). */
MAXK_API int maxk_peer_zero_partial(const maxk_peer_ctx *ctx, void *stream);
/* Pull reduce-scatter: waits until every peer has published `epoch` in flags[.][1], then
 * out[i, :] = sum over ranks r of partial_r[row_start[rank] + i, :] (dense pairs: all rows, in
 * rank order; halo pairs: only the send_rows, added with reductions), out float32[rows_of[rank], k]. */
MAXK_API int maxk_reduce_scatter_pull(const maxk_peer_ctx *ctx, uint32_t epoch, float *out, void *stream);

/* ---- reference-named launch shims (ref: cuda_kernel_wrappers.cu:38-91, declared in
 * cuda_kernel_bindings.cpp:11-35).  Same argument lists, so the reference's binding
 * file links against this library unchanged.  grid/block/shared_size are accepted and
 * IGNORED (the library picks its own launch shape); the kernels run on the legacy
 * default stream like the reference's; vout is overwritten, not accumulated.  Errors are
 * printed (reference behaviour) and also retrievable through maxk_last_error().        */
MAXK_API void spmm_kernel_opt2_sparse_v3_wrapper(const int *warp4, const int *idx, const float *val,
                                        const float *vin_data, const uint8_t *vin_selector,
                                        float *vout, const int num_v, const int num_e,
                                        const int feat_in, const int dim_sparse,
                                        const int num_warps, maxk_dim3 grid, maxk_dim3 block,
                                        int shared_size);
MAXK_API void spmm_kernel_opt2_sparse_backward_v3_wrapper(const int *warp4, const int *idx,
                                                 const float *val, const float *vin_data,
                                                 const uint8_t *vin_selector, float *vout,
                                                 const int num_v, const int num_e,
                                                 const int feat_in, const int dim_sparse,
                                                 const int num_warps, maxk_dim3 grid,
                                                 maxk_dim3 block, int shared_size);
MAXK_API void topk_kernel_wrapper(uint8_t *data, uint8_t *value, uint8_t *index, unsigned int k, int N,
                         int dim_origin, maxk_dim3 grid, maxk_dim3 block, int shared_size);

#ifdef __cplusplus
}
#endif
#endif /* MAXK_B200_H_ */
