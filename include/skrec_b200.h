/*
 * skrec_b200.h -- C ABI of the B200-native full-ranking evaluation path.
 *
 * This is the drop-in boundary for scikit-recommender's native evaluation entry point and the
 * Python loop around it.  Each entry cites the reference interface it replaces (paths relative
 * to the reference checkout):
 *
 *   skr_eval_scores / skr_eval_scores_host
 *       replace  eval_score_matrix()            skrec/utils/py/cython/pyx_eval_matrix.pyx:22-37
 *       and      cpp_evaluate_matrix()          skrec/utils/py/cython/include/evaluate.h:57-76
 *       plus the train-item masking loop        skrec/utils/py/evaluator.py:195-200
 *   skr_eval_fused / skr_eval_fused_host
 *       replace one whole pass of RankingEvaluator.evaluate()   skrec/utils/py/evaluator.py:163-214
 *       for models whose predict() is U_b @ I^T (+ bias)         BPRMF.py:84-88, LightGCN.py:102-107,
 *                                                                MultVAE.py:138-141, SelfCF.py:235-241
 *   skr_metrics_from_topk
 *       replaces the metric functions           skrec/utils/py/cython/include/metric.h:19-118
 *   skr_set_train_csr / skr_set_test_csr
 *       replace RankingEvaluator.set_train_data / set_test_data  skrec/utils/py/evaluator.py:140-145
 *       (STL sets cannot cross a C ABI: `vector<unordered_set<int>>` becomes CSR)
 *   skr_colsum_f32_seq
 *       replaces np.mean(all_results, axis=0)   skrec/utils/py/evaluator.py:206-208 (float32, row order)
 *
 * Conventions
 *   - plain pointers and sizes only; no C++/torch types; 0 = ok, negative = error; nothing throws.
 *   - the caller owns every buffer it passes; a ctx owns only its device copies of the CSRs and
 *     its workspace.  One ctx per (device, stream user); calls on one ctx must not overlap.
 *   - `*_dev` pointers are device memory on the ctx's device; `*_host` are host memory
 *     (pinned or pageable).  `stream` is a cudaStream_t passed as void* (NULL = legacy default).
 *   - device entry points are asynchronous on `stream`; `*_host` entry points return after
 *     their results are in host memory.
 *   - rows: the CSRs passed to skr_set_*_csr have one row per EVALUATED user, in evaluation
 *     order (evaluator.py:181-184).  A call covers rows [row0, row0 + n_rows).
 *   - per-user output layout is the reference's: [n_rows, n_metrics * top_k], metric-major,
 *     K ascending (evaluate.h:47-51).  metric ids: 1 Precision, 2 Recall, 3 MAP, 4 NDCG, 5 MRR
 *     (evaluator.py:57).
 *   - ties: score descending, then item id ascending (the reference's order on ties is an
 *     artefact of libstdc++'s heap; on tie-free rows both agree bit for bit).
 *   - `sums_dev` / `sums_host`: double[n_metrics * top_k]; device entry points ADD the column
 *     sums over the rows of the call into it (zero it before the first batch).
 */
#ifndef SKREC_B200_H
#define SKREC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SKR_ABI_VERSION 1

#define SKR_OK 0
#define SKR_ERR_INVALID (-1)     /* bad argument (message in skr_last_error) */
#define SKR_ERR_CUDA (-2)        /* a CUDA runtime/driver call failed */
#define SKR_ERR_NOMEM (-3)       /* device or host allocation failed */
#define SKR_ERR_UNSUPPORTED (-4) /* shape outside what the chosen kernel supports */
#define SKR_ERR_STATE (-5)       /* call order (e.g. no test CSR set) */

/* scoring arithmetic of the fused path */
#define SKR_PREC_AUTO 0   /* tensor cores when the shape allows (3xTF32; for large catalogues TF32R at d <= 64, F16R above), else FP32 FMA */
#define SKR_PREC_FP32 1   /* FP32 FMA on CUDA cores (exact products, sequential k order) */
#define SKR_PREC_3XTF32 2 /* tcgen05 kind::tf32, hi/lo split, 3 products, FP32 accumulate */
#define SKR_PREC_1XTF32 3 /* single TF32 pass; NOT reference-grade, for measurement only */
#define SKR_PREC_TF32R 4  /* single TF32 pass to find candidates inside a rigorous error band, then exact FP32
                           * re-scoring of the survivors: results equal SKR_PREC_FP32 bit for bit */
#define SKR_PREC_F16R 5   /* the same with FP16 operands (tcgen05 kind::f16, twice the TF32 rate, half the item-panel bytes):
                           * tables scaled by powers of two into fp16's range, the same 2^-11 operand rounding as TF32,
                           * exact FP32 re-scoring of the survivors: results equal SKR_PREC_FP32 bit for bit */

typedef struct skr_ctx skr_ctx;

int skr_abi_version(void);

/* Context bound to CUDA device `device`.  Fails (SKR_ERR_CUDA) when no such device exists. */
int skr_ctx_create(int device, skr_ctx **out);
int skr_ctx_destroy(skr_ctx *ctx);
/* Message of the last failure on `ctx` (or of the last failed skr_ctx_create when ctx is NULL). */
const char *skr_last_error(const skr_ctx *ctx);

/* Train interactions to mask (evaluator.py:140-141).  Host CSR; rows need not be sorted and may
 * hold duplicates.  indptr == NULL clears masking (user_train_dict=None). */
int skr_set_train_csr(skr_ctx *ctx, const int64_t *indptr_host, const int32_t *indices_host,
                      int64_t n_rows, int64_t n_items);
/* Test interactions (evaluator.py:143-145); duplicates are dropped like the reference's set. */
int skr_set_test_csr(skr_ctx *ctx, const int64_t *indptr_host, const int32_t *indices_host,
                     int64_t n_rows, int64_t n_items);

/* Score-matrix-in: mask train items, select top-K, per-user metrics, column sums.
 * scores_dev: float32 [n_rows, ld], not modified.  Any of the four outputs may be NULL.
 * topk_idx_dev int32 [n_rows, top_k]; topk_val_dev float32 [n_rows, top_k];
 * per_user_dev float32 [n_rows, n_metrics*top_k]. */
int skr_eval_scores(skr_ctx *ctx, const float *scores_dev, int64_t n_rows, int64_t n_items, int64_t ld,
                    int64_t row0, const int32_t *metric_ids, int n_metrics, int top_k,
                    int32_t *topk_idx_dev, float *topk_val_dev, float *per_user_dev, double *sums_dev,
                    void *stream);

/* Same with a HOST score matrix and host outputs; the H2D/D2H copies are part of the call.
 * sums_host is ADDED to as well.  (What a Cython/ctypes binding of eval_score_matrix calls.) */
int skr_eval_scores_host(skr_ctx *ctx, const float *scores_host, int64_t n_rows, int64_t n_items, int64_t ld,
                         int64_t row0, const int32_t *metric_ids, int n_metrics, int top_k,
                         int32_t *topk_idx_host, float *per_user_host, double *sums_host, void *stream);

/* Fused: scores = user_vecs @ item_vecs^T (+ bias) are produced tile by tile on chip and never
 * written to HBM; masking, top-K, metrics and sums as above.
 * user_vecs_dev float32 [n_rows, ld_u] (row r = evaluated user row0 + r), item_vecs_dev float32
 * [n_items, ld_i], bias_dev float32 [n_items] or NULL, d = embedding width. */
int skr_eval_fused(skr_ctx *ctx, const float *user_vecs_dev, int64_t n_rows, int64_t ld_u,
                   const float *item_vecs_dev, int64_t n_items, int64_t ld_i, int d,
                   const float *bias_dev, int64_t row0, const int32_t *metric_ids, int n_metrics,
                   int top_k, int precision, int32_t *topk_idx_dev, float *topk_val_dev,
                   float *per_user_dev, double *sums_dev, void *stream);

/* Top-k of every row of a score block without masks or metrics: the GPU form of the reference's
 * pyx_arg_top_k / pyx_top_k (skrec/utils/py/cython/pyx_sort.pyx:104-187, sort.h:136-170).  Indices (and/or
 * values) sorted by score desc, ties by lower index; either output may be NULL.  top_k <= 512. */
int skr_topk_scores(skr_ctx *ctx, const float *scores_dev, int64_t n_rows, int64_t n_items, int64_t ld,
                    int top_k, int32_t *topk_idx_dev, float *topk_val_dev, void *stream);
int skr_topk_scores_host(skr_ctx *ctx, const float *scores_host, int64_t n_rows, int64_t n_items, int64_t ld,
                         int top_k, int32_t *topk_idx_host, float *topk_val_host, void *stream);

/* Grouped evaluation (base.py:66-71 evaluate_group runs one full evaluation per user group): after ONE
 * evaluation that kept the per-user block [*, n_cols] on the device, sums_dev[c] += sum of the rows
 * row_list_dev[0 .. n_list) -- float64, deterministic. */
int skr_colsum_rows(skr_ctx *ctx, const float *per_user_dev, int64_t n_cols, const int32_t *row_list_dev,
                    int64_t n_list, double *sums_dev, void *stream);

/* Item-sharded evaluation (catalogue beyond one HBM): the reference has no counterpart -- its
 * `predict` always returns all columns (base.py:73).  Step 1, on every rank: the rows' sorted top-K over the
 * item rows [item_offset, item_offset + n_items) this rank holds, as 64-bit rank keys
 * (ord(score) << 32 | ~global_item, larger = ranked earlier) in keys_out_dev [n_rows, top_k].  The train CSR
 * of ctx must be the column partition of that shard with shard-local item ids. */
int skr_topk_fused(skr_ctx *ctx, const float *user_vecs_dev, int64_t n_rows, int64_t ld_u,
                   const float *item_vecs_dev, int64_t n_items, int64_t ld_i, int d,
                   const float *bias_dev, int64_t row0, int64_t item_offset, int top_k, int precision,
                   uint64_t *keys_out_dev, void *stream);

/* Step 2, after the ranks exchanged their lists (NCCL all-gather): keys_all_dev is [n_shards, n_rows_total,
 * top_k]; merges the n_shards lists of rows [row_begin, row_begin + n_rows) -- exact, the shards are
 * disjoint -- and evaluates them against test CSR rows row0 .. row0 + n_rows like skr_eval_scores.  Needs
 * n_shards * top_k <= 1024. */
int skr_eval_merged_topk(skr_ctx *ctx, const uint64_t *keys_all_dev, int n_shards, int64_t n_rows_total,
                         int64_t row_begin, int64_t n_rows, int64_t row0, const int32_t *metric_ids,
                         int n_metrics, int top_k, int32_t *topk_idx_dev, float *topk_val_dev,
                         float *per_user_dev, double *sums_dev, void *stream);

/* Same with HOST embedding tables and host outputs (copies inside the call). */
int skr_eval_fused_host(skr_ctx *ctx, const float *user_vecs_host, int64_t n_rows, int64_t ld_u,
                        const float *item_vecs_host, int64_t n_items, int64_t ld_i, int d,
                        const float *bias_host, int64_t row0, const int32_t *metric_ids, int n_metrics,
                        int top_k, int precision, int32_t *topk_idx_host, float *per_user_host,
                        double *sums_host, void *stream);

/* Metrics of given rank lists (metric.h:19-118): topk_idx_dev int32 [n_rows, top_k]. */
int skr_metrics_from_topk(skr_ctx *ctx, const int32_t *topk_idx_dev, int64_t n_rows, int64_t row0,
                          const int32_t *metric_ids, int n_metrics, int top_k, float *per_user_dev,
                          double *sums_dev, void *stream);

/* acc[c] = (...((acc[c] + x[0,c]) + x[1,c]) + ...) in float32, rows in order: numpy's
 * np.sum(axis=0) on a C-ordered float32 array (evaluator.py:208).  acc_dev float32 [n_cols]. */
int skr_colsum_f32_seq(skr_ctx *ctx, const float *per_user_dev, int64_t n_rows, int64_t n_cols,
                       float *acc_dev, void *stream);

/* Number of kernels this library has launched on `ctx` since creation (bench.py's gpu_launches). */
int64_t skr_launch_count(const skr_ctx *ctx);

/* Name of the scoring kernel the last skr_eval_fused* call on ctx used: "tcgen05_3xtf32", "tcgen05_tf32r", "tcgen05_f16r",
 * "tcgen05_1xtf32", "simt_fp32" or "simt_fp32_blocks" (top_k > 128: score blocks + the score-matrix kernels). */
const char *skr_last_fused_kernel(const skr_ctx *ctx);

/* Device time, in milliseconds, of the scoring kernel launched by a recent skr_eval_fused* call on
 * ctx: back = 0 is the last call, 1 the one before, ... (CUDA events recorded on the call's stream
 * around that one launch; waits for it).  The ring holds skr_set_option("event_ring", n) calls. */
int skr_fused_kernel_ms(skr_ctx *ctx, int back, float *ms_out);
/* Same for the sampled threshold pre-pass that precedes the main scoring kernel (0 for the FP32 path). */
int skr_fused_prepass_ms(skr_ctx *ctx, int back, float *ms_out);
/* Plan and outcome of the last tcgen05 skr_eval_fused* call (its last row chunk): out[0..7) = sample tiles, sample
 * stride, threshold rank r, sub-list capacity, item chunks, TMA stages, rows re-done by the exact kernel; with n_out >= 8
 * out[7] = scoring launches timed since the last "event_ring" option; with n_out >= 9 out[8] = rows the single-pass
 * attempt of tf32r left unsettled (retried in three passes or handed to the exact kernel).  Synchronises the device. */
int skr_fused_stats(skr_ctx *ctx, int64_t *out, int n_out);
/* Host only -- no CUDA call, usable without a GPU: the work list the tcgen05 main pass runs for n_user_tiles x
 * n_item_tiles tiles (128 users x 128 items each) on n_sm SMs.  A work item is (user tile, first item tile, number of
 * item tiles, chunk index of that user tile); items come largest first (the block scheduler then runs
 * longest-processing-time-first), user tiles get `slots` or `slots - 1` chunks so that the CTA count fills whole waves.
 * cta_overhead: fixed cost of a CTA in tile times; chunks: 0 = automatic, else the "chunks" option.
 * items_out (nullable): int32 [max_items][4]; info_out (nullable): [5] = slots, min_slots, largest item, mixed, simulated
 * makespan in tile times.  Returns the number of work items, or a negative SKR_ERR_*. */
int64_t skr_plan_work_host(int n_user_tiles, int n_item_tiles, int n_sm, int cta_overhead, int chunks, int32_t *items_out,
                           int64_t max_items, int64_t *info_out);
/* Development aid: with skr_set_option("trace_cta", c >= 0) the main tcgen05 pass records, for CTA c,
 * SM-clock timestamps of its pipeline events per item tile (16 slots per tile, see k_fused_tc.cuh);
 * this copies up to n_out of them to the host.  Synchronises the device. */
int skr_fused_trace(skr_ctx *ctx, int64_t *out, int64_t n_out);

/* Tunables: "chunks" (item-range chunks per user tile of the fused path, 0 = automatic), "stages"
 * (ignored: the ring depth is fixed by the kernel instantiation), "sample_tiles" / "rank" (pre-pass size and threshold rank, 0 =
 * automatic), "event_ring" (see skr_fused_kernel_ms), "trace_cta"
 * (see skr_fused_trace; -1 = off), "dbg" (timing ablations, results invalid), "chunk_rows" (rows per internal chunk of the
 * fused pipeline, 0 = 131,072), "retry_min" (tf32r: unsettled rows from which the three-pass retry runs, -1 = cost model),
 * "no_aug" (f16r: 1 = add the item bias in the epilogue instead of inside the contraction; measurements only),
 * "full_rescore" (tf32r / f16r: 1 = exact scores for every survivor even when no top-K list is requested; measurements only). */
int skr_set_option(skr_ctx *ctx, const char *name, int64_t value);

/* ---- negative sampler ----------------------------------------------------------------------------------------------
 * GPU form of c_batch_randint_choice (skrec/utils/py/cython/include/randint.h:97-128; pyx_random.pyx:79-150): element b
 * of the batch gets out[out_indptr[b] .. out_indptr[b + 1]) integers from [0, high): uniform (cdf_dev == NULL) or by
 * probabilities given as inclusive prefix sums (`cdf_dev` float [high], or [n_batch, high] with cdf_per_row != 0), with
 * (replace != 0) or without replacement, never a member of row b of the exclusion CSR (rows sorted and unique; NULL = none).
 * Every draw is a function of (seed, position, attempt) through Philox4x32-10: reproducible, independent of the launch
 * geometry; NOT the reference's mt19937 stream (sequential by construction).  All pointers are device memory; n_out =
 * out_indptr[n_batch].  Asynchronous on `stream`; a row that cannot be filled is reported by skr_check. */
int skr_batch_randint(skr_ctx *ctx, int64_t high, const int64_t *out_indptr_dev, int64_t n_batch, int64_t n_out, int replace,
                      const float *cdf_dev, int cdf_per_row, const int64_t *excl_indptr_dev, const int32_t *excl_idx_dev,
                      uint64_t seed, int32_t *out_dev, void *stream);
/* 0, or SKR_ERR_CUDA once a kernel of ctx reported a condition it could not handle; synchronises, clears the flag. */
int skr_check(skr_ctx *ctx);

/* ---- one-shot all-reduce of the metric sums over NVLink peer memory ------------------------------------------------
 * User-sharded evaluation exchanges n_metrics * top_k + 1 doubles per evaluate; the reference has no
 * counterpart (its mean runs over one process's users, evaluator.py:206-208).  One process per GPU of ONE node:
 *   skr_comm_create   allocates this rank's inbox on `device`;
 *   skr_comm_handle   -> 64 bytes (a cudaIpcMemHandle_t) to hand to the other ranks by any host channel
 *                        (torch.distributed all_gather_object in the Python host);
 *   skr_comm_connect  maps the peers' inboxes (handles: world x 64 bytes in rank order);
 *   skr_comm_allreduce in-place SUM over the ranks of vec_dev[0 .. n), n <= 4096, asynchronous on `stream`: one kernel
 *                        that writes the vector into every peer's inbox, publishes a flag, waits (bounded, ~3 s) for the
 *                        peers' flags and adds the vectors in rank order -- every rank ends with identical bits;
 *   skr_comm_status   0 unless a rank failed to arrive inside the time limit in some earlier call. */
typedef struct skr_comm skr_comm;
int skr_comm_create(int device, int rank, int world, skr_comm **out);
int skr_comm_handle(skr_comm *comm, void *handle_out64);
int skr_comm_connect(skr_comm *comm, const void *handles);
int skr_comm_allreduce(skr_comm *comm, double *vec_dev, int n, void *stream);
int skr_comm_status(skr_comm *comm);
const char *skr_comm_last_error(const skr_comm *comm);
int skr_comm_destroy(skr_comm *comm);

#ifdef __cplusplus
}
#endif
#endif /* SKREC_B200_H */
