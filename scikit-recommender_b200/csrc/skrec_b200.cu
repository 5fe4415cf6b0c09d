// skrec_b200.cu -- C ABI (include/skrec_b200.h) over the sm_100a kernels.
//
// Host side of the hot path: context, CSR normalisation + upload (the device form of
// RankingEvaluator.set_train_data / set_test_data, evaluator.py:140-145), workspace, kernel
// selection and launch.  No torch types, no CPU compute fallback: every entry point either
// launches CUDA kernels or fails with an error code.
#include <cuda.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <functional>
#include <new>
#include <string>
#include <vector>

#include "../../include/skrec_b200.h"
#include "common.cuh"
#include "fused_common.cuh"
#include "ingest.h"
#include "k_fused_simt.cuh"
#include "k_fused_tc.cuh"
#include "k_metrics.cuh"
#include "k_sampler.cuh"
#include "k_scores.cuh"
#include "k_select.cuh"

using namespace skr;

namespace {

struct Buf {
    void *p = nullptr;
    size_t cap = 0;
};

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

std::string g_create_error;

}  // namespace

struct skr_ctx {
    int device = 0;
    int n_sm = 148;
    size_t max_smem = 0;
    std::string err;
    // train CSR (sorted, unique) + fused-path mask keys
    bool has_train = false;
    int64_t tr_rows = 0, tr_items = 0;
    int64_t *d_tr_indptr = nullptr;
    int32_t *d_tr_idx = nullptr;
    uint32_t *d_mask_keys = nullptr;
    int64_t *d_mask_tile_ptr = nullptr;
    uint32_t *d_mask_tile_off = nullptr;
    // test CSR (sorted, unique)
    bool has_test = false;
    int64_t te_rows = 0, te_items = 0;
    int64_t *d_te_indptr = nullptr;
    int32_t *d_te_idx = nullptr;
    // 1/log2(i+2) table (metric.h:78,82), host glibc
    double *d_disc = nullptr;
    float *d_idcg = nullptr;  // iDCG after n terms (metric.h:82), host-accumulated
    int disc_n = 0;
    // workspace, grow-only
    Buf trace, stats, stage_s, eps2, rs_keys, rs_cnt, keys, per_user, partial, part, thr, bhi, blo, bias, sums, stage_a, stage_b, stage_c, out_idx, samp, cand, cand_cnt, fail_list;
    Buf bh16, f16s, rscale, baug;  // precision "f16r": scaled fp16 item table; {max |item| bits, s_i, max |bias| bits, m}; per-row s_u s_i | g; augmentation table
    int *d_err = nullptr;
    double *h_pin = nullptr;  // pinned host staging for the small results ([sums | watchdog flag]), SKR_PIN_DOUBLES doubles
    int64_t launches = 0;
    const char *last_fused = "none";
    int64_t opt_chunks = 0;
    int64_t opt_stages = 0;
    int64_t opt_sample_tiles = 0;
    int64_t opt_rank = 0;
    int64_t opt_dbg = 0;
    int64_t opt_chunk_rows = 0;      // rows per chunk of the fused pipeline (0 = default)
    int64_t opt_score_fn = 0;        // 0: u . i + b   1: -||u - i|| + b (FP32 tile kernels only)
    int64_t opt_retry_min = -1;      // tf32r: unsettled rows from which the three-pass retry runs (-1 = cost model)
    int64_t opt_no_aug = 0;          // f16r: 1 = add the bias in the epilogue instead of inside the contraction (A/B measurements)
    int64_t opt_full_rescore = 0;    // tf32r / f16r: 1 = re-score every survivor and sort even when only metrics are wanted (A/B measurements)
    int64_t opt_exact_seg_rows = -1;  // development: how many failed rows are cut in segments (-1 = default)
    bool debug_sync = false;  // SKR_DEBUG_SYNC=1: synchronise and check after every kernel, naming the one that failed
    int64_t opt_trace_cta = -1;  // >= 0: record the tile timeline of that CTA of the main pass (development aid)
    struct Plan { int n_samp, stride, r, cap, S, stages; } last_plan = {0, 0, 0, 0, 0, 0};
    // cached COLLECT work lists (see plan_work): a row-chunked evaluate alternates between the shape of the full
    // chunks and the shape of the last one, and building a list costs a host simulation plus a synchronous upload
    struct WorkCache {
        int key[4] = {-1, -1, -1, -1};
        Buf buf;
        int ctas = 0, slots = 0, min_slots = 0, max_tiles = 0;
        bool mixed = false;
        int64_t used = 0;
    } work_cache[4];
    int64_t work_clock = 0;
    uint32_t *stats_cur_ptr = nullptr;  // item statistics slot the current evaluate reads (chunks after the first reuse it)
    EncodeTiledFn encode = nullptr;
    std::vector<cudaEvent_t> ev0, ev1, ev2;  // ring: ev2 before the pre-pass, ev0/ev1 around the main scoring kernel
    int64_t ev_calls = 0;
    int stats_slot = 0;  // which half of `stats` the next evaluate fills (k_split_tf32 resets the other)
};

namespace {

int fail(skr_ctx *ctx, int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf; else g_create_error = buf;
    return code;
}

#define SKR_AFTER(ctx, st, name)                                                                     \
    do {                                                                                             \
        if ((ctx)->debug_sync) {                                                                     \
            cudaError_t e__ = cudaStreamSynchronize(st);                                             \
            if (e__ == cudaSuccess) e__ = cudaGetLastError();                                        \
            if (e__ != cudaSuccess) return fail(ctx, SKR_ERR_CUDA, "after %s: %s", name, cudaGetErrorString(e__)); \
        }                                                                                            \
    } while (0)

#define SKR_CUDA(ctx, call)                                                                          \
    do {                                                                                             \
        cudaError_t e__ = (call);                                                                    \
        if (e__ != cudaSuccess)                                                                      \
            return fail(ctx, e__ == cudaErrorMemoryAllocation ? SKR_ERR_NOMEM : SKR_ERR_CUDA, "%s: %s", #call, \
                        cudaGetErrorString(e__));                                                    \
    } while (0)

// launch with programmatic stream serialisation (see pdl_wait in common.cuh)
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args)
{
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

constexpr int SKR_PIN_DOUBLES = 2048;

int ensure(skr_ctx *ctx, Buf &b, size_t bytes)
{
    if (bytes <= b.cap) return SKR_OK;
    if (b.p) SKR_CUDA(ctx, cudaFree(b.p));
    b.p = nullptr;
    b.cap = 0;
    size_t want = bytes + bytes / 8 + 256;
    SKR_CUDA(ctx, cudaMalloc(&b.p, want));
    b.cap = want;
    return SKR_OK;
}

void free_dev(void *p)
{
    if (p) cudaFree(p);
}

// score-matrix rows -> sorted top-K rank keys: one warp per row for top-K <= 128, one block per row above
void launch_topk_scores(const float *scores, int64_t ld, int n_items, int64_t n_rows, int64_t row0, const int64_t *tr_indptr,
                               const int32_t *tr_idx, int K, u64 *keys, int *err_flag, cudaStream_t st)
{
    if (K <= 128) {
        static bool attr_set = false;
        if (!attr_set) {
            cudaFuncSetAttribute(k_topk_rows<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kw_smem(8));
            cudaFuncSetAttribute(k_topk_rows<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kw_smem(16));
            attr_set = true;
        }
        const unsigned g = (unsigned)((n_rows + KW_WARPS - 1) / KW_WARPS);
        // staging area of 256 entries for K <= 64, 512 for K <= 128 (room for 128 new survivors next to the K kept ones)
        if (K <= 64) k_topk_rows<8><<<g, KW_WARPS * 32, kw_smem(8), st>>>(scores, ld, n_items, n_rows, row0, tr_indptr, tr_idx, K, keys, err_flag);
        else k_topk_rows<16><<<g, KW_WARPS * 32, kw_smem(16), st>>>(scores, ld, n_items, n_rows, row0, tr_indptr, tr_idx, K, keys, err_flag);
    } else
        k_topk_scores<<<(unsigned)n_rows, K2_THREADS, 0, st>>>(scores, ld, n_items, row0, tr_indptr, tr_idx, K, keys);
}

int check_metrics(skr_ctx *ctx, const int32_t *metric_ids, int n_metrics, int top_k, MetricIds &m)
{
    if (!metric_ids || n_metrics < 1 || n_metrics > 8) return fail(ctx, SKR_ERR_INVALID, "n_metrics=%d not in [1,8]", n_metrics);
    if (top_k < 1) return fail(ctx, SKR_ERR_INVALID, "top_k=%d", top_k);
    m.n = n_metrics;
    m.packed = 0u;
    for (int i = 0; i < n_metrics; ++i) {
        if (metric_ids[i] < 1 || metric_ids[i] > 5) return fail(ctx, SKR_ERR_INVALID, "metric id %d not in 1..5", metric_ids[i]);
        m.packed |= (uint32_t)metric_ids[i] << (4 * i);
    }
    return SKR_OK;
}

int ensure_disc(skr_ctx *ctx, int K)
{
    if (K <= ctx->disc_n) return SKR_OK;
    int n = std::max(K, 1024);
    std::vector<double> h((size_t)n);
    std::vector<float> ig((size_t)n + 1);
    float idcg = 0.0f;
    ig[0] = 0.0f;
    for (int i = 0; i < n; ++i) {
        h[(size_t)i] = 1.0 / log2((double)(unsigned)(i + 2));  // metric.h:78
        idcg = (float)((double)idcg + h[(size_t)i]);           // metric.h:82, after i + 1 terms
        ig[(size_t)i + 1] = idcg;
    }
    free_dev(ctx->d_disc);
    free_dev(ctx->d_idcg);
    ctx->d_disc = nullptr;
    ctx->d_idcg = nullptr;
    SKR_CUDA(ctx, cudaMalloc((void **)&ctx->d_disc, sizeof(double) * (size_t)n));
    SKR_CUDA(ctx, cudaMalloc((void **)&ctx->d_idcg, sizeof(float) * ((size_t)n + 1)));
    SKR_CUDA(ctx, cudaMemcpy(ctx->d_disc, h.data(), sizeof(double) * (size_t)n, cudaMemcpyHostToDevice));
    SKR_CUDA(ctx, cudaMemcpy(ctx->d_idcg, ig.data(), sizeof(float) * ((size_t)n + 1), cudaMemcpyHostToDevice));
    ctx->disc_n = n;
    return SKR_OK;
}

// keys [n_rows, K] sorted -> outputs.  Shared tail of every evaluation entry point.
int run_metrics(skr_ctx *ctx, const u64 *keys, const int32_t *idx_in, int64_t n_rows, int64_t row0, const MetricIds &m,
                int K, int32_t *topk_idx, float *topk_val, float *per_user, double *sums, cudaStream_t st)
{
    if (!ctx->has_test) return fail(ctx, SKR_ERR_STATE, "no test CSR set (skr_set_test_csr)");
    if (row0 < 0 || row0 + n_rows > ctx->te_rows)
        return fail(ctx, SKR_ERR_INVALID, "rows [%lld,%lld) outside the test CSR (%lld rows)", (long long)row0, (long long)(row0 + n_rows), (long long)ctx->te_rows);
    int rc = ensure_disc(ctx, K);
    if (rc) return rc;
    const int MK = m.n * K;
    // column sums ride along in k_metrics (per-warp float64 accumulators in shared memory) when they fit;
    // otherwise the per-user block is materialised and summed by the two-stage kernels
    const size_t acc_smem = (size_t)K4_WARPS * MK * sizeof(double);
    const bool fused_sums = sums != nullptr && acc_smem <= 96 * 1024;
    float *pu = per_user;
    if (!pu && sums && !fused_sums) {
        rc = ensure(ctx, ctx->per_user, (size_t)n_rows * MK * sizeof(float));
        if (rc) return rc;
        pu = (float *)ctx->per_user.p;
    }
    const int grid = (int)std::min<int64_t>((n_rows + K4_WARPS - 1) / K4_WARPS, 6 * ctx->n_sm);
    double *acc = nullptr;
    if (fused_sums) {
        rc = ensure(ctx, ctx->partial, (size_t)grid * MK * sizeof(double));
        if (rc) return rc;
        acc = (double *)ctx->partial.p;
        if (acc_smem > 48 * 1024) SKR_CUDA(ctx, cudaFuncSetAttribute(k_metrics, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)acc_smem));
    }
    SKR_CUDA(ctx, launch_pdl(k_metrics, dim3((unsigned)(grid)), dim3((unsigned)(K4_WARPS * 32)), (size_t)(fused_sums ? acc_smem : 0), st, keys, idx_in, K, n_rows, row0, nullptr, nullptr, ctx->d_te_indptr, ctx->d_te_idx,
                                                                    m, ctx->d_disc, ctx->d_idcg, pu, topk_idx, topk_val, acc));
    ctx->launches++;
    if (fused_sums) {
        SKR_CUDA(ctx, launch_pdl(k_colsum_fold, dim3((unsigned)(MK)), dim3((unsigned)(256)), (size_t)(0), st, acc, grid, MK, sums));
        ctx->launches++;
    } else if (sums) {
        const int nblk = (int)std::min<int64_t>(n_rows, 2 * ctx->n_sm);
        rc = ensure(ctx, ctx->partial, (size_t)nblk * MK * sizeof(double));
        if (rc) return rc;
        k_colsum_partial<<<nblk, 256, 0, st>>>(pu, n_rows, MK, (double *)ctx->partial.p);
        k_colsum_final<<<(MK + 127) / 128, 128, 0, st>>>((const double *)ctx->partial.p, nblk, MK, sums);
        ctx->launches += 2;
    }
    SKR_CUDA(ctx, cudaGetLastError());
    return SKR_OK;
}

// Tail of the tensor-core path: candidate lists -> (select + sort + metrics in one kernel), the rows that
// kernel could not settle -> exact re-scoring -> metrics of just those rows, then the column sums.
struct ExactArgs {
    const float *U; int64_t ld_u; const float *V; int64_t ld_v; int d; const float *bias; int n_items;
    const int64_t *tr_indptr; const int32_t *tr_idx;
};

// exact re-scoring of the rows on the fail list: (row, item segment) work items, then a merge per row
int run_row_exact(skr_ctx *ctx, const ExactArgs &E, int32_t *fail_list, int *fail_count, int64_t n_rows, int64_t row0, int K, u64 *keys_out,
                  cudaStream_t st)
{
    // as many item segments per row as the merge takes (4,096 keys): a segment is walked by ONE block, and a walk is
    // latency-bound (3.6 ms for a tenth of a 1M-item catalogue at d = 128)
    int n_seg = std::max(1, std::min(64, 4096 / K));
    n_seg = (int)std::max<int64_t>(1, std::min<int64_t>(n_seg, ((int64_t)E.n_items + 4095) / 4096));  // at least four chunks of work per segment
    int seg_items = (int)(((int64_t)E.n_items + n_seg - 1) / n_seg);
    seg_items = ((seg_items + K2_CHUNK - 1) / K2_CHUNK) * K2_CHUNK;  // ranges start on chunk boundaries (vector loads)
    n_seg = (E.n_items + seg_items - 1) / seg_items;
    // room for the partial lists of as many failed rows as are plausible; beyond that the rows are walked whole
    const int64_t max_fail = std::min<int64_t>(n_rows, ctx->opt_exact_seg_rows >= 0 ? ctx->opt_exact_seg_rows : 256);
    int rc = ensure(ctx, ctx->part, (size_t)max_fail * n_seg * K * sizeof(u64));
    if (rc) return rc;
    SKR_CUDA(ctx, launch_pdl(k_row_exact, dim3((unsigned)std::max(1, 4 * ctx->n_sm / n_seg), (unsigned)n_seg), dim3((unsigned)(K2_THREADS)), (size_t)(0), st, fail_list, fail_count, E.U, E.ld_u, E.V, E.ld_v, E.d, E.bias, E.n_items, row0,
                                                                  E.tr_indptr, E.tr_idx, K, n_seg, seg_items, (int)max_fail, (u64 *)ctx->part.p, keys_out));
    SKR_AFTER(ctx, st, "k_row_exact");
    const int n = n_seg * K;
    const unsigned g = (unsigned)ctx->n_sm;
    if (n <= 64) SKR_CUDA(ctx, launch_pdl(k_merge_fail<2>, dim3((unsigned)(g)), dim3((unsigned)(128)), (size_t)(0), st, fail_list, fail_count, (const u64 *)ctx->part.p, n_seg, K, (int)max_fail, keys_out));
    else if (n <= 128) SKR_CUDA(ctx, launch_pdl(k_merge_fail<4>, dim3((unsigned)(g)), dim3((unsigned)(128)), (size_t)(0), st, fail_list, fail_count, (const u64 *)ctx->part.p, n_seg, K, (int)max_fail, keys_out));
    else if (n <= 256) SKR_CUDA(ctx, launch_pdl(k_merge_fail<8>, dim3((unsigned)(g)), dim3((unsigned)(128)), (size_t)(0), st, fail_list, fail_count, (const u64 *)ctx->part.p, n_seg, K, (int)max_fail, keys_out));
    else if (n <= 512) SKR_CUDA(ctx, launch_pdl(k_merge_fail<16>, dim3((unsigned)(g)), dim3((unsigned)(128)), (size_t)(0), st, fail_list, fail_count, (const u64 *)ctx->part.p, n_seg, K, (int)max_fail, keys_out));
    else if (n <= 1024) SKR_CUDA(ctx, launch_pdl(k_merge_fail<32>, dim3((unsigned)(g)), dim3((unsigned)(128)), (size_t)(0), st, fail_list, fail_count, (const u64 *)ctx->part.p, n_seg, K, (int)max_fail, keys_out));
    else if (n <= 2048) SKR_CUDA(ctx, launch_pdl(k_merge_fail<64>, dim3((unsigned)(g)), dim3((unsigned)(128)), (size_t)(0), st, fail_list, fail_count, (const u64 *)ctx->part.p, n_seg, K, (int)max_fail, keys_out));
    else SKR_CUDA(ctx, launch_pdl(k_merge_fail<128>, dim3((unsigned)(g)), dim3((unsigned)(128)), (size_t)(0), st, fail_list, fail_count, (const u64 *)ctx->part.p, n_seg, K, (int)max_fail, keys_out));
    ctx->launches += 2;
    SKR_AFTER(ctx, st, "k_merge_fail");
    SKR_CUDA(ctx, cudaGetLastError());
    return SKR_OK;
}

// Second attempt of precision "tf32r" for the rows the first one left unsettled (sub-list overflow, band below the
// collection threshold, too many candidates inside the band -- what a few high-norm items or heavy-tailed tables do to
// the single-pass error band): the same pipeline in THREE TF32 passes, whose band is 10-20 times narrower, restricted to
// those rows.  collect() launches the lo-table build and the 3-pass COLLECT kernel (both return at once when nothing
// failed); the second selection looks only at rows with rs_cnt == 0 and puts what it cannot settle on fail list 2,
// which alone goes to the exact per-row kernel.
struct RetryPlan {
    const std::function<int()> *collect;  // null: no retry (3xTF32 / 1xTF32 requested explicitly)
    const float *eps2_3;                  // [n_rows] 2 eps of the three-pass scores
    const float *thr3;                    // [n_rows] threshold the retry collects with
    int32_t *fail_list2;
    int *fail_count2;
    int retry_min;                        // fewest unsettled rows for which the tile-wise retry beats the per-row exact kernel
};

int run_select_metrics(skr_ctx *ctx, const uint2 *cand, const uint32_t *cand_cnt, int n_sub, int cap, int64_t n_rows, int64_t row0,
                       const MetricIds &m, int K, const ExactArgs &E, int32_t *fail_list, int *fail_count, int32_t *topk_idx,
                       float *topk_val, float *per_user, double *sums, u64 *keys_only, RescoreArgs RA, const float *add_back,
                       const RetryPlan &RP, cudaStream_t st)
{
    const bool rescore = RA.U != nullptr;
    const int *fail_count_first = fail_count;
    // kernel instantiations: the sort capacity is 64 / 128 keys (K <= 64 / 128); re-scoring doubles it (error band) and
    // splits the job in two kernels: candidates -> exact keys of the survivors, then sort + metrics
    typedef void (*SelKernel)(const uint2 *, const uint32_t *, int, int, int, int, int64_t, int64_t, u64 *, int32_t *, int *, const int64_t *,
                              const int32_t *, MetricIds, const double *, const float *, float *, int32_t *, float *, double *, RescoreArgs, const float *, int,
                              const int *, int);
    typedef void (*SortKernel)(const u64 *, const int *, int, int64_t, int64_t, u64 *, const int64_t *, const int32_t *, MetricIds, const double *,
                               const float *, float *, int32_t *, float *, double *);
    const int per = rescore ? (K <= 64 ? 4 : 8) : (K <= 64 ? 2 : 4);
    // metrics without top-K lists: exact scores only for the test items among the survivors and their near-ties (k_select.cuh HITS)
    const bool hits = rescore && keys_only == nullptr && topk_idx == nullptr && topk_val == nullptr && ctx->opt_full_rescore == 0;
    SelKernel sel = hits ? (K <= 64 ? (SelKernel)k_select_cands<4, true, true> : (SelKernel)k_select_cands<8, true, true>)
                  : rescore ? (K <= 64 ? (SelKernel)k_select_cands<4, true> : (SelKernel)k_select_cands<8, true>)
                            : (K <= 64 ? (SelKernel)k_select_cands<2, false> : (SelKernel)k_select_cands<4, false>);
    SortKernel srt = (K <= 64) ? (SortKernel)k_sort_metrics<4> : (SortKernel)k_sort_metrics<8>;
    int rc;
    if (rescore) {
        if ((rc = ensure(ctx, ctx->rs_keys, (size_t)n_rows * 32 * per * sizeof(u64)))) return rc;
        if ((rc = ensure(ctx, ctx->rs_cnt, (size_t)n_rows * sizeof(int)))) return rc;
        RA.rs_keys = (u64 *)ctx->rs_keys.p;
        RA.rs_cnt = (int *)ctx->rs_cnt.p;
    }
    const int g_sel = (int)std::min<int64_t>((n_rows + SEL_WARPS - 1) / SEL_WARPS, 8 * ctx->n_sm);
    if (keys_only != nullptr) {  // per-shard lists: sorted keys out, no metrics
        SKR_CUDA(ctx, launch_pdl(sel, dim3((unsigned)(g_sel)), dim3((unsigned)(SEL_WARPS * 32)), (size_t)(0), st, cand, cand_cnt, n_sub, cap, cap, K, n_rows, row0, keys_only, fail_list, fail_count, nullptr, nullptr, m,
                                              nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, RA, add_back, 0, (const int *)nullptr, 0));
        if (rescore && RP.collect != nullptr) {
            if ((rc = (*RP.collect)())) return rc;
            RescoreArgs R2 = RA;
            R2.eps2 = RP.eps2_3;
            R2.thr_c = RP.thr3;
            SKR_CUDA(ctx, launch_pdl(sel, dim3((unsigned)(g_sel)), dim3((unsigned)(SEL_WARPS * 32)), (size_t)(0), st, cand, cand_cnt, n_sub, cap, cap, K, n_rows, row0, keys_only, RP.fail_list2, RP.fail_count2,
                                                  nullptr, nullptr, m, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, R2, nullptr, 1, (const int *)fail_count_first, RP.retry_min));
            ctx->launches += 3;
            fail_list = RP.fail_list2;
            fail_count = RP.fail_count2;
        }
        if (rescore) {
            SKR_CUDA(ctx, launch_pdl(srt, dim3((unsigned)(g_sel)), dim3((unsigned)(SEL_WARPS * 32)), (size_t)(0), st, RA.rs_keys, RA.rs_cnt, K, n_rows, row0, keys_only, nullptr, nullptr, m, nullptr, nullptr, nullptr, nullptr,
                                                  nullptr, nullptr));
            ctx->launches++;
        }
        if ((rc = run_row_exact(ctx, E, fail_list, fail_count, n_rows, row0, K, keys_only, st))) return rc;
        ctx->launches += 1;
        SKR_CUDA(ctx, cudaGetLastError());
        return SKR_OK;
    }
    if (!ctx->has_test) return fail(ctx, SKR_ERR_STATE, "no test CSR set (skr_set_test_csr)");
    if (row0 < 0 || row0 + n_rows > ctx->te_rows)
        return fail(ctx, SKR_ERR_INVALID, "rows [%lld,%lld) outside the test CSR (%lld rows)", (long long)row0, (long long)(row0 + n_rows), (long long)ctx->te_rows);
    if ((rc = ensure_disc(ctx, K))) return rc;
    const int MK = m.n * K;
    const size_t acc_sel = (size_t)SEL_WARPS * MK * sizeof(double), acc_k4 = (size_t)K4_WARPS * MK * sizeof(double);
    const bool fused_sums = sums != nullptr && acc_k4 <= 96 * 1024;
    float *pu = per_user;
    if (!pu && sums && !fused_sums) {
        if ((rc = ensure(ctx, ctx->per_user, (size_t)n_rows * MK * sizeof(float)))) return rc;
        pu = (float *)ctx->per_user.p;
    }
    if ((rc = ensure(ctx, ctx->keys, (size_t)n_rows * K * sizeof(u64)))) return rc;  // written only for re-done rows
    u64 *keys = (u64 *)ctx->keys.p;
    const int g_fix = (int)std::min<int64_t>((n_rows + K4_WARPS - 1) / K4_WARPS, ctx->n_sm);
    double *acc = nullptr;
    const int n_sel_parts = (hits && RP.collect != nullptr) ? 2 * g_sel : g_sel;  // HITS: the first selection and the retry's each hold sums
    if (fused_sums) {
        if ((rc = ensure(ctx, ctx->partial, (size_t)(n_sel_parts + g_fix) * MK * sizeof(double)))) return rc;
        acc = (double *)ctx->partial.p;
        if (acc_k4 > 48 * 1024) SKR_CUDA(ctx, cudaFuncSetAttribute(k_metrics, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)acc_k4));
    }
    // the kernel that holds the column sums: static (buffers) + dynamic (sums) may pass 48 KB, opt in for what is left
    const size_t dyn_sel = fused_sums ? acc_sel : 0;
    if (dyn_sel > 0) {
        cudaFuncAttributes fa;
        if (rescore && !hits) SKR_CUDA(ctx, cudaFuncGetAttributes(&fa, srt)); else SKR_CUDA(ctx, cudaFuncGetAttributes(&fa, sel));
        if (fa.sharedSizeBytes + dyn_sel > ctx->max_smem) return fail(ctx, SKR_ERR_UNSUPPORTED, "n_metrics * top_k = %d does not fit shared memory", MK);
        if (rescore && !hits) SKR_CUDA(ctx, cudaFuncSetAttribute(srt, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(ctx->max_smem - fa.sharedSizeBytes)));
        else SKR_CUDA(ctx, cudaFuncSetAttribute(sel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(ctx->max_smem - fa.sharedSizeBytes)));
    }
    if (hits) {
        SKR_CUDA(ctx, launch_pdl(sel, dim3((unsigned)(g_sel)), dim3((unsigned)(SEL_WARPS * 32)), (size_t)(dyn_sel), st, cand, cand_cnt, n_sub, cap, cap, K, n_rows, row0, nullptr, fail_list, fail_count, ctx->d_te_indptr,
                                 ctx->d_te_idx, m, ctx->d_disc, ctx->d_idcg, pu, nullptr, nullptr, acc, RA, add_back, 0, (const int *)nullptr, 0));
        if (RP.collect != nullptr) {
            if ((rc = (*RP.collect)())) return rc;
            RescoreArgs R2 = RA;
            R2.eps2 = RP.eps2_3;
            R2.thr_c = RP.thr3;
            SKR_CUDA(ctx, launch_pdl(sel, dim3((unsigned)(g_sel)), dim3((unsigned)(SEL_WARPS * 32)), (size_t)(dyn_sel), st, cand, cand_cnt, n_sub, cap, cap, K, n_rows, row0, nullptr, RP.fail_list2, RP.fail_count2,
                                     ctx->d_te_indptr, ctx->d_te_idx, m, ctx->d_disc, ctx->d_idcg, pu, nullptr, nullptr, acc ? acc + (size_t)g_sel * MK : nullptr, R2, nullptr, 1,
                                     (const int *)fail_count_first, RP.retry_min));
            ctx->launches += 3;
            fail_list = RP.fail_list2;
            fail_count = RP.fail_count2;
        }
    } else if (rescore) {
        SKR_CUDA(ctx, launch_pdl(sel, dim3((unsigned)(g_sel)), dim3((unsigned)(SEL_WARPS * 32)), (size_t)(0), st, cand, cand_cnt, n_sub, cap, cap, K, n_rows, row0, nullptr, fail_list, fail_count, nullptr, nullptr, m, nullptr,
                                              nullptr, nullptr, nullptr, nullptr, nullptr, RA, add_back, 0, (const int *)nullptr, 0));
        if (RP.collect != nullptr) {
            if ((rc = (*RP.collect)())) return rc;
            RescoreArgs R2 = RA;
            R2.eps2 = RP.eps2_3;
            R2.thr_c = RP.thr3;
            SKR_CUDA(ctx, launch_pdl(sel, dim3((unsigned)(g_sel)), dim3((unsigned)(SEL_WARPS * 32)), (size_t)(0), st, cand, cand_cnt, n_sub, cap, cap, K, n_rows, row0, nullptr, RP.fail_list2, RP.fail_count2,
                                                  nullptr, nullptr, m, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, R2, nullptr, 1, (const int *)fail_count_first, RP.retry_min));
            ctx->launches += 3;
            fail_list = RP.fail_list2;
            fail_count = RP.fail_count2;
        }
        SKR_CUDA(ctx, launch_pdl(srt, dim3((unsigned)(g_sel)), dim3((unsigned)(SEL_WARPS * 32)), (size_t)(dyn_sel), st, RA.rs_keys, RA.rs_cnt, K, n_rows, row0, nullptr, ctx->d_te_indptr, ctx->d_te_idx, m, ctx->d_disc,
                                                    ctx->d_idcg, pu, topk_idx, topk_val, acc));
        ctx->launches++;
    } else {
        SKR_CUDA(ctx, launch_pdl(sel, dim3((unsigned)(g_sel)), dim3((unsigned)(SEL_WARPS * 32)), (size_t)(dyn_sel), st, cand, cand_cnt, n_sub, cap, cap, K, n_rows, row0, nullptr, fail_list, fail_count, ctx->d_te_indptr,
                                                    ctx->d_te_idx, m, ctx->d_disc, ctx->d_idcg, pu, topk_idx, topk_val, acc, RA, add_back, 0, (const int *)nullptr, 0));
    }
    SKR_AFTER(ctx, st, "k_select_cands / k_sort_metrics");
    if ((rc = run_row_exact(ctx, E, fail_list, fail_count, n_rows, row0, K, keys, st))) return rc;
    SKR_CUDA(ctx, launch_pdl(k_metrics, dim3((unsigned)(g_fix)), dim3((unsigned)(K4_WARPS * 32)), (size_t)(fused_sums ? acc_k4 : 0), st, keys, nullptr, K, n_rows, row0, fail_list, fail_count, ctx->d_te_indptr,
                                                                     ctx->d_te_idx, m, ctx->d_disc, ctx->d_idcg, pu, topk_idx, topk_val,
                                                                     acc ? acc + (size_t)n_sel_parts * MK : nullptr));
    ctx->launches += 2;
    if (fused_sums) {
        SKR_CUDA(ctx, launch_pdl(k_colsum_fold, dim3((unsigned)(MK)), dim3((unsigned)(256)), (size_t)(0), st, acc, n_sel_parts + g_fix, MK, sums));
        ctx->launches++;
    } else if (sums) {
        const int nblk = (int)std::min<int64_t>(n_rows, 2 * ctx->n_sm);
        if ((rc = ensure(ctx, ctx->partial, (size_t)nblk * MK * sizeof(double)))) return rc;
        k_colsum_partial<<<nblk, 256, 0, st>>>(pu, n_rows, MK, (double *)ctx->partial.p);
        k_colsum_final<<<(MK + 127) / 128, 128, 0, st>>>((const double *)ctx->partial.p, nblk, MK, sums);
        ctx->launches += 2;
    }
    SKR_CUDA(ctx, cudaGetLastError());
    return SKR_OK;
}

// ---- work list of the tensor-core main pass ---------------------------------------------------------------
// A work item = (user tile, contiguous range of item tiles).  With the same number of chunks for every user tile
// the CTA count is rarely a multiple of the SM count (c2: 702 CTAs = 4.74 waves, the last wave a quarter empty).
// Here user tiles may get s or s + 1 chunks so that the total is k * n_sm, and the list is ordered largest chunk
// first: the block scheduler hands CTAs out in index order to whichever SM frees up, i.e. it runs the
// longest-processing-time-first heuristic for us.  The makespan of every candidate (k waves, or the plain
// uniform splits) is simulated with a fixed per-CTA overhead and the cheapest wins.
struct WorkPlan {
    std::vector<int4> items;
    int slots = 1, min_slots = 1, max_tiles = 0;
    bool mixed = false;
    long makespan = 0;
};

static long lpt_makespan(const std::vector<int4> &items, int n_sm, int overhead)
{
    std::vector<long> load((size_t)n_sm, 0);  // items are already sorted by decreasing size
    for (const int4 &w : items) {
        size_t best = 0;
        for (size_t j = 1; j < load.size(); ++j)
            if (load[j] < load[best]) best = j;
        load[best] += w.z + overhead;
    }
    return *std::max_element(load.begin(), load.end());
}

static WorkPlan make_work(int n_rt, int n_ct, int s_lo, int n_plus, int n_sm, int overhead)
{
    // n_plus user tiles get s_lo + 1 chunks, the others s_lo
    WorkPlan wp;
    wp.slots = s_lo + (n_plus > 0 ? 1 : 0);
    wp.min_slots = (n_plus >= n_rt) ? s_lo + 1 : s_lo;
    wp.mixed = n_plus > 0 && n_plus < n_rt;
    // Generated user-tile-major, which the stable sort below keeps among equal sizes: at c4 the 148 CTAs of a wave then
    // sweep 8 different item ranges.  Tried and not kept: chunk-major order (a wave sweeps ONE 62 MB range, L2-resident):
    // DRAM reads per 131,072-user chunk fall from 16.0 to 4.7 GB (L2 hit 87 -> 96 %), but 148 SMs asking the same L2 lines
    // at the same time cost more than the HBM traffic saves -- tensor pipe active 77 -> 70 %, kernel 45.5 -> 49 ms.
    // Sharing a panel between SMs wants TMA multicast inside a cluster, not coincidence in L2.
    for (int rt = 0; rt < n_rt; ++rt) {
        const int s = s_lo + (rt < n_plus ? 1 : 0);
        for (int c = 0; c < s; ++c) {
            const int t0 = (int)((long)n_ct * c / s), t1 = (int)((long)n_ct * (c + 1) / s);
            if (t1 > t0) wp.items.push_back(make_int4(rt, t0, t1 - t0, c));
        }
    }
    std::stable_sort(wp.items.begin(), wp.items.end(), [](const int4 &a, const int4 &b) { return a.z > b.z; });
    for (const int4 &w : wp.items) wp.max_tiles = std::max(wp.max_tiles, w.z);
    wp.makespan = lpt_makespan(wp.items, n_sm, overhead);
    return wp;
}

static WorkPlan plan_work(int n_sm, int64_t opt_chunks, int n_rt, int n_ct, int overhead);
static WorkPlan plan_work(const skr_ctx *ctx, int n_rt, int n_ct, int overhead) { return plan_work(ctx->n_sm, ctx->opt_chunks, n_rt, n_ct, overhead); }

static WorkPlan plan_work(int n_sm, int64_t opt_chunks, int n_rt, int n_ct, int overhead)
{
    // overhead: a CTA's fixed cost (prologue, pipeline fill, tail) in tile times -- ~5 us, i.e. 6 three-pass tiles or 10
    // single-pass tiles at d = 64
    const int smax = std::min(n_ct, 8);  // 4 sub-lists per chunk, at most 32 per row
    if (opt_chunks > 0) return make_work(n_rt, n_ct, (int)std::min<int64_t>(opt_chunks, smax), 0, n_sm, overhead);
    WorkPlan best = make_work(n_rt, n_ct, 1, 0, n_sm, overhead);
    for (int s = 2; s <= smax; ++s) {
        WorkPlan w = make_work(n_rt, n_ct, s, 0, n_sm, overhead);
        if (w.makespan < best.makespan) best = w;
    }
    for (int k = 1; k <= 16; ++k) {  // k full waves
        const long C = (long)k * n_sm;
        const int s_lo = (int)(C / n_rt), n_plus = (int)(C - (long)s_lo * n_rt);
        if (s_lo < 1 || n_plus == 0) continue;
        if (s_lo + 1 > smax) break;
        WorkPlan w = make_work(n_rt, n_ct, s_lo, n_plus, n_sm, overhead);
        if (w.makespan < best.makespan) best = w;
    }
    // Many user tiles (>= the SM count): every split fills the GPU and the makespans differ by a per cent.  More chunks mean
    // more and shorter candidate sub-lists per row, which the selection stage reads faster (measured at c4, 1,024 user
    // tiles: 8 chunks instead of 1 take k_select_cands from 5.6 to 2.7 ms and the main pass from 45.4 to 44.1 ms).
    if (best.slots < smax) {
        WorkPlan w = make_work(n_rt, n_ct, smax, 0, n_sm, overhead);
        if (w.makespan * 100 <= best.makespan * 103) best = w;
    }
    return best;
}

int pick_chunks(const skr_ctx *ctx, int n_rt, int n_ct, int K, bool lists)
{
    // heap path: S*K keys must fit one warp sort; list path: chunks are independent, only balance matters
    int smax = lists ? std::min(n_ct, 8) : std::min(std::min(n_ct, 1024 / K), 16);  // lists: 4 S sub-lists <= 32
    if (smax < 1) smax = 1;
    if (ctx->opt_chunks > 0) return (int)std::min<int64_t>(ctx->opt_chunks, smax);
    long best_cost = -1;
    int best = 1;
    for (int S = 1; S <= smax; ++S) {
        long rounds = ((long)n_rt * S + ctx->n_sm - 1) / ctx->n_sm;
        long tiles = (n_ct + S - 1) / S;
        long cost = rounds * (tiles + 6);  // +6 tiles: CTA start-up, heap warm-up and merge cost per chunk
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = S; }
    }
    return best;
}

int make_tmap(skr_ctx *ctx, CUtensorMap *map, const float *base, int64_t n_rows, int d_pad)
{
    cuuint64_t dims[2] = {(cuuint64_t)d_pad, (cuuint64_t)n_rows};
    cuuint64_t strides[1] = {(cuuint64_t)d_pad * sizeof(float)};
    cuuint32_t box[2] = {(cuuint32_t)TC_KB, (cuuint32_t)TN};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = ctx->encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void *)base, dims, strides, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(ctx, SKR_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return SKR_OK;
}


// scaled fp16 item table [n_rows, d_pad] (d_pad = 64 or 128 elements): boxes of 64 elements x 128 rows, the same 128-byte
// swizzle rows and 16 KB tiles as the TF32 tables
int make_tmap_f16(skr_ctx *ctx, CUtensorMap *map, const void *base, int64_t n_rows, int d_pad)
{
    cuuint64_t dims[2] = {(cuuint64_t)d_pad, (cuuint64_t)n_rows};
    cuuint64_t strides[1] = {(cuuint64_t)d_pad * 2};
    cuuint32_t box[2] = {(cuuint32_t)(2 * TC_KB), (cuuint32_t)TN};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = ctx->encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (void *)base, dims, strides, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(ctx, SKR_ERR_CUDA, "cuTensorMapEncodeTiled (fp16) failed (%d)", (int)r);
    return SKR_OK;
}

// augmentation table [n_rows, 16] fp16 (32-byte rows): boxes of 16 x 128 = 4 KB, SWIZZLE_32B
int make_tmap_aug(skr_ctx *ctx, CUtensorMap *map, const void *base, int64_t n_rows)
{
    cuuint64_t dims[2] = {16, (cuuint64_t)n_rows};
    cuuint64_t strides[1] = {32};
    cuuint32_t box[2] = {16, (cuuint32_t)TN};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = ctx->encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (void *)base, dims, strides, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(ctx, SKR_ERR_CUDA, "cuTensorMapEncodeTiled (augmentation table) failed (%d)", (int)r);
    return SKR_OK;
}

typedef void (*TcKernel)(const CUtensorMap, const CUtensorMap, TcArgs, FusedParams);

template <int NKB>
TcKernel tc_kernel_for(int passes, int mode)
{
    if (mode == TC_MODE_SAMPLE) return k_fused_tc<NKB, 1, TC_MODE_SAMPLE>;
    return passes == 3 ? k_fused_tc<NKB, 3, TC_MODE_COLLECT> : k_fused_tc<NKB, 1, TC_MODE_COLLECT>;
}

// half: 0 TF32 operands; 1 FP16 operands (nkb counts 64-element k-blocks: 1 or 2; single pass); 2 FP16 with the bias folded
// into the contraction (mlo then describes the augmentation table)
int launch_tc(skr_ctx *ctx, int nkb, int passes, int mode, unsigned grid, cudaStream_t st, const CUtensorMap &mhi, const CUtensorMap &mlo,
              const TcArgs &A, const FusedParams &P, int half = 0)
{
    TcKernel k = nkb == 1 ? tc_kernel_for<1>(passes, mode) : nkb == 2 ? tc_kernel_for<2>(passes, mode)
               : nkb == 3 ? tc_kernel_for<3>(passes, mode) : tc_kernel_for<4>(passes, mode);
    if (half == 1) {
        if (mode == TC_MODE_SAMPLE) k = nkb == 1 ? (TcKernel)k_fused_tc<1, 1, TC_MODE_SAMPLE, true> : (TcKernel)k_fused_tc<2, 1, TC_MODE_SAMPLE, true>;
        else k = nkb == 1 ? (TcKernel)k_fused_tc<1, 1, TC_MODE_COLLECT, true> : (TcKernel)k_fused_tc<2, 1, TC_MODE_COLLECT, true>;
    } else if (half == 2) {
        if (mode == TC_MODE_SAMPLE) k = nkb == 1 ? (TcKernel)k_fused_tc<1, 1, TC_MODE_SAMPLE, true, true> : (TcKernel)k_fused_tc<2, 1, TC_MODE_SAMPLE, true, true>;
        else k = nkb == 1 ? (TcKernel)k_fused_tc<1, 1, TC_MODE_COLLECT, true, true> : (TcKernel)k_fused_tc<2, 1, TC_MODE_COLLECT, true, true>;
    }
    const size_t smem = tc_smem_bytes();
    SKR_CUDA(ctx, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SKR_CUDA(ctx, launch_pdl(k, dim3((unsigned)(grid)), dim3((unsigned)(TC_THREADS)), (size_t)(smem), st, mhi, mlo, A, P));
    return SKR_OK;
}

template <int PER>
void launch_merge(const u64 *part, int S, int K, int64_t n_rows, int64_t row0, int64_t stride_row, int64_t stride_s, const int64_t *tp,
                  const int32_t *ti, u64 *out, cudaStream_t st)
{
    k_merge_partials<PER><<<(unsigned)((n_rows + 3) / 4), 128, 0, st>>>(part, S, K, n_rows, row0, stride_row, stride_s, tp, ti, out);
}

// S lists of K keys per row -> sorted top-K keys
int merge_lists(skr_ctx *ctx, const u64 *part, int S, int K, int64_t n_rows, int64_t row0, int64_t stride_row, int64_t stride_s,
                const int64_t *tp, const int32_t *ti, u64 *out, cudaStream_t st)
{
    const int n = S * K;
    if (n > 1024) return fail(ctx, SKR_ERR_UNSUPPORTED, "merge of %d lists x %d keys exceeds 1024 keys per row", S, K);
    if (n <= 64) launch_merge<2>(part, S, K, n_rows, row0, stride_row, stride_s, tp, ti, out, st);
    else if (n <= 128) launch_merge<4>(part, S, K, n_rows, row0, stride_row, stride_s, tp, ti, out, st);
    else if (n <= 256) launch_merge<8>(part, S, K, n_rows, row0, stride_row, stride_s, tp, ti, out, st);
    else if (n <= 512) launch_merge<16>(part, S, K, n_rows, row0, stride_row, stride_s, tp, ti, out, st);
    else launch_merge<32>(part, S, K, n_rows, row0, stride_row, stride_s, tp, ti, out, st);
    ctx->launches++;
    SKR_CUDA(ctx, cudaGetLastError());
    return SKR_OK;
}

}  // namespace

extern "C" {

int skr_abi_version(void) { return SKR_ABI_VERSION; }

const char *skr_last_error(const skr_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int skr_ctx_create(int device, skr_ctx **out)
{
    if (!out) return fail(nullptr, SKR_ERR_INVALID, "out is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) return fail(nullptr, SKR_ERR_CUDA, "no CUDA device: %s", cudaGetErrorString(e));
    if (device < 0 || device >= n) return fail(nullptr, SKR_ERR_INVALID, "device %d not in [0,%d)", device, n);
    skr_ctx *ctx = new (std::nothrow) skr_ctx();
    if (!ctx) return fail(nullptr, SKR_ERR_NOMEM, "host allocation failed");
    ctx->device = device;
    cudaDeviceProp prop;
    if ((e = cudaSetDevice(device)) != cudaSuccess || (e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) {
        delete ctx;
        return fail(nullptr, SKR_ERR_CUDA, "device %d: %s", device, cudaGetErrorString(e));
    }
    if (prop.major != 10) {
        delete ctx;
        return fail(nullptr, SKR_ERR_UNSUPPORTED, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
    }
    ctx->n_sm = prop.multiProcessorCount;
    ctx->max_smem = prop.sharedMemPerBlockOptin;
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
        delete ctx;
        return fail(nullptr, SKR_ERR_CUDA, "cuTensorMapEncodeTiled not available: %s", cudaGetErrorString(e));
    }
    ctx->encode = (EncodeTiledFn)fn;
    if ((e = cudaMalloc((void **)&ctx->d_err, sizeof(int))) != cudaSuccess || (e = cudaMemset(ctx->d_err, 0, sizeof(int))) != cudaSuccess) {
        delete ctx;
        return fail(nullptr, SKR_ERR_CUDA, "cudaMalloc: %s", cudaGetErrorString(e));
    }
    const char *dbg_env = getenv("SKR_DEBUG_SYNC");
    ctx->debug_sync = dbg_env != nullptr && dbg_env[0] == '1';
    ctx->ev0.resize(1);
    ctx->ev1.resize(1);
    ctx->ev2.resize(1);
    if ((e = cudaEventCreate(&ctx->ev0[0])) != cudaSuccess || (e = cudaEventCreate(&ctx->ev1[0])) != cudaSuccess ||
        (e = cudaEventCreate(&ctx->ev2[0])) != cudaSuccess) {
        delete ctx;
        return fail(nullptr, SKR_ERR_CUDA, "cudaEventCreate: %s", cudaGetErrorString(e));
    }
    *out = ctx;
    return SKR_OK;
}

int skr_ctx_destroy(skr_ctx *ctx)
{
    if (!ctx) return SKR_OK;
    cudaSetDevice(ctx->device);
    free_dev(ctx->d_tr_indptr); free_dev(ctx->d_tr_idx); free_dev(ctx->d_mask_keys); free_dev(ctx->d_mask_tile_ptr); free_dev(ctx->d_mask_tile_off);
    free_dev(ctx->d_te_indptr); free_dev(ctx->d_te_idx); free_dev(ctx->d_disc); free_dev(ctx->d_idcg); free_dev(ctx->d_err);
    if (ctx->h_pin) cudaFreeHost(ctx->h_pin);
    Buf *bufs[] = {&ctx->bh16, &ctx->f16s, &ctx->rscale, &ctx->baug, &ctx->keys, &ctx->per_user, &ctx->partial, &ctx->part, &ctx->thr, &ctx->bhi, &ctx->blo, &ctx->bias,
                   &ctx->sums, &ctx->stage_a, &ctx->stage_b, &ctx->stage_c, &ctx->out_idx, &ctx->samp, &ctx->cand, &ctx->cand_cnt,
                   &ctx->fail_list, &ctx->trace, &ctx->stats, &ctx->stage_s, &ctx->eps2, &ctx->rs_keys, &ctx->rs_cnt};
    for (Buf *b : bufs) free_dev(b->p);
    for (auto &w : ctx->work_cache) free_dev(w.buf.p);
    for (cudaEvent_t e : ctx->ev0) cudaEventDestroy(e);
    for (cudaEvent_t e : ctx->ev1) cudaEventDestroy(e);
    for (cudaEvent_t e : ctx->ev2) cudaEventDestroy(e);
    delete ctx;
    return SKR_OK;
}

int skr_set_option(skr_ctx *ctx, const char *name, int64_t value)
{
    if (!ctx || !name) return SKR_ERR_INVALID;
    if (!strcmp(name, "chunks")) { ctx->opt_chunks = value; return SKR_OK; }
    if (!strcmp(name, "stages")) { ctx->opt_stages = value; return SKR_OK; }
    if (!strcmp(name, "sample_tiles")) { ctx->opt_sample_tiles = value; return SKR_OK; }
    if (!strcmp(name, "rank")) { ctx->opt_rank = value; return SKR_OK; }
    if (!strcmp(name, "dbg")) { ctx->opt_dbg = value; return SKR_OK; }
    if (!strcmp(name, "score_fn")) {
        if (value != 0 && value != 1) return fail(ctx, SKR_ERR_INVALID, "score_fn=%lld not in {0, 1}", (long long)value);
        ctx->opt_score_fn = value;
        return SKR_OK;
    }
    if (!strcmp(name, "retry_min")) { ctx->opt_retry_min = value; return SKR_OK; }
    if (!strcmp(name, "no_aug")) { ctx->opt_no_aug = value; return SKR_OK; }
    if (!strcmp(name, "full_rescore")) { ctx->opt_full_rescore = value; return SKR_OK; }
    if (!strcmp(name, "chunk_rows")) { ctx->opt_chunk_rows = value < 0 ? 0 : value; return SKR_OK; }
    if (!strcmp(name, "exact_seg_rows")) { ctx->opt_exact_seg_rows = value; return SKR_OK; }
    if (!strcmp(name, "trace_cta")) { ctx->opt_trace_cta = value; return SKR_OK; }
    if (!strcmp(name, "event_ring")) {
        if (value < 1 || value > 65536) return fail(ctx, SKR_ERR_INVALID, "event_ring=%lld not in [1,65536]", (long long)value);
        cudaSetDevice(ctx->device);
        while ((int64_t)ctx->ev0.size() < value) {
            cudaEvent_t a, b, c;
            if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess || cudaEventCreate(&c) != cudaSuccess)
                return fail(ctx, SKR_ERR_CUDA, "cudaEventCreate failed");
            ctx->ev0.push_back(a);
            ctx->ev1.push_back(b);
            ctx->ev2.push_back(c);
        }
        ctx->ev_calls = 0;
        return SKR_OK;
    }
    return fail(ctx, SKR_ERR_INVALID, "unknown option '%s'", name);
}

int skr_fused_kernel_ms(skr_ctx *ctx, int back, float *ms_out)
{
    if (!ctx || !ms_out) return SKR_ERR_INVALID;
    const int64_t n = (int64_t)ctx->ev0.size();
    if (back < 0 || back >= n || back >= ctx->ev_calls) return fail(ctx, SKR_ERR_STATE, "no timing for call -%d (ring %lld, calls %lld)", back, (long long)n, (long long)ctx->ev_calls);
    const size_t i = (size_t)((ctx->ev_calls - 1 - back) % n);
    SKR_CUDA(ctx, cudaEventSynchronize(ctx->ev1[i]));
    SKR_CUDA(ctx, cudaEventElapsedTime(ms_out, ctx->ev0[i], ctx->ev1[i]));
    return SKR_OK;
}

int skr_fused_prepass_ms(skr_ctx *ctx, int back, float *ms_out)
{
    if (!ctx || !ms_out) return SKR_ERR_INVALID;
    const int64_t n = (int64_t)ctx->ev0.size();
    if (back < 0 || back >= n || back >= ctx->ev_calls) return fail(ctx, SKR_ERR_STATE, "no timing for call -%d", back);
    const size_t i = (size_t)((ctx->ev_calls - 1 - back) % n);
    SKR_CUDA(ctx, cudaEventSynchronize(ctx->ev0[i]));
    SKR_CUDA(ctx, cudaEventElapsedTime(ms_out, ctx->ev2[i], ctx->ev0[i]));
    return SKR_OK;
}

int64_t skr_plan_work_host(int n_user_tiles, int n_item_tiles, int n_sm, int cta_overhead, int chunks, int32_t *items_out, int64_t max_items,
                           int64_t *info_out)
{
    if (n_user_tiles <= 0 || n_item_tiles <= 0 || n_sm <= 0 || cta_overhead < 0 || chunks < 0) return SKR_ERR_INVALID;
    const WorkPlan wp = plan_work(n_sm, chunks, n_user_tiles, n_item_tiles, cta_overhead);
    const int64_t n = (int64_t)wp.items.size();
    if (items_out != nullptr)
        for (int64_t i = 0; i < n && i < max_items; ++i) {
            items_out[4 * i + 0] = wp.items[(size_t)i].x;
            items_out[4 * i + 1] = wp.items[(size_t)i].y;
            items_out[4 * i + 2] = wp.items[(size_t)i].z;
            items_out[4 * i + 3] = wp.items[(size_t)i].w;
        }
    if (info_out != nullptr) {
        info_out[0] = wp.slots;
        info_out[1] = wp.min_slots;
        info_out[2] = wp.max_tiles;
        info_out[3] = wp.mixed ? 1 : 0;
        info_out[4] = wp.makespan;
    }
    return n;
}

int skr_fused_stats(skr_ctx *ctx, int64_t *out, int n_out)
{
    if (!ctx || !out || n_out < 7) return SKR_ERR_INVALID;
    if (n_out >= 8) out[7] = ctx->ev_calls;  // timed scoring launches since the last "event_ring" option (one per row chunk)
    int n_fails[2] = {0, 0};
    SKR_CUDA(ctx, cudaSetDevice(ctx->device));
    if (ctx->fail_list.p) SKR_CUDA(ctx, cudaMemcpy(n_fails, ctx->fail_list.p, 2 * sizeof(int), cudaMemcpyDeviceToHost));
    // tf32r: [0] rows the single-pass attempt left to the three-pass retry, [1] rows the retry left to the exact kernel;
    // other precisions have one attempt
    const bool two = !strcmp(ctx->last_fused, "tcgen05_tf32r") || !strcmp(ctx->last_fused, "tcgen05_f16r");
    const int n_fail = two ? n_fails[1] : n_fails[0];
    if (n_out >= 9) out[8] = two ? n_fails[0] : 0;
    out[0] = ctx->last_plan.n_samp;
    out[1] = ctx->last_plan.stride;
    out[2] = ctx->last_plan.r;
    out[3] = ctx->last_plan.cap;
    out[4] = ctx->last_plan.S;
    out[5] = ctx->last_plan.stages;
    out[6] = n_fail;
    return SKR_OK;
}

int skr_fused_trace(skr_ctx *ctx, int64_t *out, int64_t n_out)
{
    if (!ctx || !out || n_out < 0) return SKR_ERR_INVALID;
    if (!ctx->trace.p) return fail(ctx, SKR_ERR_STATE, "no trace recorded (set option trace_cta)");
    SKR_CUDA(ctx, cudaSetDevice(ctx->device));
    SKR_CUDA(ctx, cudaDeviceSynchronize());
    const size_t n = std::min<size_t>((size_t)n_out * sizeof(int64_t), ctx->trace.cap);
    SKR_CUDA(ctx, cudaMemcpy(out, ctx->trace.p, n, cudaMemcpyDeviceToHost));
    return SKR_OK;
}

int64_t skr_launch_count(const skr_ctx *ctx) { return ctx ? ctx->launches : 0; }
const char *skr_last_fused_kernel(const skr_ctx *ctx) { return ctx ? ctx->last_fused : "none"; }

int skr_set_train_csr(skr_ctx *ctx, const int64_t *indptr, const int32_t *indices, int64_t n_rows, int64_t n_items)
{
    if (!ctx) return SKR_ERR_INVALID;
    SKR_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!indptr) {
        ctx->has_train = false;
        return SKR_OK;
    }
    if (n_items >= (1ll << 25)) return fail(ctx, SKR_ERR_UNSUPPORTED, "n_items=%lld >= 2^25 (mask key packing)", (long long)n_items);
    // rows -> sorted unique, plus the fused path's mask keys (per 128-row user tile, ascending (item << 7 | row_in_tile),
    // with the offset of every 128-item tile): built on the device (ingest.cu)
    IngestOut o;
    char msg[256] = "";
    const int rc = ingest_csr(indptr, indices, n_rows, n_items, true, TM, TN, &o, msg, sizeof(msg), nullptr);
    if (rc) {
        ingest_free(&o);
        return fail(ctx, rc, "%s", msg);
    }
    free_dev(ctx->d_tr_indptr); free_dev(ctx->d_tr_idx); free_dev(ctx->d_mask_keys); free_dev(ctx->d_mask_tile_ptr); free_dev(ctx->d_mask_tile_off);
    ctx->d_tr_indptr = o.indptr;
    ctx->d_tr_idx = o.idx;
    ctx->d_mask_keys = o.mask_keys;
    ctx->d_mask_tile_ptr = o.tile_ptr;
    ctx->d_mask_tile_off = o.tile_off;
    ctx->tr_rows = n_rows;
    ctx->tr_items = n_items;
    ctx->has_train = true;
    return SKR_OK;
}

int skr_set_test_csr(skr_ctx *ctx, const int64_t *indptr, const int32_t *indices, int64_t n_rows, int64_t n_items)
{
    if (!ctx) return SKR_ERR_INVALID;
    if (!indptr || n_rows <= 0) return fail(ctx, SKR_ERR_INVALID, "test CSR must not be empty (evaluator.py:144)");
    SKR_CUDA(ctx, cudaSetDevice(ctx->device));
    IngestOut o;
    char msg[256] = "";
    const int rc = ingest_csr(indptr, indices, n_rows, n_items, false, TM, TN, &o, msg, sizeof(msg), nullptr);
    if (rc) {
        ingest_free(&o);
        return fail(ctx, rc, "%s", msg);
    }
    free_dev(ctx->d_te_indptr); free_dev(ctx->d_te_idx);
    ctx->d_te_indptr = o.indptr;
    ctx->d_te_idx = o.idx;
    ctx->te_rows = n_rows;
    ctx->te_items = n_items;
    ctx->has_test = true;
    return SKR_OK;
}

int skr_metrics_from_topk(skr_ctx *ctx, const int32_t *topk_idx_dev, int64_t n_rows, int64_t row0, const int32_t *metric_ids,
                          int n_metrics, int top_k, float *per_user_dev, double *sums_dev, void *stream)
{
    if (!ctx) return SKR_ERR_INVALID;
    if (!topk_idx_dev || n_rows <= 0) return fail(ctx, SKR_ERR_INVALID, "metrics_from_topk: empty input");
    MetricIds m;
    int rc = check_metrics(ctx, metric_ids, n_metrics, top_k, m);
    if (rc) return rc;
    SKR_CUDA(ctx, cudaSetDevice(ctx->device));
    return run_metrics(ctx, nullptr, topk_idx_dev, n_rows, row0, m, top_k, nullptr, nullptr, per_user_dev, sums_dev, (cudaStream_t)stream);
}

int skr_colsum_f32_seq(skr_ctx *ctx, const float *per_user_dev, int64_t n_rows, int64_t n_cols, float *acc_dev, void *stream)
{
    if (!ctx) return SKR_ERR_INVALID;
    if (!per_user_dev || !acc_dev || n_rows < 0 || n_cols <= 0) return fail(ctx, SKR_ERR_INVALID, "colsum_f32_seq: bad arguments");
    SKR_CUDA(ctx, cudaSetDevice(ctx->device));
    k_colsum_f32_seq<<<(unsigned)((n_cols + 63) / 64), 64, 0, (cudaStream_t)stream>>>(per_user_dev, n_rows, (int)n_cols, acc_dev);
    ctx->launches++;
    SKR_CUDA(ctx, cudaGetLastError());
    return SKR_OK;
}

int skr_colsum_rows(skr_ctx *ctx, const float *per_user_dev, int64_t n_cols, const int32_t *row_list_dev, int64_t n_list, double *sums_dev,
                    void *stream)
{
    if (!ctx) return SKR_ERR_INVALID;
    if (!per_user_dev || !sums_dev || n_cols <= 0 || n_list < 0 || (n_list > 0 && !row_list_dev))
        return fail(ctx, SKR_ERR_INVALID, "colsum_rows: bad arguments");
    if (n_list == 0) return SKR_OK;
    SKR_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int nblk = (int)std::min<int64_t>(n_list, 2 * ctx->n_sm);
    int rc = ensure(ctx, ctx->partial, (size_t)nblk * n_cols * sizeof(double));
    if (rc) return rc;
    k_colsum_partial<<<nblk, 256, 0, st>>>(per_user_dev, n_list, (int)n_cols, (double *)ctx->partial.p, row_list_dev);
    SKR_CUDA(ctx, launch_pdl(k_colsum_fold, dim3((unsigned)((unsigned)n_cols)), dim3((unsigned)(256)), (size_t)(0), st, (const double *)ctx->partial.p, nblk, (int)n_cols, sums_dev));
    ctx->launches += 2;
    SKR_CUDA(ctx, cudaGetLastError());
    return SKR_OK;
}

int skr_topk_scores(skr_ctx *ctx, const float *scores_dev, int64_t n_rows, int64_t n_items, int64_t ld, int top_k, int32_t *topk_idx_dev,
                    float *topk_val_dev, void *stream)
{
    if (!ctx) return SKR_ERR_INVALID;
    if (!scores_dev || n_rows <= 0 || (!topk_idx_dev && !topk_val_dev)) return fail(ctx, SKR_ERR_INVALID, "topk_scores: empty input / no output");
    if (top_k < 1 || top_k > K2_MAX_K) return fail(ctx, SKR_ERR_UNSUPPORTED, "top_k=%d not in [1,%d]", top_k, K2_MAX_K);
    if (n_items < top_k) return fail(ctx, SKR_ERR_INVALID, "n_items=%lld < top_k=%d", (long long)n_items, top_k);
    if (n_items > 0x7fffffffll - K2_CHUNK) return fail(ctx, SKR_ERR_UNSUPPORTED, "n_items=%lld too large", (long long)n_items);
    if (ld < n_items) return fail(ctx, SKR_ERR_INVALID, "ld=%lld < n_items=%lld", (long long)ld, (long long)n_items);
    SKR_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    int rc = ensure(ctx, ctx->keys, (size_t)n_rows * top_k * sizeof(u64));
    if (rc) return rc;
    launch_topk_scores(scores_dev, ld, (int)n_items, n_rows, 0, nullptr, nullptr, top_k, (u64 *)ctx->keys.p, ctx->d_err, st);
    const int64_t nk = n_rows * top_k;
    k_unpack_keys<<<(unsigned)((nk + 255) / 256), 256, 0, st>>>((const u64 *)ctx->keys.p, nk, topk_idx_dev, topk_val_dev);
    ctx->launches += 2;
    SKR_CUDA(ctx, cudaGetLastError());
    return SKR_OK;
}

int skr_topk_scores_host(skr_ctx *ctx, const float *scores_host, int64_t n_rows, int64_t n_items, int64_t ld, int top_k,
                         int32_t *topk_idx_host, float *topk_val_host, void *stream)
{
    if (!ctx) return SKR_ERR_INVALID;
    if (!scores_host || n_rows <= 0 || n_items <= 0 || ld < n_items || top_k < 1) return fail(ctx, SKR_ERR_INVALID, "topk_scores_host: bad arguments");
    SKR_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    int rc;
    if ((rc = ensure(ctx, ctx->stage_a, (size_t)n_rows * n_items * sizeof(float)))) return rc;
    if ((rc = ensure(ctx, ctx->out_idx, (size_t)n_rows * top_k * sizeof(int32_t)))) return rc;
    if ((rc = ensure(ctx, ctx->stage_c, (size_t)n_rows * top_k * sizeof(float)))) return rc;
    SKR_CUDA(ctx, cudaMemcpy2DAsync(ctx->stage_a.p, (size_t)n_items * sizeof(float), scores_host, (size_t)ld * sizeof(float),
                                    (size_t)n_items * sizeof(float), (size_t)n_rows, cudaMemcpyHostToDevice, st));
    rc = skr_topk_scores(ctx, (const float *)ctx->stage_a.p, n_rows, n_items, n_items, top_k, (int32_t *)ctx->out_idx.p, (float *)ctx->stage_c.p, stream);
    if (rc) return rc;
    if (topk_idx_host) SKR_CUDA(ctx, cudaMemcpyAsync(topk_idx_host, ctx->out_idx.p, (size_t)n_rows * top_k * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    if (topk_val_host) SKR_CUDA(ctx, cudaMemcpyAsync(topk_val_host, ctx->stage_c.p, (size_t)n_rows * top_k * sizeof(float), cudaMemcpyDeviceToHost, st));
    SKR_CUDA(ctx, cudaStreamSynchronize(st));
    return SKR_OK;
}

int skr_eval_scores(skr_ctx *ctx, const float *scores_dev, int64_t n_rows, int64_t n_items, int64_t ld, int64_t row0,
                    const int32_t *metric_ids, int n_metrics, int top_k, int32_t *topk_idx_dev, float *topk_val_dev,
                    float *per_user_dev, double *sums_dev, void *stream)
{
    if (!ctx) return SKR_ERR_INVALID;
    MetricIds m;
    int rc = check_metrics(ctx, metric_ids, n_metrics, top_k, m);
    if (rc) return rc;
    if (!scores_dev || n_rows <= 0) return fail(ctx, SKR_ERR_INVALID, "eval_scores: empty input");
    if (top_k > K2_MAX_K) return fail(ctx, SKR_ERR_UNSUPPORTED, "top_k=%d > %d", top_k, K2_MAX_K);
    if (n_items < top_k) return fail(ctx, SKR_ERR_INVALID, "n_items=%lld < top_k=%d (evaluate.h:45 would read out of bounds)", (long long)n_items, top_k);
    if (n_items > 0x7fffffffll - K2_CHUNK) return fail(ctx, SKR_ERR_UNSUPPORTED, "n_items=%lld too large", (long long)n_items);
    if (ld < n_items) return fail(ctx, SKR_ERR_INVALID, "ld=%lld < n_items=%lld", (long long)ld, (long long)n_items);
    if (ctx->has_train && (row0 < 0 || row0 + n_rows > ctx->tr_rows))
        return fail(ctx, SKR_ERR_INVALID, "rows [%lld,%lld) outside the train CSR (%lld rows)", (long long)row0, (long long)(row0 + n_rows), (long long)ctx->tr_rows);
    if (ctx->has_train && ctx->tr_items > n_items) return fail(ctx, SKR_ERR_INVALID, "train CSR has %lld items, scores only %lld", (long long)ctx->tr_items, (long long)n_items);
    SKR_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    rc = ensure(ctx, ctx->keys, (size_t)n_rows * top_k * sizeof(u64));
    if (rc) return rc;
    launch_topk_scores(scores_dev, ld, (int)n_items, n_rows, row0, ctx->has_train ? ctx->d_tr_indptr : nullptr, ctx->has_train ? ctx->d_tr_idx : nullptr,
                       top_k, (u64 *)ctx->keys.p, ctx->d_err, st);
    ctx->launches++;
    SKR_CUDA(ctx, cudaGetLastError());
    return run_metrics(ctx, (const u64 *)ctx->keys.p, nullptr, n_rows, row0, m, top_k, topk_idx_dev, topk_val_dev, per_user_dev, sums_dev, st);
}

// The fused pipeline.  keys_only == null: metrics of the rows (skr_eval_fused).  keys_only != null: the rows'
// sorted top-K rank keys over this item table with item ids shifted by item_offset, no metrics (skr_topk_fused).
// Shapes outside the selection epilogues of the fused kernels (top-K > 128): score blocks of rows with the library's
// own FP32 tile kernel (k_fused_simt<true>: same FMA chain as the exact path) into a workspace block sized to stay in
// L2 between producer and consumer, then the score-matrix kernels (train masking, top-K <= 512, metrics).  The U x I
// matrix still never exists: one block of at most 96 MB is live at a time.
static int fused_by_blocks(skr_ctx *ctx, const float *user_vecs_dev, int64_t n_rows, int64_t ld_u, const float *item_vecs_dev, int64_t n_items,
                           int64_t ld_i, int d, const float *bias_dev, int64_t row0, const MetricIds &m, int top_k, int32_t *topk_idx_dev,
                           float *topk_val_dev, float *per_user_dev, double *sums_dev, void *stream)
{
    int rc;
    if ((d & 3) || (ld_u & 3) || (ld_i & 3) || ((uintptr_t)user_vecs_dev & 15) || ((uintptr_t)item_vecs_dev & 15))
        return fail(ctx, SKR_ERR_UNSUPPORTED, "top_k > 128 runs on the FP32 tile kernel: d, ld_u, ld_i must be multiples of 4 and the tables 16-byte aligned");
    if (n_items > 0x7fffffffll - K2_CHUNK) return fail(ctx, SKR_ERR_UNSUPPORTED, "n_items=%lld too large", (long long)n_items);
    if (ctx->has_train && (row0 < 0 || row0 + n_rows > ctx->tr_rows))
        return fail(ctx, SKR_ERR_INVALID, "rows [%lld,%lld) outside the train CSR (%lld rows)", (long long)row0, (long long)(row0 + n_rows), (long long)ctx->tr_rows);
    SKR_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t ld_s = (n_items + 3) & ~(int64_t)3;
    int64_t blk = std::max<int64_t>(TM, ((96ll << 20) / (ld_s * 4)) / TM * TM);
    blk = std::min<int64_t>(blk, ((n_rows + TM - 1) / TM) * TM);
    if ((rc = ensure(ctx, ctx->stage_s, (size_t)blk * ld_s * sizeof(float)))) return rc;
    if ((rc = ensure(ctx, ctx->keys, (size_t)blk * top_k * sizeof(u64)))) return rc;
    FusedParams P;
    memset(&P, 0, sizeof(P));
    P.n_items = (int)n_items;
    P.d = d;
    P.K = top_k;
    P.n_ct = (int)((n_items + TN - 1) / TN);
    P.bias = bias_dev;  // read only below n_items
    const size_t smem = simt_smem_bytes(0);
    typedef void (*SimtKernel)(const float *, int64_t, const float *, int64_t, FusedParams, float *, int64_t);
    SimtKernel simt = ctx->opt_score_fn == 1 ? (SimtKernel)k_fused_simt<true, 1> : (SimtKernel)k_fused_simt<true, 0>;
    SKR_CUDA(ctx, cudaFuncSetAttribute(simt, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int MK = m.n * top_k;
    const size_t slot = (size_t)(ctx->ev_calls % (int64_t)ctx->ev0.size());
    SKR_CUDA(ctx, cudaEventRecord(ctx->ev2[slot], st));
    SKR_CUDA(ctx, cudaEventRecord(ctx->ev0[slot], st));
    for (int64_t b0 = 0; b0 < n_rows; b0 += blk) {
        const int64_t nb = std::min(blk, n_rows - b0);
        P.n_rows = nb;
        P.row0 = 0;
        P.n_rt = (int)((nb + TM - 1) / TM);
        // item tile ranges per user tile: enough CTAs to fill the GPU
        P.S = (int)std::max<int64_t>(1, std::min<int64_t>(P.n_ct, (2 * ctx->n_sm + P.n_rt - 1) / P.n_rt));
        P.tiles_per_chunk = (P.n_ct + P.S - 1) / P.S;
        P.S = (P.n_ct + P.tiles_per_chunk - 1) / P.tiles_per_chunk;
        simt<<<(unsigned)(P.n_rt * P.S), SIMT_THREADS, smem, st>>>(user_vecs_dev + b0 * ld_u, ld_u, item_vecs_dev, ld_i, P, (float *)ctx->stage_s.p, ld_s);
        k_topk_scores<<<(unsigned)nb, K2_THREADS, 0, st>>>((const float *)ctx->stage_s.p, ld_s, (int)n_items, row0 + b0, ctx->has_train ? ctx->d_tr_indptr : nullptr,
                                                          ctx->has_train ? ctx->d_tr_idx : nullptr, top_k, (u64 *)ctx->keys.p);
        ctx->launches += 2;
        SKR_CUDA(ctx, cudaGetLastError());
        if ((rc = run_metrics(ctx, (const u64 *)ctx->keys.p, nullptr, nb, row0 + b0, m, top_k, topk_idx_dev ? topk_idx_dev + b0 * top_k : nullptr,
                              topk_val_dev ? topk_val_dev + b0 * top_k : nullptr, per_user_dev ? per_user_dev + b0 * MK : nullptr, sums_dev, st)))
            return rc;
    }
    SKR_CUDA(ctx, cudaEventRecord(ctx->ev1[slot], st));
    ctx->ev_calls++;
    ctx->last_fused = ctx->opt_score_fn == 1 ? "simt_fp32_blocks_negl2" : "simt_fp32_blocks";
    ctx->last_plan = {0, 0, 0, 0, 0, 0};
    return SKR_OK;
}

// One row chunk of the fused pipeline.  first_chunk: the item-side preparation (bias padding, TF32 split, item
// statistics) runs; later chunks of the same evaluate reuse it.
static int fused_chunk(skr_ctx *ctx, const float *user_vecs_dev, int64_t n_rows, int64_t ld_u, const float *item_vecs_dev,
                       int64_t n_items, int64_t ld_i, int d, const float *bias_dev, int64_t row0, const MetricIds &m, int top_k,
                       int precision, int32_t *topk_idx_dev, float *topk_val_dev, float *per_user_dev, double *sums_dev,
                       u64 *keys_only, int64_t item_offset, void *stream, bool first_chunk)
{
    int rc;
    if (!user_vecs_dev || !item_vecs_dev || n_rows <= 0 || d <= 0) return fail(ctx, SKR_ERR_INVALID, "eval_fused: empty input");
    if (ld_u < d || ld_i < d) return fail(ctx, SKR_ERR_INVALID, "ld_u=%lld / ld_i=%lld < d=%d", (long long)ld_u, (long long)ld_i, d);
    if (n_items < top_k) return fail(ctx, SKR_ERR_INVALID, "n_items=%lld < top_k=%d (evaluate.h:45 would read out of bounds)", (long long)n_items, top_k);
    if (n_items >= (1ll << 25)) return fail(ctx, SKR_ERR_UNSUPPORTED, "n_items=%lld >= 2^25", (long long)n_items);
    if (top_k > K2_MAX_K) return fail(ctx, SKR_ERR_UNSUPPORTED, "top_k=%d > %d", top_k, K2_MAX_K);
    if (top_k > 128) {
        if (keys_only != nullptr) return fail(ctx, SKR_ERR_UNSUPPORTED, "per-shard lists (skr_topk_fused) support top_k <= 128 (got %d)", top_k);
        return fused_by_blocks(ctx, user_vecs_dev, n_rows, ld_u, item_vecs_dev, n_items, ld_i, d, bias_dev, row0, m, top_k, topk_idx_dev, topk_val_dev,
                               per_user_dev, sums_dev, stream);
    }
    if (row0 % TM != 0) return fail(ctx, SKR_ERR_INVALID, "row0=%lld must be a multiple of %d", (long long)row0, TM);
    if (ctx->has_train && (row0 < 0 || row0 + n_rows > ctx->tr_rows))
        return fail(ctx, SKR_ERR_INVALID, "rows [%lld,%lld) outside the train CSR (%lld rows)", (long long)row0, (long long)(row0 + n_rows), (long long)ctx->tr_rows);
    if (ctx->has_train && ctx->tr_items != n_items) return fail(ctx, SKR_ERR_INVALID, "train CSR was built for %lld items, item table has %lld", (long long)ctx->tr_items, (long long)n_items);
    if (precision < SKR_PREC_AUTO || precision > SKR_PREC_F16R) return fail(ctx, SKR_ERR_INVALID, "precision=%d", precision);
    SKR_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int K = top_k;

    // ---- kernel choice ----
    const int nkb = (d + TC_KB - 1) / TC_KB;
    const bool tc_ok = (nkb <= 4) && (tc_smem_bytes() <= ctx->max_smem);
    bool use_tc = (precision != SKR_PREC_FP32) && tc_ok;
    if (ctx->opt_score_fn == 1) {  // -||u - i|| + b: FP32 tile kernel only
        if (precision == SKR_PREC_3XTF32 || precision == SKR_PREC_1XTF32 || precision == SKR_PREC_TF32R || precision == SKR_PREC_F16R)
            return fail(ctx, SKR_ERR_UNSUPPORTED, "score_fn = -||u - i|| + b runs on the FP32 tile kernel: precision must be auto or fp32");
        use_tc = false;
    }
    // AUTO, catalogue too small for sampled thresholds (about 160 items per requested rank: ml-1m at top-50 or
    // top-100): the tensor-core pipeline would settle every row through its exact per-row fallback.  The FP32 FMA
    // kernel with running per-row heaps is one launch and exact; the whole job is a few GFLOP.
    const bool simt_ok = !((d & 3) || (ld_u & 3) || (ld_i & 3) || ((uintptr_t)user_vecs_dev & 15) || ((uintptr_t)item_vecs_dev & 15)) &&
                         simt_smem_bytes(top_k) <= ctx->max_smem;
    if (precision == SKR_PREC_AUTO && simt_ok && n_items < std::max<int64_t>(3072, 160 * (int64_t)top_k)) use_tc = false;
    if (!use_tc && (precision == SKR_PREC_3XTF32 || precision == SKR_PREC_1XTF32 || precision == SKR_PREC_TF32R || precision == SKR_PREC_F16R))
        return fail(ctx, SKR_ERR_UNSUPPORTED, "tcgen05 path needs d <= 128 (d=%d)", d);
    if (!use_tc) {
        if ((d & 3) || (ld_u & 3) || (ld_i & 3) || ((uintptr_t)user_vecs_dev & 15) || ((uintptr_t)item_vecs_dev & 15))
            return fail(ctx, SKR_ERR_UNSUPPORTED, "FP32 path needs d, ld_u, ld_i multiples of 4 and 16-byte aligned tables");
        if (simt_smem_bytes(K) > ctx->max_smem) return fail(ctx, SKR_ERR_UNSUPPORTED, "top_k=%d does not fit shared memory", K);
    }

    FusedParams P;
    P.n_rows = n_rows;
    P.row0 = row0;
    P.n_items = (int)n_items;
    P.d = d;
    P.K = K;
    P.n_ct = (int)((n_items + TN - 1) / TN);
    P.n_rt = (int)((n_rows + TM - 1) / TM);
    P.S = pick_chunks(ctx, P.n_rt, P.n_ct, K, use_tc);
    P.tiles_per_chunk = (P.n_ct + P.S - 1) / P.S;
    P.S = (P.n_ct + P.tiles_per_chunk - 1) / P.tiles_per_chunk;  // drop empty chunks
    P.mask_keys = ctx->has_train ? ctx->d_mask_keys : nullptr;
    P.mask_tile_ptr = ctx->has_train ? ctx->d_mask_tile_ptr : nullptr;
    P.mask_tile_off = ctx->has_train ? ctx->d_mask_tile_off : nullptr;
    P.thr_g = nullptr;
    P.part = nullptr;

    if ((rc = ensure(ctx, ctx->keys, (size_t)n_rows * K * sizeof(u64)))) return rc;
    u64 *keys = (u64 *)ctx->keys.p;
    const int64_t *tp = ctx->has_train ? ctx->d_tr_indptr : nullptr;
    const int32_t *ti = ctx->has_train ? ctx->d_tr_idx : nullptr;
    P.bias = nullptr;
    if (bias_dev) {
        const int n_pad = P.n_ct * TN;
        if ((rc = ensure(ctx, ctx->bias, (size_t)n_pad * sizeof(float)))) return rc;
        if (first_chunk) {
            SKR_CUDA(ctx, launch_pdl(k_pad_bias, dim3((unsigned)((n_pad + 255) / 256)), dim3((unsigned)(256)), (size_t)(0), st, bias_dev, (int)n_items, n_pad, (float *)ctx->bias.p));
            ctx->launches++;
        }
        P.bias = (const float *)ctx->bias.p;
    }
    const unsigned grid = (unsigned)(P.n_rt * P.S);
    const size_t ring = ctx->ev0.size();
    const size_t slot = (size_t)(ctx->ev_calls % (int64_t)ring);

    if (use_tc) {
        // AUTO: re-scoring costs a fixed ~60 candidate rows per user, the two extra MMA passes of 3xTF32 grow with
        // the catalogue: measured on one B200, tf32r is 5 % faster at c2 (I d = 2.6 M) and 22 % at c3b (5.9 M); below
        // ~1 M the sampled thresholds stop working for either and the choice does not matter
        // d > 64: a tile's TF32 MMAs (13-17 x 64 cycles) outlast its epilogue; with FP16 operands (9 MMAs, half the item-panel
        // bytes) the epilogue is the limit instead: c4 (d = 128) 453.7 -> 318.6 ms per 10^6 users.  At d <= 64 both are
        // epilogue-bound and equally fast (c2: 0.660 vs 0.675 ms), so TF32 stays.
        if (precision == SKR_PREC_AUTO)
            precision = ((double)n_items * d >= 2.0e6) ? (d > 64 ? SKR_PREC_F16R : SKR_PREC_TF32R) : SKR_PREC_3XTF32;
        const bool half = (precision == SKR_PREC_F16R);  // FP16 operands for the first attempt; the retry stays 3xTF32
        const bool rescore = (precision == SKR_PREC_TF32R) || half;
        const int passes = (precision == SKR_PREC_1XTF32 || rescore) ? 1 : 3;
        // operand prep: item table -> hi (and, for three passes, lo) TF32 tables, TMA descriptors
        const int d_pad = nkb * TC_KB;
        // FP16: 64-element k-blocks; with a bias an augmentation table carries (C, C, beta_hi, beta_lo) per item: threshold and
        // bias are then applied by one MMA per tile and the epilogue is that of an unbiased model (k_fused_tc AUG)
        const bool aug = half && bias_dev != nullptr && ctx->opt_no_aug == 0;
        const int nkb_h = (d + 2 * TC_KB - 1) / (2 * TC_KB), d_pad_h = nkb_h * 2 * TC_KB;
        const int nkb_1 = half ? nkb_h : nkb;  // k-blocks of the first attempt's kernels
        const int half_mode = half ? (aug ? 2 : 1) : 0;
        const size_t tbytes = (size_t)n_items * d_pad * sizeof(float);
        if ((rc = ensure(ctx, ctx->bhi, tbytes))) return rc;  // f16r: filled on demand by the retry, like blo
        if ((passes == 3 || rescore) && (rc = ensure(ctx, ctx->blo, tbytes))) return rc;  // tf32r: filled on demand by the retry
        if (half) {
            if ((rc = ensure(ctx, ctx->bh16, (size_t)n_items * d_pad_h * 2))) return rc;
            if ((rc = ensure(ctx, ctx->f16s, 4 * sizeof(uint32_t)))) return rc;
            if (aug && (rc = ensure(ctx, ctx->baug, (size_t)n_items * 32))) return rc;
            if ((rc = ensure(ctx, ctx->rscale, (size_t)2 * n_rows * sizeof(float)))) return rc;  // s_u s_i | g
        }
        if ((rc = ensure(ctx, ctx->fail_list, (size_t)(2 * n_rows + 2) * sizeof(int32_t)))) return rc;
        if (ctx->stats.cap == 0) {  // two slots of {max ||item||^2, max |bias|}, alternating between evaluates
            if ((rc = ensure(ctx, ctx->stats, 4 * sizeof(uint32_t)))) return rc;
            SKR_CUDA(ctx, cudaMemsetAsync(ctx->stats.p, 0, 4 * sizeof(uint32_t), st));
        }
        if (first_chunk || ctx->stats_cur_ptr == nullptr) {
            uint32_t *stats_cur = (uint32_t *)ctx->stats.p + 2 * ctx->stats_slot;
            uint32_t *stats_next = (uint32_t *)ctx->stats.p + 2 * (ctx->stats_slot ^ 1);
            ctx->stats_slot ^= 1;
            ctx->stats_cur_ptr = stats_cur;
            const dim3 sgrid((unsigned)std::min<int64_t>((n_items + 31) / 32, 16 * ctx->n_sm));
            if (half) {
                SKR_CUDA(ctx, cudaMemsetAsync(ctx->f16s.p, 0, 4 * sizeof(uint32_t), st));
                SKR_CUDA(ctx, launch_pdl(k_item_absmax, sgrid, dim3(256u), (size_t)0, st, item_vecs_dev, ld_i, n_items, d, aug ? bias_dev : (const float *)nullptr,
                                         (uint32_t *)ctx->f16s.p));
                SKR_CUDA(ctx, launch_pdl(k_split_f16, sgrid, dim3(256u), (size_t)0, st, item_vecs_dev, ld_i, n_items, d, d_pad_h, (__half *)ctx->bh16.p, bias_dev,
                                         (int *)ctx->fail_list.p, stats_cur, stats_next, (uint32_t *)ctx->f16s.p, aug ? (__half *)ctx->baug.p : (__half *)nullptr));
                ctx->launches++;
            } else {
                SKR_CUDA(ctx, launch_pdl(k_split_tf32, sgrid, dim3((unsigned)(256)), (size_t)(0), st, item_vecs_dev, ld_i, n_items, d, d_pad, (float *)ctx->bhi.p,
                                         passes == 3 ? (float *)ctx->blo.p : (float *)nullptr, bias_dev, (int *)ctx->fail_list.p, stats_cur, stats_next));
            }
            ctx->launches++;
        } else {
            SKR_CUDA(ctx, cudaMemsetAsync(ctx->fail_list.p, 0, 2 * sizeof(int), st));  // the split kernel's other job: empty fail lists
        }
        uint32_t *stats_cur = ctx->stats_cur_ptr;
        CUtensorMap mhi, mlo, mh16;
        if ((rc = make_tmap(ctx, &mhi, (const float *)ctx->bhi.p, n_items, d_pad))) return rc;
        if (passes == 3 || rescore) { if ((rc = make_tmap(ctx, &mlo, (const float *)ctx->blo.p, n_items, d_pad))) return rc; }
        else mlo = mhi;  // single-pass kernels never touch the lo table: it is not even built
        if (half) { if ((rc = make_tmap_f16(ctx, &mh16, ctx->bh16.p, n_items, d_pad_h))) return rc; }
        else mh16 = mhi;
        const CUtensorMap &m1 = half ? mh16 : mhi;  // what the first attempt's kernels stream
        CUtensorMap maug = mlo;                     // ... and the second map they get: lo table (3xTF32) or augmentation table (f16r with a bias)
        if (aug && (rc = make_tmap_aug(ctx, &maug, ctx->baug.p, n_items))) return rc;

        // sampling plan (k_fused_tc.cuh header): fraction f ~ 6/K of the item tiles, threshold = r-th largest
        // sampled group maximum with r = K f + 4.5 sqrt(K f) + 8 (a ~4-sigma margin against fewer than K survivors)
        int n_samp = (int)lround(6.0 * P.n_ct / K);
        n_samp = std::max(1, std::min(n_samp, std::max(1, P.n_ct / 3)));
        if (ctx->opt_sample_tiles > 0) n_samp = (int)std::min<int64_t>(ctx->opt_sample_tiles, P.n_ct);
        const int stride = std::max(1, P.n_ct / n_samp);
        n_samp = (P.n_ct + stride - 1) / stride;
        const double f_eff = std::min(1.0, (double)n_samp * TN / (double)n_items);
        const double kf = K * f_eff;
        int r = (int)ceil(kf + 4.5 * sqrt(kf) + 8.0);
        r = std::max(1, std::min(r, TC_MAX_RANK));
        if (ctx->opt_rank > 0) r = (int)std::min<int64_t>(ctx->opt_rank, TC_MAX_RANK);
        const double expect = r / f_eff;  // candidates per row
        // work list of the main pass (cached: it depends only on the tile counts and the cost of a tile)
        const int cta_overhead = (passes == 1) ? (nkb <= 2 ? 10 : 6) : (nkb <= 2 ? 6 : 4);
        const int wkey[4] = {P.n_rt, P.n_ct, (int)ctx->opt_chunks, cta_overhead};
        skr_ctx::WorkCache *wc = nullptr, *lru = &ctx->work_cache[0];
        for (auto &w : ctx->work_cache) {
            if (w.key[0] == wkey[0] && w.key[1] == wkey[1] && w.key[2] == wkey[2] && w.key[3] == wkey[3]) wc = &w;
            if (w.used < lru->used) lru = &w;
        }
        if (wc == nullptr) {
            wc = lru;
            const WorkPlan wp = plan_work(ctx, P.n_rt, P.n_ct, cta_overhead);
            if ((rc = ensure(ctx, wc->buf, wp.items.size() * sizeof(int4)))) return rc;
            SKR_CUDA(ctx, cudaMemcpyAsync(wc->buf.p, wp.items.data(), wp.items.size() * sizeof(int4), cudaMemcpyHostToDevice, st));
            SKR_CUDA(ctx, cudaStreamSynchronize(st));  // the host vector goes away; happens once per shape
            for (int q = 0; q < 4; ++q) wc->key[q] = wkey[q];
            wc->ctas = (int)wp.items.size(); wc->slots = wp.slots; wc->min_slots = wp.min_slots;
            wc->max_tiles = wp.max_tiles; wc->mixed = wp.mixed;
        }
        wc->used = ++ctx->work_clock;
        P.S = wc->slots;
        const unsigned grid_tc = (unsigned)wc->ctas;
        const int n_sub = 4 * P.S;        // one sub-list per (item chunk, column quarter of the tile)
        // capacity of a sub-list: twice the expected share of the user tiles with the fewest chunks
        int cap = next_pow2((int)(2.0 * expect / (4 * wc->min_slots)) + 16);
        cap = std::max(16, std::min(cap, 512));

        // SAMPLE CTAs per user tile: enough to fill the GPU when there are few user tiles, at least 8 sample tiles each
        const int samp_chunks = std::max(1, std::min(std::min(4, n_samp / 8), (2 * ctx->n_sm) / std::max(1, P.n_rt)));
        if ((rc = ensure(ctx, ctx->samp, (size_t)n_rows * samp_chunks * 4 * TC_R * sizeof(float)))) return rc;
        if ((rc = ensure(ctx, ctx->cand, (size_t)n_rows * n_sub * cap * sizeof(uint2)))) return rc;
        if ((rc = ensure(ctx, ctx->thr, (size_t)4 * n_rows * sizeof(float)))) return rc;  // thr | TF32 hi | TF32 lo (threshold MMA) | retry threshold
        if ((rc = ensure(ctx, ctx->cand_cnt, (size_t)n_rows * n_sub * sizeof(uint32_t)))) return rc;
        int *fail_count = (int *)ctx->fail_list.p;             // [0] = count of the first attempt, [1] = count after the retry
        int32_t *fail_list = (int32_t *)ctx->fail_list.p + 2;  // [2 ..) rows of the first attempt, [2 + n_rows ..) rows left after the retry
        int *fail_count2 = fail_count + 1;
        int32_t *fail_list2 = fail_list + n_rows;

        TcArgs A;
        A.U = user_vecs_dev;
        A.ld_u = ld_u;
        A.err_flag = ctx->d_err;
        A.dbg = (int)ctx->opt_dbg;
        A.stride = stride;
        A.n_samp = n_samp;
        A.samp_chunks = samp_chunks;
        A.samp = (float *)ctx->samp.p;
        A.thr = (const float *)ctx->thr.p;
        A.cap = cap;
        A.cand = (uint2 *)ctx->cand.p;
        A.cand_cnt = (uint32_t *)ctx->cand_cnt.p;
        A.work = (const int4 *)wc->buf.p;
        if (wc->mixed)  // user tiles with fewer chunks leave their last sub-lists untouched: they must read as empty
            SKR_CUDA(ctx, cudaMemsetAsync(ctx->cand_cnt.p, 0, (size_t)n_rows * n_sub * sizeof(uint32_t), st));
        A.trace = nullptr;
        A.trace_cta = -1;
        A.trace_tiles = 0;
        A.retry_cnt = nullptr;
        A.retry_total = nullptr;
        A.retry_min = 0;
        const bool presub = (passes == 1);  // k_fused_tc's PRESUB: the threshold is subtracted by an extra MMA
        A.thr_hi = presub ? A.thr + n_rows : nullptr;
        A.thr_lo = presub ? A.thr + 2 * n_rows : nullptr;
        A.item_scale = half ? (const float *)ctx->f16s.p + 1 : nullptr;
        A.scale = half ? (const float *)ctx->rscale.p : nullptr;
        A.bias_shift = aug ? (const int *)ctx->f16s.p + 3 : nullptr;
        A.rowg = aug ? (const float *)ctx->rscale.p + n_rows : nullptr;
        FusedParams P1 = P;  // what the first attempt's kernels see: no bias to add when it is part of the contraction
        if (aug) P1.bias = nullptr;
        float eps_coef = 0.0f, eps3_coef = 0.0f;
        float *thr3 = nullptr, *eps2_3 = nullptr;
        if (rescore) {
            if ((rc = ensure(ctx, ctx->eps2, (size_t)2 * n_rows * sizeof(float)))) return rc;  // 2 eps | 2 eps of the 3-pass retry
            eps_coef = (float)(ldexp(1.0, -10) + (2.5 * d + 8.0) * ldexp(1.0, -22));  // fp16 rn: the same 2^-11 per operand as TF32 rna
            eps3_coef = (float)((3.25 * d + 11.0) * ldexp(1.0, -22));
            thr3 = (float *)ctx->thr.p + 3 * n_rows;
            eps2_3 = (float *)ctx->eps2.p + n_rows;
        }
        // pre-pass: thresholds from a strided sample of the item tiles, single TF32 pass
        SKR_CUDA(ctx, cudaEventRecord(ctx->ev2[slot], st));
        if ((rc = launch_tc(ctx, nkb_1, 1, TC_MODE_SAMPLE, (unsigned)(P.n_rt * samp_chunks), st, m1, maug, A, P1, half_mode))) return rc;
        SKR_AFTER(ctx, st, "k_fused_tc SAMPLE");
        {
            typedef void (*ThrKernel)(const float *, int64_t, int, float *, const float *, int64_t, int, const float *, float, float *, float *, float *, float,
                                      float *, float *, const float *, float *, const int *, float *);
            const ThrKernel thr_k = samp_chunks == 1 ? (ThrKernel)k_sample_thr<1> : samp_chunks == 2 ? (ThrKernel)k_sample_thr<2>
                                    : samp_chunks == 3 ? (ThrKernel)k_sample_thr<3> : (ThrKernel)k_sample_thr<4>;
            SKR_CUDA(ctx, launch_pdl(thr_k, dim3((unsigned)((n_rows + 7) / 8)), dim3(256u), (size_t)0, st, (const float *)ctx->samp.p, n_rows, r, (float *)ctx->thr.p, user_vecs_dev, ld_u, d,
                                     (const float *)stats_cur, eps_coef, rescore ? (float *)ctx->eps2.p : (float *)nullptr, (float *)A.thr_hi, (float *)A.thr_lo, eps3_coef, eps2_3,
                                     thr3, A.item_scale, (float *)ctx->rscale.p, A.bias_shift, (float *)A.rowg));
        }
        // main pass: every item tile, reference-grade scores, survivors to the candidate lists
        if (ctx->opt_trace_cta >= 0) {
            A.trace_tiles = wc->max_tiles;
            const size_t tb = (size_t)A.trace_tiles * TC_TRACE_SLOTS * sizeof(long long);
            if ((rc = ensure(ctx, ctx->trace, tb))) return rc;
            SKR_CUDA(ctx, cudaMemsetAsync(ctx->trace.p, 0, tb, st));
            A.trace = (long long *)ctx->trace.p;
            A.trace_cta = (int)ctx->opt_trace_cta;
        }
        SKR_CUDA(ctx, cudaEventRecord(ctx->ev0[slot], st));
        SKR_AFTER(ctx, st, "k_sample_thr");
        if ((rc = launch_tc(ctx, nkb_1, passes, TC_MODE_COLLECT, grid_tc, st, m1, maug, A, P1, half_mode))) return rc;
        SKR_AFTER(ctx, st, "k_fused_tc COLLECT");
        SKR_CUDA(ctx, cudaEventRecord(ctx->ev1[slot], st));
        ctx->launches += 3;
        SKR_CUDA(ctx, cudaGetLastError());
        ctx->last_fused = (passes == 3) ? "tcgen05_3xtf32" : (half ? "tcgen05_f16r" : (rescore ? "tcgen05_tf32r" : "tcgen05_1xtf32"));
        ctx->last_plan = {n_samp, stride, r, cap, P.S, passes == 3 ? 4 : 8};
        ctx->ev_calls++;
        const ExactArgs E = {user_vecs_dev, ld_u, item_vecs_dev, ld_i, d, bias_dev, (int)n_items, tp, ti};
        RescoreArgs RA = {nullptr, 0, nullptr, 0, 0, nullptr, nullptr, nullptr, nullptr, nullptr};
        if (rescore) RA = {user_vecs_dev, ld_u, item_vecs_dev, ld_i, d, bias_dev, (const float *)ctx->thr.p, (const float *)ctx->eps2.p, nullptr, nullptr};
        // The retry sweeps the item tiles of a user tile's work items in three passes on ONE SM each: a latency of
        // max_tiles x ~0.4 us per k-block and pass (c2: ~80 us, c4: 13 ms) however few rows it serves.  The per-row exact
        // kernel spreads a row over the GPU and costs ~10 us + 0.14 us per 10^6 item-table floats per row (measured: 11 us
        // at c2, ~28 us at c4).  The retry runs only when more rows than the ratio of the two are unsettled.
        const double sweep_us = (double)wc->max_tiles * (0.25 + 0.37 * nkb);
        const double exact_row_us = 10.0 + 1.4e-7 * (double)n_items * d;
        const int retry_min = ctx->opt_retry_min >= 0 ? (int)std::min<int64_t>(ctx->opt_retry_min, 1 << 30)
                                                      : (int)std::max(4.0, std::min(1.0e9, sweep_us / exact_row_us));
        // the retry of unsettled rows (tf32r only): lo table on demand, then the three-pass kernel over the same work list
        const std::function<int()> retry_collect = [&]() -> int {
            if (half)  // the first attempt built neither TF32 table
                SKR_CUDA(ctx, launch_pdl(k_split_hilo_if, dim3((unsigned)std::min<int64_t>((n_items + 7) / 8, 16 * ctx->n_sm)), dim3(256u), (size_t)0, st, (const int *)fail_count, retry_min,
                                         item_vecs_dev, ld_i, n_items, d, d_pad, (float *)ctx->bhi.p, (float *)ctx->blo.p));
            else
                SKR_CUDA(ctx, launch_pdl(k_split_lo_if, dim3((unsigned)std::min<int64_t>((n_items + 7) / 8, 16 * ctx->n_sm)), dim3(256u), (size_t)0, st, (const int *)fail_count, retry_min, item_vecs_dev, ld_i,
                                         n_items, d, d_pad, (float *)ctx->blo.p));
            TcArgs A2 = A;
            A2.item_scale = nullptr;
            A2.scale = nullptr;
            A2.bias_shift = nullptr;
            A2.rowg = nullptr;
            A2.thr = thr3;
            A2.thr_hi = nullptr;
            A2.thr_lo = nullptr;
            A2.retry_cnt = (const int *)ctx->rs_cnt.p;
            A2.retry_total = (const int *)fail_count;
            A2.retry_min = retry_min;
            A2.trace = nullptr;
            A2.trace_cta = -1;
            int rc2 = launch_tc(ctx, nkb, 3, TC_MODE_COLLECT, grid_tc, st, mhi, mlo, A2, P);
            if (rc2) return rc2;
            SKR_AFTER(ctx, st, "k_fused_tc COLLECT (retry)");
            return SKR_OK;
        };
        const RetryPlan RP = {rescore ? &retry_collect : nullptr, eps2_3, thr3, fail_list2, fail_count2, retry_min};
        rc = run_select_metrics(ctx, A.cand, A.cand_cnt, n_sub, cap, n_rows, row0, m, K, E, fail_list, fail_count, topk_idx_dev, topk_val_dev,
                                per_user_dev, sums_dev, keys_only, RA, presub ? A.thr : nullptr, RP, st);
        if (rc || keys_only == nullptr) return rc;
    } else {
        if ((rc = ensure(ctx, ctx->thr, (size_t)n_rows * sizeof(uint32_t)))) return rc;
        if ((rc = ensure(ctx, ctx->part, (size_t)n_rows * P.S * K * sizeof(u64)))) return rc;
        P.thr_g = (uint32_t *)ctx->thr.p;
        P.part = (u64 *)ctx->part.p;
        SKR_CUDA(ctx, cudaMemsetAsync(P.thr_g, 0, (size_t)n_rows * sizeof(uint32_t), st));
        const size_t smem = simt_smem_bytes(K);
        typedef void (*SimtKernel)(const float *, int64_t, const float *, int64_t, FusedParams, float *, int64_t);
        SimtKernel simt = ctx->opt_score_fn == 1 ? (SimtKernel)k_fused_simt<false, 1> : (SimtKernel)k_fused_simt<false, 0>;
        SKR_CUDA(ctx, cudaFuncSetAttribute(simt, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        SKR_CUDA(ctx, cudaEventRecord(ctx->ev2[slot], st));
        SKR_CUDA(ctx, cudaEventRecord(ctx->ev0[slot], st));
        simt<<<grid, SIMT_THREADS, smem, st>>>(user_vecs_dev, ld_u, item_vecs_dev, ld_i, P, nullptr, 0);
        SKR_CUDA(ctx, cudaEventRecord(ctx->ev1[slot], st));
        ctx->launches++;
        ctx->last_fused = ctx->opt_score_fn == 1 ? "simt_fp32_negl2" : "simt_fp32";
        SKR_CUDA(ctx, cudaGetLastError());
        // merge the S partial lists per row
        if ((rc = merge_lists(ctx, P.part, P.S, K, n_rows, row0, (int64_t)P.S * K, K, tp, ti, keys_only ? keys_only : keys, st))) return rc;
        ctx->last_plan = {0, 0, 0, 0, P.S, 0};
        ctx->ev_calls++;
        SKR_CUDA(ctx, cudaGetLastError());
        if (keys_only == nullptr) return run_metrics(ctx, keys, nullptr, n_rows, row0, m, K, topk_idx_dev, topk_val_dev, per_user_dev, sums_dev, st);
    }
    // keys only: shard-local item ids -> global
    if (item_offset != 0) {
        const int64_t nk = n_rows * K;
        k_offset_keys<<<(unsigned)((nk + 255) / 256), 256, 0, st>>>(keys_only, nk, (uint32_t)item_offset);
        ctx->launches++;
    }
    SKR_CUDA(ctx, cudaGetLastError());
    return SKR_OK;
}

// The fused pipeline over any number of rows: row chunks of at most `chunk_rows` (option "chunk_rows", default 131,072 =
// 1,024 user tiles) bound the workspace (candidate lists: 1 GB per chunk at c4 instead of 8 GB for 10^6 rows) while
// the item-side preparation is done once.  Chunks are near-equal multiples of 128 rows, so at most two shapes occur.
static int fused_pipeline(skr_ctx *ctx, const float *user_vecs_dev, int64_t n_rows, int64_t ld_u, const float *item_vecs_dev,
                          int64_t n_items, int64_t ld_i, int d, const float *bias_dev, int64_t row0, const MetricIds &m, int top_k,
                          int precision, int32_t *topk_idx_dev, float *topk_val_dev, float *per_user_dev, double *sums_dev,
                          u64 *keys_only, int64_t item_offset, void *stream)
{
    const int64_t max_rows = ctx->opt_chunk_rows > 0 ? ((ctx->opt_chunk_rows + TM - 1) / TM) * TM : (int64_t)131072;
    if (n_rows <= max_rows)
        return fused_chunk(ctx, user_vecs_dev, n_rows, ld_u, item_vecs_dev, n_items, ld_i, d, bias_dev, row0, m, top_k, precision, topk_idx_dev,
                           topk_val_dev, per_user_dev, sums_dev, keys_only, item_offset, stream, true);
    const int64_t n_chunks = (n_rows + max_rows - 1) / max_rows;
    const int64_t step = (((n_rows + n_chunks - 1) / n_chunks) + TM - 1) / TM * TM;
    const int MK = m.n * top_k;
    for (int64_t c0 = 0; c0 < n_rows; c0 += step) {
        const int64_t n = std::min(step, n_rows - c0);
        int rc = fused_chunk(ctx, user_vecs_dev + c0 * ld_u, n, ld_u, item_vecs_dev, n_items, ld_i, d, bias_dev, row0 + c0, m, top_k, precision,
                             topk_idx_dev ? topk_idx_dev + c0 * top_k : nullptr, topk_val_dev ? topk_val_dev + c0 * top_k : nullptr,
                             per_user_dev ? per_user_dev + c0 * MK : nullptr, sums_dev, keys_only ? keys_only + c0 * top_k : nullptr,
                             item_offset, stream, c0 == 0);
        if (rc) return rc;
    }
    return SKR_OK;
}

int skr_eval_fused(skr_ctx *ctx, const float *user_vecs_dev, int64_t n_rows, int64_t ld_u, const float *item_vecs_dev,
                   int64_t n_items, int64_t ld_i, int d, const float *bias_dev, int64_t row0, const int32_t *metric_ids,
                   int n_metrics, int top_k, int precision, int32_t *topk_idx_dev, float *topk_val_dev, float *per_user_dev,
                   double *sums_dev, void *stream)
{
    if (!ctx) return SKR_ERR_INVALID;
    MetricIds m;
    int rc = check_metrics(ctx, metric_ids, n_metrics, top_k, m);
    if (rc) return rc;
    return fused_pipeline(ctx, user_vecs_dev, n_rows, ld_u, item_vecs_dev, n_items, ld_i, d, bias_dev, row0, m, top_k, precision, topk_idx_dev,
                          topk_val_dev, per_user_dev, sums_dev, nullptr, 0, stream);
}

int skr_topk_fused(skr_ctx *ctx, const float *user_vecs_dev, int64_t n_rows, int64_t ld_u, const float *item_vecs_dev, int64_t n_items,
                   int64_t ld_i, int d, const float *bias_dev, int64_t row0, int64_t item_offset, int top_k, int precision,
                   uint64_t *keys_out_dev, void *stream)
{
    if (!ctx) return SKR_ERR_INVALID;
    if (!keys_out_dev) return fail(ctx, SKR_ERR_INVALID, "topk_fused: keys_out is NULL");
    if (item_offset < 0 || item_offset + n_items > 0xffffffffll) return fail(ctx, SKR_ERR_INVALID, "item_offset=%lld", (long long)item_offset);
    MetricIds m;
    m.n = 0;
    m.packed = 0u;
    return fused_pipeline(ctx, user_vecs_dev, n_rows, ld_u, item_vecs_dev, n_items, ld_i, d, bias_dev, row0, m, top_k, precision, nullptr, nullptr,
                          nullptr, nullptr, (u64 *)keys_out_dev, item_offset, stream);
}

int skr_eval_merged_topk(skr_ctx *ctx, const uint64_t *keys_all_dev, int n_shards, int64_t n_rows_total, int64_t row_begin, int64_t n_rows,
                         int64_t row0, const int32_t *metric_ids, int n_metrics, int top_k, int32_t *topk_idx_dev, float *topk_val_dev,
                         float *per_user_dev, double *sums_dev, void *stream)
{
    if (!ctx) return SKR_ERR_INVALID;
    MetricIds m;
    int rc = check_metrics(ctx, metric_ids, n_metrics, top_k, m);
    if (rc) return rc;
    if (!keys_all_dev || n_shards < 1 || n_rows <= 0 || row_begin < 0 || row_begin + n_rows > n_rows_total)
        return fail(ctx, SKR_ERR_INVALID, "eval_merged_topk: bad arguments");
    SKR_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    if ((rc = ensure(ctx, ctx->keys, (size_t)n_rows * top_k * sizeof(u64)))) return rc;
    // lists gathered from the shards: [n_shards][n_rows_total][K]; merge rows [row_begin, row_begin + n_rows)
    const u64 *src = (const u64 *)keys_all_dev + row_begin * (int64_t)top_k;
    if ((rc = merge_lists(ctx, src, n_shards, top_k, n_rows, row0, top_k, n_rows_total * (int64_t)top_k, nullptr, nullptr, (u64 *)ctx->keys.p, st)))
        return rc;
    return run_metrics(ctx, (const u64 *)ctx->keys.p, nullptr, n_rows, row0, m, top_k, topk_idx_dev, topk_val_dev, per_user_dev, sums_dev, st);
}

// ---- host-buffer variants: the copies are part of the call --------------------------------------
static int finish_host(skr_ctx *ctx, int64_t n_rows, int MK, int K, int32_t *topk_idx_host, float *per_user_host, double *sums_host,
                       const int32_t *d_idx, const float *d_pu, const double *d_sums, cudaStream_t st)
{
    if (topk_idx_host) SKR_CUDA(ctx, cudaMemcpyAsync(topk_idx_host, d_idx, (size_t)n_rows * K * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    if (per_user_host) SKR_CUDA(ctx, cudaMemcpyAsync(per_user_host, d_pu, (size_t)n_rows * MK * sizeof(float), cudaMemcpyDeviceToHost, st));
    // sums and the watchdog flag land in one pinned buffer: two asynchronous copies and a single synchronisation
    // (a pageable destination and a separate blocking copy of the flag cost two more round trips per call)
    if (ctx->h_pin == nullptr) SKR_CUDA(ctx, cudaMallocHost((void **)&ctx->h_pin, (size_t)SKR_PIN_DOUBLES * sizeof(double)));
    const bool pinned = MK + 1 <= SKR_PIN_DOUBLES;
    std::vector<double> tmp(pinned ? 0 : (size_t)MK);
    double *dst = pinned ? ctx->h_pin : tmp.data();
    int *flag_p = reinterpret_cast<int *>(ctx->h_pin + (SKR_PIN_DOUBLES - 1));
    if (sums_host) SKR_CUDA(ctx, cudaMemcpyAsync(dst, d_sums, (size_t)MK * sizeof(double), cudaMemcpyDeviceToHost, st));
    SKR_CUDA(ctx, cudaMemcpyAsync(flag_p, ctx->d_err, sizeof(int), cudaMemcpyDeviceToHost, st));
    SKR_CUDA(ctx, cudaStreamSynchronize(st));
    if (sums_host) for (int i = 0; i < MK; ++i) sums_host[i] += dst[i];
    if (*flag_p) return fail(ctx, SKR_ERR_CUDA, "kernel watchdog tripped (code %d)", *flag_p);
    return SKR_OK;
}

int skr_eval_scores_host(skr_ctx *ctx, const float *scores_host, int64_t n_rows, int64_t n_items, int64_t ld, int64_t row0,
                         const int32_t *metric_ids, int n_metrics, int top_k, int32_t *topk_idx_host, float *per_user_host,
                         double *sums_host, void *stream)
{
    if (!ctx) return SKR_ERR_INVALID;
    if (!scores_host || n_rows <= 0 || n_items <= 0 || ld < n_items) return fail(ctx, SKR_ERR_INVALID, "eval_scores_host: bad arguments");
    if (n_metrics < 1 || n_metrics > 8 || top_k < 1) return fail(ctx, SKR_ERR_INVALID, "eval_scores_host: bad metric list / top_k");
    SKR_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int MK = n_metrics * top_k;
    int rc;
    if ((rc = ensure(ctx, ctx->stage_a, (size_t)n_rows * n_items * sizeof(float)))) return rc;
    if ((rc = ensure(ctx, ctx->sums, (size_t)MK * sizeof(double)))) return rc;
    if (per_user_host && (rc = ensure(ctx, ctx->per_user, (size_t)n_rows * MK * sizeof(float)))) return rc;
    if ((rc = ensure(ctx, ctx->out_idx, (size_t)n_rows * top_k * sizeof(int32_t)))) return rc;
    SKR_CUDA(ctx, cudaMemcpy2DAsync(ctx->stage_a.p, (size_t)n_items * sizeof(float), scores_host, (size_t)ld * sizeof(float),
                                    (size_t)n_items * sizeof(float), (size_t)n_rows, cudaMemcpyHostToDevice, st));
    SKR_CUDA(ctx, cudaMemsetAsync(ctx->sums.p, 0, (size_t)MK * sizeof(double), st));
    rc = skr_eval_scores(ctx, (const float *)ctx->stage_a.p, n_rows, n_items, n_items, row0, metric_ids, n_metrics, top_k,
                         topk_idx_host ? (int32_t *)ctx->out_idx.p : nullptr, nullptr, per_user_host ? (float *)ctx->per_user.p : nullptr, (double *)ctx->sums.p, stream);
    if (rc) return rc;
    return finish_host(ctx, n_rows, MK, top_k, topk_idx_host, per_user_host, sums_host, (const int32_t *)ctx->out_idx.p,
                       (const float *)ctx->per_user.p, (const double *)ctx->sums.p, st);
}

int skr_eval_fused_host(skr_ctx *ctx, const float *user_vecs_host, int64_t n_rows, int64_t ld_u, const float *item_vecs_host,
                        int64_t n_items, int64_t ld_i, int d, const float *bias_host, int64_t row0, const int32_t *metric_ids,
                        int n_metrics, int top_k, int precision, int32_t *topk_idx_host, float *per_user_host, double *sums_host,
                        void *stream)
{
    if (!ctx) return SKR_ERR_INVALID;
    if (!user_vecs_host || !item_vecs_host || n_rows <= 0 || n_items <= 0 || d <= 0 || ld_u < d || ld_i < d)
        return fail(ctx, SKR_ERR_INVALID, "eval_fused_host: bad arguments");
    if (n_metrics < 1 || n_metrics > 8 || top_k < 1) return fail(ctx, SKR_ERR_INVALID, "eval_fused_host: bad metric list / top_k");
    SKR_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int MK = n_metrics * top_k;
    const int64_t dp = (d + 3) & ~3;  // device copies are packed with a 16-byte aligned row pitch
    int rc;
    if ((rc = ensure(ctx, ctx->stage_a, (size_t)n_rows * dp * sizeof(float)))) return rc;
    if ((rc = ensure(ctx, ctx->stage_b, (size_t)n_items * dp * sizeof(float)))) return rc;
    if (bias_host && (rc = ensure(ctx, ctx->stage_c, (size_t)n_items * sizeof(float)))) return rc;
    if ((rc = ensure(ctx, ctx->sums, (size_t)MK * sizeof(double)))) return rc;
    if (per_user_host && (rc = ensure(ctx, ctx->per_user, (size_t)n_rows * MK * sizeof(float)))) return rc;
    if ((rc = ensure(ctx, ctx->out_idx, (size_t)n_rows * top_k * sizeof(int32_t)))) return rc;
    if (dp != d) {
        SKR_CUDA(ctx, cudaMemsetAsync(ctx->stage_a.p, 0, (size_t)n_rows * dp * sizeof(float), st));
        SKR_CUDA(ctx, cudaMemsetAsync(ctx->stage_b.p, 0, (size_t)n_items * dp * sizeof(float), st));
    }
    // packed tables (the usual case) go as one flat copy each: a pitched copy of 256-byte rows is descriptor-bound
    if (ld_u == dp && d == dp)
        SKR_CUDA(ctx, cudaMemcpyAsync(ctx->stage_a.p, user_vecs_host, (size_t)n_rows * dp * sizeof(float), cudaMemcpyHostToDevice, st));
    else
        SKR_CUDA(ctx, cudaMemcpy2DAsync(ctx->stage_a.p, (size_t)dp * sizeof(float), user_vecs_host, (size_t)ld_u * sizeof(float),
                                        (size_t)d * sizeof(float), (size_t)n_rows, cudaMemcpyHostToDevice, st));
    if (ld_i == dp && d == dp)
        SKR_CUDA(ctx, cudaMemcpyAsync(ctx->stage_b.p, item_vecs_host, (size_t)n_items * dp * sizeof(float), cudaMemcpyHostToDevice, st));
    else
        SKR_CUDA(ctx, cudaMemcpy2DAsync(ctx->stage_b.p, (size_t)dp * sizeof(float), item_vecs_host, (size_t)ld_i * sizeof(float),
                                        (size_t)d * sizeof(float), (size_t)n_items, cudaMemcpyHostToDevice, st));
    if (bias_host) SKR_CUDA(ctx, cudaMemcpyAsync(ctx->stage_c.p, bias_host, (size_t)n_items * sizeof(float), cudaMemcpyHostToDevice, st));
    SKR_CUDA(ctx, cudaMemsetAsync(ctx->sums.p, 0, (size_t)MK * sizeof(double), st));
    rc = skr_eval_fused(ctx, (const float *)ctx->stage_a.p, n_rows, dp, (const float *)ctx->stage_b.p, n_items, dp, (int)dp,
                        bias_host ? (const float *)ctx->stage_c.p : nullptr, row0, metric_ids, n_metrics, top_k, precision,
                        topk_idx_host ? (int32_t *)ctx->out_idx.p : nullptr, nullptr, per_user_host ? (float *)ctx->per_user.p : nullptr, (double *)ctx->sums.p, stream);
    if (rc) return rc;
    return finish_host(ctx, n_rows, MK, top_k, topk_idx_host, per_user_host, sums_host, (const int32_t *)ctx->out_idx.p,
                       (const float *)ctx->per_user.p, (const double *)ctx->sums.p, st);
}

}  // extern "C"

extern "C" int skr_batch_randint(skr_ctx *ctx, int64_t high, const int64_t *out_indptr_dev, int64_t n_batch, int64_t n_out, int replace,
                                 const float *cdf_dev, int cdf_per_row, const int64_t *excl_indptr_dev, const int32_t *excl_idx_dev,
                                 uint64_t seed, int32_t *out_dev, void *stream)
{
    if (!ctx) return SKR_ERR_INVALID;
    if (high <= 1 || high > 0x7fffffffll) return fail(ctx, SKR_ERR_INVALID, "'high' must be larger than 1 (and fit int32): %lld", (long long)high);
    if (!out_indptr_dev || !out_dev || n_batch <= 0 || n_out < 0) return fail(ctx, SKR_ERR_INVALID, "batch_randint: bad arguments");
    if ((excl_indptr_dev == nullptr) != (excl_idx_dev == nullptr)) return fail(ctx, SKR_ERR_INVALID, "batch_randint: exclusion CSR needs both arrays");
    if (n_out == 0) return SKR_OK;
    SKR_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    SamplerArgs A = {high, out_indptr_dev, n_batch, cdf_dev, cdf_per_row, excl_indptr_dev, excl_idx_dev, (uint32_t)seed, (uint32_t)(seed >> 32), out_dev, ctx->d_err};
    if (replace) k_sample_with_replacement<<<(unsigned)((n_out + 255) / 256), 256, 0, st>>>(A, n_out);
    else k_sample_without_replacement<<<(unsigned)((n_batch + 3) / 4), 128, 0, st>>>(A);
    ctx->launches++;
    SKR_CUDA(ctx, cudaGetLastError());
    return SKR_OK;
}

/* 0, or SKR_ERR_CUDA once a kernel of this ctx has reported a condition it could not handle (a watchdog, a sampler row
 * that cannot be filled); synchronises the device and clears the flag. */
extern "C" int skr_check(skr_ctx *ctx)
{
    if (!ctx) return SKR_ERR_INVALID;
    int v = 0;
    SKR_CUDA(ctx, cudaSetDevice(ctx->device));
    SKR_CUDA(ctx, cudaMemcpy(&v, ctx->d_err, sizeof(int), cudaMemcpyDeviceToHost));
    if (v) {
        cudaMemset(ctx->d_err, 0, sizeof(int));
        return fail(ctx, SKR_ERR_CUDA, v == 41 ? "sampler: a row asks for more values than exist outside its exclusion set (code 41)" : "a kernel reported error code %d", v);
    }
    return SKR_OK;
}

// ---- one-shot all-reduce of the metric sums over NVLink (user-sharded evaluation exchanges M*K + 1 doubles) ----
// Every rank owns an inbox [2 parities][world][SKR_COMM_MAX doubles] + flags in its own HBM, exported with cudaIpc and
// mapped by its peers.  One kernel per rank: push my vector into slot `rank` of every rank's inbox (plain stores over
// NVLink), system fence, publish a sequence number in the peers' flag words, wait (bounded) for all `world` flags of my
// own inbox, then add the world vectors IN RANK ORDER -- every rank gets the same bits, no reduction tree, no NCCL
// launch: ~5 us instead of ~25-35 us for a 1-4 KB ncclAllReduce.  Two parities: a rank can be at most one call ahead
// of the slowest (it needs that rank's flag of the current call to finish it).
#define SKR_COMM_MAX 4096
#define SKR_COMM_MAX_WORLD 16

struct skr_comm {
    int device = 0, rank = 0, world = 1;
    unsigned char *local = nullptr;                 // my inbox + flags
    unsigned char *peer[SKR_COMM_MAX_WORLD] = {};   // everybody's (peer[rank] == local)
    bool opened[SKR_COMM_MAX_WORLD] = {};
    uint32_t seq = 0;
    int *d_err = nullptr;
    std::string err;
};

namespace {

__host__ __device__ constexpr size_t comm_data_bytes(int world) { return (size_t)2 * world * SKR_COMM_MAX * sizeof(double); }
__host__ __device__ constexpr size_t comm_total_bytes(int world) { return comm_data_bytes(world) + (size_t)2 * world * sizeof(uint32_t) + 256; }

struct CommPeers { unsigned char *p[SKR_COMM_MAX_WORLD]; };

__global__ void __launch_bounds__(256)
k_allreduce_oneshot(double *__restrict__ vec, int n, CommPeers P, int rank, int world, int parity, uint32_t seq, int *err)
{
    pdl_wait();
    pdl_trigger();
    const int tid = threadIdx.x;
    const size_t slot = ((size_t)parity * world + rank) * SKR_COMM_MAX;
    for (int p = 0; p < world; ++p) {
        double *dst = reinterpret_cast<double *>(P.p[p]) + slot;
        for (int i = tid; i < n; i += 256) dst[i] = vec[i];
    }
    __threadfence_system();
    __syncthreads();
    if (tid < world) {
        uint32_t *flag = reinterpret_cast<uint32_t *>(P.p[tid] + comm_data_bytes(world)) + (size_t)parity * world + rank;
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(seq) : "memory");
        const uint32_t *mine = reinterpret_cast<const uint32_t *>(P.p[rank] + comm_data_bytes(world)) + (size_t)parity * world + tid;
        const long long t0 = clock64();
        for (;;) {
            uint32_t v;
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
            if (v == seq) break;
            if (clock64() - t0 > 6000000000ll) {  // ~3 s: a rank that never arrives must not hang the GPU
                atomicExch(err, 31);
                break;
            }
        }
    }
    __syncthreads();
    const double *in = reinterpret_cast<const double *>(P.p[rank]) + (size_t)parity * world * SKR_COMM_MAX;
    for (int i = tid; i < n; i += 256) {
        double s = 0.0;
        for (int r = 0; r < world; ++r) s += __ldcg(in + (size_t)r * SKR_COMM_MAX + i);
        vec[i] = s;
    }
}

}  // namespace

extern "C" {

int skr_comm_create(int device, int rank, int world, skr_comm **out)
{
    if (!out) return SKR_ERR_INVALID;
    *out = nullptr;
    if (world < 1 || world > SKR_COMM_MAX_WORLD || rank < 0 || rank >= world) return fail(nullptr, SKR_ERR_INVALID, "comm: rank %d / world %d", rank, world);
    skr_comm *c = new (std::nothrow) skr_comm();
    if (!c) return fail(nullptr, SKR_ERR_NOMEM, "host allocation failed");
    c->device = device; c->rank = rank; c->world = world;
    cudaError_t e;
    if ((e = cudaSetDevice(device)) != cudaSuccess || (e = cudaMalloc((void **)&c->local, comm_total_bytes(world))) != cudaSuccess ||
        (e = cudaMemset(c->local, 0, comm_total_bytes(world))) != cudaSuccess || (e = cudaMalloc((void **)&c->d_err, sizeof(int))) != cudaSuccess ||
        (e = cudaMemset(c->d_err, 0, sizeof(int))) != cudaSuccess) {
        if (c->local) cudaFree(c->local);
        delete c;
        return fail(nullptr, SKR_ERR_CUDA, "comm: %s", cudaGetErrorString(e));
    }
    c->peer[rank] = c->local;
    *out = c;
    return SKR_OK;
}

int skr_comm_handle(skr_comm *c, void *handle_out64)
{
    if (!c || !handle_out64) return SKR_ERR_INVALID;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "ipc handle size");
    cudaSetDevice(c->device);
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, c->local);
    if (e != cudaSuccess) { c->err = std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(e); return SKR_ERR_CUDA; }
    memcpy(handle_out64, &h, 64);
    return SKR_OK;
}

int skr_comm_connect(skr_comm *c, const void *handles /* world x 64 bytes, rank order */)
{
    if (!c || !handles) return SKR_ERR_INVALID;
    cudaSetDevice(c->device);
    for (int r = 0; r < c->world; ++r) {
        if (r == c->rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, (const unsigned char *)handles + (size_t)r * 64, 64);
        void *p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) { c->err = std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e); cudaGetLastError(); return SKR_ERR_CUDA; }
        c->peer[r] = (unsigned char *)p;
        c->opened[r] = true;
    }
    return SKR_OK;
}

int skr_comm_allreduce(skr_comm *c, double *vec_dev, int n, void *stream)
{
    if (!c || !vec_dev || n < 1 || n > SKR_COMM_MAX) return SKR_ERR_INVALID;
    for (int r = 0; r < c->world; ++r)
        if (!c->peer[r]) { c->err = "comm: not connected"; return SKR_ERR_STATE; }
    cudaSetDevice(c->device);
    CommPeers P;
    for (int r = 0; r < SKR_COMM_MAX_WORLD; ++r) P.p[r] = c->peer[r < c->world ? r : 0];
    c->seq++;
    cudaError_t e = launch_pdl(k_allreduce_oneshot, dim3(1), dim3(256), (size_t)0, (cudaStream_t)stream, vec_dev, n, P, c->rank, c->world, (int)(c->seq & 1u),
                               c->seq, c->d_err);
    if (e != cudaSuccess) { c->err = std::string("k_allreduce_oneshot: ") + cudaGetErrorString(e); return SKR_ERR_CUDA; }
    return SKR_OK;
}

/* 0 = no rank ever timed out in a collective of this communicator (reads a device word: synchronises the stream's device) */
int skr_comm_status(skr_comm *c)
{
    if (!c) return SKR_ERR_INVALID;
    int v = 0;
    cudaSetDevice(c->device);
    if (cudaMemcpy(&v, c->d_err, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return SKR_ERR_CUDA;
    if (v) { c->err = "comm: a rank did not arrive within the time limit"; return SKR_ERR_CUDA; }
    return SKR_OK;
}

const char *skr_comm_last_error(const skr_comm *c) { return c ? c->err.c_str() : ""; }

int skr_comm_destroy(skr_comm *c)
{
    if (!c) return SKR_OK;
    cudaSetDevice(c->device);
    for (int r = 0; r < c->world; ++r)
        if (c->opened[r]) cudaIpcCloseMemHandle(c->peer[r]);
    if (c->local) cudaFree(c->local);
    if (c->d_err) cudaFree(c->d_err);
    delete c;
    return SKR_OK;
}

}  // extern "C"
