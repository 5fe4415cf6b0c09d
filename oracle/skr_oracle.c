/*
 * skr_oracle.c -- CPU restatement of scikit-recommender's full-ranking evaluation path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it,
 * and only as the checker.  The product path (scikit-recommender_b200/) never imports it.
 *
 * Parity pin: this restatement is checked (tests/test_oracle.py) against
 *   - the hand-checked known-answer vector of SURVEY.md App. A.7,
 *   - tests/golden/*.npz, produced by the *compiled, unmodified* reference
 *     (oracle/build_ref.py -> oracle/_ref, script tests/golden/make_golden.py),
 *   - oracle/_ref itself wherever it is present (this container and the GPU box).
 * The reference's own test-suite holds no vector for this path.
 *
 * What each function follows (paths relative to /root/reference):
 *   skr_oracle_topk_row      skrec/utils/py/cython/include/evaluate.h:24-45 (selection)
 *   metric_* / skr_oracle_metrics_row   .../include/metric.h:19-109 (formulas, op order, dtypes)
 *   skr_oracle_mask_rows     skrec/utils/py/evaluator.py:195-200 (train items -> -inf)
 *   skr_oracle_eval_scores   .../pyx_eval_matrix.pyx:22-37 + evaluate.h:57-76 (row loop, layout)
 *   skr_oracle_mean_f32      skrec/utils/py/evaluator.py:206-208 (np.mean axis 0, float32)
 *   skr_oracle_scores        the `predict` forms of SURVEY.md section 2.3 (U_b @ I^T + b)
 *
 * One deliberate difference, documented in SURVEY.md App. A.4: on *tied* scores the reference
 * order is an artefact of libstdc++'s heap; the contract here is score desc, item id asc.
 * On rows whose top-2K scores are pairwise distinct both give the same list, bit for bit.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ---- ordering: a is ranked before b iff score higher, or equal score and lower id -------- */
static inline int ranks_before(float sa, int32_t ia, float sb, int32_t ib)
{
    if (sa > sb) return 1;
    if (sa < sb) return 0;
    return ia < ib;
}

/* Min-heap over (score, id) whose root is the *worst* kept element. */
static void sift_down(float *hs, int32_t *hi, int n, int p)
{
    for (;;) {
        int l = 2 * p + 1, r = l + 1, w = p;
        if (l < n && ranks_before(hs[w], hi[w], hs[l], hi[l])) w = l;
        if (r < n && ranks_before(hs[w], hi[w], hs[r], hi[r])) w = r;
        if (w == p) return;
        float ts = hs[p]; hs[p] = hs[w]; hs[w] = ts;
        int32_t ti = hi[p]; hi[p] = hi[w]; hi[w] = ti;
        p = w;
    }
}

/* evaluate.h:24-45 -- ids of the K best of `scores[0..n)`, best first.  Returns 0, or -1 if n<k
 * (the reference reads out of bounds there, evaluate.h:45; the restatement refuses). */
int skr_oracle_topk_row(const float *scores, int64_t n, int k, int32_t *out_idx, float *out_val)
{
    if (n < k || k <= 0) return -1;
    float *hs = (float *)malloc(sizeof(float) * (size_t)k);
    int32_t *hi = (int32_t *)malloc(sizeof(int32_t) * (size_t)k);
    int cnt = 0;
    for (int64_t j = 0; j < n; ++j) {
        float s = scores[j];
        if (s != s) s = -INFINITY;            /* NaN is undefined in the reference; rank it last */
        if (cnt < k) {
            hs[cnt] = s; hi[cnt] = (int32_t)j; ++cnt;
            if (cnt == k) for (int p = k / 2 - 1; p >= 0; --p) sift_down(hs, hi, k, p);
        } else if (ranks_before(s, (int32_t)j, hs[0], hi[0])) {
            hs[0] = s; hi[0] = (int32_t)j;
            sift_down(hs, hi, k, 0);
        }
    }
    /* heap-sort: repeatedly move the worst to the end -> best first */
    for (int m = k - 1; m > 0; --m) {
        float ts = hs[0]; hs[0] = hs[m]; hs[m] = ts;
        int32_t ti = hi[0]; hi[0] = hi[m]; hi[m] = ti;
        sift_down(hs, hi, m, 0);
    }
    for (int i = 0; i < k; ++i) {
        out_idx[i] = hi[i];
        if (out_val) out_val[i] = hs[i];
    }
    free(hs); free(hi);
    return 0;
}

/* ---- truth set: sorted, de-duplicated int32 array (unordered_set<int> in the reference) --- */
static int in_truth(const int32_t *t, int n, int32_t x)
{
    int lo = 0, hi = n;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (t[mid] < x) lo = mid + 1; else hi = mid;
    }
    return lo < n && t[lo] == x;
}

static int cmp_i32(const void *a, const void *b)
{
    int32_t x = *(const int32_t *)a, y = *(const int32_t *)b;
    return (x > y) - (x < y);
}

/* metric.h:19-30 */
static void metric_precision(const int32_t *rank, int k, const int32_t *t, int nt, float *out)
{
    float hits = 0.0f;
    for (int i = 0; i < k; ++i) {
        if (in_truth(t, nt, rank[i])) hits += 1.0f;
        out[i] = hits / (float)(i + 1);
    }
}

/* metric.h:33-45 */
static void metric_recall(const int32_t *rank, int k, const int32_t *t, int nt, float *out)
{
    float hits = 0.0f;
    float truth_len = (float)(nt > 1 ? nt : 1);
    for (int i = 0; i < k; ++i) {
        if (in_truth(t, nt, rank[i])) hits += 1.0f;
        out[i] = hits / truth_len;
    }
}

/* metric.h:48-66 */
static void metric_ap(const int32_t *rank, int k, const int32_t *t, int nt, float *out)
{
    float hits = 0.0f, pre = 0.0f, sum_pre = 0.0f, denominator = 1.0f;
    int truth_len = nt > 1 ? nt : 1;
    for (int i = 0; i < k; ++i) {
        if (in_truth(t, nt, rank[i])) {
            hits += 1.0f;
            pre = hits / (float)(i + 1);
            sum_pre += pre;
        }
        denominator = (float)(truth_len < i + 1 ? truth_len : i + 1);
        out[i] = sum_pre / denominator;
    }
}

/* metric.h:69-86 -- `DCG += 1.0/log2(i+2)` is float <- (double)float + double */
static void metric_ndcg(const int32_t *rank, int k, const int32_t *t, int nt, float *out)
{
    float iDCG = 0.0f, DCG = 0.0f;
    unsigned truth_len = (unsigned)(nt > 1 ? nt : 1);
    for (unsigned i = 0; i < (unsigned)k; ++i) {
        if (in_truth(t, nt, rank[i])) DCG = (float)((double)DCG + 1.0 / log2((double)(i + 2)));
        if (i < truth_len) iDCG = (float)((double)iDCG + 1.0 / log2((double)(i + 2)));
        out[i] = DCG / iDCG;
    }
}

/* metric.h:89-109 */
static void metric_mrr(const int32_t *rank, int k, const int32_t *t, int nt, float *out)
{
    for (int i = 0; i < k; ++i) {
        if (in_truth(t, nt, rank[i])) {
            float rr = (float)(1.0 / (double)(i + 1));
            for (int j = i; j < k; ++j) out[j] = rr;
            return;
        }
        out[i] = 0.0f;
    }
}

/* metric.h:112-118 dispatch + evaluate.h:47-51 layout [m0@1..K | m1@1..K | ...].
 * `truth` may be unsorted and hold duplicates (the reference's set de-duplicates). */
int skr_oracle_metrics_row(const int32_t *rank, int k, const int32_t *truth, int n_truth,
                           const int32_t *metric_ids, int n_metrics, float *out)
{
    int32_t *t = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n_truth > 0 ? n_truth : 1));
    int nt = 0;
    if (n_truth > 0) {
        memcpy(t, truth, sizeof(int32_t) * (size_t)n_truth);
        qsort(t, (size_t)n_truth, sizeof(int32_t), cmp_i32);
        for (int i = 0; i < n_truth; ++i)
            if (nt == 0 || t[i] != t[nt - 1]) t[nt++] = t[i];
    }
    for (int m = 0; m < n_metrics; ++m) {
        float *o = out + (size_t)m * (size_t)k;
        switch (metric_ids[m]) {
        case 1: metric_precision(rank, k, t, nt, o); break;
        case 2: metric_recall(rank, k, t, nt, o); break;
        case 3: metric_ap(rank, k, t, nt, o); break;
        case 4: metric_ndcg(rank, k, t, nt, o); break;
        case 5: metric_mrr(rank, k, t, nt, o); break;
        default: free(t); return -2;
        }
    }
    free(t);
    return 0;
}

/* evaluator.py:195-200 -- scores[r, train_items[r]] = -inf, in place. */
int skr_oracle_mask_rows(float *scores, int64_t n_rows, int64_t n_items, int64_t ld,
                         const int64_t *indptr, const int32_t *indices)
{
    for (int64_t r = 0; r < n_rows; ++r)
        for (int64_t p = indptr[r]; p < indptr[r + 1]; ++p) {
            int32_t j = indices[p];
            if (j < 0 || j >= n_items) return -3;
            scores[r * ld + j] = -INFINITY;
        }
    return 0;
}

/* pyx_eval_matrix.pyx:22-37 + evaluate.h:57-76 -- one row per user; out is [B, M*K]. */
int skr_oracle_eval_scores(const float *scores, int64_t n_rows, int64_t n_items, int64_t ld,
                           const int64_t *test_indptr, const int32_t *test_indices,
                           const int32_t *metric_ids, int n_metrics, int top_k,
                           float *out, int32_t *topk_idx_or_null)
{
    int32_t *rank = (int32_t *)malloc(sizeof(int32_t) * (size_t)top_k);
    for (int64_t r = 0; r < n_rows; ++r) {
        int rc = skr_oracle_topk_row(scores + r * ld, n_items, top_k, rank, NULL);
        if (rc) { free(rank); return rc; }
        if (topk_idx_or_null) memcpy(topk_idx_or_null + r * top_k, rank, sizeof(int32_t) * (size_t)top_k);
        rc = skr_oracle_metrics_row(rank, top_k, test_indices + test_indptr[r],
                                    (int)(test_indptr[r + 1] - test_indptr[r]),
                                    metric_ids, n_metrics, out + r * (int64_t)n_metrics * top_k);
        if (rc) { free(rank); return rc; }
    }
    free(rank);
    return 0;
}

/* `predict` of the dot-product models (BPRMF.py:84-88, LightGCN.py:102-107, ...):
 * S = U_b I^T (+ b).  Products are accumulated in double and rounded once to float32
 * ("FP64-exact -> f32" in SURVEY.md App. A.6); the bias is then added in float32 like torch. */
int skr_oracle_scores(const float *user_vecs, int64_t n_rows, int64_t ld_u,
                      const float *item_vecs, int64_t n_items, int64_t ld_i, int d,
                      const float *bias_or_null, float *scores, int64_t ld)
{
    for (int64_t r = 0; r < n_rows; ++r) {
        const float *u = user_vecs + r * ld_u;
        for (int64_t j = 0; j < n_items; ++j) {
            const float *v = item_vecs + j * ld_i;
            double acc = 0.0;
            for (int k = 0; k < d; ++k) acc += (double)u[k] * (double)v[k];
            float s = (float)acc;
            if (bias_or_null) s = s + bias_or_null[j];
            scores[r * ld + j] = s;
        }
    }
    return 0;
}

/* evaluator.py:206-208 -- np.mean(all_results, axis=0) on a C-ordered float32 [U, C] array:
 * numpy reduces axis 0 by adding row after row into a float32 vector, then divides by
 * float32(U) (checked against numpy 2.3 in tests/test_oracle.py). */
int skr_oracle_mean_f32(const float *per_user, int64_t n_rows, int64_t n_cols, float *out)
{
    for (int64_t c = 0; c < n_cols; ++c) out[c] = 0.0f;
    for (int64_t r = 0; r < n_rows; ++r)
        for (int64_t c = 0; c < n_cols; ++c) out[c] = out[c] + per_user[r * n_cols + c];
    for (int64_t c = 0; c < n_cols; ++c) out[c] = out[c] / (float)n_rows;
    return 0;
}

/* float64 column sums: what the GPU path reduces (and all-reduces across ranks). */
int skr_oracle_sums_f64(const float *per_user, int64_t n_rows, int64_t n_cols, double *out)
{
    for (int64_t c = 0; c < n_cols; ++c) out[c] = 0.0;
    for (int64_t r = 0; r < n_rows; ++r)
        for (int64_t c = 0; c < n_cols; ++c) out[c] += (double)per_user[r * n_cols + c];
    return 0;
}
