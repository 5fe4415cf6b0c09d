// k_metrics.cuh -- K3 (partial top-K merge) and K4 (hit test, per-user metrics, column sums).
//
// The metric recurrences restate metric.h:19-109 operation by operation in float32 (NDCG adds
// a double term into a float, MRR rounds a double quotient), so per-user vectors are
// bit-identical to the reference's C++ for the same rank list.  1/log2(i+2) comes from a table
// computed on the host with the same libm the reference links (glibc log2).
#pragma once
#include "common.cuh"

namespace skr {

struct MetricIds {
    int n;
    int id[8];
};

constexpr int K4_WARPS = 4;

// ---- K4: one warp per user ---------------------------------------------------------------
// keys: sorted rank keys [n_rows, K] (or null), idx_in: int32 rank lists [n_rows, K] (or null).
__global__ void __launch_bounds__(K4_WARPS * 32)
k_metrics(const u64 *__restrict__ keys, const int32_t *__restrict__ idx_in, int K, int64_t n_rows,
          int64_t row0, const int64_t *__restrict__ te_indptr, const int32_t *__restrict__ te_idx,
          MetricIds mids, const double *__restrict__ disc, float *__restrict__ per_user,
          int32_t *__restrict__ topk_idx_out, float *__restrict__ topk_val_out)
{
    extern __shared__ unsigned char k4_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int M = mids.n;
    const int hit_words = (K + 31) >> 5;
    float *outs = reinterpret_cast<float *>(k4_smem) + (size_t)warp * (size_t)(M * K);
    uint32_t *hits = reinterpret_cast<uint32_t *>(reinterpret_cast<float *>(k4_smem) + (size_t)K4_WARPS * (size_t)(M * K)) +
                     warp * hit_words;

    const int64_t row = (int64_t)blockIdx.x * K4_WARPS + warp;
    if (row >= n_rows) return;

    const int64_t tb = __ldg(te_indptr + row0 + row);
    const int nt = (int)(__ldg(te_indptr + row0 + row + 1) - tb);
    const int32_t *truth = te_idx + tb;

    for (int i0 = 0; i0 < K; i0 += 32) {
        const int i = i0 + lane;
        bool hit = false;
        if (i < K) {
            int32_t item;
            if (keys != nullptr) {
                u64 k = keys[row * K + i];
                item = (k == 0) ? -1 : (int32_t)key_item(k);
                if (topk_val_out != nullptr) topk_val_out[row * K + i] = (k == 0) ? -__int_as_float(0x7f800000) : key_score(k);
            } else {
                item = idx_in[row * K + i];
            }
            if (topk_idx_out != nullptr) topk_idx_out[row * K + i] = item;
            hit = (item >= 0) && sorted_contains(truth, nt, item);
        }
        unsigned bal = __ballot_sync(0xffffffffu, hit);
        if (lane == 0) hits[i0 >> 5] = bal;
    }
    __syncwarp();

    if (lane < M) {
        float *o = outs + lane * K;
        const int L = nt > 1 ? nt : 1;
        const int id = mids.id[lane];
        if (id == 1) {  // metric.h:19-30
            float h = 0.0f;
            for (int i = 0; i < K; ++i) {
                if ((hits[i >> 5] >> (i & 31)) & 1u) h += 1.0f;
                o[i] = h / (float)(i + 1);
            }
        } else if (id == 2) {  // metric.h:33-45
            float h = 0.0f;
            const float tl = (float)L;
            for (int i = 0; i < K; ++i) {
                if ((hits[i >> 5] >> (i & 31)) & 1u) h += 1.0f;
                o[i] = h / tl;
            }
        } else if (id == 3) {  // metric.h:48-66
            float h = 0.0f, sum_pre = 0.0f;
            for (int i = 0; i < K; ++i) {
                if ((hits[i >> 5] >> (i & 31)) & 1u) {
                    h += 1.0f;
                    float pre = h / (float)(i + 1);
                    sum_pre += pre;
                }
                float den = (float)(L < i + 1 ? L : i + 1);
                o[i] = sum_pre / den;
            }
        } else if (id == 4) {  // metric.h:69-86
            float idcg = 0.0f, dcg = 0.0f;
            for (int i = 0; i < K; ++i) {
                const double t = disc[i];
                if ((hits[i >> 5] >> (i & 31)) & 1u) dcg = (float)((double)dcg + t);
                if (i < L) idcg = (float)((double)idcg + t);
                o[i] = dcg / idcg;
            }
        } else {  // id == 5, metric.h:89-109
            float rr = 0.0f;
            bool found = false;
            for (int i = 0; i < K; ++i) {
                if (!found && ((hits[i >> 5] >> (i & 31)) & 1u)) {
                    rr = (float)(1.0 / (double)(i + 1));
                    found = true;
                }
                o[i] = rr;
            }
        }
    }
    __syncwarp();
    const int MK = M * K;
    for (int c = lane; c < MK; c += 32) per_user[row * MK + c] = outs[c];
}

// ---- column sums, deterministic two-stage float64 -------------------------------------------
__global__ void k_colsum_partial(const float *__restrict__ per_user, int64_t n_rows, int n_cols,
                                 double *__restrict__ partial)
{
    for (int c = threadIdx.x; c < n_cols; c += blockDim.x) {
        double acc = 0.0;
        for (int64_t r = blockIdx.x; r < n_rows; r += gridDim.x) acc += (double)per_user[r * n_cols + c];
        partial[(size_t)blockIdx.x * n_cols + c] = acc;
    }
}

__global__ void k_colsum_final(const double *__restrict__ partial, int n_blk, int n_cols, double *__restrict__ sums)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_cols) return;
    double acc = 0.0;
    for (int b = 0; b < n_blk; ++b) acc += partial[(size_t)b * n_cols + c];
    sums[c] += acc;
}

// numpy's float32 np.sum(axis=0): row after row into a float32 accumulator (evaluator.py:208).
__global__ void k_colsum_f32_seq(const float *__restrict__ per_user, int64_t n_rows, int n_cols, float *__restrict__ acc_io)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_cols) return;
    float acc = acc_io[c];
    int64_t r = 0;
    for (; r + 8 <= n_rows; r += 8) {
        float x[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) x[q] = per_user[(r + q) * n_cols + c];
#pragma unroll
        for (int q = 0; q < 8; ++q) acc = acc + x[q];
    }
    for (; r < n_rows; ++r) acc = acc + per_user[r * n_cols + c];
    acc_io[c] = acc;
}

// ---- K3: merge S partial lists of K keys per row into one sorted top-K -----------------------
// part: [n_rows, S, K] keys in any order, 0 = empty.  One warp per row, 32*PER >= S*K.
// Rows that end up with fewer than K valid keys (fewer than K unmasked items) are completed
// with the row's train items in ascending id order at score -inf: those are exactly the
// candidates the reference ranks last (evaluator.py:195-200).
template <int PER>
__global__ void __launch_bounds__(128)
k_merge_partials(const u64 *__restrict__ part, int S, int K, int64_t n_rows, int64_t row0,
                 const int64_t *__restrict__ tr_indptr, const int32_t *__restrict__ tr_idx,
                 u64 *__restrict__ out_keys)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * 4 + warp;
    if (row >= n_rows) return;
    const int n = S * K;
    const u64 *src = part + row * (int64_t)n;
    u64 v[PER];
#pragma unroll
    for (int e = 0; e < PER; ++e) {
        int i = e * 32 + lane;
        v[e] = (i < n) ? src[i] : 0ull;
    }
    warp_bitonic_desc<PER>(v, lane);
    u64 *dst = out_keys + row * (int64_t)K;
    int n_valid = 0;
#pragma unroll
    for (int e = 0; e < PER; ++e) {
        int i = e * 32 + lane;
        if (i < K) dst[i] = v[e];
        unsigned bal = __ballot_sync(0xffffffffu, v[e] != 0ull);
        n_valid += __popc(bal);
    }
    if (n_valid < K && tr_indptr != nullptr) {
        __syncwarp();
        const int64_t tb = __ldg(tr_indptr + row0 + row), te = __ldg(tr_indptr + row0 + row + 1);
        const float ninf = -__int_as_float(0x7f800000);
        for (int64_t p = tb + lane; p < te && n_valid + (p - tb) < K; p += 32)
            dst[n_valid + (p - tb)] = make_key(ninf, (uint32_t)__ldg(tr_idx + p));
    }
}

}  // namespace skr
