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
    uint32_t packed;  // metric id of column block mi in bits [4 mi, 4 mi + 4): a register, not an indexed array
    __host__ __device__ __forceinline__ int id(int mi) const { return (int)((packed >> (4 * mi)) & 15u); }
};

constexpr int K4_WARPS = 8;

// ---- per-user metrics, one warp per user, every lane owns rank positions i = lane, lane + 32, ... ---
//
// The reference's recurrences run i = 0..K-1 with float accumulators.  Restated per position without
// changing a single rounding:
//   hits(i)    = number of hits at positions <= i        (small integer, exact in float)
//   Precision  = hits(i) / (float)(i+1);  Recall = hits(i) / (float)L
//   MAP        : sum_pre changes only AT hits, by pre = hits/(i+1) = Precision(i): the float adds are
//                replayed in ascending hit order (a warp-uniform loop over the set bits of the hit
//                mask, usually a handful); position i keeps the value after the last hit <= i.
//   NDCG       : DCG changes only at hits (same loop, double add rounded to float); iDCG after i
//                depends only on min(i+1, L): table idcg[n] built on the host with the same operations.
//   MRR        = (float)(1.0 / (double)(p+1)) from the first hit p on.
struct RowMetrics {
    // per row
    const int32_t *truth;
    int nt, L;
    float Lf;
    int hits_c, first;  // carried across 32-position chunks
    float sum_pre_c, dcg_c;

    __device__ __forceinline__ void begin(const int64_t *__restrict__ te_indptr, const int32_t *__restrict__ te_idx, int64_t csr_row)
    {
        const int64_t tb = __ldg(te_indptr + csr_row);
        nt = (int)(__ldg(te_indptr + csr_row + 1) - tb);
        truth = te_idx + tb;
        L = nt > 1 ? nt : 1;
        Lf = (float)L;
        hits_c = 0;
        first = -1;
        sum_pre_c = 0.0f;
        dcg_c = 0.0f;
    }

    // positions i0 + lane of the rank list; item < 0 = no item.  Writes per_user_row[mi * K + i] and adds
    // into acc[mi * K + i] (either may be null).  Must be called by the whole warp, chunks in order.
    __device__ __forceinline__ void chunk(int i0, int lane, int K, int32_t item, const MetricIds &mids, const double *__restrict__ disc,
                                          const float *__restrict__ idcg, float *__restrict__ per_user_row, double *acc)
    {
        const bool hit = (i0 + lane < K) && (item >= 0) && sorted_contains(truth, nt, item);
        chunk_hit(i0, lane, K, hit, mids, disc, idcg, per_user_row, acc);
    }

    // the same when the caller already knows which positions hold test items (hit: position i0 + lane < K is one)
    __device__ __forceinline__ void chunk_hit(int i0, int lane, int K, bool hit, const MetricIds &mids, const double *__restrict__ disc,
                                              const float *__restrict__ idcg, float *__restrict__ per_user_row, double *acc)
    {
        const int i = i0 + lane;
        const bool valid = i < K;
        const uint32_t mask = __ballot_sync(0xffffffffu, hit && valid);
        const uint32_t le_mask = 0xffffffffu >> (31 - lane);  // lanes <= mine
        // 0 / x is exactly +0: skipping those divisions keeps every bit and avoids the IEEE division's slow path,
        // which the compiler's fast-path check sends every zero numerator through (most positions have no hit yet)
        const float hf = (float)(hits_c + __popc(mask & le_mask));
        const float prec = (hf == 0.0f) ? 0.0f : hf / (float)(i + 1);
        float my_sum = sum_pre_c, my_dcg = dcg_c;
        for (uint32_t m = mask; m != 0u; m &= m - 1u) {  // hits of this chunk, ascending
            const int b = __ffs(m) - 1;
            sum_pre_c = sum_pre_c + __shfl_sync(0xffffffffu, prec, b);  // metric.h:57-59
            dcg_c = (float)((double)dcg_c + __ldg(disc + i0 + b));      // metric.h:78
            if (lane >= b) { my_sum = sum_pre_c; my_dcg = dcg_c; }
        }
        if (first < 0 && mask != 0u) first = i0 + __ffs(mask) - 1;
        hits_c += __popc(mask);
        if (valid) {
            const int M = mids.n;
            for (int mi = 0; mi < M; ++mi) {
                const int id = mids.id(mi);
                float val;
                if (id == 1) val = prec;                                                 // metric.h:19-30
                else if (id == 2) val = (hf == 0.0f) ? 0.0f : hf / Lf;                                            // metric.h:33-45
                else if (id == 3) val = (my_sum == 0.0f) ? 0.0f : my_sum / (float)(L < i + 1 ? L : i + 1);        // metric.h:48-66
                else if (id == 4) val = (my_dcg == 0.0f) ? 0.0f : my_dcg / __ldg(idcg + (L < i + 1 ? L : i + 1)); // metric.h:69-86
                else val = (first >= 0 && i >= first) ? (float)(1.0 / (double)(first + 1)) : 0.0f;  // metric.h:89-109
                if (per_user_row != nullptr) per_user_row[mi * K + i] = val;
                if (acc != nullptr) acc[mi * K + i] += (double)val;
            }
        }
    }
};

// block-level tail of the fused column sums: every warp kept float64 sums of its rows per column in
// shared memory [n_warps][MK]; fold them in warp order into one partial row (deterministic, no atomics)
__device__ __forceinline__ void fold_block_sums(const double *acc_all, int n_warps, int MK, double *__restrict__ out_row)
{
    __syncthreads();
    for (int c = threadIdx.x; c < MK; c += blockDim.x) {
        double t = 0.0;
        for (int w = 0; w < n_warps; ++w) t += acc_all[(size_t)w * MK + c];
        out_row[c] = t;
    }
}

// ---- K4: metrics from sorted rank keys [n_rows, K] or int32 rank lists [n_rows, K] ---------------------
// Rows are dealt round-robin to warps.  row_list / row_count (both or neither): evaluate only rows
// row_list[0 .. *row_count) (the rows the exact kernel re-did).  acc_out: [gridDim.x][M*K] partial sums.
__global__ void __launch_bounds__(K4_WARPS * 32)
k_metrics(const u64 *__restrict__ keys, const int32_t *__restrict__ idx_in, int K, int64_t n_rows,
          int64_t row0, const int32_t *__restrict__ row_list, const int *__restrict__ row_count,
          const int64_t *__restrict__ te_indptr, const int32_t *__restrict__ te_idx,
          MetricIds mids, const double *__restrict__ disc, const float *__restrict__ idcg, float *__restrict__ per_user,
          int32_t *__restrict__ topk_idx_out, float *__restrict__ topk_val_out, double *__restrict__ acc_out)
{
    pdl_wait();
    pdl_trigger();
    extern __shared__ double k4_acc[];  // [K4_WARPS][M*K] when acc_out != null
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int MK = mids.n * K;
    double *acc = (acc_out != nullptr) ? k4_acc + (size_t)warp * MK : nullptr;
    if (acc != nullptr)
        for (int c = lane; c < MK; c += 32) acc[c] = 0.0;
    const int64_t n_list = (row_list != nullptr) ? (int64_t)*row_count : n_rows;
    const int64_t n_warps = (int64_t)gridDim.x * K4_WARPS;
    for (int64_t j = (int64_t)blockIdx.x * K4_WARPS + warp; j < n_list; j += n_warps) {
        const int64_t row = (row_list != nullptr) ? (int64_t)row_list[j] : j;
        RowMetrics rm;
        rm.begin(te_indptr, te_idx, row0 + row);
        for (int i0 = 0; i0 < K; i0 += 32) {
            const int i = i0 + lane;
            int32_t item = -1;
            if (i < K) {
                if (keys != nullptr) {
                    const u64 k = keys[row * K + i];
                    item = (k == 0) ? -1 : (int32_t)key_item(k);
                    if (topk_val_out != nullptr) topk_val_out[row * K + i] = (k == 0) ? -__int_as_float(0x7f800000) : key_score(k);
                } else {
                    item = idx_in[row * K + i];
                }
                if (topk_idx_out != nullptr) topk_idx_out[row * K + i] = item;
            }
            rm.chunk(i0, lane, K, item, mids, disc, idcg, per_user != nullptr ? per_user + row * MK : nullptr, acc);
        }
    }
    if (acc_out != nullptr) fold_block_sums(k4_acc, K4_WARPS, MK, acc_out + (size_t)blockIdx.x * MK);
}

// sums[c] += sum over blocks of partial[b][c]: one block per column, fixed reduction tree
__global__ void __launch_bounds__(256)
k_colsum_fold(const double *__restrict__ partial, int n_blk, int n_cols, double *__restrict__ sums)
{
    pdl_wait();
    pdl_trigger();
    __shared__ double s_t[8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c = blockIdx.x;
    double t = 0.0;
    for (int b = threadIdx.x; b < n_blk; b += 256) t += partial[(size_t)b * n_cols + c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (lane == 0) s_t[warp] = t;
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) a += s_t[w];
        sums[c] += a;
    }
}

// row_list != null: sum rows row_list[0..n_rows) of per_user instead of rows 0..n_rows (grouped evaluation)
__global__ void k_colsum_partial(const float *__restrict__ per_user, int64_t n_rows, int n_cols,
                                 double *__restrict__ partial, const int32_t *__restrict__ row_list = nullptr)
{
    for (int c = threadIdx.x; c < n_cols; c += blockDim.x) {
        double acc = 0.0;
        for (int64_t r = blockIdx.x; r < n_rows; r += gridDim.x) {
            const int64_t row = (row_list != nullptr) ? (int64_t)row_list[r] : r;
            acc += (double)per_user[row * n_cols + c];
        }
        partial[(size_t)blockIdx.x * n_cols + c] = acc;
    }
}

// sorted rank keys -> item ids / scores (the pyx_sort-style top-k entry points)
__global__ void k_unpack_keys(const u64 *__restrict__ keys, int64_t n, int32_t *__restrict__ idx, float *__restrict__ val)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const u64 k = keys[i];
    if (idx != nullptr) idx[i] = (k == 0) ? -1 : (int32_t)key_item(k);
    if (val != nullptr) val[i] = (k == 0) ? -__int_as_float(0x7f800000) : key_score(k);
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
k_merge_partials(const u64 *__restrict__ part, int S, int K, int64_t n_rows, int64_t row0, int64_t stride_row, int64_t stride_s,
                 const int64_t *__restrict__ tr_indptr, const int32_t *__restrict__ tr_idx,
                 u64 *__restrict__ out_keys)
{
    // list s of row r starts at part[r * stride_row + s * stride_s]: [rows][S][K] for the chunk lists of one
    // GPU, [S][rows][K] for per-shard lists gathered from S GPUs
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * 4 + warp;
    if (row >= n_rows) return;
    const int n = S * K;
    const u64 *src = part + row * stride_row;
    u64 v[PER];
#pragma unroll
    for (int e = 0; e < PER; ++e) {
        const int i = e * 32 + lane;
        const int s = i / K;
        v[e] = (i < n) ? src[(int64_t)s * stride_s + (i - s * K)] : 0ull;
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

// item ids of rank keys: shard-local -> global (key = ord << 32 | ~item, so item + off is key - off)
__global__ void k_offset_keys(u64 *__restrict__ keys, int64_t n, uint32_t item_offset)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && keys[i] != 0ull) keys[i] -= (u64)item_offset;
}

}  // namespace skr
