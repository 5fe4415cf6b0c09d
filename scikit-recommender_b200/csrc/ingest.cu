// ingest.cu -- device-side ingestion of interaction CSRs (set-up path, not the hot path).
//
// skr_set_train_csr / skr_set_test_csr receive what `ImplicitFeedback.to_user_dict()` / `to_csr_matrix()` hold
// (dataset.py:131-156): per-user item lists in file order, possibly with duplicates.  The kernels want
//   * rows sorted by item and unique (binary search in k_metrics / k_topk_scores, cursor walks in the FP32 paths),
//   * for the fused tensor-core path, per 128-user tile the keys (item << 7 | row_in_tile) ascending, with the
//     offset of every 128-item tile inside the user tile's keys.
// Round 1 built both on one host thread (a std::vector copy and std::sort per row, a std::sort per user tile, a
// two-level loop for the offsets: seconds at 5*10^7 interactions).  Here the raw arrays are uploaded once and
// everything is built on the GPU: one kernel expands (row, item) pairs, two CUB radix sorts + CUB unique order them,
// binary-search kernels cut row pointers, tile pointers and tile offsets.  CUB is library code; it is used on this
// set-up path only.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <cub/cub.cuh>

#include "ingest.h"

namespace skr {
namespace {

typedef unsigned long long u64;

// first index in [0, n) with a[i] >= x
__device__ __forceinline__ int64_t lower_bound_u64(const u64 *__restrict__ a, int64_t n, u64 x)
{
    int64_t lo = 0, hi = n;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (a[mid] < x) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// flags: bit 0 indptr not monotone / negative, bit 1 item out of range
__global__ void k_check_indptr(const int64_t *__restrict__ indptr, int64_t n_rows, int *__restrict__ flags)
{
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n_rows && indptr[r + 1] < indptr[r]) atomicOr(flags, 1);
}

// pair p of the raw CSR -> (row << 32 | item); the row comes from a binary search of p in indptr
__global__ void k_expand_pairs(const int64_t *__restrict__ indptr, const int32_t *__restrict__ idx, int64_t n_rows, int64_t nnz, int64_t base,
                               int64_t n_items, u64 *__restrict__ keys, int *__restrict__ flags)
{
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= nnz) return;
    int64_t lo = 0, hi = n_rows;  // last row r with indptr[r] <= base + p
    while (hi - lo > 1) {
        const int64_t mid = (lo + hi) >> 1;
        if (indptr[mid] <= base + p) lo = mid; else hi = mid;
    }
    const int32_t it = idx[p];
    if (it < 0 || it >= n_items) atomicOr(flags, 2);
    keys[p] = ((u64)lo << 32) | (u64)(uint32_t)it;
}

__global__ void k_row_ptr(const u64 *__restrict__ keys, int64_t n, int64_t n_rows, int64_t *__restrict__ indptr_out)
{
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r <= n_rows) indptr_out[r] = lower_bound_u64(keys, n, (u64)r << 32);
}

__global__ void k_split_items(const u64 *__restrict__ keys, int64_t n, int32_t *__restrict__ idx_out, u64 *__restrict__ mkeys, int tm)
{
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const u64 k = keys[p];
    const uint32_t item = (uint32_t)k;
    const u64 row = k >> 32;
    idx_out[p] = (int32_t)item;
    if (mkeys != nullptr) mkeys[p] = ((row / (u64)tm) << 32) | (u64)((item << 7) | (uint32_t)(row % (u64)tm));
}

__global__ void k_mask_low(const u64 *__restrict__ mkeys, int64_t n, uint32_t *__restrict__ out)
{
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p < n) out[p] = (uint32_t)mkeys[p];
}

__global__ void k_tile_ptr(const u64 *__restrict__ mkeys, int64_t n, int64_t n_rt, int64_t *__restrict__ tile_ptr)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t <= n_rt) tile_ptr[t] = lower_bound_u64(mkeys, n, (u64)t << 32);
}

// tile_off[rt][ct] = number of keys of user tile rt whose item lies below ct * tn (ct = 0 .. n_ct)
__global__ void k_tile_off(const u64 *__restrict__ mkeys, const int64_t *__restrict__ tile_ptr, int64_t n_rt, int64_t n_ct, int64_t n_items, int tn,
                           uint32_t *__restrict__ tile_off)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rt * (n_ct + 1)) return;
    const int64_t rt = i / (n_ct + 1), ct = i - rt * (n_ct + 1);
    const int64_t b = tile_ptr[rt], e = tile_ptr[rt + 1];
    int64_t lim_item = ct * tn;
    if (lim_item > n_items) lim_item = n_items;
    const u64 x = ((u64)rt << 32) | ((u64)lim_item << 7);
    tile_off[i] = (uint32_t)(lower_bound_u64(mkeys + b, e - b, x));
}

struct Scratch {
    void *p[12] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    int n = 0;
    ~Scratch() { for (int i = 0; i < n; ++i) if (p[i]) cudaFree(p[i]); }
    cudaError_t get(void **out, size_t bytes)
    {
        cudaError_t e = cudaMalloc(out, bytes ? bytes : 1);
        if (e == cudaSuccess && n < 12) p[n++] = *out;
        return e;
    }
};

#define ING_CUDA(call)                                                                               \
    do {                                                                                             \
        cudaError_t e__ = (call);                                                                    \
        if (e__ != cudaSuccess) {                                                                    \
            snprintf(err, errlen, "%s: %s", #call, cudaGetErrorString(e__));                         \
            return e__ == cudaErrorMemoryAllocation ? -3 : -2;                                       \
        }                                                                                            \
    } while (0)

int bits_for(u64 max_value)
{
    int b = 1;
    while (b < 64 && (max_value >> b) != 0) ++b;
    return b;
}

}  // namespace

void ingest_free(IngestOut *o)
{
    if (!o) return;
    if (o->indptr) cudaFree(o->indptr);
    if (o->idx) cudaFree(o->idx);
    if (o->mask_keys) cudaFree(o->mask_keys);
    if (o->tile_ptr) cudaFree(o->tile_ptr);
    if (o->tile_off) cudaFree(o->tile_off);
    *o = IngestOut();
}

int ingest_csr(const int64_t *h_indptr, const int32_t *h_idx, int64_t n_rows, int64_t n_items, bool want_mask, int tm, int tn,
               IngestOut *out, char *err, size_t errlen, cudaStream_t st)
{
    *out = IngestOut();
    if (n_rows < 0 || n_items <= 0) { snprintf(err, errlen, "csr: n_rows=%lld n_items=%lld", (long long)n_rows, (long long)n_items); return -1; }
    if (h_indptr[0] < 0) { snprintf(err, errlen, "csr: indptr[0] < 0"); return -1; }
    if (h_indptr[n_rows] < h_indptr[0]) { snprintf(err, errlen, "csr: indptr not monotone"); return -1; }
    const int64_t base = h_indptr[0], nnz = h_indptr[n_rows] - base;
    Scratch S;
    int64_t *d_raw_ptr = nullptr;
    int32_t *d_raw_idx = nullptr;
    u64 *d_keys = nullptr, *d_keys2 = nullptr;
    int *d_flags = nullptr;
    int64_t *d_count = nullptr;
    ING_CUDA(S.get((void **)&d_raw_ptr, sizeof(int64_t) * (size_t)(n_rows + 1)));
    ING_CUDA(S.get((void **)&d_raw_idx, sizeof(int32_t) * (size_t)nnz));
    ING_CUDA(S.get((void **)&d_keys, sizeof(u64) * (size_t)nnz));
    ING_CUDA(S.get((void **)&d_keys2, sizeof(u64) * (size_t)nnz));
    ING_CUDA(S.get((void **)&d_flags, sizeof(int)));
    ING_CUDA(S.get((void **)&d_count, sizeof(int64_t)));
    ING_CUDA(cudaMemcpyAsync(d_raw_ptr, h_indptr, sizeof(int64_t) * (size_t)(n_rows + 1), cudaMemcpyHostToDevice, st));
    if (nnz > 0) ING_CUDA(cudaMemcpyAsync(d_raw_idx, h_idx + base, sizeof(int32_t) * (size_t)nnz, cudaMemcpyHostToDevice, st));
    ING_CUDA(cudaMemsetAsync(d_flags, 0, sizeof(int), st));
    const int T = 256;
    if (n_rows > 0) k_check_indptr<<<(unsigned)((n_rows + T - 1) / T), T, 0, st>>>(d_raw_ptr, n_rows, d_flags);
    int64_t n_unique = 0;
    u64 *d_sorted = d_keys;
    if (nnz > 0) {
        k_expand_pairs<<<(unsigned)((nnz + T - 1) / T), T, 0, st>>>(d_raw_ptr, d_raw_idx, n_rows, nnz, base, n_items, d_keys, d_flags);
        // (row, item) ascending, then unique: rows sorted by item without duplicates
        const int end_bit = 32 + bits_for((u64)(n_rows > 0 ? n_rows - 1 : 0));
        size_t tmp_bytes = 0, tmp2 = 0;
        cub::DoubleBuffer<u64> db(d_keys, d_keys2);
        ING_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, db, (int64_t)nnz, 0, end_bit, st));
        ING_CUDA(cub::DeviceSelect::Unique(nullptr, tmp2, d_keys, d_keys2, d_count, (int64_t)nnz, st));
        if (tmp2 > tmp_bytes) tmp_bytes = tmp2;
        void *d_tmp = nullptr;
        ING_CUDA(S.get(&d_tmp, tmp_bytes));
        ING_CUDA(cub::DeviceRadixSort::SortKeys(d_tmp, tmp_bytes, db, (int64_t)nnz, 0, end_bit, st));
        u64 *srt = db.Current(), *alt = db.Alternate();
        ING_CUDA(cub::DeviceSelect::Unique(d_tmp, tmp_bytes, srt, alt, d_count, (int64_t)nnz, st));
        d_sorted = alt;
        int h_flags = 0;
        ING_CUDA(cudaMemcpyAsync(&n_unique, d_count, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
        ING_CUDA(cudaMemcpyAsync(&h_flags, d_flags, sizeof(int), cudaMemcpyDeviceToHost, st));
        ING_CUDA(cudaStreamSynchronize(st));
        if (h_flags & 1) { snprintf(err, errlen, "csr: indptr not monotone"); return -1; }
        if (h_flags & 2) { snprintf(err, errlen, "csr: an item id lies outside [0,%lld)", (long long)n_items); return -1; }
        // outputs of the row-major form
        ING_CUDA(cudaMalloc((void **)&out->indptr, sizeof(int64_t) * (size_t)(n_rows + 1)));
        ING_CUDA(cudaMalloc((void **)&out->idx, sizeof(int32_t) * (size_t)(n_unique ? n_unique : 1)));
        u64 *d_mkeys = want_mask ? srt : nullptr;  // the sorted-with-duplicates buffer is free again
        k_row_ptr<<<(unsigned)((n_rows + 1 + T - 1) / T), T, 0, st>>>(d_sorted, n_unique, n_rows, out->indptr);
        if (n_unique > 0) k_split_items<<<(unsigned)((n_unique + T - 1) / T), T, 0, st>>>(d_sorted, n_unique, out->idx, d_mkeys, tm);
        if (want_mask) {
            const int64_t n_rt = (n_rows + tm - 1) / tm, n_ct = (n_items + tn - 1) / tn;
            // keys of one user tile ascending by (item, row in tile): sort by (user tile, item << 7 | row)
            // the unique pairs (d_sorted) have been consumed by k_row_ptr / k_split_items, which are stream-ordered before
            // this sort: their buffer is its alternate
            cub::DoubleBuffer<u64> dm(d_mkeys, d_sorted);
            const int mbits = 32 + bits_for((u64)(n_rt > 0 ? n_rt - 1 : 0));
            size_t need = 0;
            ING_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, need, dm, (int64_t)n_unique, 0, mbits, st));
            void *d_tmp2 = d_tmp;
            if (need > tmp_bytes) ING_CUDA(S.get(&d_tmp2, need));
            if (n_unique > 0) ING_CUDA(cub::DeviceRadixSort::SortKeys(d_tmp2, need, dm, (int64_t)n_unique, 0, mbits, st));
            const u64 *msorted = dm.Current();
            ING_CUDA(cudaMalloc((void **)&out->mask_keys, sizeof(uint32_t) * (size_t)(n_unique ? n_unique : 1)));
            ING_CUDA(cudaMalloc((void **)&out->tile_ptr, sizeof(int64_t) * (size_t)(n_rt + 1)));
            ING_CUDA(cudaMalloc((void **)&out->tile_off, sizeof(uint32_t) * (size_t)(n_rt * (n_ct + 1) ? n_rt * (n_ct + 1) : 1)));
            if (n_unique > 0) k_mask_low<<<(unsigned)((n_unique + T - 1) / T), T, 0, st>>>(msorted, n_unique, out->mask_keys);
            k_tile_ptr<<<(unsigned)((n_rt + 1 + T - 1) / T), T, 0, st>>>(msorted, n_unique, n_rt, out->tile_ptr);
            const int64_t n_off = n_rt * (n_ct + 1);
            if (n_off > 0) k_tile_off<<<(unsigned)((n_off + T - 1) / T), T, 0, st>>>(msorted, out->tile_ptr, n_rt, n_ct, n_items, tn, out->tile_off);
        }
    } else {
        int h_flags = 0;
        ING_CUDA(cudaMemcpyAsync(&h_flags, d_flags, sizeof(int), cudaMemcpyDeviceToHost, st));
        ING_CUDA(cudaStreamSynchronize(st));
        if (h_flags & 1) { snprintf(err, errlen, "csr: indptr not monotone"); return -1; }
        ING_CUDA(cudaMalloc((void **)&out->indptr, sizeof(int64_t) * (size_t)(n_rows + 1)));
        ING_CUDA(cudaMalloc((void **)&out->idx, sizeof(int32_t)));
        ING_CUDA(cudaMemsetAsync(out->indptr, 0, sizeof(int64_t) * (size_t)(n_rows + 1), st));
        if (want_mask) {
            const int64_t n_rt = (n_rows + tm - 1) / tm, n_ct = (n_items + tn - 1) / tn;
            const size_t n_off = (size_t)(n_rt * (n_ct + 1) ? n_rt * (n_ct + 1) : 1);
            ING_CUDA(cudaMalloc((void **)&out->mask_keys, sizeof(uint32_t)));
            ING_CUDA(cudaMalloc((void **)&out->tile_ptr, sizeof(int64_t) * (size_t)(n_rt + 1)));
            ING_CUDA(cudaMalloc((void **)&out->tile_off, sizeof(uint32_t) * n_off));
            ING_CUDA(cudaMemsetAsync(out->tile_ptr, 0, sizeof(int64_t) * (size_t)(n_rt + 1), st));
            ING_CUDA(cudaMemsetAsync(out->tile_off, 0, sizeof(uint32_t) * n_off, st));
        }
    }
    out->nnz = n_unique;
    ING_CUDA(cudaGetLastError());
    ING_CUDA(cudaStreamSynchronize(st));  // the scratch buffers go away with this frame
    return 0;
}

}  // namespace skr
