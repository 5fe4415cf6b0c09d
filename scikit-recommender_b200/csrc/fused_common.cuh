// fused_common.cuh -- pieces shared by the two fused score+mask+top-K kernels.
//
// Work decomposition: the U x I score plane is cut into user tiles of TM=128 rows and item
// tiles of TN=128 columns.  A work item is (chunk c, user tile rt): one CTA sweeps the item
// tiles of chunk c in ascending item order for the 128 users of rt, keeping a K-entry min-heap
// of rank keys per user in shared memory.  Work items are numbered chunk-major so that chunk 0
// of every user tile is scheduled first and its thresholds (thr_g) are visible to later chunks.
//
// Per-row running threshold.  thr = max(score of the row's heap root once the heap is full,
// best threshold published by any CTA for that row).  A score strictly below thr can never be
// in the row's top-K (K better items are known), so it is dropped after ONE compare; survivors
// are checked against the train-item bitmap of the tile and pushed.  Equal scores pass the
// filter and are resolved exactly by the 64-bit key (score desc, item id asc).
#pragma once
#include "common.cuh"

namespace skr {

constexpr int TM = 128;  // users per tile (= TMEM lanes = UMMA M)
constexpr int TN = 128;  // items per tile (= UMMA N)

struct FusedParams {
    int64_t n_rows;   // rows in this call
    int64_t row0;     // first CSR row (multiple of TM)
    int n_items;
    int d;
    int K;
    int S;                // item chunks per user tile
    int tiles_per_chunk;  // item tiles per chunk
    int n_ct;             // item tiles
    int n_rt;             // user tiles in this call
    const float *bias;                 // [n_ct*TN] padded, or null
    const uint32_t *mask_keys;         // per user tile, ascending (item << 7 | row_in_tile), or null
    const int64_t *mask_tile_ptr;      // [total user tiles + 1]
    const uint32_t *mask_tile_off;     // [total user tiles, n_ct + 1] key offsets of each item tile within the user tile
    uint32_t *thr_g;                   // [n_rows] ord(score) thresholds, 0 = none
    u64 *part;                         // [n_rows, S, K] partial lists
};

// ---- K-entry min-heap of keys, one per row, interleaved: slot s of row r at heap[s*TM + r] ----
__device__ __forceinline__ void heap_push(u64 *h, int &n, int K, u64 key)
{
    if (n < K) {
        int p = n++;
        while (p > 0) {
            int q = (p - 1) >> 1;
            u64 v = h[q * TM];
            if (v <= key) break;
            h[p * TM] = v;
            p = q;
        }
        h[p * TM] = key;
    } else if (key > h[0]) {
        int p = 0;
        for (;;) {
            int l = 2 * p + 1;
            if (l >= K) break;
            u64 vl = h[l * TM];
            int w = l;
            u64 vw = vl;
            if (l + 1 < K) {
                u64 vr = h[(l + 1) * TM];
                if (vr < vl) { w = l + 1; vw = vr; }
            }
            if (vw >= key) break;
            h[p * TM] = vw;
            p = w;
        }
        h[p * TM] = key;
    }
}

// first index in [lo, hi) whose key >= x
__device__ __forceinline__ int64_t lower_bound_u32(const uint32_t *__restrict__ a, int64_t lo, int64_t hi, uint32_t x)
{
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (__ldg(a + mid) < x) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// bits of word w (columns col0 + 32w .. +31) that lie at or beyond n_items
__device__ __forceinline__ uint32_t oob_bits(int col0, int w, int n_items)
{
    int first = n_items - (col0 + 32 * w);  // number of valid columns in this word
    if (first >= 32) return 0u;
    if (first <= 0) return 0xffffffffu;
    return 0xffffffffu << first;
}

}  // namespace skr
