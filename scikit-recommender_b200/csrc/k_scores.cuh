// k_scores.cuh -- K2: per-row streaming top-K with train-item masking, one CTA per row.
//
// k_topk_scores: score-matrix-in.  Replaces, for a device-resident float32 [B, I] score block, the
//   reference's masking loop (evaluator.py:195-200) and the per-row selection of eval_one_user
//   (evaluate.h:27-45).  HBM-bound: each row is streamed exactly once (4*I bytes per user).
// k_row_exact:   the same selection over scores computed on the fly in FP32 (u . item_j + bias_j),
//   for the few rows the fused tensor-core path could not settle from its candidate lists.
//
// A running threshold (the K-th key so far) filters the stream, survivors go to a shared-memory
// buffer that is folded into the sorted top-K by a block bitonic sort of only the power-of-two prefix
// in use.  Masked (train) items become -inf and stay candidates, like the reference; keys order
// them by ascending item id.
#pragma once
#include "common.cuh"

namespace skr {

constexpr int K2_THREADS = 256;
constexpr int K2_CHUNK = 1024;  // elements per pass: four per thread
constexpr int K2_P = 2048;      // key slots in shared memory (sorted top-K + survivor buffer)
constexpr int K2_MAX_K = 512;
constexpr int K2_MAX_D = 1024;  // k_row_exact keeps the user vector in shared memory

__device__ __forceinline__ float4 ldg_stream_f4(const float *p)
{
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}

// score source: FP32 dot products of one user vector (shared memory) with item rows, k ascending
struct RowDot {
    const float *u;  // shared memory, d floats
    const float *V;
    int64_t ld_v;
    int d;
    const float *bias;  // or null
    bool vec_ok;
    __device__ __forceinline__ void get4(int j0, int n_items, float (&v)[4]) const
    {
        // four items at once: four independent FMA chains (each still k ascending, the reference order of the exact
        // path) with their loads in flight together -- one chain at a time left this kernel waiting on every load
        // (4 ms to walk one 1M-item row at d = 128)
        float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        const float *it[4];
        bool ok[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            ok[q] = j0 + q < n_items;
            it[q] = V + (int64_t)(ok[q] ? j0 + q : 0) * ld_v;
        }
        int k = 0;
        if (vec_ok) {
            // two k-steps per trip: eight 16-byte loads in flight per thread (a failed row of a 10^6-item catalogue is
            // walked by 40 blocks only -- the merge takes 4,096 keys -- and every trip waits one DRAM round trip)
            for (; k + 8 <= d; k += 8) {
                float4 x[4], y[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) x[q] = __ldg(reinterpret_cast<const float4 *>(it[q] + k));
#pragma unroll
                for (int q = 0; q < 4; ++q) y[q] = __ldg(reinterpret_cast<const float4 *>(it[q] + k + 4));
                const float u0 = u[k], u1 = u[k + 1], u2 = u[k + 2], u3 = u[k + 3];
                const float u4 = u[k + 4], u5 = u[k + 5], u6 = u[k + 6], u7 = u[k + 7];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    acc[q] = fmaf(u0, x[q].x, acc[q]);
                    acc[q] = fmaf(u1, x[q].y, acc[q]);
                    acc[q] = fmaf(u2, x[q].z, acc[q]);
                    acc[q] = fmaf(u3, x[q].w, acc[q]);
                    acc[q] = fmaf(u4, y[q].x, acc[q]);
                    acc[q] = fmaf(u5, y[q].y, acc[q]);
                    acc[q] = fmaf(u6, y[q].z, acc[q]);
                    acc[q] = fmaf(u7, y[q].w, acc[q]);
                }
            }
            for (; k + 4 <= d; k += 4) {
                float4 x[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) x[q] = __ldg(reinterpret_cast<const float4 *>(it[q] + k));
                const float u0 = u[k], u1 = u[k + 1], u2 = u[k + 2], u3 = u[k + 3];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    acc[q] = fmaf(u0, x[q].x, acc[q]);
                    acc[q] = fmaf(u1, x[q].y, acc[q]);
                    acc[q] = fmaf(u2, x[q].z, acc[q]);
                    acc[q] = fmaf(u3, x[q].w, acc[q]);
                }
            }
        }
        for (; k < d; ++k) {
            const float uk = u[k];
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[q] = fmaf(uk, __ldg(it[q] + k), acc[q]);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (bias != nullptr && ok[q]) acc[q] += __ldg(bias + j0 + q);
            v[q] = ok[q] ? acc[q] : 0.0f;
        }
    }
};

// Whole-block routine: exact top-K keys of one row over the items [j_begin, j_end) (j_begin a multiple of
// K2_CHUNK; 0 .. n_items for the whole row), sorted, to out[0..K) (0 = empty slot when the range holds fewer).
template <class Src>
__device__ __forceinline__ void topk_row_block(const Src &src, int n_items, int64_t tr_begin, int64_t tr_end,
                                               const int32_t *__restrict__ tr_idx, int K, u64 *__restrict__ out,
                                               u64 *keys, uint32_t *bitmap, int *s_cnt, int j_begin = 0, int j_end = 0x7fffffff)
{
    const int tid = threadIdx.x, lane = tid & 31;
    const int n_end = min(n_items, j_end);
    int64_t cur = tr_begin;
    const int64_t te = tr_end;
    if (j_begin > 0) {  // train items below the range are not ours: first index with item >= j_begin
        int64_t lo = tr_begin, hi = tr_end;
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if (__ldg(tr_idx + mid) < j_begin) lo = mid + 1; else hi = mid;
        }
        cur = lo;
    }
    int next_train = (cur < te) ? __ldg(tr_idx + cur) : 0x7fffffff;

    int base = 0;  // keys[0..base) hold the sorted best-so-far
    u64 thr_key = 0;
    float thr_f = -__int_as_float(0x7f800000);
    if (tid == 0) *s_cnt = 0;
    __syncthreads();

    float v[4], nxt[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    src.get4(j_begin + tid * 4, n_end, v);

    for (int c0 = j_begin; c0 < n_end; c0 += K2_CHUNK) {
        const int j0 = c0 + tid * 4;
        if (c0 + K2_CHUNK < n_end) src.get4(j0 + K2_CHUNK, n_end, nxt);

        // train items that fall into this chunk -> bitmap (sorted CSR row, cursor moves forward)
        uint32_t mbits = 0;
        if (next_train < c0 + K2_CHUNK) {
            if (tid < K2_CHUNK / 32) bitmap[tid] = 0;
            __syncthreads();
            for (;;) {
                int64_t p = cur + tid;
                int t = (p < te) ? __ldg(tr_idx + p) : 0x7fffffff;
                bool in = t < c0 + K2_CHUNK;
                if (in) atomicOr(&bitmap[(t - c0) >> 5], 1u << ((t - c0) & 31));
                int n_in = __syncthreads_count(in);
                cur += n_in;
                if (n_in < K2_THREADS) break;
            }
            next_train = (cur < te) ? __ldg(tr_idx + cur) : 0x7fffffff;
            mbits = (bitmap[(tid * 4) >> 5] >> ((tid * 4) & 31)) & 0xFu;
        }

#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float s = v[q];
            if ((mbits >> q) & 1u) s = -__int_as_float(0x7f800000);
            bool pass = (j0 + q < n_end) && !(s < thr_f);
            u64 key = 0;
            if (pass) {
                key = make_key(s, (uint32_t)(j0 + q));
                pass = key > thr_key;
            }
            unsigned bal = __ballot_sync(0xffffffffu, pass);
            if (bal) {
                int leader = __ffs(bal) - 1;
                int pos = 0;
                if (lane == leader) pos = atomicAdd(s_cnt, __popc(bal));
                pos = __shfl_sync(0xffffffffu, pos, leader);
                if (pass) keys[base + pos + __popc(bal & ((1u << lane) - 1u))] = key;
            }
        }
        __syncthreads();
        const int cnt = *s_cnt;
        // every thread must have read the count before anyone appends survivors of the next chunk: a late reader
        // would see a larger count, could decide to sort on its own and meet barriers the others never reach
        __syncthreads();
        const bool last = (c0 + K2_CHUNK >= n_end);
        if (last || base + cnt + K2_CHUNK > K2_P) {
            const int tot = base + cnt;
            int n_sort = next_pow2(tot);
            if (n_sort < 2) n_sort = 2;
            for (int i = tot + tid; i < n_sort; i += K2_THREADS) keys[i] = 0;
            __syncthreads();
            block_bitonic_desc(keys, n_sort, tid, K2_THREADS);
            base = tot < K ? tot : K;
            if (base == K) {
                thr_key = keys[K - 1];
                thr_f = key_score(thr_key);
            }
            __syncthreads();
            if (tid == 0) *s_cnt = 0;
            __syncthreads();
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) v[q] = nxt[q];
    }
    for (int i = tid; i < K; i += K2_THREADS) out[i] = (i < base) ? keys[i] : 0ull;
}

// Streaming form for a row of a score matrix (the HBM-bound kernel of the `predict` path: 4 I bytes per user is all
// it must read).  The generic routine above meets 2-3 block barriers per 1,024 scores and keeps one float4 per
// thread in flight -- measured 0.77 TB/s at c2.  Here a chunk is 1, 2, 4, then 8 float4 per thread (8,192 scores,
// 32 KB per block in flight), all loads issued before the first compare, and the work is split in two so that no
// warp ever walks a slow path for the sake of one lane:
//   filter   one compare per score against the threshold; survivors are appended RAW (score bits, item) to the shared
//            buffer with one atomic per warp and score position (ballot + popc);
//   fix-up   after the chunk's barrier all threads share the new entries: train items (binary search in the row's
//            sorted CSR; evaluator.py:195-200) become -inf and stay candidates, the entry becomes a rank key and is
//            kept only if it beats the K-th key.  (Doing this inside the filter cost ~2,000 cycles per score position
//            whenever ANY lane of the warp had a survivor -- 80 % of the positions of a young row.)
//   fold     block bitonic of [sorted top-K | buffer].  A threshold taken from the first m scores lets K / m of
//            everything after it through until it is tightened, so chunks grow geometrically and are folded right away
//            while the row is young (about K survivors per doubling); after 65,536 scores a chunk yields a handful and
//            the buffer is folded only when half full.  A chunk that overflows the buffer (adversarial ascending rows)
//            is dropped and redone 1,024 scores at a time.
// Rows of a matrix whose pitch is not a multiple of 4 floats start at any 4-byte offset: up to 3 head scores are
// offered one by one and the float4 stream starts at the first 16-byte boundary.
constexpr int K2_SUPER = 8;  // float4 per thread in a full chunk

__device__ __forceinline__ void k2_fold(u64 *keys, int *s_cnt, int &base, int K, u64 &thr_key, float &thr_f, int tid)
{
    // callers have passed a barrier since the last append; every thread reads the same count
    const int room = K2_P - base;
    const int cnt = min(*s_cnt, room);
    __syncthreads();
    const int tot = base + cnt;
    int n_sort = next_pow2(tot);
    if (n_sort < 2) n_sort = 2;
    for (int i = tot + tid; i < n_sort; i += K2_THREADS) keys[i] = 0;
    __syncthreads();
    block_bitonic_desc(keys, n_sort, tid, K2_THREADS);
    base = tot < K ? tot : K;
    if (base == K && keys[K - 1] != 0ull) {
        thr_key = keys[K - 1];
        thr_f = key_score(thr_key);
    }
    __syncthreads();
    if (tid == 0) *s_cnt = 0;
    __syncthreads();
}

__device__ __forceinline__ void topk_row_stream(const float *__restrict__ row, int n_items, int64_t tb, int64_t te,
                                                const int32_t *__restrict__ tr_idx, int K, u64 *__restrict__ out, u64 *keys, int *s_cnt)
{
    const int tid = threadIdx.x, lane = tid & 31;
    const uint32_t lt_mask = (1u << lane) - 1u;
    int base = 0;          // keys[0 .. base) sorted best-so-far; the buffer starts at keys[base]
    u64 thr_key = 0;       // K-th key once K are known
    float thr_f = -__int_as_float(0x7f800000);
    if (tid == 0) *s_cnt = 0;
    __syncthreads();
    const int a0 = (int)((4u - (unsigned)((reinterpret_cast<uintptr_t>(row) >> 2) & 3u)) & 3u);
    const float *rowa = row + a0;       // 16-byte aligned
    const int n_body = n_items - a0;    // scores from the first aligned one on (<= 0 for a row of 1-3 scores)

    // raw append of one score position across the warp (every lane calls this, `pass` says who has a survivor)
    auto offer = [&](bool pass, float sc, int j) {
        const unsigned bal = __ballot_sync(0xffffffffu, pass);
        if (bal) {
            int pos = 0;
            if (lane == __ffs(bal) - 1) pos = atomicAdd(s_cnt, __popc(bal));
            pos = __shfl_sync(0xffffffffu, pos, __ffs(bal) - 1) + __popc(bal & lt_mask);
            if (pass && pos < K2_P - base) keys[base + pos] = ((u64)__float_as_uint(sc) << 32) | (u64)(uint32_t)j;
        }
    };
    // raw entries [from, to) of the buffer -> rank keys (0 = rejected); all threads
    auto fixup = [&](int from, int to) {
        for (int i = from + tid; i < to; i += K2_THREADS) {
            const u64 raw = keys[base + i];
            float sc = __uint_as_float((uint32_t)(raw >> 32));
            const int j = (int)(uint32_t)raw;
            if (te > tb && sorted_contains(tr_idx + tb, (int)(te - tb), (int32_t)j)) sc = -__int_as_float(0x7f800000);
            const u64 key = make_key(sc, (uint32_t)j);
            keys[base + i] = (key > thr_key) ? key : 0ull;
        }
    };
    // one chunk of nf4 float4 per thread starting at body score c0; force_fold: fold whatever it appended.
    // -> false when the chunk overflowed the buffer (nothing of it was kept; the caller redoes it in pieces)
    int pending = 0;  // entries in the buffer (already fixed up) that are not folded yet; block-uniform
    auto chunk = [&](int c0, int nf4, bool force_fold, bool head) -> bool {
        float v[K2_SUPER][4];
        if (c0 + nf4 * K2_CHUNK <= n_body) {
#pragma unroll
            for (int q = 0; q < K2_SUPER; ++q)
                if (q < nf4) {
                    const float4 t = ldg_stream_f4(rowa + c0 + q * K2_CHUNK + tid * 4);
                    v[q][0] = t.x; v[q][1] = t.y; v[q][2] = t.z; v[q][3] = t.w;
                }
        } else {  // the last, partial chunk: guarded loads, missing scores are never offered
#pragma unroll
            for (int q = 0; q < K2_SUPER; ++q)
                if (q < nf4) {
                    const int jj = c0 + q * K2_CHUNK + tid * 4;
                    if (jj + 3 < n_body) {
                        const float4 t = ldg_stream_f4(rowa + jj);
                        v[q][0] = t.x; v[q][1] = t.y; v[q][2] = t.z; v[q][3] = t.w;
                    } else {
#pragma unroll
                        for (int e = 0; e < 4; ++e) v[q][e] = (jj + e < n_body) ? __ldg(rowa + jj + e) : 0.0f;
                    }
                }
        }
        if (head) {  // the up to 3 scores before the first 16-byte boundary (first chunk only; warp 0 holds lanes 0-2)
            if (tid < 32) {
                const bool have = tid < a0 && tid < n_items;
                offer(have, have ? __ldg(row + tid) : 0.0f, tid);
            }
        }
#pragma unroll
        for (int q = 0; q < K2_SUPER; ++q) {
            if (q < nf4) {
                const int jj = c0 + q * K2_CHUNK + tid * 4;
#pragma unroll
                for (int e = 0; e < 4; ++e) offer(jj + e < n_body && !(v[q][e] < thr_f), v[q][e], a0 + jj + e);
            }
        }
        __syncthreads();
        const int cnt = *s_cnt;
        const int room = K2_P - base;
        __syncthreads();  // everyone has the count before anyone appends again or resets it
        if (cnt > room) {
            // positions are handed out in order: everything at or beyond `pending` is this chunk's -- drop it, fold the rest
            if (tid == 0) *s_cnt = pending;
            __syncthreads();
            k2_fold(keys, s_cnt, base, K, thr_key, thr_f, tid);
            pending = 0;
            return false;
        }
        fixup(pending, cnt);
        if (cnt > pending && (force_fold || 2 * cnt > room)) {
            __syncthreads();
            k2_fold(keys, s_cnt, base, K, thr_key, thr_f, tid);
            pending = 0;
        } else {
            pending = cnt;
        }
        return true;
    };

    int c0 = 0, nf4 = 1;
    bool first = true;
    do {  // threshold first: chunks of 1,024 until K keys are known (one, unless the row is shorter or full of train items)
        chunk(c0, 1, true, first);  // cannot overflow: the buffer is empty and takes 1,024 + 3
        first = false;
        c0 += K2_CHUNK;
    } while (c0 < n_body && thr_key == 0ull);
    while (c0 < n_body) {
        if (!chunk(c0, nf4, c0 < 65536, false))
            for (int p = 0; p < nf4 && c0 + p * K2_CHUNK < n_body; ++p) chunk(c0 + p * K2_CHUNK, 1, true, false);
        c0 += nf4 * K2_CHUNK;
        if (nf4 < K2_SUPER) nf4 *= 2;
    }
    if (pending > 0) {
        __syncthreads();
        k2_fold(keys, s_cnt, base, K, thr_key, thr_f, tid);
    }
    for (int i = tid; i < K; i += K2_THREADS) out[i] = (i < base) ? keys[i] : 0ull;
}

__global__ void __launch_bounds__(K2_THREADS)
k_topk_scores(const float *__restrict__ scores, int64_t ld, int n_items, int64_t row0,
              const int64_t *__restrict__ tr_indptr, const int32_t *__restrict__ tr_idx, int K,
              u64 *__restrict__ out_keys)
{
    __shared__ u64 keys[K2_P];
    __shared__ int s_cnt;
    const int64_t r = blockIdx.x;
    int64_t tb = 0, te = 0;
    if (tr_indptr != nullptr) { tb = __ldg(tr_indptr + row0 + r); te = __ldg(tr_indptr + row0 + r + 1); }
    topk_row_stream(scores + r * ld, n_items, tb, te, tr_idx, K, out_keys + r * K, keys, &s_cnt);
}

// ---- one WARP per row, top-K <= 128: the form the common evaluations (K <= 100, run_config.py:16) take -------------
// ncu on the block-per-row kernel (c1: 6,040 rows of 3,706 scores, 0.50 ms; c2: 8,192 rows of 40,981, 1.34 ms): issue
// slots 61 % busy, DRAM 2 % -- it is bound by the INSTRUCTIONS of its block-wide bitonic sorts, not by memory.  A first
// warp-per-row version that kept a sorted best-128 in registers and merged every batch of survivors through a 256-key
// bitonic network still executed 40,000 warp instructions per 41 K-item row (ncu, 1.03 ms).  Sorting is not needed to
// stream: only the K-th best matters.  Now a warp owns a row and
//   * full steps of 1,024 scores arrive through a per-warp ring of four 4 KB bulk copies (cp.async.bulk + mbarrier:
//     16 KB in flight per warp, no stream data in registers);
//   * a float4 position costs a NaN-propagating max of four, one compare and one vote when nothing passes; survivors are
//     appended RAW (score bits, item) to the warp's staging area in shared memory (positions from ballot + popc);
//   * PRUNE (when the area may overflow, and after 256, 1,024, 4,096, 16,384, 65,536 scores -- a threshold from m scores
//     lets K / m of what follows through): new entries are fixed up by all lanes (train items -> -inf by binary search in
//     the row's sorted CSR, evaluator.py:195-200; rank keys), then the K-th largest score is found by a 32-step bit
//     search over the monotone integer image of the scores (one compare per entry and step, one REDUX per step), the
//     entries at or above it are compacted to the front and it becomes the threshold.  No sort.  Only when more than 32
//     entries tie at the cut (constant rows) a 64-step search over the full keys (score, then lower item id) runs;
//   * at the end of the row the survivors are cut to exactly K by that exact search and sorted once.
constexpr int KW_WARPS = 8;
constexpr int KW_STEP = 512;              // scores per step and warp: one 2 KB bulk copy
constexpr int KW_RING = 2;                // steps in flight per warp: 4 KB (occupancy matters more than depth: the row loop is latency-bound)
__host__ __device__ constexpr size_t kw_smem(int capq) { return (size_t)KW_WARPS * (KW_RING * KW_STEP * 4 + capq * 32 * 8 + KW_RING * 8) + 16; }

__device__ __forceinline__ float max_nan(float a, float b)
{
    float r;
    asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
}
// global -> shared bulk copy (TMA, 1-D), completion counted in bytes on an mbarrier
__device__ __forceinline__ void kw_bulk_load(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src), "r"(bytes), "r"((uint32_t)__cvta_generic_to_shared(bar))
                 : "memory");
}
__device__ __forceinline__ bool kw_try_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}

// stage[0 .. n_keep) rank keys kept by the last prune, stage[n_keep .. cnt) raw entries.  Afterwards stage[0 .. cnt)
// holds the survivors' rank keys: every key that can still be among the row's best K (exactly the best K when `exact`).
template <int CAPQ>
__device__ __forceinline__ void kw_prune(u64 *stage, int &cnt, int &n_keep, const int32_t *__restrict__ tr, int n_tr, int K, int lane,
                                         uint32_t lt_mask, u64 &thr_key, float &thr_f, bool exact)
{
    __syncwarp();
    for (int i = n_keep + lane; i < cnt; i += 32) {  // fix-up of the new entries
        const u64 raw = stage[i];
        float sc = __uint_as_float((uint32_t)(raw >> 32));
        const uint32_t j = (uint32_t)raw;
        if (n_tr > 0 && sorted_contains(tr, n_tr, (int32_t)j)) sc = -__int_as_float(0x7f800000);
        const u64 key = make_key(sc, j);
        stage[i] = (key > thr_key) ? key : 0ull;
    }
    __syncwarp();
    uint32_t hi[CAPQ];  // ord(score) of my entries (0: empty / rejected; a valid key has ord >= 0x007fffff > 0)
    int nv = 0;
#pragma unroll
    for (int q = 0; q < CAPQ; ++q) {
        const int i = q * 32 + lane;
        hi[q] = (i < cnt) ? (uint32_t)(stage[i] >> 32) : 0u;
        nv += hi[q] != 0u;
    }
    nv = __reduce_add_sync(0xffffffffu, nv);
    uint32_t T = 1u;        // keep everything valid
    u64 T64 = 0ull;
    bool use64 = false;
    if (nv > K) {
        T = 0u;
#pragma unroll 1
        for (int bit = 31; bit >= 0; --bit) {
            const uint32_t cand = T | (1u << bit);
            int c = 0;
#pragma unroll
            for (int q = 0; q < CAPQ; ++q) c += hi[q] >= cand;
            if (__reduce_add_sync(0xffffffffu, c) >= K) T = cand;
        }
        int c = 0;
#pragma unroll
        for (int q = 0; q < CAPQ; ++q) c += hi[q] >= T;
        const int n_ge = __reduce_add_sync(0xffffffffu, c);
        if (exact ? n_ge > K : n_ge > K + 32) {
            // ties at the cut: the K-th largest full key (score desc, item id asc); 64 steps, rare
            use64 = true;
#pragma unroll 1
            for (int bit = 63; bit >= 0; --bit) {
                const u64 cand = T64 | (1ull << bit);
                int c2 = 0;
#pragma unroll
                for (int q = 0; q < CAPQ; ++q) {
                    const int i = q * 32 + lane;
                    c2 += (i < cnt) && stage[i] >= cand;
                }
                if (__reduce_add_sync(0xffffffffu, c2) >= K) T64 = cand;
            }
        }
        // the new threshold: nothing below the K-th best score (or key) can enter the best K any more
        if (use64) { thr_key = T64 - 1ull; thr_f = key_score(T64); }
        else { thr_key = ((u64)T << 32) - 1ull; thr_f = unord_f32(T); }
    }
    // compaction (stable): survivors to the front
    int base = 0;
#pragma unroll
    for (int q = 0; q < CAPQ; ++q) {
        const int i = q * 32 + lane;
        const u64 k = (i < cnt) ? stage[i] : 0ull;
        const bool keep = use64 ? (k >= T64 && k != 0ull) : (hi[q] >= T && hi[q] != 0u);
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        __syncwarp();  // all lanes have read position q before anyone overwrites entries (base + ... <= i always)
        if (keep) stage[base + __popc(bal & lt_mask)] = k;
        base += __popc(bal);
    }
    __syncwarp();
    cnt = n_keep = base;
}

template <int CAPQ>
__global__ void __launch_bounds__(KW_WARPS * 32, CAPQ <= 8 ? 4 : 3)
k_topk_rows(const float *__restrict__ scores, int64_t ld, int n_items, int64_t n_rows, int64_t row0,
            const int64_t *__restrict__ tr_indptr, const int32_t *__restrict__ tr_idx, int K, u64 *__restrict__ out_keys, int *err_flag)
{
    constexpr int CAP = CAPQ * 32;
    extern __shared__ __align__(128) unsigned char kw_smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float *ring = reinterpret_cast<float *>(kw_smem_raw) + (size_t)warp * KW_RING * KW_STEP;
    u64 *stage = reinterpret_cast<u64 *>(kw_smem_raw + (size_t)KW_WARPS * KW_RING * KW_STEP * 4) + (size_t)warp * CAP;
    uint64_t *bars = reinterpret_cast<uint64_t *>(kw_smem_raw + (size_t)KW_WARPS * (KW_RING * KW_STEP * 4 + CAP * 8)) + warp * KW_RING;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const int64_t r = (int64_t)blockIdx.x * KW_WARPS + warp;
    if (r >= n_rows) return;  // warps are independent: no block-wide barrier below
    const float *row = scores + r * ld;
    int64_t tb = 0, te = 0;
    if (tr_indptr != nullptr) { tb = __ldg(tr_indptr + row0 + r); te = __ldg(tr_indptr + row0 + r + 1); }
    const int n_tr = (int)(te - tb);
    const int32_t *tr = tr_idx + tb;
    const int a0 = (int)((4u - (unsigned)((reinterpret_cast<uintptr_t>(row) >> 2) & 3u)) & 3u);
    const float *rowa = row + a0;     // 16-byte aligned
    const int n_body = n_items - a0;
    const int n_steps = n_body > 0 ? n_body / KW_STEP : 0;  // full steps

    if (lane == 0) {
        for (int s = 0; s < KW_RING; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(bars + s)));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (int s = 0; s < KW_RING && s < n_steps; ++s) kw_bulk_load(ring + s * KW_STEP, rowa + (size_t)s * KW_STEP, KW_STEP * 4, bars + s);
    }
    __syncwarp();

    u64 thr_key = 0;   // a key must be strictly larger to matter
    float thr_f = -__int_as_float(0x7f800000);
    int cnt = 0, n_keep = 0;  // warp-uniform

    // one score position across the warp
    auto offer = [&](bool pass, float sc, int j) {
        const unsigned bal = __ballot_sync(0xffffffffu, pass);
        if (bal) {
            if (pass) stage[cnt + __popc(bal & lt_mask)] = ((u64)__float_as_uint(sc) << 32) | (u64)(uint32_t)j;
            cnt += __popc(bal);
        }
    };
    // one float4 position: 4 x 32 scores, item ids jj .. jj + 3 per lane
    auto offer4 = [&](float x0, float x1, float x2, float x3, int jj) {
        const bool p0 = jj < n_items && !(x0 < thr_f), p1 = jj + 1 < n_items && !(x1 < thr_f);
        const bool p2 = jj + 2 < n_items && !(x2 < thr_f), p3 = jj + 3 < n_items && !(x3 < thr_f);
        if (__any_sync(0xffffffffu, p0 | p1 | p2 | p3)) {
            if (cnt > CAP - 128) kw_prune<CAPQ>(stage, cnt, n_keep, tr, n_tr, K, lane, lt_mask, thr_key, thr_f, false);
            offer(p0, x0, jj);
            offer(p1, x1, jj + 1);
            offer(p2, x2, jj + 2);
            offer(p3, x3, jj + 3);
        }
    };
    {   // head scores before the first 16-byte boundary
        const bool have = lane < a0 && lane < n_items;
        offer(have, have ? __ldg(row + lane) : 0.0f, lane);
    }
    int next_prune = 256;
    for (int st = 0; st < n_steps; ++st) {
        const int slot = st % KW_RING;
        const uint32_t parity = (uint32_t)((st / KW_RING) & 1);
        if (!kw_try_wait(bars + slot, parity)) {
            const long long t0 = clock64();
            while (!kw_try_wait(bars + slot, parity)) {
                if (clock64() - t0 > 4000000000ll) {  // ~2 s: a lost copy must not hang the GPU
                    if (err_flag != nullptr) atomicExch(err_flag, 21);
                    __threadfence_system();
                    __trap();
                }
            }
        }
        const float4 *buf = reinterpret_cast<const float4 *>(ring + slot * KW_STEP);
        const int c0 = st * KW_STEP;
#pragma unroll
        for (int q = 0; q < KW_STEP / 128; ++q) {
            const float4 v = buf[q * 32 + lane];
            const float m = max_nan(max_nan(v.x, v.y), max_nan(v.z, v.w));
            if (__any_sync(0xffffffffu, !(m < thr_f))) offer4(v.x, v.y, v.z, v.w, a0 + c0 + q * 128 + lane * 4);
            if (c0 + (q + 1) * 128 >= next_prune && next_prune <= 65536) {
                if (cnt > n_keep) kw_prune<CAPQ>(stage, cnt, n_keep, tr, n_tr, K, lane, lt_mask, thr_key, thr_f, false);
                next_prune *= 4;
            }
        }
        __syncwarp();  // every lane has read the slot: it may be refilled
        if (lane == 0 && st + KW_RING < n_steps)
            kw_bulk_load(ring + slot * KW_STEP, rowa + (size_t)(st + KW_RING) * KW_STEP, KW_STEP * 4, bars + slot);
    }
    // tail: guarded positions read directly
    for (int c0 = n_steps * KW_STEP; c0 < n_body; c0 += 128) {
        const int jj = c0 + lane * 4;
        float x[4];
        if (jj + 3 < n_body) {
            const float4 t = ldg_stream_f4(rowa + jj);
            x[0] = t.x; x[1] = t.y; x[2] = t.z; x[3] = t.w;
        } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) x[e] = (jj + e < n_body) ? __ldg(rowa + jj + e) : 0.0f;
        }
        offer4(x[0], x[1], x[2], x[3], a0 + jj);
        if (c0 + 128 >= next_prune && next_prune <= 65536) {
            if (cnt > n_keep) kw_prune<CAPQ>(stage, cnt, n_keep, tr, n_tr, K, lane, lt_mask, thr_key, thr_f, false);
            next_prune *= 4;
        }
    }
    // exactly the best K, sorted once
    kw_prune<CAPQ>(stage, cnt, n_keep, tr, n_tr, K, lane, lt_mask, thr_key, thr_f, true);
    constexpr int PER = CAPQ / 4;  // K <= 64: 64 keys, K <= 128: 128 keys
    u64 v[PER];
#pragma unroll
    for (int e = 0; e < PER; ++e) {
        const int i = e * 32 + lane;
        v[e] = (i < cnt) ? stage[i] : 0ull;
    }
    warp_bitonic_desc<PER>(v, lane);
#pragma unroll
    for (int e = 0; e < PER; ++e) {
        const int i = e * 32 + lane;
        if (i < K) out_keys[r * (int64_t)K + i] = v[e];
    }
}

// Rows on the fail list of k_select_cands: exact FP32 scores on the fly, same selection.  A work item is
// (failed row, one of n_seg item ranges): a handful of failed rows then spreads over the whole GPU instead of
// occupying one CTA each for the time it takes to walk the catalogue (0.6 ms per row at 92 K items).  The n_seg
// partial lists of a row are merged by k_merge_fail.
__global__ void __launch_bounds__(K2_THREADS)
k_row_exact(const int32_t *__restrict__ fail_list, const int *__restrict__ fail_count, const float *__restrict__ U, int64_t ld_u,
            const float *__restrict__ V, int64_t ld_v, int d, const float *__restrict__ bias, int n_items, int64_t row0,
            const int64_t *__restrict__ tr_indptr, const int32_t *__restrict__ tr_idx, int K, int n_seg, int seg_items,
            int max_seg_rows, u64 *__restrict__ part, u64 *__restrict__ out_keys)
{
    pdl_wait();
    pdl_trigger();
    __shared__ u64 keys[K2_P];
    __shared__ uint32_t bitmap[K2_CHUNK / 32];
    __shared__ int s_cnt;
    __shared__ float u_s[K2_MAX_D];
    const int n_fail = *fail_count;
    // grid (x, n_seg): block (x, g) walks segment g of the failed rows x, x + gridDim.x, ...  Only the first
    // max_seg_rows failed rows are cut in segments -- that spreads a handful of rows over the GPU; with many failed
    // rows there is parallelism enough and whole rows are cheaper (measured 13 ms against 38 ms for 4,494 rows at c2):
    // the rest is dealt whole, round-robin over all blocks, straight into out_keys.
    const int n_segd = min(n_fail, max_seg_rows);
    const int g = blockIdx.y;
    const int n_blocks = gridDim.x * gridDim.y, b_lin = blockIdx.y * gridDim.x + blockIdx.x;
    // iterations: first my share of the segmented rows (x, x + gridDim.x, ...), then my share of the whole rows
    const int n_it_seg = (n_segd > (int)blockIdx.x) ? (n_segd - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    const int n_whole = n_fail - n_segd;
    const int n_it_whole = (n_whole > b_lin) ? (n_whole - b_lin + n_blocks - 1) / n_blocks : 0;
    for (int it = 0; it < n_it_seg + n_it_whole; ++it) {
        const bool whole = it >= n_it_seg;
        const int i = whole ? n_segd + b_lin + (it - n_it_seg) * n_blocks : (int)blockIdx.x + it * (int)gridDim.x;
        const int64_t w = (int64_t)i * n_seg + g;
        const int64_t r = fail_list[i];
        __syncthreads();
        for (int k = threadIdx.x; k < d; k += K2_THREADS) u_s[k] = U[r * ld_u + k];
        __syncthreads();
        RowDot src;
        src.u = u_s;
        src.V = V;
        src.ld_v = ld_v;
        src.d = d;
        src.bias = bias;
        src.vec_ok = ((ld_v & 3) == 0) && ((reinterpret_cast<uintptr_t>(V) & 15) == 0);
        int64_t tb = 0, te = 0;
        if (tr_indptr != nullptr) { tb = __ldg(tr_indptr + row0 + r); te = __ldg(tr_indptr + row0 + r + 1); }
        u64 *dst = whole ? out_keys + r * K : part + w * (int64_t)K;
        const int j_begin = whole ? 0 : g * seg_items, j_end = whole ? 0x7fffffff : (g + 1) * seg_items;
        topk_row_block(src, n_items, tb, te, tr_idx, K, dst, keys, bitmap, &s_cnt, j_begin, j_end);
    }
}

// n_seg partial lists of every failed row -> its sorted top-K in out_keys[row].  One block per row at a time: the
// lists (<= 1,024 keys) are sorted by the block bitonic network in shared memory (a 1,024-key network in one warp's
// registers -- the round-1 version -- took 111 us for a handful of rows).
template <int PER>
__global__ void __launch_bounds__(128)
k_merge_fail(const int32_t *__restrict__ fail_list, const int *__restrict__ fail_count, const u64 *__restrict__ part, int n_seg, int K,
             int max_seg_rows, u64 *__restrict__ out_keys)
{
    pdl_wait();
    pdl_trigger();
    __shared__ u64 s_keys[32 * PER];
    const int n_fail = min(*fail_count, max_seg_rows);
    const int n = n_seg * K;
    const int n_sort = next_pow2(n < 2 ? 2 : n);
    for (int i = blockIdx.x; i < n_fail; i += gridDim.x) {
        const u64 *src = part + (int64_t)i * n;
        __syncthreads();
        for (int j = threadIdx.x; j < n_sort; j += 128) s_keys[j] = (j < n) ? src[j] : 0ull;
        __syncthreads();
        block_bitonic_desc(s_keys, n_sort, threadIdx.x, 128);
        u64 *dst = out_keys + (int64_t)fail_list[i] * K;
        for (int j = threadIdx.x; j < K; j += 128) dst[j] = s_keys[j];
    }
}

}  // namespace skr
