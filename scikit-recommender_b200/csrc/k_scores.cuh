// k_scores.cuh -- K2: score-matrix-in top-K with train-item masking.
//
// Replaces, for a device-resident float32 [B, I] score block, the reference's masking loop
// (evaluator.py:195-200) and the per-row selection of eval_one_user (evaluate.h:27-45).
// One CTA streams one row exactly once (HBM-bound: 4*I bytes per user); a running threshold
// (the K-th key so far) filters the stream, survivors go to a shared-memory buffer that is
// folded into the sorted top-K by a block bitonic sort whenever it could overflow.
// Masked (train) items become -inf and stay candidates, like the reference.
#pragma once
#include "common.cuh"

namespace skr {

constexpr int K2_THREADS = 256;
constexpr int K2_CHUNK = 1024;  // elements per pass: one float4 per thread
constexpr int K2_P = 2048;      // key slots in shared memory (sorted top-K + survivor buffer)
constexpr int K2_MAX_K = 512;

__device__ __forceinline__ float4 ldg_stream_f4(const float *p)
{
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}

__device__ __forceinline__ void k2_load4(const float *__restrict__ row, int j0, int n_items, bool vec_ok, float (&v)[4])
{
    if (vec_ok && j0 + 3 < n_items) {
        float4 t = ldg_stream_f4(row + j0);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) v[q] = (j0 + q < n_items) ? __ldg(row + j0 + q) : 0.0f;
    }
}

__global__ void __launch_bounds__(K2_THREADS)
k_topk_scores(const float *__restrict__ scores, int64_t ld, int n_items, int64_t row0,
              const int64_t *__restrict__ tr_indptr, const int32_t *__restrict__ tr_idx, int K,
              u64 *__restrict__ out_keys)
{
    __shared__ u64 keys[K2_P];
    __shared__ uint32_t bitmap[K2_CHUNK / 32];
    __shared__ int s_cnt;

    const int tid = threadIdx.x, lane = tid & 31;
    const int64_t r = blockIdx.x;
    const float *row = scores + r * ld;
    const bool vec_ok = ((reinterpret_cast<uintptr_t>(row) & 15) == 0);

    int64_t cur = 0, te = 0;
    if (tr_indptr != nullptr) { cur = __ldg(tr_indptr + row0 + r); te = __ldg(tr_indptr + row0 + r + 1); }
    int next_train = (cur < te) ? __ldg(tr_idx + cur) : 0x7fffffff;

    int base = 0;  // keys[0..base) hold the sorted best-so-far
    u64 thr_key = 0;
    float thr_f = -__int_as_float(0x7f800000);
    if (tid == 0) s_cnt = 0;
    __syncthreads();

    float v[4], nxt[4];
    k2_load4(row, tid * 4, n_items, vec_ok, v);

    for (int c0 = 0; c0 < n_items; c0 += K2_CHUNK) {
        const int j0 = c0 + tid * 4;
        if (c0 + K2_CHUNK < n_items) k2_load4(row, j0 + K2_CHUNK, n_items, vec_ok, nxt);

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
            bool pass = (j0 + q < n_items) && !(s < thr_f);
            u64 key = 0;
            if (pass) {
                key = make_key(s, (uint32_t)(j0 + q));
                pass = key > thr_key;
            }
            unsigned bal = __ballot_sync(0xffffffffu, pass);
            if (bal) {
                int leader = __ffs(bal) - 1;
                int pos = 0;
                if (lane == leader) pos = atomicAdd(&s_cnt, __popc(bal));
                pos = __shfl_sync(0xffffffffu, pos, leader);
                if (pass) keys[base + pos + __popc(bal & ((1u << lane) - 1u))] = key;
            }
        }
        __syncthreads();
        const int cnt = s_cnt;
        const bool last = (c0 + K2_CHUNK >= n_items);
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
            if (tid == 0) s_cnt = 0;
            __syncthreads();
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) v[q] = nxt[q];
    }
    for (int i = tid; i < K; i += K2_THREADS) out_keys[r * K + i] = keys[i];
}

}  // namespace skr
