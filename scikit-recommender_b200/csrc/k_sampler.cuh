// k_sampler.cuh -- GPU form of the reference's negative sampler.
//
// Reference: c_batch_randint_choice / c_randint_choice / _random_int (skrec/utils/py/cython/include/randint.h:22-128;
// callers data_iterator.py:81-94 `randint_choice(num_items, size=n_pos*num_neg, exclusion=user_pos_dict[user])`,
// SASRec.py:358): for every batch element b, size[b] integers from [0, high), uniformly or by a probability vector,
// with or without replacement, never a member of the element's exclusion set.  The reference draws from ONE
// std::mt19937 shared by all rows (sequential; its thread-pool variant races on it); its stream cannot be reproduced
// in parallel, so parity here is distributional: range, exclusion, distinctness, uniformity (tests).
//
// Here every draw is a pure function of (seed, output position, attempt) through the counter-based Philox4x32-10
// generator: results do not depend on the launch geometry and an evaluation can be replayed.
//   replace = true : one thread per output; rejection against the row's sorted exclusion CSR by binary search.
//   replace = false: one warp per batch element; rounds of 32 candidates, rejected when excluded, already chosen, or
//                    drawn twice in the round (lowest lane wins), appended in lane order.
// Probabilities: an inclusive prefix-sum table (one shared row or one row per element); a draw is a binary search.
#pragma once
#include "common.cuh"

namespace skr {

struct Philox {
    uint32_t k0, k1;
    __device__ __forceinline__ static void round(uint32_t (&c)[4], uint32_t k0, uint32_t k1)
    {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
        const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    }
    // 4 x 32 random bits for counter (a, b, c, d)
    __device__ __forceinline__ void draw(uint32_t a, uint32_t b, uint32_t c_, uint32_t d, uint32_t (&out)[4]) const
    {
        uint32_t c[4] = {a, b, c_, d};
        uint32_t x0 = k0, x1 = k1;
#pragma unroll
        for (int r = 0; r < 10; ++r) {
            round(c, x0, x1);
            x0 += 0x9E3779B9u;
            x1 += 0xBB67AE85u;
        }
        out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
    }
};

struct SamplerArgs {
    int64_t high;
    const int64_t *out_indptr;   // [n_batch + 1]: element b owns out[out_indptr[b] .. out_indptr[b + 1])
    int64_t n_batch;
    const float *cdf;            // inclusive prefix sums of the probabilities, [high] or [n_batch, high]; null = uniform
    int cdf_per_row;
    const int64_t *excl_indptr;  // exclusion CSR (rows sorted, unique) or null
    const int32_t *excl_idx;
    uint32_t seed_lo, seed_hi;
    int32_t *out;
    int *err;                    // set to 41 when a row cannot be filled (more requested than can exist)
};

// candidate for (row b, position pos, attempt): 32 random bits -> [0, high)
__device__ __forceinline__ int32_t sampler_candidate(const SamplerArgs &A, int64_t b, uint32_t r)
{
    if (A.cdf == nullptr) return (int32_t)(((unsigned long long)r * (unsigned long long)A.high) >> 32);  // bias <= high / 2^32
    const float *cdf = A.cdf + (A.cdf_per_row ? b * A.high : 0);
    const float total = __ldg(cdf + A.high - 1);
    const float u = ((float)(r >> 8) + 0.5f) * (1.0f / 16777216.0f) * total;  // (0, total)
    int64_t lo = 0, hi = A.high - 1;  // first index with cdf > u (an index with zero probability is never returned)
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (__ldg(cdf + mid) > u) hi = mid; else lo = mid + 1;
    }
    return (int32_t)lo;
}

__device__ __forceinline__ bool sampler_excluded(const SamplerArgs &A, int64_t b, int32_t x)
{
    if (A.excl_indptr == nullptr) return false;
    const int64_t e0 = __ldg(A.excl_indptr + b), e1 = __ldg(A.excl_indptr + b + 1);
    return sorted_contains(A.excl_idx + e0, (int)(e1 - e0), x);
}

// replace = true: one thread per output position
__global__ void __launch_bounds__(256)
k_sample_with_replacement(SamplerArgs A, int64_t n_out)
{
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_out) return;
    int64_t lo = 0, hi = A.n_batch;  // row of position p: last b with out_indptr[b] <= p
    while (hi - lo > 1) {
        const int64_t mid = (lo + hi) >> 1;
        if (__ldg(A.out_indptr + mid) <= p) lo = mid; else hi = mid;
    }
    const int64_t b = lo;
    const Philox g = {A.seed_lo, A.seed_hi};
    for (uint32_t round = 0; round < (1u << 16); ++round) {
        uint32_t r[4];
        g.draw((uint32_t)p, (uint32_t)(p >> 32), round, 0x5eedu, r);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int32_t x = sampler_candidate(A, b, r[q]);
            if (!sampler_excluded(A, b, x)) { A.out[p] = x; return; }
        }
    }
    A.out[p] = -1;
    atomicExch(A.err, 41);
}

// replace = false: one warp per batch element
__global__ void __launch_bounds__(128)
k_sample_without_replacement(SamplerArgs A)
{
    const int lane = threadIdx.x & 31;
    const int64_t b = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (b >= A.n_batch) return;
    const int64_t o0 = __ldg(A.out_indptr + b), o1 = __ldg(A.out_indptr + b + 1);
    const int64_t want = o1 - o0;
    int32_t *out = A.out + o0;
    const Philox g = {A.seed_lo, A.seed_hi};
    int64_t have = 0;
    for (uint32_t round = 0; have < want; ++round) {
        if (round >= (1u << 20)) {  // the row asks for more distinct values than exist outside its exclusion set
            if (lane == 0) atomicExch(A.err, 41);
            for (int64_t i = have + lane; i < want; i += 32) out[i] = -1;
            return;
        }
        uint32_t r[4];
        g.draw((uint32_t)b, (uint32_t)(b >> 32), round, 0x0d15u + (uint32_t)lane, r);
        const int32_t x = sampler_candidate(A, b, r[0]);
        bool ok = !sampler_excluded(A, b, x);
        for (int64_t i = 0; ok && i < have; ++i) ok = out[i] != x;            // already chosen in an earlier round
        const unsigned same = __match_any_sync(0xffffffffu, x);               // drawn twice in this round: the lowest lane keeps it
        ok = ok && (__ffs(same) - 1 == lane);
        const unsigned bal = __ballot_sync(0xffffffffu, ok);
        const int64_t pos = have + __popc(bal & ((1u << lane) - 1u));
        if (ok && pos < want) out[pos] = x;
        have += __popc(bal);
        __syncwarp();
    }
}

}  // namespace skr
