// common.cuh -- shared device helpers: rank keys, block bitonic sort, binary search.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace skr {

// Programmatic dependent launch: kernels of one evaluate are launched with programmatic stream serialisation, so
// the next kernel's blocks are set up while this one drains.  pdl_wait() must precede the first access to anything
// an earlier kernel of the stream wrote (it is a no-op for a plain launch); pdl_trigger() lets the dependent grid be
// scheduled once every block of this grid has got this far.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

typedef unsigned long long u64;

// ---------------------------------------------------------------------------------------------
// Rank key.  key = ord(score) << 32 | ~item  so that "larger key" == "ranked earlier" under the
// contract score desc, item id asc (the comparator of evaluate.h:43 plus the documented tie rule).
// ord() is the usual monotone float->uint map; -0.0 is folded onto +0.0 (the reference's `>`
// treats them as equal) and NaN is ranked like -inf (undefined in the reference, App. A.4).
// A valid key is never 0, so 0 is the "empty slot" value.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t ord_f32(float s)
{
    s = s + 0.0f;                                  // -0.0 -> +0.0
    if (s != s) s = -__int_as_float(0x7f800000);   // NaN -> -inf
    uint32_t b = __float_as_uint(s);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float unord_f32(uint32_t o)
{
    uint32_t b = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
    return __uint_as_float(b);
}
__device__ __forceinline__ u64 make_key(float s, uint32_t item) { return ((u64)ord_f32(s) << 32) | (u64)(~item); }
__device__ __forceinline__ uint32_t key_item(u64 k) { return ~(uint32_t)k; }
__device__ __forceinline__ float key_score(u64 k) { return unord_f32((uint32_t)(k >> 32)); }

// Block-wide bitonic sort of n (power of two) keys in shared memory, descending.
__device__ __forceinline__ void block_bitonic_desc(u64 *keys, int n, int tid, int nthreads)
{
    for (int k = 2; k <= n; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int p = tid; p < (n >> 1); p += nthreads) {
                int i = ((p & ~(j - 1)) << 1) | (p & (j - 1));
                int x = i | j;
                u64 a = keys[i], b = keys[x];
                bool desc = ((i & k) == 0);
                if (desc ? (a < b) : (a > b)) { keys[i] = b; keys[x] = a; }
            }
            __syncthreads();
        }
    }
}

// Warp-wide bitonic sort, descending, of 32*PER keys held PER per lane (element e of lane l is
// logical index e*32 + l).  Used by the partial-list merge.
template <int PER>
__device__ __forceinline__ void warp_bitonic_desc(u64 (&v)[PER], int lane)
{
    constexpr int N = 32 * PER;
#pragma unroll
    for (int k = 2; k <= N; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            if (j >= 32) {
                const int je = j >> 5;  // partner differs in the element index
#pragma unroll
                for (int e = 0; e < PER; ++e) {
                    if ((e & je) == 0) {
                        int i = e * 32 + lane;
                        bool desc = ((i & k) == 0);
                        u64 a = v[e], b = v[e | je];
                        if (desc ? (a < b) : (a > b)) { v[e] = b; v[e | je] = a; }
                    }
                }
            } else {
#pragma unroll
                for (int e = 0; e < PER; ++e) {
                    int i = e * 32 + lane;
                    u64 other = __shfl_xor_sync(0xffffffffu, v[e], j);
                    bool lower = ((lane & j) == 0);
                    bool desc = ((i & k) == 0);
                    // the lower index keeps the larger key when the run is descending
                    bool take_max = (lower == desc);
                    u64 mx = v[e] > other ? v[e] : other;
                    u64 mn = v[e] > other ? other : v[e];
                    v[e] = take_max ? mx : mn;
                }
            }
        }
    }
}

// Same network on 32-bit keys: a compare-exchange is one shuffle, one min/max and one select instead of
// the ~10 instructions a 64-bit key costs.
template <int PER>
__device__ __forceinline__ void warp_bitonic_desc32(uint32_t (&v)[PER], int lane)
{
    constexpr int N = 32 * PER;
#pragma unroll
    for (int k = 2; k <= N; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            if (j >= 32) {
                const int je = j >> 5;
#pragma unroll
                for (int e = 0; e < PER; ++e) {
                    if ((e & je) == 0) {
                        const int i = e * 32 + lane;
                        const bool desc = ((i & k) == 0);
                        const uint32_t a = v[e], b = v[e | je];
                        const uint32_t mx = max(a, b), mn = min(a, b);
                        v[e] = desc ? mx : mn;
                        v[e | je] = desc ? mn : mx;
                    }
                }
            } else {
#pragma unroll
                for (int e = 0; e < PER; ++e) {
                    const int i = e * 32 + lane;
                    const uint32_t other = __shfl_xor_sync(0xffffffffu, v[e], j);
                    const bool take_max = (((lane & j) == 0) == ((i & k) == 0));
                    v[e] = take_max ? max(v[e], other) : min(v[e], other);
                }
            }
        }
    }
}

// true iff x is in the sorted int32 array a[0..n)
__device__ __forceinline__ bool sorted_contains(const int32_t *__restrict__ a, int n, int32_t x)
{
    int lo = 0, hi = n;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (__ldg(a + mid) < x) lo = mid + 1; else hi = mid;
    }
    return lo < n && __ldg(a + lo) == x;
}

__host__ __device__ __forceinline__ int next_pow2(int x)
{
    int p = 1;
    while (p < x) p <<= 1;
    return p;
}

}  // namespace skr
