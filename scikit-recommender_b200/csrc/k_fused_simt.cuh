// k_fused_simt.cuh -- fused score + bias + train mask + running top-K on the FP32 FMA pipe.
//
// precision = SKR_PREC_FP32: exact FP32 products, k accumulated in ascending order per output.
// It is the reference-grade path for shapes the tcgen05 kernel does not take (d > 128, d not a
// multiple of 32, large K) and the on-GPU cross-check of the 3xTF32 kernel.  Same work
// decomposition, heaps, thresholds and partial-list output as k_fused_tc.cuh.
//
// CTA = 256 threads, tile 128 users x 128 items, 8x8 outputs per thread (rows ty+16i, columns
// tx+16j), operands staged in shared memory in k-chunks of 32 with cp.async double buffering.
#pragma once
#include "fused_common.cuh"

namespace skr {

constexpr int SIMT_THREADS = 256;
constexpr int SIMT_KC = 32;   // k chunk
constexpr int SIMT_LD = 36;   // padded smem row (floats): conflict-free LDS.128
constexpr int SIMT_CAP = 16;  // staged survivors per row per column step (16 columns per step)

// K = 0: the score-block variant (no heaps; the other small arrays are laid out but unused)
__host__ __device__ inline size_t simt_smem_bytes(int K)
{
    return (size_t)(2 * 2 * TM * SIMT_LD) * 4   // As[2], Bs[2]
           + (size_t)K * TM * 8                  // heaps
           + (size_t)SIMT_CAP * TM * 8           // staging
           + (size_t)TM * 4 * 3                  // scnt, thr_row, hcnt
           + (size_t)4 * TM * 4                  // bitmap
           + 64;
}

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem, int src_bytes)
{
    uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

// loads rows [r0, r0+128) x k [k0, k0+32) of X (row stride ld floats) into dst[128][SIMT_LD]
__device__ __forceinline__ void simt_load_chunk(float *dst, const float *__restrict__ X, int64_t ld, int64_t r0,
                                                int64_t n_rows, int k0, int d, int tid)
{
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        int idx = tid + q * SIMT_THREADS;  // 0..1023
        int m = idx >> 3, k4 = (idx & 7) * 4;
        int64_t r = r0 + m;
        bool ok = (r < n_rows) && (k0 + k4 < d);
        const float *src = ok ? (X + r * ld + k0 + k4) : X;
        cp_async16(dst + m * SIMT_LD + k4, src, ok ? 16 : 0);
    }
}

// SCORES = true: the same FP32 tile main loop, but the tile (+ bias) is written to a score block
// scores_out[row, item] (row pitch ld_out) instead of going through the heaps -- the library's own GEMM for the
// shapes the selection epilogues do not take (top-K > 128): the block is consumed by k_topk_scores, which masks the
// train items itself, so no bitmap is built here.  One CTA per (user tile, item tile range).
// SMODE = 1: score = -||u - i|| (+ bias) instead of u . i (+ bias): the translation scorers of the reference
// (TransRec.py:86-93 `-l2_distance(u + g + last, I) + b`, SGAT.py:300) -- a square root next to a per-item bias is not
// a monotone image of a dot product, so it cannot go through the tensor-core contraction.  The tile loop accumulates
// (u_k - i_k)^2, k ascending (the reference's torch.norm(a - b): no ||u||^2 - 2 u.i + ||i||^2 cancellation), the epilogue
// takes -sqrt.
template <bool SCORES, int SMODE>
__global__ void __launch_bounds__(SIMT_THREADS, 1)
k_fused_simt(const float *__restrict__ U, int64_t ld_u, const float *__restrict__ V, int64_t ld_v, FusedParams P,
             float *__restrict__ scores_out, int64_t ld_out)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *As = reinterpret_cast<float *>(smem_raw);
    float *Bs = As + 2 * TM * SIMT_LD;
    u64 *heap = reinterpret_cast<u64 *>(Bs + 2 * TN * SIMT_LD);
    u64 *stage = heap + (size_t)P.K * TM;
    int *scnt = reinterpret_cast<int *>(stage + SIMT_CAP * TM);
    float *thr_row = reinterpret_cast<float *>(scnt + TM);
    int *hcnt = reinterpret_cast<int *>(thr_row + TM);
    uint32_t *bitmap = reinterpret_cast<uint32_t *>(hcnt + TM);
    __shared__ long long s_mcur;

    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int c = blockIdx.x / P.n_rt, rt = blockIdx.x % P.n_rt;
    const int t0 = c * P.tiles_per_chunk;
    const int t1 = min(t0 + P.tiles_per_chunk, P.n_ct);
    const int64_t row_base = (int64_t)rt * TM;
    const int K = P.K;
    const float NINF = -__int_as_float(0x7f800000);

    // owner state: thread r < TM owns row r's heap
    int hn = 0;
    float thr = NINF;
    uint32_t published = 0;
    const int64_t my_row = row_base + tid;
    const bool owner = tid < TM;
    const bool my_valid = owner && my_row < P.n_rows;
    if (!SCORES && owner) { scnt[tid] = 0; thr_row[tid] = my_valid ? NINF : -NINF; }  // rows beyond n_rows collect nothing

    // cursor into this user tile's mask keys
    int64_t mcur = 0, mend = 0;
    if (!SCORES && P.mask_keys != nullptr) {
        const int64_t rt_abs = (P.row0 / TM) + rt;
        if (tid == 0) {
            int64_t lo = P.mask_tile_ptr[rt_abs], hi = P.mask_tile_ptr[rt_abs + 1];
            s_mcur = lower_bound_u32(P.mask_keys, lo, hi, ((uint32_t)(t0 * TN)) << 7);
        }
        mend = P.mask_tile_ptr[rt_abs + 1];
    }
    __syncthreads();
    if (!SCORES && P.mask_keys != nullptr) mcur = s_mcur;

    const int n_kc = (P.d + SIMT_KC - 1) / SIMT_KC;

    for (int t = t0; t < t1; ++t) {
        const int col0 = t * TN;
        // kick off the first operand chunk
        simt_load_chunk(As, U, ld_u, row_base, P.n_rows, 0, P.d, tid);
        simt_load_chunk(Bs, V, ld_v, col0, P.n_items, 0, P.d, tid);
        cp_async_commit();

        // thresholds published by other CTAs for these rows
        if (!SCORES && my_valid) {
            uint32_t g = P.thr_g[my_row];
            if (g != 0) {
                float gf = unord_f32(g);
                if (gf > thr) thr = gf;
            }
            thr_row[tid] = thr;
        }
        // bitmap: out-of-range columns, then this tile's train items
        if (!SCORES) for (int i = tid; i < 4 * TM; i += SIMT_THREADS) bitmap[i] = oob_bits(col0, i / TM, P.n_items);
        __syncthreads();
        if (!SCORES && P.mask_keys != nullptr) {
            const uint32_t lim = ((uint32_t)(col0 + TN)) << 7;
            for (;;) {
                int64_t p = mcur + tid;
                uint32_t key = (p < mend) ? __ldg(P.mask_keys + p) : 0xffffffffu;
                bool in = key < lim;
                if (in) {
                    int cc = (int)(key >> 7) - col0;
                    atomicOr(&bitmap[(cc >> 5) * TM + (int)(key & 127u)], 1u << (cc & 31));
                }
                int n_in = __syncthreads_count(in);
                mcur += n_in;
                if (n_in < SIMT_THREADS) break;
            }
        }

        float acc[8][8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;

        for (int kc = 0; kc < n_kc; ++kc) {
            const int buf = kc & 1;
            if (kc + 1 < n_kc) {
                simt_load_chunk(As + (buf ^ 1) * TM * SIMT_LD, U, ld_u, row_base, P.n_rows, (kc + 1) * SIMT_KC, P.d, tid);
                simt_load_chunk(Bs + (buf ^ 1) * TN * SIMT_LD, V, ld_v, col0, P.n_items, (kc + 1) * SIMT_KC, P.d, tid);
                cp_async_commit();
                cp_async_wait<1>();
            } else {
                cp_async_wait<0>();
            }
            __syncthreads();
            const float *a_s = As + buf * TM * SIMT_LD;
            const float *b_s = Bs + buf * TN * SIMT_LD;
#pragma unroll
            for (int kk = 0; kk < SIMT_KC; kk += 4) {
                float4 a4[8], b4[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) a4[i] = *reinterpret_cast<const float4 *>(a_s + (ty + 16 * i) * SIMT_LD + kk);
#pragma unroll
                for (int j = 0; j < 8; ++j) b4[j] = *reinterpret_cast<const float4 *>(b_s + (tx + 16 * j) * SIMT_LD + kk);
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        if (SMODE == 0) {
                            acc[i][j] = fmaf(a4[i].x, b4[j].x, acc[i][j]);
                            acc[i][j] = fmaf(a4[i].y, b4[j].y, acc[i][j]);
                            acc[i][j] = fmaf(a4[i].z, b4[j].z, acc[i][j]);
                            acc[i][j] = fmaf(a4[i].w, b4[j].w, acc[i][j]);
                        } else {
                            const float t0 = a4[i].x - b4[j].x, t1 = a4[i].y - b4[j].y, t2 = a4[i].z - b4[j].z, t3 = a4[i].w - b4[j].w;
                            acc[i][j] = fmaf(t0, t0, acc[i][j]);
                            acc[i][j] = fmaf(t1, t1, acc[i][j]);
                            acc[i][j] = fmaf(t2, t2, acc[i][j]);
                            acc[i][j] = fmaf(t3, t3, acc[i][j]);
                        }
                    }
            }
            __syncthreads();
        }

        if (SCORES) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int gcol = col0 + tx + 16 * j;
                if (gcol >= P.n_items) continue;
                const float b = (P.bias != nullptr) ? __ldg(P.bias + gcol) : 0.0f;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int64_t row = row_base + ty + 16 * i;
                    if (row < P.n_rows) scores_out[row * ld_out + gcol] = (SMODE == 0 ? acc[i][j] : -sqrtf(acc[i][j])) + b;
                }
            }
            continue;
        }
        // epilogue: 8 steps of 16 columns; survivors are staged, owners fold them into the heaps
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int cc = tx + 16 * j;
            const int gcol = col0 + cc;
            const float b = (P.bias != nullptr) ? __ldg(P.bias + gcol) : 0.0f;
            bool any = false;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int r = ty + 16 * i;
                const float s = (SMODE == 0 ? acc[i][j] : -sqrtf(acc[i][j])) + b;
                if (s >= thr_row[r]) {
                    if (((bitmap[(cc >> 5) * TM + r] >> (cc & 31)) & 1u) == 0u) {
                        int slot = atomicAdd(&scnt[r], 1);
                        stage[slot * TM + r] = make_key(s, (uint32_t)gcol);
                        any = true;
                    }
                }
            }
            if (__syncthreads_or(any)) {
                if (owner) {
                    const int n = scnt[tid];
                    if (n > 0) {
                        for (int q = 0; q < n; ++q) heap_push(heap + tid, hn, K, stage[q * TM + tid]);
                        scnt[tid] = 0;
                        if (hn == K) {
                            float rs = key_score(heap[tid]);
                            if (rs > thr) thr = rs;
                            thr_row[tid] = thr;
                        }
                    }
                }
                __syncthreads();
            }
        }
        // publish the row threshold for the other chunks of this user tile
        if (my_valid && hn == K) {
            uint32_t o = (uint32_t)(heap[tid] >> 32);
            if (o > published) { atomicMax(P.thr_g + my_row, o); published = o; }
        }
    }

    if (SCORES) return;
    // partial lists out: [row, c, K], coalesced along K
    if (owner) hcnt[tid] = hn;
    __syncthreads();
    for (int idx = tid; idx < TM * K; idx += SIMT_THREADS) {
        const int r = idx / K, i = idx - r * K;
        const int64_t row = row_base + r;
        if (row < P.n_rows) P.part[(row * P.S + c) * K + i] = (i < hcnt[r]) ? heap[i * TM + r] : 0ull;
    }
}

}  // namespace skr
