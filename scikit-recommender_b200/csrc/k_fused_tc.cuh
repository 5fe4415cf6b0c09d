// k_fused_tc.cuh -- fused score + bias + train mask + candidate selection on tcgen05 / TMEM (sm_100a).
//
// scores = U_tile (128 users) x I_tile^T (128 items) are produced by tcgen05.mma kind::tf32 with
// FP32 accumulation in TMEM and consumed straight out of TMEM by the epilogue warps: the U x I
// score matrix never exists in shared or global memory.
//
// Reference-grade scores from TF32 tensor cores (3xTF32): x = hi + lo with hi = rna_tf32(x),
// lo = rna_tf32(x - hi); score = sum lo_u*hi_i + hi_u*lo_i + hi_u*hi_i, all three products
// accumulated into the same TMEM tile.  hi and lo are exactly representable in TF32, so the
// tensor core's operand rounding mode does not matter.  (1xTF32 breaks the
// 1e-5 metric contract, 3xTF32 does not.)
//
// Selection (round-1 ncu finding: per-row heaps in shared memory made the epilogue, not the MMA,
// the bottleneck by 20x).  The kernel now runs in two modes over the same pipeline:
//   SAMPLE  - a strided ~f = 6/K fraction of the item tiles is scored once in 1xTF32; each row keeps
//             the R largest 32-column group maxima (train items masked) in registers.  The r-th
//             largest of them is a threshold T0 that, with overwhelming probability, has between K
//             and a few K catalogue items above it.  It only steers work: results never depend on it.
//   COLLECT - every item tile is scored in 3xTF32; a max tree over 32 scores and ONE compare against
//             T0 rejects almost every 32-column group; survivors that are not train items are
//             appended (rank key = ord(score) << 32 | ~item) to the row's candidate list in HBM.
// k_select_cands then sorts each row's ~3-5 K candidates; a row whose list is short (< K) or
// overflowed is re-done exactly by k_row_exact (k_scores.cuh).  No heaps, no thresholds to update,
// no shared memory per row: K is limited only by the list capacity.
//
// Operands: A (users) lives in TMEM for the whole work item -- each epilogue thread of the first
// warpgroup loads its user's row from global memory, splits it in registers and tcgen05.st's hi/lo
// into TMEM lanes (TS-mode MMA; no shared memory for A).  B (items) is pre-split by k_split_tf32
// into hi/lo tables and streamed by TMA (SWIZZLE_128B, 128 rows x 32 floats per box) through an
// mbarrier ring of k-block stages.
//
// Warp roles (384 threads): warps 0-3 epilogue of columns 0-63, warps 4-7 epilogue of columns 64-127,
// warp 8 TMA producer, warp 9 TMEM allocator + MMA issuer of the even tiles (one lane), warp 10
// train-mask bitmap builder, warp 11 MMA issuer of the odd tiles (thread <-> TMEM lane <-> user row; two warps per SM sub-partition so
// the TMEM-load and compare latencies of one hide behind the other).  Accumulators and bitmaps are
// double buffered so the epilogue of tile n overlaps the MMAs of tile n+1.
//
// TMEM columns: [0, 32*nkb) A_hi, [32*nkb, 64*nkb) A_lo, [256, 384) acc 0, [384, 512) acc 1.
#pragma once
#include <cuda.h>
#include "fused_common.cuh"

namespace skr {

constexpr int TC_THREADS = 384;
constexpr int TC_EPI_THREADS = 256;
constexpr int TC_KB = 32;                             // floats per k-block (one 128-byte swizzle row)
constexpr int TC_TILE_BYTES = TN * TC_KB * 4;         // 16 KB: one operand tile of one k-block
constexpr int TC_STAGE_BYTES = 2 * TC_TILE_BYTES;     // hi + lo
constexpr int TC_MAX_STAGES = 4;
constexpr int TC_ACC_COL = 256;                       // first accumulator column
constexpr int TC_R = 32;                              // group maxima kept per row and warpgroup (SAMPLE)
constexpr long long TC_TIMEOUT_CYCLES = 4000000000ll; // watchdog: ~2 s

enum { TC_MODE_COLLECT = 0, TC_MODE_SAMPLE = 1 };

__host__ __device__ inline size_t tc_smem_bytes(int stages)
{
    return (size_t)1024                          // alignment slack
           + (size_t)stages * TC_STAGE_BYTES
           + (size_t)2 * 4 * TM * 4              // two bitmaps
           + (size_t)8 * TC_EPI_THREADS * 16     // score staging: one 32-float row per epilogue thread, [8][256] float4
           + 256;                                // barriers + tmem pointer
}

// ---- PTX wrappers ------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a pipeline bug must not hang the GPU box; it traps with a flag set instead.
__device__ __noinline__ void mbar_wait_slow(uint64_t *bar, uint32_t parity, int *err_flag, int code)
{
    const long long t0 = clock64();
    for (;;) {
#pragma unroll 1
        for (int i = 0; i < 512; ++i)
            if (mbar_try_wait(bar, parity)) return;
        if (clock64() - t0 > TC_TIMEOUT_CYCLES) {
            if (err_flag != nullptr) atomicExch(err_flag, code);
            __threadfence_system();
            __trap();
        }
    }
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity, int *err_flag, int code)
{
    if (mbar_try_wait(bar, parity)) return;
    mbar_wait_slow(bar, parity, err_flag, code);
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int x, int y, uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(x), "r"(y), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]^T, kind::tf32, M=128 (TS mode)
__device__ __forceinline__ void tc_mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ uint32_t to_tf32(float x)
{
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}

#define SKR_R32(a, o) \
    "=r"(a[o + 0]), "=r"(a[o + 1]), "=r"(a[o + 2]), "=r"(a[o + 3]), "=r"(a[o + 4]), "=r"(a[o + 5]), "=r"(a[o + 6]), "=r"(a[o + 7])
#define SKR_W32(a, o) \
    "r"(a[o + 0]), "r"(a[o + 1]), "r"(a[o + 2]), "r"(a[o + 3]), "r"(a[o + 4]), "r"(a[o + 5]), "r"(a[o + 6]), "r"(a[o + 7])

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : SKR_R32(r, 0), SKR_R32(r, 8), SKR_R32(r, 16), SKR_R32(r, 24)
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32])
{
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%32], "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31};"
        ::SKR_W32(r, 0), SKR_W32(r, 8), SKR_W32(r, 16), SKR_W32(r, 24), "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, sm_100):
// start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) | version=1 [46,48) | layout=2 (SW128) [61,64).
// Rows are 128 bytes, 8-row swizzle atoms are 1024 bytes apart (SBO); LBO is unused here.
__device__ __forceinline__ uint64_t make_b_desc(uint32_t smem_addr)
{
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3fffu);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// cute::UMMA::InstrDescriptor: c_format F32 (1) [4,6) | a_format TF32 (2) [7,10) | b_format TF32 (2)
// [10,13) | a/b K-major | N>>3 [17,23) | M>>4 [24,29)
constexpr uint32_t TC_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);

struct TcArgs {
    const float *U;   // user vectors [n_rows, ld_u]
    int64_t ld_u;
    int nkb;          // k-blocks of 32 (d padded)
    int stages;
    int passes;       // 3 = 3xTF32, 1 = single TF32 pass
    int mode;         // TC_MODE_*
    int *err_flag;    // device int, set before a watchdog trap
    int dbg;          // timing experiments only (results invalid): 1 no epilogue work, 2 no MMA, 4 no appends, 8 no TMA
    // SAMPLE: item tiles 0, stride, 2*stride, ...; out: samp[row][2][TC_R] group maxima, descending
    int stride;
    int n_samp;
    float *samp;
    // COLLECT: per-row thresholds (k_sample_thr) and candidate lists
    const float *thr;        // [n_rows]
    int cap;                 // entries per (row, chunk, column half) sub-list
    int sub_stride;          // storage stride in entries: power of two >= cap + 32 (room for one group past cap)
    uint2 *cand;             // [n_rows, S*2, sub_stride] (score bits, item), base aligned to the stride
    uint32_t *cand_cnt;      // [n_rows, S*2] entries written (cap + 1 means overflow)
};

// descending insertion of x into v[0..TC_R): branch-free compare-exchange chain
__device__ __forceinline__ void sorted_insert(float (&v)[TC_R], float x)
{
#pragma unroll
    for (int i = 0; i < TC_R; ++i) {
        const float hi = fmaxf(v[i], x);
        x = fminf(v[i], x);
        v[i] = hi;
    }
}

__global__ void __launch_bounds__(TC_THREADS, 1)
k_fused_tc(const __grid_constant__ CUtensorMap tm_bhi, const __grid_constant__ CUtensorMap tm_blo, TcArgs A, FusedParams P)
{
    extern __shared__ unsigned char tc_smem_raw[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(tc_smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char *b_tiles = smem;
    uint32_t *bitmap = reinterpret_cast<uint32_t *>(smem + (size_t)A.stages * TC_STAGE_BYTES);  // [2][4][TM]
    float4 *stage_buf = reinterpret_cast<float4 *>(bitmap + 2 * 4 * TM);  // [8][TC_EPI_THREADS]
    uint64_t *bars = reinterpret_cast<uint64_t *>(stage_buf + 8 * TC_EPI_THREADS);
    uint64_t *full = bars;                            // [TC_MAX_STAGES]
    uint64_t *empty = bars + TC_MAX_STAGES;           // [TC_MAX_STAGES]
    uint64_t *tmem_full = bars + 2 * TC_MAX_STAGES;   // [2]
    uint64_t *tmem_empty = tmem_full + 2;             // [2]
    uint64_t *bm_full = tmem_empty + 2;               // [2]
    uint64_t *bm_empty = bm_full + 2;                 // [2]
    uint64_t *a_ready = bm_empty + 2;                 // [1]
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(a_ready + 1);

    // Epilogue = warps 0-7, helpers = warps 8-11: the warp scheduler favours higher warp ids, and the
    // latency-critical single-thread roles (TMA producer, MMA issuers) must not queue behind the epilogue.
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int role = warp - 8;  // 0 TMA producer, 1 MMA issuer (even tiles) + TMEM allocator, 2 mask builder, 3 MMA issuer (odd tiles)
    const int c = blockIdx.x / P.n_rt, rt = blockIdx.x % P.n_rt;
    const bool sample = (A.mode == TC_MODE_SAMPLE);
    // tiles of this work item: COLLECT t0 + i, SAMPLE i * stride
    const int t0 = sample ? 0 : c * P.tiles_per_chunk;
    const int n_tiles = sample ? A.n_samp : (min(t0 + P.tiles_per_chunk, P.n_ct) - t0);
    const int t_step = sample ? A.stride : 1;
    const int64_t row_base = (int64_t)rt * TM;
    const int nkb = A.nkb;

    if (tid == 0) {
        for (int s = 0; s < TC_MAX_STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        for (int b = 0; b < 2; ++b) {
            mbar_init(tmem_full + b, 1);
            mbar_init(tmem_empty + b, TC_EPI_THREADS / 32);
            mbar_init(bm_full + b, 1);
            mbar_init(bm_empty + b, TC_EPI_THREADS / 32);
        }
        mbar_init(a_ready, TM / 32);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (role == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_bhi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_blo) : "memory");
    }
    if (role == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_ptr)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (role == 0) {
        // ===== TMA producer: item k-block tiles (hi, lo) through the stage ring ==================
        if (lane == 0) {
            const uint32_t tx_bytes = (A.passes == 3) ? TC_STAGE_BYTES : TC_TILE_BYTES;
            int it = 0;
            for (int i = 0; i < n_tiles; ++i) {
                const int t = t0 + i * t_step;
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const int s = it % A.stages;
                    const uint32_t ph = (uint32_t)((it / A.stages) & 1);
                    mbar_wait(empty + s, ph ^ 1u, A.err_flag, 1);
                    if (A.dbg & 8) { mbar_arrive(full + s); continue; }
                    mbar_expect_tx(full + s, tx_bytes);
                    unsigned char *dst = b_tiles + (size_t)s * TC_STAGE_BYTES;
                    tma_load_2d(dst, &tm_bhi, kb * TC_KB, t * TN, full + s);
                    if (A.passes == 3) tma_load_2d(dst + TC_TILE_BYTES, &tm_blo, kb * TC_KB, t * TN, full + s);
                }
            }
        }
    } else if (role == 1 || role == 3) {
        // ===== MMA issuers: warp 1 takes the even tiles (accumulator 0), warp 3 the odd ones (accumulator 1).
        // One thread can issue a tcgen05.mma only every ~80 cycles (measured: MMA-only time = 1434 + 78 n
        // cycles per tile for n MMAs), which is slower than an M128 N128 K8 MMA executes (64 cycles); two
        // issuers keep the tensor pipe fed and hide each other's per-tile barrier latencies.
        if (lane == 0) {
            const int p = (role == 1) ? 0 : 1;
            mbar_wait(a_ready, 0, A.err_flag, 2);
            tc_fence_after();
            const uint32_t a_hi0 = tmem_base;
            const uint32_t a_lo0 = tmem_base + (uint32_t)(nkb * TC_KB);
            const uint32_t d_tmem = tmem_base + (uint32_t)(TC_ACC_COL + p * TN);
            for (int i = p; i < n_tiles; i += 2) {
                const uint32_t u = (uint32_t)((i >> 1) & 1);
                mbar_wait(tmem_empty + p, u ^ 1u, A.err_flag, 3);
                tc_fence_after();
                uint32_t acc = 0;
                int it = i * nkb;
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const int s = it % A.stages;
                    const uint32_t ph = (uint32_t)((it / A.stages) & 1);
                    mbar_wait(full + s, ph, A.err_flag, 4);
                    tc_fence_after();
                    const uint64_t d0 = make_b_desc(smem_u32(b_tiles + (size_t)s * TC_STAGE_BYTES));
#pragma unroll
                    for (int k8 = 0; k8 < ((A.dbg & 2) ? 0 : 4); ++k8) {  // UMMA K = 8 tf32 = 32 bytes
                        const uint32_t acol = (uint32_t)(kb * TC_KB + k8 * 8);
                        const uint64_t dhi = d0 + (uint64_t)(k8 * 2);  // start-address field counts 16-byte units
                        if (A.passes == 3) {
                            const uint64_t dlo = dhi + (uint64_t)(TC_TILE_BYTES >> 4);
                            tc_mma_ts(d_tmem, a_lo0 + acol, dhi, TC_IDESC, acc);
                            tc_mma_ts(d_tmem, a_hi0 + acol, dlo, TC_IDESC, 1u);
                            tc_mma_ts(d_tmem, a_hi0 + acol, dhi, TC_IDESC, 1u);
                        } else {
                            tc_mma_ts(d_tmem, a_hi0 + acol, dhi, TC_IDESC, acc);
                        }
                        acc = 1u;
                    }
                    tc_commit(empty + s);  // stage reusable once these MMAs have read it
                }
                tc_commit(tmem_full + p);  // accumulator p complete
            }
        }
    } else if (role == 2) {
        // ===== train-mask bitmap builder ===========================================================
        // Keys of this user tile are sorted by item; mask_tile_off gives, per item tile, where its keys
        // start, so nothing is searched or skipped and every load address is known tiles ahead: the
        // offsets and first 32 keys of tile i+1 are fetched while tile i is being built.
        const uint32_t *keys = nullptr;
        const uint32_t *offs = nullptr;
        if (P.mask_keys != nullptr) {
            const int64_t rt_abs = (P.row0 / TM) + rt;
            keys = P.mask_keys + __ldg(P.mask_tile_ptr + rt_abs);
            offs = P.mask_tile_off + rt_abs * (int64_t)(P.n_ct + 1);
        }
        uint32_t nb = 0, ne = 0, nkey = 0xffffffffu;  // next tile: key range and first batch
        if (keys != nullptr && n_tiles > 0) {
            nb = __ldg(offs + t0);
            ne = __ldg(offs + t0 + 1);
            if (nb + lane < ne) nkey = __ldg(keys + nb + lane);
        }
        for (int i = 0; i < n_tiles; ++i) {
            const int b = i & 1;
            const uint32_t u = (uint32_t)((i >> 1) & 1);
            const int col0 = (t0 + i * t_step) * TN;
            const uint32_t kb0 = nb, ke0 = ne;
            uint32_t key = nkey;
            if (keys != nullptr && i + 1 < n_tiles) {  // prefetch for tile i + 1
                const int tn = t0 + (i + 1) * t_step;
                nb = __ldg(offs + tn);
                ne = __ldg(offs + tn + 1);
            }
            mbar_wait(bm_empty + b, u ^ 1u, A.err_flag, 5);
            uint32_t *bm = bitmap + b * 4 * TM;
            for (int q = lane; q < 4 * TM; q += 32) bm[q] = oob_bits(col0, q / TM, P.n_items);
            __syncwarp();
            if (keys != nullptr) {
                for (uint32_t p = kb0; p < ke0; p += 32) {
                    if (p != kb0) key = (p + lane < ke0) ? __ldg(keys + p + lane) : 0xffffffffu;
                    if (p + lane < ke0) {
                        const int cc = (int)(key >> 7) - col0;
                        atomicOr(&bm[(cc >> 5) * TM + (int)(key & 127u)], 1u << (cc & 31));
                    }
                }
                nkey = 0xffffffffu;
                if (i + 1 < n_tiles && nb + lane < ne) nkey = __ldg(keys + nb + lane);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(bm_full + b);
        }
    } else if (role < 0) {
        // ===== epilogue: thread <-> user row, warpgroup wg <-> columns [64 wg, 64 wg + 64) ==========
        const int wg = warp >> 2;
        const int r = tid & 127;  // TMEM lane
        const int64_t my_row = row_base + r;
        const bool my_valid = my_row < P.n_rows;
        const uint32_t lane_addr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
        const float NINF = -__int_as_float(0x7f800000);
        const float PINF = __int_as_float(0x7f800000);

        if (wg == 0) {  // A: my user's vector -> hi/lo TF32 -> TMEM
            const float *urow = A.U + (my_valid ? my_row : 0) * A.ld_u;
            const bool vec = ((A.ld_u & 3) == 0) && ((reinterpret_cast<uintptr_t>(A.U) & 15) == 0);
            for (int kb = 0; kb < nkb; ++kb) {
                uint32_t hi[32], lo[32];
                float x[32];
                if (vec && my_valid && kb * TC_KB + 32 <= P.d) {
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const float4 f = __ldg(reinterpret_cast<const float4 *>(urow + kb * TC_KB) + q);
                        x[4 * q + 0] = f.x; x[4 * q + 1] = f.y; x[4 * q + 2] = f.z; x[4 * q + 3] = f.w;
                    }
                } else {
#pragma unroll
                    for (int q = 0; q < 32; ++q) {
                        const int k = kb * TC_KB + q;
                        x[q] = (my_valid && k < P.d) ? __ldg(urow + k) : 0.0f;
                    }
                }
#pragma unroll
                for (int q = 0; q < 32; ++q) {
                    const uint32_t h = to_tf32(x[q]);
                    hi[q] = h;
                    lo[q] = to_tf32(x[q] - __uint_as_float(h));
                }
                tmem_st32(lane_addr + (uint32_t)(kb * TC_KB), hi);
                if (A.passes == 3) tmem_st32(lane_addr + (uint32_t)(nkb * TC_KB + kb * TC_KB), lo);
            }
            tmem_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(a_ready);
        }

        // COLLECT: fixed per-row threshold from the sampled pre-pass; survivors go to this thread's sub-list
        float thr = PINF;  // rows beyond n_rows collect nothing
        uint2 *wbase = nullptr, *wp = nullptr, *wend = nullptr;
        float v[TC_R];  // SAMPLE: largest group maxima so far, descending
#pragma unroll
        for (int q = 0; q < TC_R; ++q) v[q] = NINF;
        if (!sample && my_valid) {
            thr = __ldg(A.thr + my_row);
            wbase = A.cand + ((my_row * P.S + c) * 2 + wg) * (int64_t)A.sub_stride;
            wp = wbase;
            wend = wbase + A.cap;
        }
        const float QNAN = __int_as_float(0x7fffffff);  // masked score: fails every >=, ignored by fmaxf
        float4 *my_stage = stage_buf + tid;  // element q of my row: float (q & 3) of my_stage[(q >> 2) * TC_EPI_THREADS]

        for (int i = 0; i < n_tiles; ++i) {
            const int b = i & 1;
            const uint32_t u = (uint32_t)((i >> 1) & 1);
            const int col0 = (t0 + i * t_step) * TN + wg * 64;
            mbar_wait(tmem_full + b, u, A.err_flag, 6);
            tc_fence_after();
            mbar_wait(bm_full + b, u, A.err_flag, 7);
            const uint32_t *bm = bitmap + b * 4 * TM + (wg * 2) * TM + r;
            const uint32_t acc_addr = lane_addr + (uint32_t)(TC_ACC_COL + b * TN + wg * 64);
#pragma unroll 1
            for (int gq = 0; gq < ((A.dbg & 1) ? 0 : 2); ++gq) {
                uint32_t raw[32];
                __syncwarp();
                tmem_ld32(acc_addr + (uint32_t)(gq * 32), raw);
                const uint32_t mword = bm[gq * TM];
                float s[32];
                if (P.bias != nullptr) {
                    const float4 *b4 = reinterpret_cast<const float4 *>(P.bias + col0 + gq * 32);
                    float bv[32];
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const float4 x = __ldg(b4 + q);
                        bv[4 * q + 0] = x.x; bv[4 * q + 1] = x.y; bv[4 * q + 2] = x.z; bv[4 * q + 3] = x.w;
                    }
                    tmem_wait_ld();
#pragma unroll
                    for (int q = 0; q < 32; ++q) s[q] = __uint_as_float(raw[q]) + bv[q];
                } else {
                    tmem_wait_ld();
#pragma unroll
                    for (int q = 0; q < 32; ++q) s[q] = __uint_as_float(raw[q]);
                }
                if (sample && mword != 0u) {  // train items (and columns past the catalogue) never count
#pragma unroll
                    for (int q = 0; q < 32; ++q)
                        if ((mword >> q) & 1u) s[q] = QNAN;
                }
                if (sample) {
                    float m1[11];
#pragma unroll
                    for (int q = 0; q < 10; ++q) m1[q] = fmaxf(fmaxf(s[3 * q], s[3 * q + 1]), s[3 * q + 2]);
                    m1[10] = fmaxf(s[30], s[31]);
                    float mx = fmaxf(fmaxf(m1[0], m1[1]), m1[2]);
                    mx = fmaxf(mx, fmaxf(fmaxf(m1[3], m1[4]), m1[5]));
                    mx = fmaxf(mx, fmaxf(fmaxf(m1[6], m1[7]), m1[8]));
                    mx = fmaxf(mx, fmaxf(m1[9], m1[10]));
                    if (my_valid && mx > v[TC_R - 1]) sorted_insert(v, mx);
                } else if (!(A.dbg & 4)) {
                    // Detection costs two instructions per score on two different pipes and no
                    // predicates: d = s - T0 (FMA pipe), then a funnel shift (ALU pipe) collects the sign
                    // bit of d; bit q of `pass` ends up set iff s[q] >= T0 and item q is not masked.
                    uint32_t m0 = 0, m1 = 0, m2 = 0, m3 = 0;  // four independent chains of 8 for ILP
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        m0 = __funnelshift_l(__float_as_uint(s[q] - thr), m0, 1);
                        m1 = __funnelshift_l(__float_as_uint(s[q + 8] - thr), m1, 1);
                        m2 = __funnelshift_l(__float_as_uint(s[q + 16] - thr), m2, 1);
                        m3 = __funnelshift_l(__float_as_uint(s[q + 24] - thr), m3, 1);
                    }
                    // chain j holds items 8j..8j+7 with item 8j in bit 7: assemble so that item 0 lands in
                    // bit 31, then reverse
                    const uint32_t m = (m0 << 24) | (m1 << 16) | (m2 << 8) | m3;
                    uint32_t pass = ~__brev(m) & ~mword;
                    if (pass != 0u) {
                        // rare per lane: park my 32 scores in shared memory so they can be indexed,
                        // then append each survivor as (score bits, item) to my list in HBM
#pragma unroll
                        for (int q = 0; q < 8; ++q)
                            my_stage[q * TC_EPI_THREADS] = make_float4(s[4 * q], s[4 * q + 1], s[4 * q + 2], s[4 * q + 3]);
                        const uint32_t cb = (uint32_t)(col0 + gq * 32);
                        const float *row_f = reinterpret_cast<const float *>(my_stage);
                        do {
                            const int q = __ffs(pass) - 1;
                            pass &= pass - 1u;
                            const float sc = row_f[(q >> 2) * (TC_EPI_THREADS * 4) + (q & 3)];
                            if (wp < wend) *wp = make_uint2(__float_as_uint(sc), cb + (uint32_t)q);
                            ++wp;
                        } while (pass != 0u);
                    }
                }
            }
            // this thread is done with accumulator b and bitmap b
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(tmem_empty + b);
                mbar_arrive(bm_empty + b);
            }
        }

        if (my_valid) {
            if (sample) {
                float *dst = A.samp + (my_row * 2 + wg) * TC_R;
#pragma unroll
                for (int q = 0; q < TC_R; ++q) dst[q] = v[q];
            } else {
                A.cand_cnt[(my_row * P.S + c) * 2 + wg] = (uint32_t)(wp - wbase);  // > cap means overflow
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (role == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    }
}

// ---- sampled group maxima -> per-row threshold -------------------------------------------------------
// thr[row] = r-th largest of the row's two descending lists (one per column half) = r-th largest
// sampled group maximum.  With fewer than r finite entries it is -inf (everything is a candidate).
__global__ void k_sample_thr(const float *__restrict__ samp, int64_t n_rows, int r, float *__restrict__ thr)
{
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n_rows) return;
    const float *la = samp + row * (2 * TC_R), *lb = la + TC_R;
    const float NINF = -__int_as_float(0x7f800000);
    int ia = 0, ib = 0;
    float t = NINF;
    for (int q = 0; q < r; ++q) {
        const float xa = (ia < TC_R) ? la[ia] : NINF;
        const float xb = (ib < TC_R) ? lb[ib] : NINF;
        if (xa >= xb) { t = xa; ++ia; } else { t = xb; ++ib; }
    }
    thr[row] = t;
}

// ---- operand preparation ---------------------------------------------------------------------------
// item table -> hi/lo TF32 tables [n, d_pad] (zero padded in k), one thread per output element
__global__ void k_split_tf32(const float *__restrict__ X, int64_t ld, int64_t n, int d, int d_pad,
                             float *__restrict__ hi, float *__restrict__ lo)
{
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * d_pad) return;
    const int64_t r = idx / d_pad;
    const int k = (int)(idx - r * d_pad);
    const float x = (k < d) ? X[r * ld + k] : 0.0f;
    const uint32_t h = to_tf32(x);
    hi[idx] = __uint_as_float(h);
    lo[idx] = __uint_as_float(to_tf32(x - __uint_as_float(h)));
}

__global__ void k_pad_bias(const float *__restrict__ bias, int n, int n_pad, float *__restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_pad) out[i] = (i < n) ? bias[i] : 0.0f;
}

}  // namespace skr
