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
// Selection.  The kernel is one pipeline instantiated in two modes:
//   SAMPLE  - a strided ~f = 6/K fraction of the item tiles is scored once in 1xTF32; each epilogue
//             thread keeps the TC_R largest 32-column group maxima (train items masked) of its row and
//             column quarter in registers.  The r-th largest of the row's 4 TC_R values is a threshold
//             T0 that, with overwhelming probability, has between K and a few K catalogue items above
//             it.  It only steers work: results never depend on it.
//   COLLECT - every item tile is scored in 3xTF32; two instructions per score and no predicates
//             (d = s - T0 on the FMA pipe, a funnel shift collecting the sign bits on the ALU pipe)
//             give a 32-bit survivor mask per thread; survivors that are not train items are appended
//             (score bits, item) to the thread's own candidate sub-list in HBM.
// k_select_cands (k_select.cuh) then picks and sorts each row's K best out of its ~3-5 K candidates;
// a row whose lists are short (< K) or overflowed is re-done exactly by k_row_exact (k_scores.cuh).
//
// Operands: A (users) lives in TMEM for the whole work item -- epilogue threads load their user's
// row from global memory, split it in registers and tcgen05.st hi/lo into TMEM lanes (TS-mode MMA; no
// shared memory for A).  B (items) is pre-split by k_split_tf32 into hi/lo tables and streamed by TMA
// (SWIZZLE_128B, 128 rows x 32 floats per box) through an mbarrier ring of k-block stages.
//
// Round-1 measurements that shaped this version (c2, ablation switches A.dbg): the bare pipeline
// skeleton cost 1270 cycles per tile (a mask builder with exposed global-load latency and four barrier
// round trips per tile), one thread issued a tcgen05.mma only every ~78 cycles (run-time loop bounds and
// div/mod by the stage count in the issue loop) and 8 epilogue warps exposed every TMEM-load and
// barrier latency.  Hence: k-blocks and passes are template parameters (the issue loop is straight-line
// code), 16 epilogue warps each own one 32-column group of the tile (thread <-> TMEM lane <-> user row;
// four warps per SM sub-partition hide each other's latencies), the accumulator-ready and mask-ready
// signals share one barrier per buffer, and the mask builders keep tile offsets in registers (one
// coalesced load per 32 tiles) and fetch keys four of their tiles ahead.
//
// Later measurements (single-pass modes, c2): the epilogue was bound by shared-memory store traffic, not by
// instruction issue -- parking all 32 scores of every lane that has a survivor moved ~450 store wavefronts per
// tile; predicating the stores per 4-score group cut the kernel from 0.355 to 0.304 ms.  One bitmap builder
// (~600-750 cycles per tile, an L2 round trip for its keys with a one-tile lookahead) paced the SAMPLE pass and the
// MMA-only ablation; two builders on alternate tiles with a deeper prefetch removed that.  Subtracting the
// threshold inside the tensor core (PRESUB below) takes 32 FADDs per thread and tile out of the epilogue.
//
// Warp roles (672 threads): warps 0-15 epilogue (lane quarter = warp & 3, column quarter = warp >> 2),
// warp 16 TMA producer, warp 17 TMEM allocator + MMA issuer of the even tiles, warp 18 train-mask bitmap
// builder of the even tiles, warp 19 MMA issuer of the odd tiles, warp 20 bitmap builder of the odd tiles.
// Accumulators and bitmaps are multi-buffered so the epilogue of tile n overlaps the MMAs of tile n+1.
// (21 warps put six on one SM sub-partition: 80 registers per thread instead of 96.)
//
// TMEM columns: [0, 32*NKB) A_hi, [32*NKB, 64*NKB) A_lo (3xTF32 only), 16 columns for the threshold operand
// (single-pass COLLECT), then NBUF accumulators of 128 columns at the top: three when A fits 128 columns
// (d <= 64 in three passes, d <= 96 in one; SAMPLE up to d = 128), otherwise two.  The third
// buffer lets the issuers run a tile further ahead of the epilogue.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include "fused_common.cuh"

namespace skr {

constexpr int TC_EPI_WARPS = 16;
constexpr int TC_EPI_THREADS = TC_EPI_WARPS * 32;
constexpr int TC_THREADS = TC_EPI_THREADS + 160;      // + 5 helper warps
constexpr int TC_KB = 32;                             // floats per k-block (one 128-byte swizzle row)
constexpr int TC_TILE_BYTES = TN * TC_KB * 4;         // 16 KB: one operand tile of one k-block
constexpr int TC_RING_BYTES = 8 * TC_TILE_BYTES;      // 4 stages of hi+lo (3xTF32) or 8 stages of hi (1xTF32)
constexpr int TC_MAX_STAGES = 8;
constexpr int TC_MAX_BUF = 3;                         // accumulator / bitmap buffers
// bias tiles live one slot longer than accumulators: the epilogue hands an accumulator back BEFORE it works on the scores
// (and reads the bias), so the builder of tile i + 4, which has waited for every epilogue warp to take tile i + 4 - NBUF >= i + 1
// (a warp takes tile i + 1 only after it is done with tile i), may overwrite the bias of tile i
constexpr int TC_BIAS_RING = 4;
constexpr int TC_R = 16;                              // group maxima kept per row and column quarter (SAMPLE)
constexpr int TC_MAX_RANK = 32;                       // largest threshold rank the 4 x TC_R lists support
constexpr long long TC_TIMEOUT_CYCLES = 4000000000ll; // watchdog: ~2 s

enum { TC_MODE_COLLECT = 0, TC_MODE_SAMPLE = 1 };

__host__ __device__ inline size_t tc_smem_bytes()
{
    return (size_t)1024                          // alignment slack
           + (size_t)TC_RING_BYTES
           + (size_t)TC_MAX_BUF * 4 * TM * 4     // train-mask bitmaps, one per accumulator buffer
           + (size_t)8 * TC_EPI_THREADS * 16     // score staging: one 32-float row per epilogue thread, [8][512] float4
           + (size_t)TC_TILE_BYTES               // constant B tile of the threshold MMA (single-pass COLLECT)
           + (size_t)TC_BIAS_RING * TN * 4       // bias of the last tiles (staged by the mask builders)
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
// one lane of a converged warp (always the same one for a full mask: tcgen05.commit must come from the
// thread that issued the MMAs it tracks)
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
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
// the same with FP16 operands (kind::f16: K = 16 per instruction, twice the TF32 rate), FP32 accumulation
__device__ __forceinline__ void tc_mma_ts_f16(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
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

// load + wait in one statement: the registers are defined only once the data has landed
__device__ __forceinline__ void tmem_ld32_wait(uint32_t taddr, uint32_t (&r)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : SKR_R32(r, 0), SKR_R32(r, 8), SKR_R32(r, 16), SKR_R32(r, 24)
        : "r"(taddr)
        : "memory");
}
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
// K-major SWIZZLE_32B (layout 6): rows are 32 bytes (one K = 16 fp16 slice), 8-row atoms 256 bytes apart -- the 4 KB
// augmentation tiles of the FP16 pipeline (cute: Swizzle<1,4,3> o ((8,n),2):((2,SBO),1) in 16-byte units)
__device__ __forceinline__ uint64_t make_b_desc_sw32(uint32_t smem_addr)
{
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3fffu);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(256 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)6 << 61;
    return d;
}
// cute::UMMA::InstrDescriptor: c_format F32 (1) [4,6) | a_format TF32 (2) [7,10) | b_format TF32 (2)
// [10,13) | a/b K-major | N>>3 [17,23) | M>>4 [24,29)
constexpr uint32_t TC_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
// a_format / b_format F16 (0)
constexpr uint32_t TC_IDESC_F16 = (1u << 4) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);

// ---- FP16 operands (precision "f16r") -------------------------------------------------------------------------------
// fp16 keeps 11 significant bits like TF32 but only 5 exponent bits, so both tables are scaled by powers of two (exact)
// before the conversion: the item table by one factor s_i for the whole catalogue, every user row by its own s_u, each
// chosen so that the largest magnitude lands in [2^9, 2^10).  Elements down to 2^-23 of the largest keep a relative error
// of 2^-11 (fp16 normals reach down to 2^-14); smaller ones are off by at most 2^-25 in scaled units, which the error
// band accounts for (k_sample_thr).  Scaled scores stay below d 2^20 <= 2^27.  The threshold operand of the threshold
// MMA is (hi + lo) x C with C = 2^12 in the B' tile (constant, or the augmentation tile), hi and lo fp16 values of
// magnitude < 2^15.
constexpr int TC_F16_TARGET = 9;     // floor(log2(scaled maximum))
constexpr int TC_F16_THR_SHIFT = 12;  // log2 C
__host__ __device__ inline int f16_scale_exp(float amax)
{
    // exponent e with amax 2^e in [2^9, 2^10); 0 for an all-zero, infinite or NaN maximum (garbage in: the rows fail over to the exact kernel)
    if (!(amax > 0.0f) || !(amax < 3.0e38f)) return 0;
#ifdef __CUDA_ARCH__
    const int ex = (int)((__float_as_uint(amax) >> 23) & 0xffu) - 127;
#else
    int ex;
    frexpf(amax, &ex);
    ex -= 1;
#endif
    int e = TC_F16_TARGET - ex;
    return e < -60 ? -60 : (e > 60 ? 60 : e);
}
__device__ __forceinline__ float exp2i(int e) { return __int_as_float((uint32_t)(e + 127) << 23); }  // 2^e, -126 <= e <= 127
__device__ __forceinline__ uint32_t pack_h2(float lo, float hi)  // (lo, hi) -> fp16 pair, lo in bits 0-15 (the lower k index)
{
    const __half2 h = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<const uint32_t *>(&h);
}

struct TcArgs {
    const float *U;   // user vectors [n_rows, ld_u]
    int64_t ld_u;
    int *err_flag;    // device int, set before a watchdog trap
    int dbg;          // timing experiments only (results invalid): 1 no epilogue work, 2 no MMA, 4 no appends, 8 no TMA, 16 no bitmaps
    // SAMPLE: item tiles 0, stride, 2*stride, ...; out: samp[row][4][TC_R] group maxima, descending
    int stride;
    int n_samp;
    int samp_chunks;  // CTAs per user tile in SAMPLE mode: each scores a contiguous share of the sample tiles (fills the GPU when
                      // there are fewer user tiles than SMs: a rank's share of a strong-scaled run) and writes its own 4 x TC_R maxima
    float *samp;
    // COLLECT: per-row thresholds (k_sample_thr) and candidate lists
    const float *thr;        // [n_rows]
    const float *thr_hi;     // [n_rows] single-pass COLLECT: thr = fl(hi + lo), hi and lo TF32 values (k_sample_thr)
    const float *thr_lo;
    // FP16 operands: s_i of the item table (k_split_f16); COLLECT: per-row s_u s_i from k_sample_thr, 0 = the row collects
    // nothing (no usable threshold); SAMPLE derives s_u from the row itself
    const float *item_scale;
    const float *scale;
    // FP16 operands with the bias folded into the contraction (AUG): an augmentation table (second tensor map) carries
    // (C, C, beta_hi, beta_lo, 0...) per item, beta = bias s_i 2^m, the user operand (-T_hi, -T_lo, g, g); bias_shift -> m
    // (k_split_f16), rowg: per-row g = s_u 2^-m from k_sample_thr (COLLECT; SAMPLE derives it from the row)
    const int *bias_shift;
    const float *rowg;
    int cap;                 // entries per (row, chunk, column quarter) sub-list
    uint2 *cand;             // [n_rows, S*4, cap] (score bits, item); PRESUB kernels store score - thr[row]
    uint32_t *cand_cnt;      // [n_rows, S*4] entries offered (> cap means overflow)
    // COLLECT work list: CTA b scores item tiles [t0, t0 + n) of user tile rt into sub-list slot `slot` of its rows;
    // built on the host with unequal chunk counts per user tile and ordered largest first, so that the hardware's
    // in-order block scheduler packs the SMs evenly (LPT) instead of leaving a quarter of them idle in the last wave
    const int4 *work;        // {rt, t0, n, slot}
    // second attempt of precision "tf32r" (see run_select_metrics): only rows whose retry_cnt[row] == 0 (not settled by
    // the first attempt) collect, with the threshold in `thr`; CTAs whose user tile has no such row exit at once
    const int *retry_cnt;    // [n_rows] or null
    const int *retry_total;  // number of unsettled rows; below retry_min the retry does not pay (see fused_chunk) and every CTA exits
    int retry_min;
    // development aid: per-tile clock64 timestamps of one CTA (null = off), [tile][TC_TRACE_SLOTS]
    long long *trace;
    int trace_cta;
    int trace_tiles;
};

constexpr int TC_TRACE_SLOTS = 16;
constexpr int TC_MASK_PF = 4;  // mask builders: key prefetch distance in own tiles (2 * TC_MASK_PF < 32)
// slots: 0 producer got the stage of kb 0, 1 producer issued the last TMA of the tile, 2 issuer: accumulator free,
// 3 issuer: first stage full, 4 issuer: tile committed, 5 mask: buffer free, 6 mask: bitmap ready,
// 7/10 epilogue warp 0/15: tile full, 8/11: accumulator in registers (released), 9/12: tile processed
#ifndef SKR_TC_TRACE
#define SKR_TC_TRACE 0
#endif
// ablation switches (TcArgs::dbg) are compiled in only for tools/dbg_timing.py and tools/trace_tiles.py
// (SKR_NVCC_EXTRA="-DSKR_TC_DBG=1"): the tests on them cost ~10 instructions per tile in the epilogue loop
#ifndef SKR_TC_DBG
#define SKR_TC_DBG 0
#endif
#define TC_DBG(word, bit) (SKR_TC_DBG && ((word) & (bit)))
__device__ __forceinline__ void tc_trace(const TcArgs &A, int tile, int slot)
{
    if (SKR_TC_TRACE && A.trace != nullptr && (int)blockIdx.x == A.trace_cta && tile < A.trace_tiles)
        A.trace[(size_t)tile * TC_TRACE_SLOTS + slot] = clock64();
}

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16])
{
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%16], "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15};"
        ::SKR_W32(r, 0), SKR_W32(r, 8), "r"(taddr)
        : "memory");
}

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

// One thread's share of a tile: 32 scores of its row.  SAMPLE: fold the group maximum into the sorted
// list v.  COLLECT: survivor mask against the row threshold, survivors appended to the sub-list.
// HALF (FP16 operands): the accumulator is in scaled units, inv = 1 / (s_u s_i) brings it back (a power of two: exact);
// with a bias that is the multiplier of the FMA that adds the bias, without one only survivors / sampled maxima are
// rescaled.  dead: all ones for a row that must not collect (no usable threshold), else 0.
template <bool SAMPLE, bool BIAS, bool PRESUB, bool HALF>
__device__ __forceinline__ void tc_process(const uint32_t (&raw)[32], const float *__restrict__ bias32, uint32_t mword, float thr, int col0,
                                           bool my_valid, int cap, int dbg, float4 *my_stage, uint2 *wbase, int &wn, float (&v)[TC_R],
                                           float inv, uint32_t dead)
{
    float s[32];
    if (BIAS) {
        const float4 *b4 = reinterpret_cast<const float4 *>(bias32);  // shared memory: this tile's bias, my 32 columns
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const float4 x = b4[q];
            if (HALF) {
                s[4 * q + 0] = fmaf(__uint_as_float(raw[4 * q + 0]), inv, x.x);
                s[4 * q + 1] = fmaf(__uint_as_float(raw[4 * q + 1]), inv, x.y);
                s[4 * q + 2] = fmaf(__uint_as_float(raw[4 * q + 2]), inv, x.z);
                s[4 * q + 3] = fmaf(__uint_as_float(raw[4 * q + 3]), inv, x.w);
            } else {
                s[4 * q + 0] = __uint_as_float(raw[4 * q + 0]) + x.x;
                s[4 * q + 1] = __uint_as_float(raw[4 * q + 1]) + x.y;
                s[4 * q + 2] = __uint_as_float(raw[4 * q + 2]) + x.z;
                s[4 * q + 3] = __uint_as_float(raw[4 * q + 3]) + x.w;
            }
        }
    } else {
#pragma unroll
        for (int q = 0; q < 32; ++q) s[q] = __uint_as_float(raw[q]);
    }
    if (SAMPLE) {
        if (mword != 0u) {  // train items (and columns past the catalogue) never count
            const float QNAN = __int_as_float(0x7fffffff);  // ignored by fmaxf
#pragma unroll
            for (int q = 0; q < 32; ++q)
                if ((mword >> q) & 1u) s[q] = QNAN;
        }
        float m1[11];
#pragma unroll
        for (int q = 0; q < 10; ++q) m1[q] = fmaxf(fmaxf(s[3 * q], s[3 * q + 1]), s[3 * q + 2]);
        m1[10] = fmaxf(s[30], s[31]);
        float mx = fmaxf(fmaxf(m1[0], m1[1]), m1[2]);
        mx = fmaxf(mx, fmaxf(fmaxf(m1[3], m1[4]), m1[5]));
        mx = fmaxf(mx, fmaxf(fmaxf(m1[6], m1[7]), m1[8]));
        mx = fmaxf(mx, fmaxf(m1[9], m1[10]));
        if (my_valid && mx > v[TC_R - 1]) sorted_insert(v, mx);
    } else if (!TC_DBG(dbg, 4)) {
        // Detection costs two instructions per score on two different pipes and no predicates:
        // d = s - T0 (FMA pipe), then a funnel shift (ALU pipe) collects the sign bit of d; bit q of
        // `pass` ends up set iff s[q] >= T0 and item q is not masked.  PRESUB: the accumulator already
        // holds score - T0 (threshold MMA, see the kernel), so the subtraction is gone.
        uint32_t m0 = 0, m1 = 0, m2 = 0, m3 = 0;  // four independent chains of 8 for ILP
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            m0 = __funnelshift_l(__float_as_uint(PRESUB ? s[q] : s[q] - thr), m0, 1);
            m1 = __funnelshift_l(__float_as_uint(PRESUB ? s[q + 8] : s[q + 8] - thr), m1, 1);
            m2 = __funnelshift_l(__float_as_uint(PRESUB ? s[q + 16] : s[q + 16] - thr), m2, 1);
            m3 = __funnelshift_l(__float_as_uint(PRESUB ? s[q + 24] : s[q + 24] - thr), m3, 1);
        }
        // chain j holds items 8j..8j+7 with item 8j in bit 7: assemble so that item 0 lands in bit 31,
        // then reverse
        const uint32_t m = (m0 << 24) | (m1 << 16) | (m2 << 8) | m3;
        uint32_t pass = ~__brev(m) & ~mword & ~dead;  // one LOP3
        if (TC_DBG(dbg, 32)) pass &= (uint32_t)(wn >> 30);  // timing experiment: detection only
        if (pass != 0u) {
            // rare per lane: park my 32 scores in shared memory so they can be indexed, then append
            // each survivor as (score bits, item) to my list in HBM
            // only the groups of 4 scores that hold a survivor (usually one of the eight): the staging stores are
            // the epilogue's main shared-memory traffic, and a store whose lanes are all predicated off moves nothing
            // (measured: staging all 32 scores cost 9 % of the single-pass kernel at c2, 13 % at c3b)
#pragma unroll
            for (int g = 0; g < 8; ++g)
                if (pass & (0xfu << (4 * g)))
                    my_stage[g * TC_EPI_THREADS] = make_float4(s[4 * g], s[4 * g + 1], s[4 * g + 2], s[4 * g + 3]);
            const float *row_f = reinterpret_cast<const float *>(my_stage);
            do {
                const int q = 31 - __clz(pass);  // highest first: one FLO instead of BREV + FLO; the lists are unordered
                pass ^= 1u << q;
                // PRESUB: the sub-lists hold margins (score - T0); k_select_cands adds T0 back
                float sc = row_f[(q >> 2) * (TC_EPI_THREADS * 4) + (q & 3)];
                if (HALF && !BIAS) sc *= inv;
                if (wn < cap && !TC_DBG(dbg, 64)) wbase[wn] = make_uint2(__float_as_uint(sc), (uint32_t)(col0 + q));
                if (TC_DBG(dbg, 64)) wn += (int)(__float_as_uint(sc) >> 31);  // timing experiment: no global store
                ++wn;
            } while (pass != 0u);
        }
    }
}

// HALF: FP16 operands (tm_bhi describes the scaled fp16 item table; a k-block is 64 elements = the same 128-byte swizzle
// row, 32 TMEM columns of packed pairs and four K = 16 MMAs, so the pipeline below is unchanged), single pass only.
// AUG (with HALF, models with an item bias): the B' tile of the threshold MMA is no longer a constant but a 4 KB slice
// of an augmentation table streamed with the item tiles (tm_blo; SWIZZLE_32B rows of 16 fp16: C, C, beta_hi, beta_lo, 0...,
// beta = bias s_i 2^m), and the user side holds (-T_hi, -T_lo, g, g, 0...) with g = s_u 2^-m: ONE K = 16 MMA per tile both
// subtracts the row threshold and adds the item bias inside the tensor core, and the epilogue of a biased model becomes
// the unbiased one (no bias loads, no FMAs: ~40 of its ~155 instructions per tile).  (A first version carried the four
// columns in a whole extra 16 KB k-block: 48 KB per tile put the kernel on the L2 -> SM limit, 13 of ~14.5 TB/s.)
template <int NKB, int PASSES, int MODE, bool HALF = false, bool AUG = false>
__global__ void __launch_bounds__(TC_THREADS, 1)
k_fused_tc(const __grid_constant__ CUtensorMap tm_bhi, const __grid_constant__ CUtensorMap tm_blo, TcArgs A, FusedParams P)
{
    pdl_trigger();  // the wait comes after the prologue, which touches nothing an earlier kernel wrote
    constexpr bool SAMPLE = (MODE == TC_MODE_SAMPLE);
    constexpr int KB_ELEMS = HALF ? 2 * TC_KB : TC_KB;  // operand elements per k-block
    constexpr int NKB_REAL = NKB;                        // k-blocks of embedding dimensions
    constexpr int AUG_TILE_BYTES = TN * 32;              // 128 item rows x 16 fp16
    constexpr int AUG_SLOTS = 4;                         // in the 16 KB the constant threshold tile occupies otherwise
    static_assert(!HALF || PASSES == 1, "FP16 operands: single pass");
    static_assert(!AUG || HALF, "the augmentation tile rides on the FP16 pipeline");
    // AUG: the producer may run at most AUG_SLOTS tiles ahead of the MMAs (slot i & 3 is reloaded for tile i + 4 once the
    // stage of tile i's first k-block is free, i.e. after tile i's augmentation MMA, which is issued first)
    constexpr int STAGES = (PASSES == 3) ? 4 : ((AUG && NKB == 1) ? 4 : 8);
    constexpr int STAGE_BYTES = (PASSES == 3) ? 2 * TC_TILE_BYTES : TC_TILE_BYTES;
    static_assert(NKB >= 1 && NKB <= 4 && (PASSES == 1 || PASSES == 3), "tile shape");
    static_assert(STAGES * STAGE_BYTES <= TC_RING_BYTES, "ring size");
    static_assert(!AUG || STAGES / NKB <= AUG_SLOTS, "augmentation slots");
    // Single-pass COLLECT (precision tf32r / 1xtf32): one extra K = 8 MMA per tile starts the accumulator at -T0[row]
    // (A' = [-hi, -lo, 0 x 6] per row in TMEM, B' = [1, 1, 0 x 6] for every column, a constant shared-memory tile),
    // so the epilogue tests sign bits instead of subtracting: the single-pass modes are bound by the epilogue's
    // instruction issue, not by the tensor pipe.  Needs 16 more TMEM columns; at d = 128 (NKB = 4) that leaves two
    // accumulator buffers instead of three, which still measured 4 % faster than three buffers with the FADDs.
    constexpr bool PRESUB = !SAMPLE && PASSES == 1;
    constexpr int A_THR_COL = TC_KB * NKB * (PASSES == 3 ? 2 : 1);  // first column of A'
    constexpr int A_COLS = A_THR_COL + ((PRESUB || AUG) ? 16 : 0);
    constexpr int NBUF = (A_COLS <= 512 - 3 * TN) ? 3 : 2;
    constexpr int ACC_COL = 512 - NBUF * TN;  // first accumulator column

    extern __shared__ unsigned char tc_smem_raw[];
    // 1024-byte alignment for SWIZZLE_128B, computed on the shared-window address so that the compiler
    // keeps every pointer below in the shared address space (LDS/STS, not generic LD/ST)
    unsigned char *smem = tc_smem_raw + ((1024u - (smem_u32(tc_smem_raw) & 1023u)) & 1023u);
    unsigned char *b_tiles = smem;
    uint32_t *bitmap = reinterpret_cast<uint32_t *>(smem + TC_RING_BYTES);        // [NBUF][4][TM]
    float4 *stage_buf = reinterpret_cast<float4 *>(bitmap + TC_MAX_BUF * 4 * TM);  // [8][TC_EPI_THREADS]
    unsigned char *thr_tile = reinterpret_cast<unsigned char *>(stage_buf + 8 * TC_EPI_THREADS);  // [TN][128 B], SWIZZLE_128B, 1024-aligned
    float *bias_buf = reinterpret_cast<float *>(thr_tile + TC_TILE_BYTES);                        // [TC_BIAS_RING][TN]
    uint64_t *bars = reinterpret_cast<uint64_t *>(bias_buf + TC_BIAS_RING * TN);
    uint64_t *full = bars;                            // [TC_MAX_STAGES] TMA landed
    uint64_t *empty = bars + TC_MAX_STAGES;           // [TC_MAX_STAGES] MMAs that read the stage are done
    uint64_t *tile_full = bars + 2 * TC_MAX_STAGES;   // [NBUF] accumulator complete (MMA commit) + bitmap built (mask warp)
    uint64_t *tile_empty = tile_full + TC_MAX_BUF;    // [NBUF] all 16 epilogue warps are done with accumulator and bitmap
    uint64_t *a_ready = tile_empty + TC_MAX_BUF;      // [1] A operand is in TMEM
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(a_ready + 1);

    // Epilogue = warps 0-15, helpers = warps 16-19: the warp scheduler favours higher warp ids, and the
    // latency-critical single-thread roles (TMA producer, MMA issuers) must not queue behind the epilogue.
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int role = warp - TC_EPI_WARPS;  // 0 TMA producer, 1 MMA issuer (even tiles) + TMEM allocator, 2 mask builder (even), 3 MMA issuer (odd tiles), 4 mask builder (odd)
    // tiles of this work item: COLLECT t0 + i (work list), SAMPLE i * stride
    int4 wk;
    if (SAMPLE) {
        const int C = A.samp_chunks, rt_ = (int)blockIdx.x / C, c_ = (int)blockIdx.x % C;
        const int s0 = (int)((long)A.n_samp * c_ / C), s1 = (int)((long)A.n_samp * (c_ + 1) / C);
        wk = make_int4(rt_, s0 * A.stride, s1 - s0, c_);
    } else {
        wk = __ldg(A.work + blockIdx.x);
    }
    const int rt = wk.x, t0 = wk.y, n_tiles = wk.z, c = wk.w;
    const int t_step = SAMPLE ? A.stride : 1;
    const int64_t row_base = (int64_t)rt * TM;
    if (!SAMPLE && A.retry_cnt != nullptr) {  // kernel argument: the branch is uniform over the grid
        pdl_wait();                           // the flags are the previous kernel's output
        if (__ldg(A.retry_total) < A.retry_min) return;  // a handful of rows: the per-row exact kernel is cheaper than a tile sweep
        const bool mine = tid < TM && row_base + tid < P.n_rows && __ldg(A.retry_cnt + row_base + tid) == 0;
        if (!__syncthreads_or(mine)) return;  // nothing to redo in this user tile (the usual case)
    }

    if (tid == 0) {
        for (int s = 0; s < TC_MAX_STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        for (int b = 0; b < TC_MAX_BUF; ++b) {
            mbar_init(tile_full + b, 2);
            mbar_init(tile_empty + b, TC_EPI_WARPS);
        }
        mbar_init(a_ready, 4 * NKB_REAL + ((PRESUB || AUG) ? 4 : 0));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (PRESUB && !AUG) {
        // B' in the K-major SWIZZLE_128B layout the item tiles use: logical 16-byte chunk c of row r sits at
        // chunk position c ^ (r & 7); chunk 0 holds k = 0..3 = (1, 1, 0, 0), everything else is zero
        uint4 *t4 = reinterpret_cast<uint4 *>(thr_tile);
        for (int j = tid; j < TC_TILE_BYTES / 16; j += TC_THREADS) {
            const int r8 = (j >> 3) & 7, pos = j & 7;
            // TF32: k = 0, 1 are 1.0f; FP16: k = 0, 1 are C = 2^12 (0x6c00), one packed pair
            const uint4 one = HALF ? make_uint4(0x6c006c00u, 0u, 0u, 0u) : make_uint4(0x3f800000u, 0x3f800000u, 0u, 0u);
            t4[j] = ((pos ^ r8) == 0) ? one : make_uint4(0u, 0u, 0u, 0u);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> visible to the tensor core
    }
    if (role == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_bhi) : "memory");
        if (PASSES == 3) asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_blo) : "memory");
    }
    if (role == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_ptr)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    // barriers, TMEM, the constant threshold tile and the work item (host-written, constant) are set up: everything
    // below reads what the previous kernels of this evaluate produced (item tables, thresholds, bitmaps' keys)
    pdl_wait();

    if (role == 0) {
        // ===== TMA producer: item k-block tiles (hi, lo) through the stage ring ==================
        // The whole warp stays converged (every operand is warp-uniform and lives in uniform registers);
        // one elected lane issues the copies.
        int s = 0;
        uint32_t ph = 0;
        for (int i = 0; i < n_tiles; ++i) {
            const int t = t0 + i * t_step;
#pragma unroll
            for (int kb = 0; kb < NKB; ++kb) {
                mbar_wait(empty + s, ph ^ 1u, A.err_flag, 1);
                if (elect_one()) {
                    if (kb == 0) tc_trace(A, i, 0);
                    if (TC_DBG(A.dbg, 8)) {
                        mbar_arrive(full + s);
                    } else {
                        mbar_expect_tx(full + s, STAGE_BYTES + ((AUG && kb == 0) ? AUG_TILE_BYTES : 0));
                        unsigned char *dst = b_tiles + (size_t)s * STAGE_BYTES;
                        if (AUG && kb == 0) tma_load_2d(thr_tile + (i & (AUG_SLOTS - 1)) * AUG_TILE_BYTES, &tm_blo, 0, t * TN, full + s);
                        tma_load_2d(dst, &tm_bhi, kb * KB_ELEMS, t * TN, full + s);
                        if (PASSES == 3) tma_load_2d(dst + TC_TILE_BYTES, &tm_blo, kb * TC_KB, t * TN, full + s);
                    }
                    if (kb == NKB - 1) tc_trace(A, i, 1);
                }
                __syncwarp();
                if (++s == STAGES) { s = 0; ph ^= 1u; }
            }
        }
    } else if (role == 1 || role == 3) {
        // ===== MMA issuers: warp 17 takes the even tiles (accumulator 0), warp 19 the odd ones (accumulator 1).
        // Two issuers hide each other's per-tile barrier latencies; the tensor pipe executes in issue order.
        // Converged warp, one elected lane issues: inside an `if (lane == 0)` region the compiler wraps every
        // tcgen05.mma in an ELECT / 3 x R2UR / branch loop (~100 cycles per MMA, measured) because it cannot
        // prove the descriptors uniform; here they are uniform registers and an MMA is a single instruction.
        const int p = (role == 1) ? 0 : 1;
        const bool do_mma = !TC_DBG(A.dbg, 2);
        mbar_wait(a_ready, 0, A.err_flag, 2);
        tc_fence_after();
        const uint32_t a_hi0 = tmem_base;
        const uint32_t a_lo0 = tmem_base + (uint32_t)(NKB * TC_KB);
        const uint64_t desc0 = make_b_desc(smem_u32(b_tiles));
        const uint64_t desc_thr = make_b_desc(smem_u32(thr_tile));
        const uint64_t desc_aug = make_b_desc_sw32(smem_u32(thr_tile));
        for (int i = p; i < n_tiles; i += 2) {
            const int b = i % NBUF;
            const uint32_t d_tmem = tmem_base + (uint32_t)(ACC_COL + b * TN);
            mbar_wait(tile_empty + b, (uint32_t)(((i / NBUF) & 1) ^ 1), A.err_flag, 3);
            tc_fence_after();
            if (elect_one()) tc_trace(A, i, 2);
#pragma unroll
            for (int kb = 0; kb < NKB; ++kb) {
                const int it = i * NKB + kb;
                const int s = it & (STAGES - 1);
                const uint32_t ph = (uint32_t)((it / STAGES) & 1);
                mbar_wait(full + s, ph, A.err_flag, 4);
                tc_fence_after();
                const uint64_t ds = desc0 + (uint64_t)(s * (STAGE_BYTES >> 4));  // start-address field counts 16-byte units
                if (elect_one()) {
                    if (kb == 0) tc_trace(A, i, 3);
                    if (do_mma) {
                        if ((PRESUB || AUG) && kb == 0) {  // acc = -T0[row] (AUG: + bias[col])
                            if (AUG) tc_mma_ts_f16(d_tmem, tmem_base + (uint32_t)A_THR_COL, desc_aug + (uint64_t)((i & (AUG_SLOTS - 1)) * (AUG_TILE_BYTES >> 4)), TC_IDESC_F16, 0u);
                            else if (HALF) tc_mma_ts_f16(d_tmem, tmem_base + (uint32_t)A_THR_COL, desc_thr, TC_IDESC_F16, 0u);
                            else tc_mma_ts(d_tmem, tmem_base + (uint32_t)A_THR_COL, desc_thr, TC_IDESC, 0u);
                        }
#pragma unroll
                        for (int k8 = 0; k8 < 4; ++k8) {  // UMMA K = 8 tf32 / 16 fp16 = 32 bytes of B, 8 TMEM columns of A
                            const uint32_t acol = (uint32_t)(kb * TC_KB + k8 * 8);
                            const uint64_t dhi = ds + (uint64_t)(k8 * 2);
                            const uint32_t acc = (PRESUB || AUG || (kb | k8)) ? 1u : 0u;
                            if (PASSES == 3) {
                                const uint64_t dlo = dhi + (uint64_t)(TC_TILE_BYTES >> 4);
                                tc_mma_ts(d_tmem, a_lo0 + acol, dhi, TC_IDESC, acc);
                                tc_mma_ts(d_tmem, a_hi0 + acol, dlo, TC_IDESC, 1u);
                                tc_mma_ts(d_tmem, a_hi0 + acol, dhi, TC_IDESC, 1u);
                            } else if (HALF) {
                                tc_mma_ts_f16(d_tmem, a_hi0 + acol, dhi, TC_IDESC_F16, acc);
                            } else {
                                tc_mma_ts(d_tmem, a_hi0 + acol, dhi, TC_IDESC, acc);
                            }
                        }
                    }
                    tc_commit(empty + s);  // stage reusable once these MMAs have read it
                    if (kb == NKB - 1) {
                        tc_commit(tile_full + b);  // accumulator b complete
                        tc_trace(A, i, 4);
                    }
                }
                __syncwarp();
            }
        }
    } else if (role == 2 || role == 4) {
        // ===== train-mask bitmap builder ===========================================================
        // Keys of this user tile are sorted by item; mask_tile_off gives, per item tile, where its keys
        // start.  Lane l holds the key range of tile (32-tile batch start + l): one coalesced load per
        // 32 tiles, then shuffles; the first 32 keys of tile i+1 are fetched while tile i is built.
        const uint32_t *keys = nullptr;
        const uint32_t *offs = nullptr;
        if (P.mask_keys != nullptr) {
            const int64_t rt_abs = (P.row0 / TM) + rt;
            keys = P.mask_keys + __ldg(P.mask_tile_ptr + rt_abs);
            offs = P.mask_tile_off + rt_abs * (int64_t)(P.n_ct + 1);
        }
        // Two builders take alternate tiles: one tile of bitmap work (zero, scatter, arrive: ~600 cycles with the
        // epilogue competing for issue slots) is longer than the single-pass MMA of a tile.
        // Lane l holds the key range of tile (32-tile batch + l) for the current batch (off_*) and the next (nxt_*).
        // The first 32 keys of a tile are fetched TC_MASK_PF of this warp's tiles ahead: one tile of bitmap work is
        // shorter than an L2 round trip, and with a one-tile lookahead this warp paced the whole pipeline.
        const int mp = (role == 2) ? 0 : 1;
        uint32_t off_b = 0, off_e = 0, nxt_b = 0, nxt_e = 0;
        auto load_offs = [&](int base, uint32_t &ob, uint32_t &oe) {
            const int ti = base + lane;
            ob = oe = 0;
            if (keys != nullptr && ti < n_tiles) {
                const int t = t0 + ti * t_step;
                ob = __ldg(offs + t);
                oe = __ldg(offs + t + 1);
            }
        };
        // first keys of tile x, x in the batch of tile i or the one after it
        auto fetch = [&](int x, int i) -> uint32_t {
            if (keys == nullptr || x >= n_tiles) return 0xffffffffu;
            const bool nx = (x >> 5) != (i >> 5);
            const uint32_t nb = __shfl_sync(0xffffffffu, nx ? nxt_b : off_b, x & 31);
            const uint32_t ne = __shfl_sync(0xffffffffu, nx ? nxt_e : off_e, x & 31);
            return (nb + lane < ne) ? __ldg(keys + nb + lane) : 0xffffffffu;
        };
        load_offs(0, off_b, off_e);
        load_offs(32, nxt_b, nxt_e);
        int batch = 0;
        uint32_t nk[TC_MASK_PF];
#pragma unroll
        for (int j = 0; j < TC_MASK_PF; ++j) nk[j] = fetch(mp + 2 * j, 0);
        // The tile's 128 bias values travel the same way: fetched TC_MASK_PF own tiles ahead (one float4 per lane), stored
        // next to the bitmap, published by the same arrive.  The epilogue then reads them with shared-memory latency (its
        // per-tile global loads were the long-scoreboard stalls of the c4 profile: 57 % of the kernel's samples).
        auto fetch_bias = [&](int x) -> float4 {
            if (P.bias == nullptr || x >= n_tiles) return make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            return __ldg(reinterpret_cast<const float4 *>(P.bias + (size_t)(t0 + x * t_step) * TN) + lane);
        };
        float4 nb4[TC_MASK_PF];
#pragma unroll
        for (int j = 0; j < TC_MASK_PF; ++j) nb4[j] = fetch_bias(mp + 2 * j);
        for (int i0 = mp; i0 < n_tiles; i0 += 2 * TC_MASK_PF) {
#pragma unroll
          for (int j = 0; j < TC_MASK_PF; ++j) {
            const int i = i0 + 2 * j;
            if (i >= n_tiles) break;
            const int b = i % NBUF;
            const int col0 = (t0 + i * t_step) * TN;
            if ((i >> 5) != batch) {
                batch = i >> 5;
                off_b = nxt_b;
                off_e = nxt_e;
                load_offs(batch * 32 + 32, nxt_b, nxt_e);
            }
            const uint32_t kb0 = __shfl_sync(0xffffffffu, off_b, i & 31), ke0 = __shfl_sync(0xffffffffu, off_e, i & 31);
            uint32_t key = nk[j];
            nk[j] = fetch(i + 2 * TC_MASK_PF, i);
            const float4 bias4 = nb4[j];
            nb4[j] = fetch_bias(i + 2 * TC_MASK_PF);
            mbar_wait(tile_empty + b, (uint32_t)(((i / NBUF) & 1) ^ 1), A.err_flag, 5);
            if (P.bias != nullptr) reinterpret_cast<float4 *>(bias_buf + (i & (TC_BIAS_RING - 1)) * TN)[lane] = bias4;
            if (lane == 0) tc_trace(A, i, 5);
            uint32_t *bm = bitmap + b * 4 * TM;
            if (TC_DBG(A.dbg, 16)) {  // timing ablation: no bitmap work at all
                if (lane == 0) mbar_arrive(tile_full + b);
                continue;
            }
            if (lane == 0) tc_trace(A, i, 13);
            if (col0 + TN <= P.n_items) {
                uint4 *bm4 = reinterpret_cast<uint4 *>(bm);
#pragma unroll
                for (int q = 0; q < 4; ++q) bm4[q * 32 + lane] = make_uint4(0u, 0u, 0u, 0u);
            } else {  // last tile: columns past the catalogue are masked
                for (int q = lane; q < 4 * TM; q += 32) bm[q] = oob_bits(col0, q / TM, P.n_items);
            }
            __syncwarp();
            if (lane == 0) tc_trace(A, i, 14);
            // 256 keys per step, their loads issued together: dense interaction data (ml-1m: ~580 train items per
            // 128-user x 128-item tile) otherwise pays one L2 round trip per 32 keys -- 7 us per tile, which made this warp
            // the pace of the whole kernel at c1 (77 us for a 10-tile sample pass)
            if (ke0 - kb0 <= 32u) {  // the usual case (c2: ~11 keys per tile): the prefetched keys are all there is
                if (kb0 + lane < ke0) {
                    const int cc = (int)(key >> 7) - col0;
                    atomicOr(&bm[(cc >> 5) * TM + (int)(key & 127u)], 1u << (cc & 31));
                }
            } else
            for (uint32_t p = kb0; p < ke0; p += 256) {
                uint32_t k4[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const uint32_t idx = p + 32u * q + lane;
                    k4[q] = (q == 0 && p == kb0) ? key : ((idx < ke0) ? __ldg(keys + idx) : 0xffffffffu);
                }
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    if (p + 32u * q + lane < ke0) {
                        const int cc = (int)(k4[q] >> 7) - col0;
                        atomicOr(&bm[(cc >> 5) * TM + (int)(k4[q] & 127u)], 1u << (cc & 31));
                    }
                }
            }
            __syncwarp();
            if (lane == 0) tc_trace(A, i, 15);
            if (lane == 0) { mbar_arrive(tile_full + b); tc_trace(A, i, 6); }
          }
        }
    } else {
        // ===== epilogue: thread <-> user row (TMEM lane), warp <-> 32 columns of every tile ========
        const int lq = warp & 3, cq = warp >> 2;
        const int r = lq * 32 + lane;  // TMEM lane
        const int64_t my_row = row_base + r;
        const bool my_valid = my_row < P.n_rows;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(lq * 32) << 16);
        const float NINF = -__int_as_float(0x7f800000);
        const float PINF = __int_as_float(0x7f800000);

        // FP16 operands: scale of my row.  su: what the row is multiplied by before the conversion; inv = 1 / (su s_i)
        // brings accumulator values back to true units; dead: the row has no usable threshold and collects nothing.
        float su = 1.0f, inv = 1.0f, g_aug = 0.0f;
        uint32_t dead = 0u;
        if (HALF) {
            const float si = __ldg(A.item_scale);
            if (SAMPLE) {
                float amax = 0.0f;  // fmaxf drops NaNs: a NaN element poisons its own products only
                if (my_valid) {
                    const float *urow = A.U + my_row * A.ld_u;
                    if (((A.ld_u & 3) == 0) && ((reinterpret_cast<uintptr_t>(A.U) & 15) == 0) && (P.d & 3) == 0) {
                        for (int k = 0; k < P.d; k += 4) {
                            const float4 f = __ldg(reinterpret_cast<const float4 *>(urow + k));
                            amax = fmaxf(fmaxf(amax, fmaxf(fabsf(f.x), fabsf(f.y))), fmaxf(fabsf(f.z), fabsf(f.w)));
                        }
                    } else {
                        for (int k = 0; k < P.d; ++k) amax = fmaxf(amax, fabsf(__ldg(urow + k)));
                    }
                }
                int eu = f16_scale_exp(amax);
                if (AUG) {  // g = s_u 2^-m must be an fp16 power of two: rows dwarfed by the bias scale less, giants drop it
                    const int m = __ldg(A.bias_shift);
                    eu = min(eu, m + 15);
                    g_aug = (eu - m < -24) ? 0.0f : exp2i(eu - m);
                }
                su = exp2i(eu);
                inv = 1.0f / (su * si);  // powers of two with exponents within +-60 each: exact
            } else {
                const float S = my_valid ? __ldg(A.scale + my_row) : 0.0f;
                if (S > 0.0f) {
                    su = S / si;
                    inv = 1.0f / S;
                    if (AUG) g_aug = __ldg(A.rowg + my_row);
                } else {
                    dead = 0xffffffffu;
                }
            }
        }
        if (HALF && cq < NKB_REAL) {  // A: k-block cq (64 elements) of my user's vector -> scaled fp16 pairs -> 32 TMEM columns
            const int kb = cq;
            const float *urow = A.U + (my_valid ? my_row : 0) * A.ld_u;
            const bool vec = ((A.ld_u & 3) == 0) && ((reinterpret_cast<uintptr_t>(A.U) & 15) == 0);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int k0 = kb * KB_ELEMS + h * 32;
                uint32_t pk[16];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    float4 f = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                    const int k = k0 + 4 * q;
                    if (my_valid) {
                        if (vec && k + 4 <= P.d) {
                            f = __ldg(reinterpret_cast<const float4 *>(urow + k));
                        } else {
                            if (k + 0 < P.d) f.x = __ldg(urow + k + 0);
                            if (k + 1 < P.d) f.y = __ldg(urow + k + 1);
                            if (k + 2 < P.d) f.z = __ldg(urow + k + 2);
                            if (k + 3 < P.d) f.w = __ldg(urow + k + 3);
                        }
                    }
                    pk[2 * q + 0] = pack_h2(f.x * su, f.y * su);
                    pk[2 * q + 1] = pack_h2(f.z * su, f.w * su);
                }
                tmem_st16(lane_addr + (uint32_t)(kb * TC_KB + h * 16), pk);
            }
            tmem_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(a_ready);
        }
        if (!HALF && cq < NKB) {  // A: k-block cq of my user's vector -> hi/lo TF32 -> TMEM
            const int kb = cq;
            const float *urow = A.U + (my_valid ? my_row : 0) * A.ld_u;
            const bool vec = ((A.ld_u & 3) == 0) && ((reinterpret_cast<uintptr_t>(A.U) & 15) == 0);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int k0 = kb * TC_KB + h * 16;
                float x[16];
                if (vec && my_valid && k0 + 16 <= P.d) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float4 f = __ldg(reinterpret_cast<const float4 *>(urow + k0) + q);
                        x[4 * q + 0] = f.x; x[4 * q + 1] = f.y; x[4 * q + 2] = f.z; x[4 * q + 3] = f.w;
                    }
                } else {
#pragma unroll
                    for (int q = 0; q < 16; ++q) x[q] = (my_valid && k0 + q < P.d) ? __ldg(urow + k0 + q) : 0.0f;
                }
                uint32_t hi[16], lo[16];
#pragma unroll
                for (int q = 0; q < 16; ++q) {
                    const uint32_t hh = to_tf32(x[q]);
                    hi[q] = hh;
                    lo[q] = to_tf32(x[q] - __uint_as_float(hh));
                }
                tmem_st16(lane_addr + (uint32_t)k0, hi);
                if (PASSES == 3) tmem_st16(lane_addr + (uint32_t)(NKB * TC_KB + k0), lo);
            }
            tmem_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(a_ready);
        }

        // COLLECT: fixed per-row threshold from the sampled pre-pass; survivors go to this thread's sub-list
        // rows beyond n_rows collect nothing (PRESUB: 2^126, a TF32 value; inf would make the threshold MMA produce NaN)
        float thr = PRESUB ? __int_as_float(0x7e800000) : PINF;
        float thr_hi = thr, thr_lo = 0.0f;
        uint2 *wbase = nullptr;
        int wn = 0;
        float v[TC_R];  // SAMPLE: largest group maxima so far, descending
#pragma unroll
        for (int q = 0; q < TC_R; ++q) v[q] = NINF;
        if (!SAMPLE && my_valid && (A.retry_cnt == nullptr || __ldg(A.retry_cnt + my_row) == 0)) {
            thr = __ldg(A.thr + my_row);
            if (PRESUB) {
                thr_hi = __ldg(A.thr_hi + my_row);
                thr_lo = __ldg(A.thr_lo + my_row);
            }
        }
        if (!SAMPLE && my_valid) wbase = A.cand + ((my_row * P.S + c) * 4 + cq) * (int64_t)A.cap;
        if ((PRESUB || AUG) && cq == 3) {  // A': -T0 as TF32 hi + lo (k_sample_thr), 16 columns reserved, 8 read
            uint32_t x[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) x[q] = 0u;
            if (HALF) {  // k_sample_thr wrote hi, lo as fp16 values of T0 su s_i / C; a dead row's operand is irrelevant
                x[0] = (SAMPLE || dead) ? 0u : pack_h2(-thr_hi, -thr_lo);
                if (AUG) x[1] = pack_h2(g_aug, g_aug);
            } else {
                x[0] = __float_as_uint(-thr_hi);
                x[1] = __float_as_uint(-thr_lo);
            }
            tmem_st16(lane_addr + (uint32_t)A_THR_COL, x);
            tmem_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(a_ready);
        }
        float4 *my_stage = stage_buf + tid;  // element q of my row: float (q & 3) of my_stage[(q >> 2) * TC_EPI_THREADS]
        const uint32_t *my_bm = bitmap + cq * TM + r;
        const uint32_t my_acc = lane_addr + (uint32_t)(ACC_COL + cq * 32);

        int b = 0;
        uint32_t bph = 0;  // buffer and its phase for tile i
        for (int i = 0; i < n_tiles; ++i) {
            const int col0 = (t0 + i * t_step) * TN + cq * 32;
            mbar_wait(tile_full + b, bph, A.err_flag, 6);
            tc_fence_after();
            const int tslot = (warp == 0) ? 7 : 10;
            const bool tr_me = (lane == 0) && (warp == 0 || warp == TC_EPI_WARPS - 1);
            if (tr_me) tc_trace(A, i, tslot);
            uint32_t raw[32];
            uint32_t mword = 0u;
            if (!TC_DBG(A.dbg, 1)) {
                mword = my_bm[b * 4 * TM];
                tmem_ld32_wait(my_acc + (uint32_t)(b * TN), raw);
            }
            // accumulator b and bitmap b are in registers: hand both back before working on them
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tile_empty + b);
            if (tr_me) tc_trace(A, i, tslot + 1);
            if (++b == NBUF) { b = 0; bph ^= 1u; }
            if (TC_DBG(A.dbg, 1)) continue;

            // two copies of the per-tile work, so that without a bias the scores are consumed in the very
            // registers tcgen05.ld wrote (one shared copy costs 32 register moves per tile)
            if (P.bias != nullptr)
                tc_process<SAMPLE, true, PRESUB, HALF>(raw, bias_buf + (i & (TC_BIAS_RING - 1)) * TN + cq * 32, mword, thr, col0, my_valid, A.cap, A.dbg, my_stage,
                                                       wbase, wn, v, inv, dead);
            else
                tc_process<SAMPLE, false, PRESUB, HALF>(raw, nullptr, mword, thr, col0, my_valid, A.cap, A.dbg, my_stage, wbase, wn, v, inv, dead);
            if (tr_me) tc_trace(A, i, tslot + 2);
        }

        if (my_valid) {
            if (SAMPLE) {
                float4 *dst = reinterpret_cast<float4 *>(A.samp + ((my_row * A.samp_chunks + c) * 4 + cq) * TC_R);
                if (HALF && P.bias == nullptr) {  // the maxima were kept in scaled units
#pragma unroll
                    for (int q = 0; q < TC_R; ++q) v[q] *= inv;
                }
#pragma unroll
                for (int q = 0; q < TC_R / 4; ++q) dst[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
            } else {
                A.cand_cnt[(my_row * P.S + c) * 4 + cq] = (uint32_t)wn;  // > cap means overflow
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
// thr[row] = r-th largest of the row's 4 x TC_R kept group maxima (one warp per row, two values per lane):
// the largest T with #{v >= T} >= r, built bit by bit over the monotone integer image of the floats.
// With fewer than r finite entries it is -inf (everything is a candidate).  A list that was truncated at
// TC_R can only lower the estimate, i.e. admit more candidates.
//
// precision "tf32r" (eps2_out != null): the main pass scores in ONE TF32 pass and k_select_cands re-scores
// the survivors exactly, so the threshold is lowered by 2 eps, eps bounding |s_tf32 - s_fp32| for every item
// of the row:
//   operands rounded to TF32 (rna): |hi(x) - x| <= 2^-11 |x|, so each product is off by <= (2^-10 + 2^-22) |u_k i_k|;
//   FP32 accumulation in the tensor core, order and rounding mode unspecified: <= d 2^-22 sum |u_k i_k| (a
//   very loose bound: every one of <= d additions may lose a full ulp of the running sum, twice);
//   the exact kernels' own FMA chain: <= d 2^-24 sum |u_k i_k|; the two bias additions: 2^-23 |score|;
//   sum |u_k i_k| <= ||u|| max_j ||i_j|| (Cauchy-Schwarz), |score| <= that + max |b|.
//   threshold MMA (single-pass kernel, PRESUB): the accumulator starts at -T0, |T0| <= ||u|| N_max + B_max, so the
//   partial sums are at most twice as large, and the epilogue's add-back rounds once more: covered by doubling
//   the accumulation term (and a few ulps);
// eps = 1.25 [(2^-10 + (2.5 d + 8) 2^-22) ||u|| N_max + 2^-22 B_max]: the 1.25 covers the FP32 evaluation of
// the norms and leaves slack (tests assert the observed error stays below eps / 2).  eps_coef carries the
// bracket's first factor, stats = {N_max^2, B_max} from k_split_tf32.
template <int NQ>  // sampling chunks per row (1..4): 64 NQ values per row, 2 NQ per lane
__global__ void __launch_bounds__(256)
k_sample_thr(const float *__restrict__ samp, int64_t n_rows, int r, float *__restrict__ thr, const float *__restrict__ U, int64_t ld_u,
             int d, const float *__restrict__ stats, float eps_coef, float *__restrict__ eps2_out, float *__restrict__ thr_hi_out,
             float *__restrict__ thr_lo_out, float eps3_coef, float *__restrict__ eps2_3_out, float *__restrict__ thr3_out,
             const float *__restrict__ item_scale, float *__restrict__ scale_out, const int *__restrict__ bias_shift, float *__restrict__ g_out)
{
    // bias_shift != null: the bias rides in the contraction (AUG, k_fused_tc).  beta = bias s_i 2^m is carried as two fp16
    // values (error <= 2^-22 |b|), g = s_u 2^-m is a power of two and exact; the accumulator's partial sums grow by the
    // bias, so the accumulation term of the band is taken on ||u|| N_max + B_max.  g must be an fp16 value: rows so small
    // that s_u 2^-m > 2^15 are scaled less (their scores are all bias anyway), rows so large that it would drop below
    // 2^-24 lose the bias in the candidate pass and get B_max added to their band.
    // item_scale != null: FP16 operands (precision "f16r").  The band gets one more term -- elements too small for an fp16
    // normal are off by up to 2^-25 in scaled units instead of 2^-11 relative: sum_k |du'_k i'_k| + |u'_k di'_k| <=
    // 2^-25 sqrt(d) (N'_max + ||u'||), i.e. 2^-25 sqrt(d) (N_max / s_u + ||u|| / s_i) in true units (fp16 round-to-nearest
    // has the same 2^-11 relative bound as TF32's rna, and products of two fp16 values are exact in FP32) -- and the
    // row's scale s_u is chosen here: largest magnitude to [2^9, 2^10), lowered if need be until the threshold operand
    // T0 s_u s_i / C fits fp16 (|.| < 2^15; only when |T0| exceeds ~256 max|u| max|i|, i.e. the bias dominates).  Rows
    // whose scale would fall more than 14 binades short, or without a usable threshold, are marked dead (scale 0):
    // they collect nothing and the exact kernel settles them.
    pdl_wait();
    pdl_trigger();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * 8 + warp;
    if (row >= n_rows) return;
    // 64 values per sampling chunk of the row: 2 per lane and chunk
    constexpr int n_vals = 64 * NQ;
    uint32_t v[2 * NQ];
    uint32_t vlo = 0xffffffffu, vhi = 0u;
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        const float2 x = __ldg(reinterpret_cast<const float2 *>(samp + row * n_vals + q * 64) + lane);
        v[2 * q] = ord_f32(x.x);
        v[2 * q + 1] = ord_f32(x.y);
        vlo = min(vlo, min(v[2 * q], v[2 * q + 1]));
        vhi = max(vhi, max(v[2 * q], v[2 * q + 1]));
    }
    const uint32_t lo = __reduce_min_sync(0xffffffffu, vlo), hi = __reduce_max_sync(0xffffffffu, vhi);
    // bits above the highest bit in which min and max differ are common to every value
    const int nb = 32 - __clz((lo ^ hi) | 1u);
    uint32_t T = (nb >= 32) ? 0u : (hi >> nb) << nb;
    for (int bit = nb - 1; bit >= 0; --bit) {
        const uint32_t cand = T | (1u << bit);
        int c = 0;
#pragma unroll
        for (int q = 0; q < 2 * NQ; ++q) c += (int)(v[q] >= cand);
        if (__reduce_add_sync(0xffffffffu, c) >= r) T = cand;
    }
    float t0 = unord_f32(T);
    int eu = 0, ei = 0;  // FP16 operands: exponents of s_u, s_i
    bool half_dead = false;
    if (eps2_out != nullptr) {
        float ss = 0.0f, amax = 0.0f;
        for (int k = lane; k < d; k += 32) {
            const float v = __ldg(U + row * ld_u + k);
            ss = fmaf(v, v, ss);
            amax = fmaxf(amax, fabsf(v));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            ss += __shfl_xor_sync(0xffffffffu, ss, o);
            amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
        }
        float eps = 1.25f * (eps_coef * sqrtf(ss) * sqrtf(__ldg(stats)) + 2.384185791015625e-07f * __ldg(stats + 1));
        if (item_scale != nullptr) {
            ei = (int)((__float_as_uint(__ldg(item_scale)) >> 23) & 0xffu) - 127;
            const int eu0 = f16_scale_exp(amax);
            eu = eu0;
            int m_aug = 0;
            if (bias_shift != nullptr) {
                m_aug = __ldg(bias_shift);
                eu = min(eu, m_aug + 15);
                eps += 1.25f * (((2.5f * (float)d + 8.0f) * 2.384185791015625e-07f + 4.76837158203125e-07f) * __ldg(stats + 1));
            }
            // |T0 - 2 eps| 2^(eu + ei - 12) < 2^14 (one binade of headroom for the terms added below)
            const float tmag = fabsf(t0) + 2.0f * eps;
            if (tmag < 3.0e38f && tmag > 0.0f) {
                const int et = (int)((__float_as_uint(tmag) >> 23) & 0xffu) - 127 + 1;  // tmag < 2^et
                eu = min(eu, 14 + TC_F16_THR_SHIFT - ei - et);
            }
            half_dead = !(tmag < 3.0e38f) || eu < eu0 - 14 || eu + ei < -100 || eu + ei > 100;
            if (!half_dead)
                eps += 1.25f * 2.98023223876953125e-08f * sqrtf((float)d) * (sqrtf(__ldg(stats)) * exp2i(-eu) + sqrtf(ss) * exp2i(-ei));
            if (bias_shift != nullptr) {
                float g = 0.0f;
                if (!half_dead) {
                    if (eu - m_aug < -24) eps += 1.25f * __ldg(stats + 1);  // the candidate pass scores this row without the bias
                    else g = exp2i(eu - m_aug);
                }
                if (lane == 0) g_out[row] = g;
            }
        }
        float e2 = 2.0f * eps;
        const float INF = __int_as_float(0x7f800000);
        if (!(e2 < INF)) e2 = INF;  // overflow / NaN operands: collect everything, the exact kernel settles the row
        if (thr3_out != nullptr) {
            // the retry of rows this band was too wide for scores in three passes (3xTF32): operands are then exact to
            // 2^-22 each and the dropped lo*lo term is 2^-22 |u_k i_k|, the accumulator takes 3 d terms:
            // eps3 = 1.25 [(3.25 d + 11) 2^-22 ||u|| N_max + 2^-22 B_max] -- 10 to 20 times narrower than eps
            float e3 = 2.0f * 1.25f * (eps3_coef * sqrtf(ss) * sqrtf(__ldg(stats)) + 2.384185791015625e-07f * __ldg(stats + 1));
            if (!(e3 < INF)) e3 = INF;
            if (lane == 0) {
                eps2_3_out[row] = e3;
                thr3_out[row] = (e3 < INF) ? __fsub_rd(t0, e3) : -INF;
            }
        }
        t0 = (e2 < INF) ? __fsub_rd(t0, e2) : -INF;
        if (lane == 0) eps2_out[row] = e2;
    }
    if (thr_hi_out != nullptr) {
        // The single-pass main kernel subtracts the threshold inside the tensor core (threshold MMA): it needs
        // T = hi + lo with TF32 hi, lo and T <= t0 (collect a superset).  hi = rna(t0); t0 - hi is exact; lo is that
        // remainder rounded towards -inf on the TF32 grid.  thr = fl(hi + lo) <= t0 is what the epilogue adds back and
        // what k_select_cands compares against.  Without a usable threshold (-inf: fewer than r finite samples;
        // NaN) the row collects nothing and k_select_cands hands it to the exact kernel.
        float hi = __int_as_float(0x7e800000), lo = 0.0f;  // 2^126
        if (item_scale != nullptr) {
            // FP16: T0 s_u s_i / C = hi + lo, fp16 values, lo rounded towards -inf; hi + lo spans at most 22 bits, so
            // thr = (hi + lo) C / (s_u s_i) is exact in FP32 -- the same construction as the TF32 one below
            float S = 0.0f;
            hi = 0.0f;
            if (!half_dead && fabsf(t0) < 1.0e37f) {
                const float t = t0 * exp2i(eu + ei - TC_F16_THR_SHIFT);
                if (fabsf(t) < 32000.0f) {
                    hi = __half2float(__float2half_rn(t));
                    lo = __half2float(__float2half_rd(t - hi));
                    S = exp2i(eu + ei);
                    t0 = (hi + lo) * exp2i(TC_F16_THR_SHIFT - eu - ei);
                }
            }
            if (S == 0.0f) { hi = 0.0f; lo = 0.0f; t0 = __int_as_float(0x7e800000); }
            if (lane == 0) {
                scale_out[row] = S;
                thr_hi_out[row] = hi;
                thr_lo_out[row] = lo;
                thr[row] = t0;
            }
            return;
        }
        if (fabsf(t0) < 1.0e37f) {
            hi = __uint_as_float(to_tf32(t0));
            const float rem = t0 - hi;
            uint32_t lb = __float_as_uint(rem);
            const bool inexact = (lb & 0x1fffu) != 0u;
            lb &= ~0x1fffu;
            if (rem < 0.0f && inexact) lb += 0x2000u;  // magnitude up = value down
            lo = __uint_as_float(lb);
        }
        t0 = hi + lo;
        if (lane == 0) {
            thr_hi_out[row] = hi;
            thr_lo_out[row] = lo;
        }
    }
    if (lane == 0) thr[row] = t0;
}

// ---- operand preparation ---------------------------------------------------------------------------
// item table -> hi/lo TF32 tables [n, d_pad] (zero padded in k), one warp per item row (lo == null: the single-pass
// kernels read only hi, the lo table is neither allocated nor written).  The same pass takes
// max_j ||item_j||^2 and max_j |bias_j| for the tf32r error band (non-negative floats order like their bit
// patterns: atomicMax on uint) into stats_cur, and resets the words later kernels count into: the fail-list
// length, and the statistics slot of the NEXT evaluate (the two slots alternate, so nothing races with the
// atomics of this launch).  A stale slot after a failed call can only enlarge the band.
__global__ void __launch_bounds__(256)
k_split_tf32(const float *__restrict__ X, int64_t ld, int64_t n, int d, int d_pad, float *__restrict__ hi, float *__restrict__ lo,
             const float *__restrict__ bias, int *__restrict__ zero_a, uint32_t *__restrict__ stats_cur, uint32_t *__restrict__ stats_next)
{
    pdl_wait();
    pdl_trigger();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        if (zero_a != nullptr) { zero_a[0] = 0; zero_a[1] = 0; }  // fail counters of the first attempt and of the retry
        if (stats_next != nullptr) { stats_next[0] = 0u; stats_next[1] = 0u; }
    }
    float best = 0.0f, bb = 0.0f;
    // four rows per warp and step, all loads issued before the first store (d_pad <= 128: at most 4 values per lane and row)
    for (int64_t row0 = ((int64_t)blockIdx.x * 8 + warp) * 4; row0 < n; row0 += (int64_t)gridDim.x * 32) {
        float x[4][4];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int k = lane + 32 * q;
                x[r][q] = (row0 + r < n && k < d) ? __ldg(X + (row0 + r) * ld + k) : 0.0f;
            }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            if (row0 + r >= n) break;
            float ss = 0.0f;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int k = lane + 32 * q;
                if (k < d_pad) {
                    const uint32_t h = to_tf32(x[r][q]);
                    hi[(row0 + r) * d_pad + k] = __uint_as_float(h);
                    if (lo != nullptr) lo[(row0 + r) * d_pad + k] = __uint_as_float(to_tf32(x[r][q] - __uint_as_float(h)));
                    ss = fmaf(x[r][q], x[r][q], ss);
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
            best = fmaxf(best, (ss == ss) ? ss : __int_as_float(0x7f800000));
            if (bias != nullptr && lane == 0) {
                const float v = fabsf(__ldg(bias + row0 + r));
                bb = fmaxf(bb, (v == v) ? v : __int_as_float(0x7f800000));
            }
        }
    }
    if (lane == 0 && stats_cur != nullptr) {
        atomicMax(stats_cur, __float_as_uint(best));
        if (bias != nullptr) atomicMax(stats_cur + 1, __float_as_uint(bb));
    }
}

// ---- FP16 operands: largest item magnitude, then the scaled fp16 item table ---------------------------------------------
// amax_bits: zeroed by the host before the launch; non-negative floats order like their bit patterns.
// f16s words: [0] max |item element| bits, [1] s_i (float), [2] max |bias| bits, [3] m (int): beta = bias s_i 2^m
__global__ void __launch_bounds__(256)
k_item_absmax(const float *__restrict__ X, int64_t ld, int64_t n, int d, const float *__restrict__ bias, uint32_t *__restrict__ f16s)
{
    pdl_wait();
    pdl_trigger();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float best = 0.0f, bb = 0.0f;
    for (int64_t row0 = ((int64_t)blockIdx.x * 8 + warp) * 4; row0 < n; row0 += (int64_t)gridDim.x * 32) {
#pragma unroll
        for (int r = 0; r < 4; ++r)
            if (row0 + r < n)
                for (int k = lane; k < d; k += 32) best = fmaxf(best, fabsf(__ldg(X + (row0 + r) * ld + k)));  // NaNs are skipped
        if (bias != nullptr && lane < 4 && row0 + lane < n) bb = fmaxf(bb, fabsf(__ldg(bias + row0 + lane)));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        best = fmaxf(best, __shfl_xor_sync(0xffffffffu, best, o));
        bb = fmaxf(bb, __shfl_xor_sync(0xffffffffu, bb, o));
    }
    if (lane == 0 && best > 0.0f) atomicMax(f16s, __float_as_uint(best));
    if (lane == 0 && bb > 0.0f) atomicMax(f16s + 2, __float_as_uint(bb));
}

// item table -> fp16 table [n, d_pad] (d_pad = 64 or 128, zero padded in k), every element multiplied by s_i = 2^e with
// e = f16_scale_exp(max |item element|) first (exact); s_i is published in *si_out for the kernels that follow.  Same
// side jobs as k_split_tf32: max ||item||^2 and max |bias| of the UNscaled table for the error band, counters reset.
__global__ void __launch_bounds__(256)
k_split_f16(const float *__restrict__ X, int64_t ld, int64_t n, int d, int d_pad, __half *__restrict__ out, const float *__restrict__ bias,
            int *__restrict__ zero_a, uint32_t *__restrict__ stats_cur, uint32_t *__restrict__ stats_next, uint32_t *__restrict__ f16s,
            __half *__restrict__ aug)
{
    // aug != null: augmentation table [n, 16] fp16, row = (C, C, beta_hi, beta_lo, 0 ...) with beta = bias s_i 2^m and m such that
    // the largest |beta| lies in [2^13, 2^14) (see k_fused_tc AUG); words [1] and [3] of f16s are written here, read by the
    // kernels that follow ([0], [2] come from k_item_absmax: every block reads them before any later kernel may start)
    pdl_wait();
    pdl_trigger();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ei = f16_scale_exp(__uint_as_float(__ldg(f16s)));
    const float si = exp2i(ei);
    int m = 0;
    {
        const float bmax = __uint_as_float(__ldg(f16s + 2));
        if (bmax > 0.0f && bmax < 3.0e38f) m = 13 - ((int)((__float_as_uint(bmax) >> 23) & 0xffu) - 127) - ei;
        m = max(-100 - min(ei, 0), min(100 - max(ei, 0), m));  // 2^(ei + m) stays a normal float
    }
    const float bscale = exp2i(ei + m);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        if (zero_a != nullptr) { zero_a[0] = 0; zero_a[1] = 0; }
        if (stats_next != nullptr) { stats_next[0] = 0u; stats_next[1] = 0u; }
        reinterpret_cast<float *>(f16s)[1] = si;
        reinterpret_cast<int *>(f16s)[3] = m;
    }
    const int d_emb = d_pad;
    float best = 0.0f, bb = 0.0f;
    // four rows per warp and step; lane l owns the element pairs (2 l, 2 l + 1) and (64 + 2 l, 65 + 2 l) of a row
    for (int64_t row0 = ((int64_t)blockIdx.x * 8 + warp) * 4; row0 < n; row0 += (int64_t)gridDim.x * 32) {
        float x[4][4];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int k = 2 * lane + 64 * (q >> 1) + (q & 1);
                x[r][q] = (row0 + r < n && k < d) ? __ldg(X + (row0 + r) * ld + k) : 0.0f;
            }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            if (row0 + r >= n) break;
            float ss = 0.0f;
#pragma unroll
            for (int q = 0; q < 4; q += 2) {
                const int k = 2 * lane + 64 * (q >> 1);
                if (k < d_emb) {
                    *reinterpret_cast<uint32_t *>(out + (row0 + r) * d_pad + k) = pack_h2(x[r][q] * si, x[r][q + 1] * si);
                    ss = fmaf(x[r][q], x[r][q], ss);
                    ss = fmaf(x[r][q + 1], x[r][q + 1], ss);
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
            best = fmaxf(best, (ss == ss) ? ss : __int_as_float(0x7f800000));
            float bv = 0.0f;
            if (bias != nullptr) {
                bv = __ldg(bias + row0 + r);
                const float v = fabsf(bv);
                if (lane == 0) bb = fmaxf(bb, (v == v) ? v : __int_as_float(0x7f800000));
            }
            if (aug != nullptr && lane < 8) {  // 32 bytes per row: lane 0 (C, C), lane 1 (beta_hi, beta_lo), zeros
                uint32_t w = 0u;
                if (lane == 0) w = 0x6c006c00u;  // C = 2^12 twice
                if (lane == 1) {
                    const float beta = bv * bscale;
                    const float bh = __half2float(__float2half_rn(beta));
                    w = pack_h2(bh, beta - bh);
                }
                *reinterpret_cast<uint32_t *>(aug + (row0 + r) * 16 + 2 * lane) = w;
            }
        }
    }
    if (lane == 0 && stats_cur != nullptr) {
        atomicMax(stats_cur, __float_as_uint(best));
        if (bias != nullptr) atomicMax(stats_cur + 1, __float_as_uint(bb));
    }
}

// hi AND lo TF32 tables for the three-pass retry of an "f16r" evaluate (the first attempt built neither), on demand
__global__ void __launch_bounds__(256)
k_split_hilo_if(const int *__restrict__ need, int need_min, const float *__restrict__ X, int64_t ld, int64_t n, int d, int d_pad,
                float *__restrict__ hi, float *__restrict__ lo)
{
    pdl_wait();
    pdl_trigger();
    if (*need < need_min) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int64_t row = (int64_t)blockIdx.x * 8 + warp; row < n; row += (int64_t)gridDim.x * 8) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int k = lane + 32 * q;
            if (k < d_pad) {
                const float x = (k < d) ? __ldg(X + row * ld + k) : 0.0f;
                const uint32_t h = to_tf32(x);
                hi[row * d_pad + k] = __uint_as_float(h);
                lo[row * d_pad + k] = __uint_as_float(to_tf32(x - __uint_as_float(h)));
            }
        }
    }
}

// The lo table of the three-pass retry, built only when the first attempt left rows unsettled (*need != 0): the
// common case costs one block-wide early exit per CTA.
__global__ void __launch_bounds__(256)
k_split_lo_if(const int *__restrict__ need, int need_min, const float *__restrict__ X, int64_t ld, int64_t n, int d, int d_pad, float *__restrict__ lo)
{
    pdl_wait();
    pdl_trigger();
    if (*need < need_min) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int64_t row = (int64_t)blockIdx.x * 8 + warp; row < n; row += (int64_t)gridDim.x * 8) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int k = lane + 32 * q;
            if (k < d_pad) {
                const float x = (k < d) ? __ldg(X + row * ld + k) : 0.0f;
                const uint32_t h = to_tf32(x);
                lo[row * d_pad + k] = __uint_as_float(to_tf32(x - __uint_as_float(h)));
            }
        }
    }
}

__global__ void k_pad_bias(const float *__restrict__ bias, int n, int n_pad, float *__restrict__ out)
{
    pdl_wait();
    pdl_trigger();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_pad) out[i] = (i < n) ? bias[i] : 0.0f;
}

}  // namespace skr
