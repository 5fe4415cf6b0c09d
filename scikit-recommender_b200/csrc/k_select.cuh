// k_select.cuh -- candidate lists of the fused tensor-core pass -> sorted top-K rank keys.
//
// A row arrives with n ~ 3-5 K candidates (score bits, item) spread over its sub-lists.  Sorting all
// of them (round-1 version: a 256-512 element warp bitonic per row, 4 000 instructions per row and the
// second most expensive kernel of the step) is wasted work: only the K best are wanted.  One warp per
// row now
//   1. gathers the sub-lists into shared memory as (ord(score), item),
//   2. finds a cut T with K <= #{ord >= T} <= 32 PER by a most-significant-digit radix search over
//      the span [min, max] of the row's own values (8 bits per pass through a 256-bin histogram; the
//      search stops as soon as the bin holding the K-th value is small enough to be taken whole --
//      after one pass for almost every row),
//   3. compacts the survivors into 64-bit rank keys and sorts just those (warp bitonic, PER = 2 or 4).
//   4. with the sorted keys still in registers (element e of lane l is rank e*32 + l) runs the metric
//      recurrences of k_metrics.cuh on them and adds the row into the warp's float64 column sums -- the
//      rank keys of a settled row never travel to HBM unless the caller asked for the top-K lists.
// RESCORE (precision "tf32r"): the candidates' scores came from ONE TF32 pass and carry an error of at most
// eps per score (bound derived in k_fused_tc.cuh at k_sample_thr).  The cut then only says where to look:
// with >= K candidates scoring >= cut in TF32, the true K-th best exact score is >= cut - eps, so every
// member of the true top-K has a TF32 score >= cut - 2 eps.  All candidates down to that bound (complete as
// long as the bound is not below the collection threshold, else the row is re-done exactly) are re-scored
// with the FP32 FMA chain of the exact kernels (k ascending, then the bias), and only those exact scores
// decide membership, order, ties and the reported values: results equal precision="fp32" bit for bit.
// Rows that cannot be settled here -- a sub-list overflowed, fewer than K candidates (threshold
// estimate too high or fewer than K unmasked items), more candidates than the buffer, or more than
// 32 PER values tied at the cut -- go on the fail list and are re-done exactly by k_row_exact.
#pragma once
#include "common.cuh"
#include "k_metrics.cuh"

namespace skr {

constexpr int SEL_WARPS = 4;
// candidates a row may carry into the selection: ~4 K are expected (r / f of the sampling plan), so the
// buffer follows the sort capacity (K <= 64: 512, K <= 128: 1024)
__host__ __device__ constexpr int sel_max(int per, bool rescore) { return (rescore ? per / 2 : per) * 256; }
constexpr int SEL_MAX_D = 128;  // RESCORE keeps the user vector in shared memory (the tensor-core path has d <= 128)

struct RescoreArgs {
    const float *U; int64_t ld_u;   // user vectors of the call's rows
    const float *V; int64_t ld_v;   // item vectors
    int d;
    const float *bias;              // or null
    const float *thr_c;             // [n_rows] threshold the main pass collected with
    const float *eps2;              // [n_rows] 2 eps of the row
    // out: exact rank keys of the re-scored survivors, unsorted, and how many (0 = the row is on the fail list);
    // sorting and metrics run in k_sort_metrics -- one kernel doing both needs 128 registers and runs at 25 %
    // occupancy on a latency-bound job (measured 260 us at c2 against ~150 us for the pair)
    u64 *rs_keys;                   // [n_rows, 32 PER]
    int *rs_cnt;                    // [n_rows]
};

struct SelOut {
    u64 *out_keys; int32_t *topk_idx; float *topk_val; float *per_user;
    const int64_t *te_indptr; const int32_t *te_idx; const double *disc; const float *idcg;
};

// skey[0 .. 32 PER): rank keys of the survivors (0 = empty) -> sorted; emits the K best and their metrics.
// cut / vmax: smallest / largest ord(score) among them.
template <int PER>
__device__ __forceinline__ void sel_sort_emit(const u64 *skey, uint32_t cut, uint32_t vmax, int lane, int K, int64_t row, int64_t row0,
                                              const MetricIds &mids, const SelOut &O, double *acc)
{
    // Sort.  Fast path: the survivors' scores span less than 2^(32 - BITS) float steps above the cut, so
    // (steps above the cut + 1) << BITS | slot is a 32-bit key with the same order as long as no two
    // survivors have equal scores; the 64-bit rank keys are fetched back by slot afterwards.  Rows with
    // a wider span or with tied scores take the 64-bit network (item id decides ties).
    constexpr int CAP = 32 * PER;
    constexpr int BITS = (PER <= 2) ? 6 : (PER <= 4 ? 7 : 8);
    static_assert((1 << BITS) >= CAP, "slot bits");
    u64 v[PER];
    bool fast = (vmax - cut) < ((1u << (32 - BITS)) - 2u);
    if (fast) {
        uint32_t v32[PER];
#pragma unroll
        for (int e = 0; e < PER; ++e) {
            const int i = e * 32 + lane;
            const u64 k = skey[i];
            v32[e] = (k != 0ull) ? ((((uint32_t)(k >> 32) - cut + 1u) << BITS) | (uint32_t)i) : 0u;
        }
        warp_bitonic_desc32<PER>(v32, lane);
        bool tie = false;
#pragma unroll
        for (int e = 0; e < PER; ++e) {
            const uint32_t mine = v32[e] >> BITS;
            uint32_t next = __shfl_down_sync(0xffffffffu, mine, 1);
            const uint32_t wrap = (e + 1 < PER) ? __shfl_sync(0xffffffffu, v32[(e + 1 < PER) ? e + 1 : e] >> BITS, 0) : 0u;
            if (lane == 31) next = wrap;
            tie |= (mine != 0u) && (mine == next);
            v[e] = (v32[e] != 0u) ? skey[v32[e] & (uint32_t)(CAP - 1)] : 0ull;
        }
        fast = !__any_sync(0xffffffffu, tie);
    }
    if (!fast) {
#pragma unroll
        for (int e = 0; e < PER; ++e) v[e] = skey[e * 32 + lane];
        warp_bitonic_desc<PER>(v, lane);
    }

// outputs: top-K lists on request, metrics straight from the registers
    const bool do_metrics = O.te_indptr != nullptr;  // null: keys only (per-shard lists of item-sharded evaluation)
    RowMetrics rm;
    if (do_metrics) rm.begin(O.te_indptr, O.te_idx, row0 + row);
#pragma unroll
    for (int e = 0; e < PER; ++e) {
        const int i = e * 32 + lane;
        if (e * 32 < K) {
            if (i < K) {
                if (O.out_keys != nullptr) O.out_keys[row * (int64_t)K + i] = v[e];
                if (O.topk_idx != nullptr) O.topk_idx[row * (int64_t)K + i] = (int32_t)key_item(v[e]);
                if (O.topk_val != nullptr) O.topk_val[row * (int64_t)K + i] = key_score(v[e]);
            }
            if (do_metrics)
                rm.chunk(e * 32, lane, K, (int32_t)key_item(v[e]), mids, O.disc, O.idcg,
                         O.per_user != nullptr ? O.per_user + row * (int64_t)(mids.n * K) : nullptr, acc);
        }
    }
}

// HITS (with RESCORE, when the caller wants metrics but no top-K lists): the metrics depend on the rank list only through
// the positions of the row's test items.  A test item h among the survivors has rank 1 + #{survivors j: s_j > s_h}; against
// a survivor whose approximate score differs from h's by more than 2 eps the comparison of the approximate scores is
// already the comparison of the exact ones, so only the test items found among the survivors and the survivors within
// 2 eps of one of them are re-scored exactly (c4: ~25 item rows per test item in the top 100 instead of ~115 per user;
// a user without one among the survivors needs none), everything else is counted.  Items that did not survive cannot
// outrank a test item that ends up in the top K (they lie below the K-th exact score).  Ranks come from one pass over
// the survivors' keys -- exact where re-scored, approximate elsewhere -- per chunk of 32 survivors holding a test item:
// no sort.  Per-user metric vectors and sums are bit-identical to the full re-scoring path.
template <int PER, bool RESCORE, bool HITS = false>
__global__ void __launch_bounds__(SEL_WARPS * 32, 8)
k_select_cands(const uint2 *__restrict__ cand, const uint32_t *__restrict__ cand_cnt, int n_sub, int cap, int sub_stride, int K,
               int64_t n_rows, int64_t row0, u64 *__restrict__ out_keys, int32_t *__restrict__ fail_list, int *__restrict__ fail_count,
               const int64_t *__restrict__ te_indptr, const int32_t *__restrict__ te_idx, MetricIds mids, const double *__restrict__ disc,
               const float *__restrict__ idcg, float *__restrict__ per_user, int32_t *__restrict__ topk_idx_out,
               float *__restrict__ topk_val_out, double *__restrict__ acc_out, RescoreArgs R, const float *__restrict__ add_back,
               int retry_only, const int *__restrict__ retry_total, int retry_min)
{
    pdl_wait();
    pdl_trigger();
    // add_back (nullable): the lists hold margins score - add_back[row] (single-pass main kernel with the threshold MMA)
    constexpr int CAP = 32 * PER;
    constexpr int CAPS = RESCORE ? CAP / 2 : CAP;  // size at which the search for the cut stops
    constexpr int SEL_MAX = sel_max(PER, RESCORE);
    __shared__ __align__(16) float s_u[RESCORE ? SEL_WARPS : 1][RESCORE ? SEL_MAX_D : 4];
    __shared__ uint2 s_ent[SEL_WARPS][SEL_MAX];
    __shared__ uint32_t s_hist[SEL_WARPS][256];
    __shared__ int s_off[SEL_WARPS][33];
    __shared__ u64 s_key[SEL_WARPS][CAP];
    extern __shared__ double sel_acc[];  // [SEL_WARPS][M*K] when acc_out != null
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint2 *ent = s_ent[warp];
    uint32_t *hist = s_hist[warp];
    u64 *skey = s_key[warp];
    const uint32_t lt_mask = (1u << lane) - 1u;
    const int MK = mids.n * K;
    double *acc = (acc_out != nullptr) ? sel_acc + (size_t)warp * MK : nullptr;
    if (acc != nullptr)
        for (int c = lane; c < MK; c += 32) acc[c] = 0.0;

    const int64_t n_warps = (int64_t)gridDim.x * SEL_WARPS;
    for (int64_t row = (int64_t)blockIdx.x * SEL_WARPS + warp; row < n_rows; row += n_warps) {
        __syncwarp();
        // second attempt (RESCORE): only the rows the first one left unsettled (rs_cnt == 0) are looked at
        if (RESCORE && retry_only) {
            if (R.rs_cnt[row] != 0) continue;
            if (*retry_total < retry_min) {  // no retry pass ran (too few rows to pay for a tile sweep): straight to the exact kernel
                if (lane == 0) fail_list[atomicAdd(fail_count, 1)] = (int32_t)row;
                continue;
            }
        }
        // ---- 1. sub-list sizes (n_sub <= 32: one per lane), exclusive scan, gather -----------------
        const int c_mine = (lane < n_sub) ? (int)__ldg(cand_cnt + row * n_sub + lane) : 0;
        int incl = c_mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        const int n = __shfl_sync(0xffffffffu, incl, 31);
        const bool over = __any_sync(0xffffffffu, c_mine > cap);
        if (over || n > SEL_MAX || n < K) {
            if (lane == 0) {
                fail_list[atomicAdd(fail_count, 1)] = (int32_t)row;
                if (RESCORE) R.rs_cnt[row] = 0;
            }
            continue;
        }
        // flat gather: candidate j of the row lives in sub-list s(j) = last s with off[s] <= j (binary search over the
        // <= 32 offsets in shared memory); every lane has useful work in every step, unlike a loop over sub-lists
        // whose ~17 entries leave half a warp idle (measured: the old gather was a third of the kernel's instructions)
        s_off[warp][lane] = (lane < n_sub) ? incl - c_mine : 0x7fffffff;
        __syncwarp();
        const int *off = s_off[warp];
        uint32_t vmin = 0xffffffffu, vmax = 0u;
        const uint2 *row_src = cand + row * n_sub * (int64_t)sub_stride;
        const float t_add = (add_back != nullptr) ? __ldg(add_back + row) : 0.0f;
        for (int j0 = lane; j0 < n; j0 += 128) {  // four independent loads in flight per lane
            uint2 e[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int j = j0 + 32 * q;
                int s = 0;
#pragma unroll
                for (int step = 16; step > 0; step >>= 1)
                    if (off[s + step] <= j) s += step;
                e[q] = (j < n) ? row_src[(int64_t)s * sub_stride + (j - off[s])] : make_uint2(0u, 0u);
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int j = j0 + 32 * q;
                if (j < n) {
                    e[q].x = ord_f32((add_back != nullptr) ? __uint_as_float(e[q].x) + t_add : __uint_as_float(e[q].x));
                    vmin = min(vmin, e[q].x);
                    vmax = max(vmax, e[q].x);
                    ent[j] = e[q];
                }
            }
        }
        vmin = __reduce_min_sync(0xffffffffu, vmin);
        vmax = __reduce_max_sync(0xffffffffu, vmax);
        __syncwarp();

        // ---- 2. cut: smallest-known T (as offset from vmin) with K <= #{w >= T} <= CAP --------------
        uint32_t T = 0;  // n <= CAP: everything is sorted
        bool ok = true;
        if (RESCORE || n > CAP) {  // RESCORE always needs a cut with >= K candidates above it
            const uint32_t range = vmax - vmin;
            int width_bits = 32 - __clz(range | 1u);  // values w = ord - vmin lie in [0, 2^width_bits)
            uint32_t base = 0;                         // current bucket: [base, base + 2^width_bits)
            int above = 0;                             // values at or beyond the bucket's end (all selected)
            ok = false;
            for (;;) {
                const int shift = width_bits > 8 ? width_bits - 8 : 0;
#pragma unroll
                for (int q = 0; q < 8; ++q) hist[q * 32 + lane] = 0u;
                __syncwarp();
                for (int i = lane; i < n; i += 32) {
                    const uint32_t w = ent[i].x - vmin;
                    const uint32_t rel = (w - base) >> shift;  // w < base wraps to a huge value
                    if (w >= base && rel < 256u) atomicAdd(&hist[rel], 1u);
                }
                __syncwarp();
                // lane l owns bins 8 l .. 8 l + 7; counts from the top bin down
                uint32_t h[8];
                int t_l = 0;
#pragma unroll
                for (int q = 0; q < 8; ++q) { h[q] = hist[lane * 8 + q]; t_l += (int)h[q]; }
                int suf = t_l;  // inclusive suffix sum over lanes >= mine
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int t = __shfl_down_sync(0xffffffffu, suf, o);
                    if (lane + o < 32) suf += t;
                }
                const int with_me = above + suf, without_me = with_me - t_l;
                const bool mine = (with_me >= K) && (without_me < K);
                const int owner = __ffs(__ballot_sync(0xffffffffu, mine)) - 1;  // exactly one lane (n >= K)
                int j = 0, c_above = without_me, c_with = without_me;
                if (mine) {
#pragma unroll
                    for (int q = 7; q >= 0; --q) {
                        if (c_with < K) { c_above = c_with; c_with += (int)h[q]; j = lane * 8 + q; }
                    }
                }
                j = __shfl_sync(0xffffffffu, j, owner);
                c_above = __shfl_sync(0xffffffffu, c_above, owner);
                c_with = __shfl_sync(0xffffffffu, c_with, owner);
                base += (uint32_t)j << shift;
                if (c_with <= CAPS) { ok = true; break; }
                if (shift == 0) break;  // more than CAP values tied around the K-th: exact path
                above = c_above;
                width_bits = shift;
                __syncwarp();
            }
            T = base;
        }
        if (!ok) {
            if (lane == 0) {
                fail_list[atomicAdd(fail_count, 1)] = (int32_t)row;
                if (RESCORE) R.rs_cnt[row] = 0;
            }
            continue;
        }

        // ---- 3. compact the survivors into rank keys, sort ------------------------------------------
        uint32_t cut = vmin + T;
        if (RESCORE && ok) {
            // every candidate down to cut - 2 eps must be looked at; complete only if the main pass collected that far
            const float lb = __fsub_rd(unord_f32(cut), __ldg(R.eps2 + row));
            ok = lb > __ldg(R.thr_c + row);  // strict: a margin that rounds to -0/+0 in the threshold MMA may go either way
            const uint32_t ol = ord_f32(lb);
            T = (ol > vmin) ? ol - vmin : 0u;
            if (!ok) {
                if (lane == 0) {
                    fail_list[atomicAdd(fail_count, 1)] = (int32_t)row;
                    R.rs_cnt[row] = 0;
                }
                continue;
            }
        }
#pragma unroll
        for (int e = 0; e < PER; ++e) skey[e * 32 + lane] = 0ull;
        __syncwarp();
        int m = 0;
        for (int i0 = 0; i0 < n; i0 += 32) {
            const int i = i0 + lane;
            uint2 e = make_uint2(0u, 0u);
            bool take = false;
            if (i < n) {
                e = ent[i];
                take = (e.x - vmin) >= T;
            }
            const uint32_t bal = __ballot_sync(0xffffffffu, take);
            const int pos = m + __popc(bal & lt_mask);
            if (take && pos < CAP) skey[pos] = ((u64)e.x << 32) | (u64)(~e.y);
            m += __popc(bal);
        }
        if (RESCORE) {
            if (m > CAP) {  // too many candidates inside the error band
                if (lane == 0) {
                    fail_list[atomicAdd(fail_count, 1)] = (int32_t)row;
                    R.rs_cnt[row] = 0;
                }
                continue;
            }
            // exact scores of the survivors: the FMA chain of RowDot::get4 (k_scores.cuh), operation for operation.
            // (Staging the item rows through shared memory with coalesced 128-byte segments was tried and measured
            // 7 % slower than letting every lane walk its own row: the gather is latency, not L1-tag, bound.)
            float *u = s_u[warp];
            const bool vec_ok = ((R.ld_v & 3) == 0) && ((reinterpret_cast<uintptr_t>(R.V) & 15) == 0);
            uint32_t omin = 0xffffffffu, omax = 0u;
            // HITS: the slots to re-score are listed in `hist` (free after the cut search); else every survivor
            int n_rs = m;
            RowMetrics rm;
            uint32_t hb[PER];  // HITS: ballot of "survivor e * 32 + lane is a test item"
            if (HITS) {
                __syncwarp();
                rm.begin(te_indptr, te_idx, row0 + row);
                int n_hit = 0;
#pragma unroll
                for (int e = 0; e < PER; ++e) {
                    const int i = e * 32 + lane;
                    const bool is = (e * 32 < m) && i < m && sorted_contains(rm.truth, rm.nt, (int32_t)key_item(skey[i]));
                    hb[e] = (e * 32 < m) ? __ballot_sync(0xffffffffu, is) : 0u;
                    n_hit += __popc(hb[e]);
                }
                if (n_hit == 0) {  // no test item can be in this user's top K: every metric of the row is 0
                    if (per_user != nullptr) {
                        float *pr = per_user + row * (int64_t)MK;
                        for (int c = lane; c < MK; c += 32) pr[c] = 0.0f;
                    }
                    if (lane == 0) R.rs_cnt[row] = m;
                    continue;
                }
                // survivors within 2 eps of a test item's approximate score (the test items themselves included)
                const float e2 = __ldg(R.eps2 + row);
                uint32_t need = 0u;  // bit e: my survivor e * 32 + lane gets an exact score
#pragma unroll
                for (int e2i = 0; e2i < PER; ++e2i) {
                    for (uint32_t mb = hb[e2i]; mb != 0u; mb &= mb - 1u) {  // warp-uniform: the row's test items among the survivors
                        const float ah = key_score(skey[e2i * 32 + (__ffs(mb) - 1)]);
                        const float lo = __fsub_rd(ah, e2), hi = __fadd_ru(ah, e2);
#pragma unroll
                        for (int e = 0; e < PER; ++e) {
                            const int i = e * 32 + lane;
                            if (i < m) {
                                const float a = key_score(skey[i]);
                                if (a >= lo && a <= hi) need |= 1u << e;
                            }
                        }
                    }
                }
                n_rs = 0;
#pragma unroll
                for (int e = 0; e < PER; ++e) {
                    const bool f = (need >> e) & 1u;
                    const uint32_t bal = __ballot_sync(0xffffffffu, f);
                    if (f) hist[n_rs + __popc(bal & lt_mask)] = (uint32_t)(e * 32 + lane);
                    n_rs += __popc(bal);
                }
            }
            for (int k = lane; k < R.d; k += 32) u[k] = __ldg(R.U + row * R.ld_u + k);
            __syncwarp();
            for (int i0 = 0; i0 < n_rs; i0 += 32) {
                // one candidate per lane; its row is fetched 8 float4 at a time, all loads issued before the first FMA
                // (with the loads inside the FMA loop every 4 k-values waited a full L2 round trip: 16 trips per row)
                const bool va = i0 + lane < n_rs;
                const int ia = HITS ? (va ? (int)hist[i0 + lane] : 0) : i0 + lane;
                const uint32_t item_a = va ? ~(uint32_t)skey[ia] : 0u;
                const float *pa = R.V + (int64_t)item_a * R.ld_v;
                float a = 0.0f;
                int k = 0;
                if (vec_ok) {
                    for (; k + 32 <= R.d; k += 32) {
                        float4 x[8];
#pragma unroll
                        for (int q = 0; q < 8; ++q) x[q] = __ldg(reinterpret_cast<const float4 *>(pa + k) + q);
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            const float4 w = *reinterpret_cast<const float4 *>(u + k + 4 * q);
                            a = fmaf(w.x, x[q].x, a);
                            a = fmaf(w.y, x[q].y, a);
                            a = fmaf(w.z, x[q].z, a);
                            a = fmaf(w.w, x[q].w, a);
                        }
                    }
                    for (; k + 4 <= R.d; k += 4) {
                        const float4 x = __ldg(reinterpret_cast<const float4 *>(pa + k));
                        const float4 w = *reinterpret_cast<const float4 *>(u + k);
                        a = fmaf(w.x, x.x, a);
                        a = fmaf(w.y, x.y, a);
                        a = fmaf(w.z, x.z, a);
                        a = fmaf(w.w, x.w, a);
                    }
                }
                for (; k < R.d; ++k) a = fmaf(u[k], __ldg(pa + k), a);
                if (R.bias != nullptr) a += __ldg(R.bias + item_a);
                if (va) {
                    const u64 key = make_key(a, item_a);
                    skey[ia] = key;
                    omin = min(omin, (uint32_t)(key >> 32));
                    omax = max(omax, (uint32_t)(key >> 32));
                }
            }
            __syncwarp();
            if (HITS) {
                // rank of every test item among the survivors = number of keys above its own (exact against exact where both
                // were re-scored, else the approximate one decides: they differ by more than 2 eps) -> bit mask of hit positions
                uint32_t *pos = reinterpret_cast<uint32_t *>(s_off[warp]);
                if (lane < 8) pos[lane] = 0u;
                __syncwarp();
#pragma unroll
                for (int e = 0; e < PER; ++e) {
                    if (hb[e] == 0u) continue;  // warp-uniform
                    if ((hb[e] >> lane) & 1u) {
                        const u64 kh = skey[e * 32 + lane];
                        int above = 0;
                        for (int j = 0; j < m; ++j) above += (skey[j] > kh) ? 1 : 0;
                        if (above < K) atomicOr(&pos[above >> 5], 1u << (above & 31));
                    }
                    __syncwarp();
                }
                float *pr = (per_user != nullptr) ? per_user + row * (int64_t)MK : nullptr;
                for (int i0 = 0; i0 < K; i0 += 32)
                    rm.chunk_hit(i0, lane, K, (pos[i0 >> 5] >> lane) & 1u, mids, disc, idcg, pr, acc);
                if (lane == 0) R.rs_cnt[row] = m;
                continue;
            }
            // hand the exact keys to k_sort_metrics
            u64 *dst = R.rs_keys + row * (int64_t)CAP;
            for (int i = lane; i < m; i += 32) dst[i] = skey[i];
            if (lane == 0) R.rs_cnt[row] = m;
            continue;
        }
        __syncwarp();
        // ---- 4. sort, outputs, metrics -----------------------------------------------------------------------
        const SelOut O = {out_keys, topk_idx_out, topk_val_out, per_user, te_indptr, te_idx, disc, idcg};
        sel_sort_emit<PER>(skey, cut, vmax, lane, K, row, row0, mids, O, acc);
    }
    if (acc_out != nullptr) fold_block_sums(sel_acc, SEL_WARPS, MK, acc_out + (size_t)blockIdx.x * MK);
}

// ---- second half of the re-scoring path: exact keys of the survivors -> sorted top-K -> metrics -----------------
// keys [n_rows, 32 PER] unsorted with cnt[row] valid entries (0: the row is on the fail list and is handled by
// k_row_exact + k_metrics).  One warp per row, rows dealt round-robin.
template <int PER>
__global__ void __launch_bounds__(SEL_WARPS * 32, 8)
k_sort_metrics(const u64 *__restrict__ rs_keys, const int *__restrict__ rs_cnt, int K, int64_t n_rows, int64_t row0, u64 *__restrict__ out_keys,
               const int64_t *__restrict__ te_indptr, const int32_t *__restrict__ te_idx, MetricIds mids, const double *__restrict__ disc,
               const float *__restrict__ idcg, float *__restrict__ per_user, int32_t *__restrict__ topk_idx_out,
               float *__restrict__ topk_val_out, double *__restrict__ acc_out)
{
    pdl_wait();
    pdl_trigger();
    constexpr int CAP = 32 * PER;
    __shared__ u64 s_key[SEL_WARPS][CAP];
    extern __shared__ double sel_acc[];  // [SEL_WARPS][M*K] when acc_out != null
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    u64 *skey = s_key[warp];
    const int MK = mids.n * K;
    double *acc = (acc_out != nullptr) ? sel_acc + (size_t)warp * MK : nullptr;
    if (acc != nullptr)
        for (int c = lane; c < MK; c += 32) acc[c] = 0.0;
    const SelOut O = {out_keys, topk_idx_out, topk_val_out, per_user, te_indptr, te_idx, disc, idcg};
    const int64_t n_warps = (int64_t)gridDim.x * SEL_WARPS;
    for (int64_t row = (int64_t)blockIdx.x * SEL_WARPS + warp; row < n_rows; row += n_warps) {
        const int m = __ldg(rs_cnt + row);
        if (m == 0) continue;
        __syncwarp();
        uint32_t omin = 0xffffffffu, omax = 0u;
#pragma unroll
        for (int e = 0; e < PER; ++e) {
            const int i = e * 32 + lane;
            const u64 k = (i < m) ? rs_keys[row * (int64_t)CAP + i] : 0ull;
            skey[i] = k;
            if (i < m) {
                omin = min(omin, (uint32_t)(k >> 32));
                omax = max(omax, (uint32_t)(k >> 32));
            }
        }
        omin = __reduce_min_sync(0xffffffffu, omin);
        omax = __reduce_max_sync(0xffffffffu, omax);
        __syncwarp();
        if (PER >= 4 && m <= 16 * PER && K <= 16 * PER)  // the survivors usually fit the smaller network
            sel_sort_emit<(PER >= 4 ? PER / 2 : PER)>(skey, omin, omax, lane, K, row, row0, mids, O, acc);
        else
            sel_sort_emit<PER>(skey, omin, omax, lane, K, row, row0, mids, O, acc);
    }
    if (acc_out != nullptr) fold_block_sums(sel_acc, SEL_WARPS, MK, acc_out + (size_t)blockIdx.x * MK);
}

}  // namespace skr
