/*
 * prophylo_b200/csrc/ppp_sweep.cu -- the sm_100a PPP sweep kernel (bit-plane targets).
 *
 * Replaces, for a tile of (query, target) pairs, the per-pair Perl loop
 *   PhyloProf::Algorithm::ppp::get_match_score     /root/reference/lib/PhyloProf/Algorithm/ppp.pm:130-245
 * under the canonical-order embedding of SURVEY.md section 8d: with target bits T and query planes P (present), Y (yes),
 * at genome column g with T[g] = 1:   depth = popc(T[0..g]),  number = popc((T&P)[0..g]),  yes = popc((T&Y)[0..g]),
 * and the reference's running `answer < score` minimum of bdtrc(yes-1, number, p) can only move at a column where
 * yes increments (the tail is strictly increasing in number at fixed yes), i.e. at the set bits of T & Y.
 *
 * One warp owns one target row (kept in registers) and walks a 32-query tile held in shared memory:
 *   1. AND + POPC per word, one packed warp prefix scan -> (yes, number) offsets per lane; depth offsets once per target.
 *   2. every set bit of T&Y becomes a candidate (yes = its rank, number, genome) in a per-warp shared-memory queue,
 *      so that the 32 lanes then evaluate candidates evenly regardless of where the bits fell.
 *   3. tier A (FP32 + 3 FP64 table reads): an ENCLOSURE [lo, hi] of ln P(X >= yes | number, p) from the leading
 *      binomial term (log-factorial table in shared memory) and 8 series terms with a geometric upper and a
 *      Bahadur lower bound on the remainder.  U = min hi;  only candidates with lo <= U can be the minimum.
 *   4. tier B (FP64, warp-cooperative): survivors get ln p to ~1e-11 by summing the ratio series, 32 terms a round.
 *   5. pairs whose outcome could hinge on Cephes' own rounding (near-ties, saturation at 1, underflow to 0, odd p)
 *      are queued for the exact tier (ppp_exact.cu), which replays the reference loop with a bit-exact bdtrc.
 * The query tile, its constants and the log-factorial table are staged with TMA bulk copies (cp.async.bulk +
 * mbarrier).  Tensor cores are not used: the sweep is a prefix recurrence, not a contraction.
 */
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include <algorithm>

#include "../../include/ppp_b200.h"
#include "ppp_common.cuh"

#include "ppp_pair.cuh"

/* ---------------------------------------------------------------- tile kernel: every pair of (query tile x targets)
 * One warp owns one target row (registers) and walks a 32-query tile staged in shared memory by TMA. */
template <int WPL>
__global__ void __launch_bounds__(512, PPP_MINBLOCKS_OF(WPL)) ppp_sweep_kernel(const SweepParams prm)
{
    constexpr uint32_t PPP_QCAP = PPP_QCAP_OF(WPL);
    constexpr int W = 32 * WPL;  /* words per row */
    constexpr int GD = 32 * W;   /* padded genome count */
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int nwarps = blockDim.x >> 5;

    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw);                /* [0]: lf table, [1]: query tile */
    double *s_lf = reinterpret_cast<double *>(smem_raw + 128);              /* GD + 2 doubles */
    constexpr int LF_BYTES = (((GD + 2) * 8) + 127) & ~127;
    uint32_t *s_q = reinterpret_cast<uint32_t *>(smem_raw + 128 + LF_BYTES);    /* PPP_QTILE * 2 * W words */
    QConst *s_qc = reinterpret_cast<QConst *>(reinterpret_cast<unsigned char *>(s_q) + PPP_QTILE * 2 * W * 4);
    uint32_t *s_off = reinterpret_cast<uint32_t *>(reinterpret_cast<unsigned char *>(s_qc) + PPP_QTILE * sizeof(QConst)); /* [32] */
    uint32_t *s_queue = s_off + PPP_QTILE;
    uint32_t *my_queue = s_queue + warp * (3 * PPP_QCAP);                   /* [QCAP] n | g<<16, [QCAP] yes, [QCAP] float lo */
    uint32_t *my_y = my_queue + PPP_QCAP;
    float *my_lo = reinterpret_cast<float *>(my_queue + 2 * PPP_QCAP);

    if (threadIdx.x == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        fence_barrier_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t bytes = (uint32_t)((GD + 2) * 8);
        mbar_expect_tx(&bars[0], bytes);
        tma_load_1d(s_lf, prm.lf, bytes, &bars[0]);
    }

    /* Work items = (query tile, super-block of PPP_SB targets).  CTAs draw items from a global counter and, inside an
     * item, warps draw targets from a shared one: a pair can cost 50x another (query |Y| 20 .. 1200, target density
     * 0.02 .. 0.9), so any static split leaves most of the chip waiting for the slowest part. */
    const uint32_t nTl = prm.t1 - prm.t0;                                  /* targets in this launch */
    const uint32_t nQtiles = (prm.nQ + PPP_QTILE - 1) / PPP_QTILE;
    const uint32_t nSB = (nTl + PPP_SB - 1) / PPP_SB;
    const uint64_t nItems = (uint64_t)nQtiles * nSB;
    uint32_t *s_next = reinterpret_cast<uint32_t *>(smem_raw + 64);        /* per-item target counter */
    unsigned long long *s_item = reinterpret_cast<unsigned long long *>(smem_raw + 72); /* item drawn by thread 0 */
    (void)nwarps;

    WarpCounters wc;
    uint32_t tile_parity = 0;
    uint32_t cur_qt = 0xffffffffu;
    mbar_wait(&bars[0], 0);

    for (;;) {
        __syncthreads(); /* everyone is done with the previous item (and tile) */
        if (threadIdx.x == 0) *s_item = atomicAdd(&prm.counters[CNT_ITEM], 1ull);
        __syncthreads();
        const uint64_t item = *s_item;
        if (item >= nItems) break;
        const uint32_t qt = (uint32_t)(item / nSB);
        const uint32_t sb = (uint32_t)(item % nSB);
        const uint32_t q0 = qt * PPP_QTILE;
        const uint32_t nq = min((uint32_t)PPP_QTILE, prm.nQ - q0);
        if (threadIdx.x == 0) *s_next = 0;
        if (qt != cur_qt) { /* block-uniform */
            if (threadIdx.x == 0) {
                fence_proxy_async();
                const uint32_t b1 = nq * 2 * W * 4, b2 = nq * (uint32_t)sizeof(QConst);
                mbar_expect_tx(&bars[1], b1 + b2);
                tma_load_1d(s_q, prm.Q + (size_t)q0 * 2 * W, b1, &bars[1]);
                tma_load_1d(s_qc, prm.qc + q0, b2, &bars[1]);
            }
            if (prm.append && threadIdx.x < PPP_QTILE)
                s_off[threadIdx.x] = (threadIdx.x < nq && prm.nmax_off) ? prm.nmax_off[q0 + threadIdx.x] : PPP_NO_TABLE;
            mbar_wait(&bars[1], tile_parity);
            tile_parity ^= 1;
            cur_qt = qt;
        }
        __syncthreads();
        const uint32_t sb_count = min((uint32_t)PPP_SB, nTl - sb * PPP_SB);
      for (;;) { /* targets of the item, drawn dynamically */
        uint32_t pick = 0;
        if (lane == 0) pick = atomicAdd(s_next, 1u);
        pick = __shfl_sync(FULL, pick, 0);
        if (pick >= sb_count) break;
        const uint32_t tl = sb * PPP_SB + pick;
        const uint32_t t = prm.t0 + tl;

        uint32_t tw[WPL];
        load_row<WPL>(prm.T + (size_t)t * W + lane * WPL, tw);
        uint32_t dlane = 0;
#pragma unroll
        for (int w = 0; w < WPL; ++w) dlane += __popc(tw[w]);
        uint32_t dpre, dtotal;
        depth_prefix(dlane, lane, dpre, dtotal);

        PairOut mine;
        mine.score = 1.0; mine.yes = 0; mine.number = 0; mine.depth = 0; mine.margin = INFINITY;
        uint32_t mine_flags = 0;
        for (uint32_t qi = 0; qi < nq; ++qi) {
            const uint32_t *pw_ = s_q + (size_t)qi * 2 * W + lane * WPL;
            const uint32_t *yw_ = pw_ + W;
            uint32_t tp[WPL], ty[WPL];
#pragma unroll
            for (int w = 0; w < WPL; ++w) {
                tp[w] = tw[w] & pw_[w];
                ty[w] = tp[w] & yw_[w];
            }
            const uint16_t *stair = nullptr;
            if (prm.append) {
                const uint32_t off = s_off[qi];
                if (off != PPP_NO_TABLE) stair = prm.nmax + off;
            }
            PairOut res;
            const uint32_t fl = evaluate_pair<WPL>(tw, tp, ty, dpre, dtotal, s_qc[qi], s_lf, my_queue, my_y, my_lo, stair, prm, lane, wc, &res);
            if (lane == (int)qi) {
                mine = res;
                mine_flags = fl;
            }
        }
        /* lane i holds the result of query q0 + i */
        if ((uint32_t)lane < nq && !(mine_flags & PF_REJECTED)) {
            const uint32_t q = q0 + lane;
            if (mine_flags & PF_EXACT) {
                push_exact(prm, (uint32_t)((uint64_t)q * nTl + tl), (mine_flags & PF_UNDERFLOW) != 0);
                if (!prm.append) ppp_emit(prm, q, t, 1.0, 0, 0, 0, -1.f);
            } else {
                ppp_emit(prm, q, t, mine.score, mine.yes, mine.number, mine.depth, mine.margin);
            }
        }
      } /* targets of the item */
    }
    flush_counters(prm, wc, lane);
}

/* ---------------------------------------------------------------- refine kernel: warp per surviving pair */
#ifndef PPP_REFINE_WARPS
#define PPP_REFINE_WARPS 16 /* warps per CTA the refine kernel is compiled for (register budget = 65536 / threads) */
#endif
template <int WPL>
__global__ void __launch_bounds__(PPP_REFINE_WARPS * 32, PPP_MINBLOCKS_OF(WPL)) ppp_refine_kernel(const SweepParams prm, const uint32_t *__restrict__ surv,
                                                         const unsigned long long *__restrict__ surv_count, unsigned long long surv_cap)
{
    constexpr uint32_t PPP_QCAP = PPP_QCAP_OF(WPL);
    constexpr int W = 32 * WPL;
    constexpr int GD = 32 * W;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw);
    double *s_lf = reinterpret_cast<double *>(smem_raw + 128);
    constexpr int LF_BYTES = (((GD + 2) * 8) + 127) & ~127;
    uint32_t *s_queue = reinterpret_cast<uint32_t *>(smem_raw + 128 + LF_BYTES);
    uint32_t *my_queue = s_queue + warp * (3 * PPP_QCAP);
    uint32_t *my_y = my_queue + PPP_QCAP;
    float *my_lo = reinterpret_cast<float *>(my_queue + 2 * PPP_QCAP);
    if (threadIdx.x == 0) {
        mbar_init(&bars[0], 1);
        fence_barrier_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t bytes = (uint32_t)((GD + 2) * 8);
        mbar_expect_tx(&bars[0], bytes);
        tma_load_1d(s_lf, prm.lf, bytes, &bars[0]);
    }
    mbar_wait(&bars[0], 0);
    const unsigned long long n = min(*surv_count, surv_cap);
    const uint32_t nTl = prm.t1 - prm.t0;
    WarpCounters wc;
    (void)nwarps;
    /* Survivors are drawn from a global counter (their costs differ by orders of magnitude), 8 at a time: one atomic
     * and one coalesced load of the ids per batch, and the rows of the whole batch (target, presence and yes planes) are
     * prefetched into L1 before the first pair is evaluated -- the kernel runs at 16 warps per SM, so the row loads'
     * latency is otherwise exposed at the start of every pair. */
    constexpr uint32_t BATCH = 8;
    for (;;) {
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(&prm.counters[CNT_ITEM], (unsigned long long)BATCH);
        base = __shfl_sync(FULL, base, 0);
        if (base >= n) break;
        const uint32_t m = (uint32_t)min((unsigned long long)BATCH, n - base);
        const uint32_t myid = ((uint32_t)lane < m) ? surv[base + lane] : 0u;
        for (uint32_t e0 = 0; e0 < m; e0 += 32 / WPL) { /* one 128-byte line per lane: WPL lines per row */
            const uint32_t e = e0 + (uint32_t)lane / WPL, line = (uint32_t)lane % WPL;
            const uint32_t pid = __shfl_sync(FULL, myid, e & 31);
            if (e < m) {
                const uint32_t pq = pid / nTl, pt = prm.t0 + pid % nTl;
                const uint32_t *rt = prm.T + (size_t)pt * W + line * 32, *rq = prm.Q + (size_t)pq * 2 * W + line * 32;
                asm volatile("prefetch.global.L1 [%0];" ::"l"(rt));
                asm volatile("prefetch.global.L1 [%0];" ::"l"(rq));
                asm volatile("prefetch.global.L1 [%0];" ::"l"(rq + W));
            }
        }
        for (uint32_t j = 0; j < m; ++j) {
            const uint32_t id = __shfl_sync(FULL, myid, j);
            const uint32_t q = id / nTl, tl = id % nTl, t = prm.t0 + tl;
            uint32_t tw[WPL], pw[WPL], yw[WPL], tp[WPL], ty[WPL];
            load_row<WPL>(prm.T + (size_t)t * W + lane * WPL, tw);
            load_row<WPL>(prm.Q + (size_t)q * 2 * W + lane * WPL, pw);
            load_row<WPL>(prm.Q + (size_t)q * 2 * W + W + lane * WPL, yw);
            uint32_t dlane = 0;
#pragma unroll
            for (int w = 0; w < WPL; ++w) {
                tp[w] = tw[w] & pw[w];
                ty[w] = tp[w] & yw[w];
                dlane += __popc(tw[w]);
            }
            uint32_t dpre, dtotal;
            depth_prefix(dlane, lane, dpre, dtotal);
            const QConst qc = prm.qc[q];
            const uint32_t off = prm.nmax_off[q];
            const uint16_t *stair = (off != PPP_NO_TABLE) ? prm.nmax + off : nullptr;
            PairOut res;
            const uint32_t fl = evaluate_pair<WPL>(tw, tp, ty, dpre, dtotal, qc, s_lf, my_queue, my_y, my_lo, stair, prm, lane, wc, &res);
            if (lane == 0 && !(fl & PF_REJECTED)) {
                if (fl & PF_EXACT) push_exact(prm, id, (fl & PF_UNDERFLOW) != 0);
                else ppp_emit(prm, q, t, res.score, res.yes, res.number, res.depth, res.margin);
            }
        }
    }
    flush_counters(prm, wc, lane);
}

/* ---------------------------------------------------------------- re-screen kernel: exact staircase test of the flagged pairs
 * The tensor-core chunk screen (ppp_prefilter_tc.cu, pair-level mode) flags every pair with a passing 128-genome CHUNK -- a
 * superset (about 40 x) of the pairs with a passing CANDIDATE -- as one column mask per (query tile, target).  This kernel
 * applies the candidate-level test of evaluate_pair() to each flagged pair -- number at the candidate <= nmax[its rank] for
 * some set bit of T & Y (ppp.pm:207-237) -- and appends the pairs that have one to the survivor list of the refine kernel.
 * One CTA per (query tile, block of targets): the tile's query planes are staged in shared memory once, every warp takes a
 * target row (registers) and walks the set bits of its mask, so that a flagged pair costs shared-memory reads, a packed warp
 * scan and a few staircase lookups instead of three row loads from L2. */
#define RS_ROWS 256 /* targets per work item */
template <int WPL>
__device__ __forceinline__ void load_words_smem(const uint32_t *src, uint32_t (&v)[WPL])
{
    if constexpr (WPL == 4) {
        const uint4 a = *reinterpret_cast<const uint4 *>(src);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
    } else if constexpr (WPL == 8) {
        const uint4 a = *reinterpret_cast<const uint4 *>(src), b = *(reinterpret_cast<const uint4 *>(src) + 1);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else if constexpr (WPL == 2) {
        const uint2 a = *reinterpret_cast<const uint2 *>(src);
        v[0] = a.x; v[1] = a.y;
    } else {
        v[0] = *src;
    }
}
template <int WPL, int NQ>
__global__ void __launch_bounds__(512) ppp_rescreen_masks_kernel(const SweepParams prm, const uint32_t *__restrict__ qorder, uint32_t nqk, bool pall,
                                                                 const uint32_t *__restrict__ masks, uint32_t *__restrict__ out,
                                                                 unsigned long long *__restrict__ out_count, unsigned long long out_cap)
{
    constexpr int W = 32 * WPL;
    constexpr int MW = NQ / 32;
    extern __shared__ __align__(16) unsigned char rs_smem[];
    uint32_t *s_qid = reinterpret_cast<uint32_t *>(rs_smem);   /* [NQ] */
    uint32_t *s_off = s_qid + NQ;                               /* [NQ] offset of the column's staircase in s_stair, PPP_NO_TABLE = accepts all */
    uint32_t *s_y = s_off + NQ;                                 /* [NQ][W] yes planes (already ANDed with the presence planes) */
    uint32_t *s_p = s_y + NQ * W;                               /* [NQ][W] presence planes (general launch only) */
    uint16_t *s_stair = reinterpret_cast<uint16_t *>(s_p + (pall ? 0 : NQ * W)); /* the tile's staircases, back to back */
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const uint32_t nTl = prm.t1 - prm.t0;
    const uint32_t nTb = (nTl + RS_ROWS - 1) / RS_ROWS, nQt = (nqk + NQ - 1) / NQ;
    const uint64_t nItems = (uint64_t)nQt * nTb;
    const uint64_t span = (nItems + gridDim.x - 1) / gridDim.x;
    const uint64_t i0 = (uint64_t)blockIdx.x * span, i1 = min(nItems, i0 + span);
    uint32_t cur_qt = 0xffffffffu;
    for (uint64_t item = i0; item < i1; ++item) {
        const uint32_t qt = (uint32_t)(item / nTb), tb = (uint32_t)(item % nTb);
        if (qt != cur_qt) {
            __syncthreads();
            const uint32_t nq = min((uint32_t)NQ, nqk - qt * NQ);
            if (warp == 0) { /* staircase offsets inside shared memory: a warp scan over NQ / 32 columns per lane */
                uint32_t len[MW], qid[MW], sum = 0;
#pragma unroll
                for (int m = 0; m < MW; ++m) {
                    const uint32_t j = (uint32_t)(MW * lane + m);
                    qid[m] = j < nq ? qorder[qt * NQ + j] : 0xffffffffu;
                    const uint32_t off = qid[m] != 0xffffffffu ? prm.nmax_off[qid[m]] : PPP_NO_TABLE;
                    len[m] = (qid[m] != 0xffffffffu && off != PPP_NO_TABLE) ? prm.qc[qid[m]].ny + 1 : 0;
                    sum += len[m];
                }
                uint32_t incl = sum;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t v = __shfl_up_sync(FULL, incl, o);
                    if (lane >= o) incl += v;
                }
                uint32_t run = incl - sum;
#pragma unroll
                for (int m = 0; m < MW; ++m) {
                    const uint32_t j = (uint32_t)(MW * lane + m);
                    s_qid[j] = qid[m];
                    s_off[j] = len[m] ? run : PPP_NO_TABLE;
                    run += len[m];
                }
            }
            __syncthreads();
            for (uint32_t j = warp; j < nq; j += nwarps) {
                if (s_off[j] == PPP_NO_TABLE) continue;
                const uint32_t qid = s_qid[j], len = prm.qc[qid].ny + 1;
                const uint16_t *src = prm.nmax + prm.nmax_off[qid];
                for (uint32_t i = lane; i < len; i += 32) s_stair[s_off[j] + i] = __ldg(src + i);
            }
            for (uint32_t i = threadIdx.x; i < (uint32_t)(NQ * W); i += blockDim.x) {
                const uint32_t j = i / W, w = i % W;
                uint32_t y = 0, p = 0;
                if (j < nq) {
                    const uint32_t qid = qorder[qt * NQ + j];
                    p = pall ? 0xffffffffu : __ldg(prm.Q + (size_t)qid * 2 * W + w);
                    y = __ldg(prm.Q + (size_t)qid * 2 * W + W + w) & p;
                }
                s_y[i] = y;
                if (!pall) s_p[i] = p;
            }
            cur_qt = qt;
            __syncthreads();
        }
        for (uint32_t r = warp; r < RS_ROWS; r += nwarps) {
            const uint32_t tl = tb * RS_ROWS + r;
            if (tl >= nTl) break;
            uint32_t mw = lane < MW ? __ldg(masks + ((size_t)qt * nTl + tl) * MW + lane) : 0u;
            if (!__any_sync(FULL, mw != 0)) continue;
            uint32_t tw[WPL];
            load_row<WPL>(prm.T + (size_t)(prm.t0 + tl) * W + lane * WPL, tw);
            uint32_t passed[MW];
#pragma unroll
            for (int m = 0; m < MW; ++m) {
                uint32_t bits = __shfl_sync(FULL, mw, m);
                uint32_t okbits = 0;
                while (bits) { /* warp-uniform walk over the flagged columns of this row */
                    const uint32_t bit = __ffs(bits) - 1;
                    bits &= bits - 1;
                    const uint32_t j = 32u * m + bit;
                    const uint32_t off = s_off[j];
                    const bool real = s_qid[j] != 0xffffffffu;  /* padding columns are never flagged */
                    bool pass = real && off == PPP_NO_TABLE;    /* a query without a staircase accepts every pair */
                    if (real && !pass) {
                        uint32_t aw[WPL], yw[WPL], cy = 0, cn = 0;
                        load_words_smem<WPL>(s_y + j * W + lane * WPL, yw);
                        if (!pall) load_words_smem<WPL>(s_p + j * W + lane * WPL, aw);
#pragma unroll
                        for (int w = 0; w < WPL; ++w) {
                            aw[w] = pall ? tw[w] : (tw[w] & aw[w]);
                            yw[w] &= tw[w];
                            cy += __popc(yw[w]);
                            cn += __popc(aw[w]);
                        }
                        const uint32_t packed = cy | (cn << 16);
                        uint32_t incl = packed;
#pragma unroll
                        for (int o = 1; o < 32; o <<= 1) {
                            const uint32_t v = __shfl_up_sync(FULL, incl, o);
                            if (lane >= o) incl += v;
                        }
                        const uint16_t *stair = s_stair + off;
                        uint32_t yrun = (incl - packed) & 0xffffu, nbase = (incl - packed) >> 16;
#pragma unroll
                        for (int w = 0; w < WPL; ++w) {
                            const uint32_t cyw = __popc(yw[w]);
                            if (cyw && (nbase + 1u) <= (uint32_t)stair[yrun + cyw]) { /* word-level screen, one lookup */
                                uint32_t mbits = yw[w], yy = yrun;
                                while (mbits) {
                                    const uint32_t b = __ffs(mbits) - 1;
                                    mbits &= mbits - 1;
                                    ++yy;
                                    pass |= nbase + __popc(aw[w] & ((2u << b) - 1u)) <= (uint32_t)stair[yy];
                                }
                            }
                            yrun += cyw;
                            nbase += __popc(aw[w]);
                        }
                        pass = __any_sync(FULL, pass);
                    }
                    if (pass) okbits |= 1u << bit;
                }
                passed[m] = okbits;
            }
            uint32_t total = 0;
#pragma unroll
            for (int m = 0; m < MW; ++m) total += __popc(passed[m]);
            if (total) {
                unsigned long long base = 0;
                if (lane == 0) base = atomicAdd(out_count, (unsigned long long)total);
                base = __shfl_sync(FULL, base, 0);
                uint32_t before = 0;
#pragma unroll
                for (int m = 0; m < MW; ++m) {
                    if ((passed[m] >> lane) & 1u) {
                        const unsigned long long slot = base + before + __popc(passed[m] & ((1u << lane) - 1u));
                        if (slot < out_cap) out[slot] = (uint32_t)((uint64_t)s_qid[32u * m + lane] * nTl + tl);
                    }
                    before += __popc(passed[m]);
                }
            }
        }
    }
}

static size_t rescreen_fixed_smem(int wpl, int pall)
{
    const size_t NQ = pall ? 128 : 64, W = 32 * (size_t)wpl;
    return 2 * NQ * 4 + (pall ? 1 : 2) * NQ * W * 4;
}
/* largest staircase footprint (entries) of a query tile the re-screen kernel can stage */
extern "C" uint32_t ppp_rescreen_max_stair(int wpl, int pall)
{
    const size_t fixed = rescreen_fixed_smem(wpl, pall);
    if (fixed + 1024 > PPP_MAX_DYN_SMEM) return 0;
    return (uint32_t)((PPP_MAX_DYN_SMEM - 512 - fixed) / 2);
}

extern "C" cudaError_t ppp_launch_rescreen_masks(const SweepParams *prm, int wpl, int pall, const uint32_t *qorder, uint32_t nqk, uint32_t stair_cap,
                                                 const uint32_t *masks, uint32_t *out, unsigned long long *out_count, unsigned long long out_cap,
                                                 int nsm, cudaStream_t st)
{
    if (!nqk) return cudaSuccess;
    const int NQ = pall ? 128 : 64;
    const size_t smem = rescreen_fixed_smem(wpl, pall) + (((size_t)stair_cap * 2 + 15) & ~(size_t)15);
    const uint32_t nTl = prm->t1 - prm->t0;
    const uint64_t items = (uint64_t)((nqk + NQ - 1) / NQ) * ((nTl + RS_ROWS - 1) / RS_ROWS);
    const unsigned grid = (unsigned)std::min<uint64_t>(items, (uint64_t)nsm * 2);
    cudaError_t e = cudaSuccess;
#define RS_LAUNCH(WPL_, NQ_)                                                                                                                   \
    do {                                                                                                                                       \
        e = cudaFuncSetAttribute(ppp_rescreen_masks_kernel<WPL_, NQ_>, cudaFuncAttributeMaxDynamicSharedMemorySize, PPP_MAX_DYN_SMEM);                \
        if (e != cudaSuccess) return e;                                                                                                        \
        ppp_rescreen_masks_kernel<WPL_, NQ_><<<grid, 512, smem, st>>>(*prm, qorder, nqk, pall != 0, masks, out, out_count, out_cap);           \
    } while (0)
    if (smem > PPP_MAX_DYN_SMEM) return cudaErrorInvalidValue;
    switch (wpl * 2 + (pall ? 1 : 0)) {
    case 2: RS_LAUNCH(1, 64); break;
    case 3: RS_LAUNCH(1, 128); break;
    case 4: RS_LAUNCH(2, 64); break;
    case 5: RS_LAUNCH(2, 128); break;
    case 8: RS_LAUNCH(4, 64); break;
    case 9: RS_LAUNCH(4, 128); break;
    case 16: RS_LAUNCH(8, 64); break;
    case 17: RS_LAUNCH(8, 128); break;
    default: return cudaErrorInvalidValue;
    }
#undef RS_LAUNCH
    return cudaGetLastError();
}

/* transpose row-major planes T[nT][W] into chunk-major Tt[W/4][nT] of uint4 (coalesced per-thread row streaming) */
__global__ void ppp_transpose_kernel(const uint4 *__restrict__ T, uint4 *__restrict__ Tt, uint32_t nT, uint32_t C, uint32_t t_begin,
                                     uint32_t t_end)
{
    __shared__ uint4 tile[32][33];
    const uint32_t t0 = t_begin + blockIdx.x * 32, c0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const uint32_t t = t0 + r, c = c0 + threadIdx.x;
        if (t < t_end && c < C) tile[r][threadIdx.x] = T[(size_t)t * C + c];
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const uint32_t c = c0 + r, t = t0 + threadIdx.x;
        if (t < t_end && c < C) Tt[(size_t)c * nT + t] = tile[threadIdx.x][r];
    }
}

/* ---------------------------------------------------------------- launchers */
static size_t sweep_smem_bytes(int wpl, int nwarps)
{
    const size_t W = 32 * (size_t)wpl, GD = 32 * W;
    return 128 + ((((GD + 2) * 8) + 127) & ~(size_t)127) + PPP_QTILE * 2 * W * 4 + PPP_QTILE * sizeof(QConst) + PPP_QTILE * 4 +
           (size_t)nwarps * 3 * PPP_QCAP_OF(wpl) * 4 + 64;
}
static size_t refine_smem_bytes(int wpl, int nwarps)
{
    const size_t W = 32 * (size_t)wpl, GD = 32 * W;
    return 128 + ((((GD + 2) * 8) + 127) & ~(size_t)127) + (size_t)nwarps * 3 * PPP_QCAP_OF(wpl) * 4 + 64;
}


template <int WPL>
static cudaError_t launch_t(const SweepParams &prm, int nwarps, int nsm, cudaStream_t st)
{
    while (nwarps > 1 && sweep_smem_bytes(WPL, nwarps) > PPP_MAX_DYN_SMEM) --nwarps; /* G = 8192: the tile + table leave room for 15 warps */
    const size_t smem = sweep_smem_bytes(WPL, nwarps);
    cudaError_t e = cudaFuncSetAttribute(ppp_sweep_kernel<WPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, PPP_MAX_DYN_SMEM);
    if (e != cudaSuccess) return e;
    int per_sm = 1;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ppp_sweep_kernel<WPL>, nwarps * 32, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    const uint32_t nTl = prm.t1 - prm.t0;
    const uint64_t items = (uint64_t)((prm.nQ + PPP_QTILE - 1) / PPP_QTILE) * ((nTl + PPP_SB - 1) / PPP_SB);
    uint64_t grid = (uint64_t)nsm * per_sm;
    if (grid > items) grid = items;
    if (grid < 1) grid = 1;
    ppp_sweep_kernel<WPL><<<(unsigned)grid, nwarps * 32, smem, st>>>(prm);
    return cudaGetLastError();
}

extern "C" cudaError_t ppp_launch_sweep(const SweepParams *prm, int wpl, int nwarps, int nsm, cudaStream_t st)
{
    switch (wpl) {
    case 1: return launch_t<1>(*prm, nwarps, nsm, st);
    case 2: return launch_t<2>(*prm, nwarps, nsm, st);
    case 4: return launch_t<4>(*prm, nwarps, nsm, st);
    case 8: return launch_t<8>(*prm, nwarps, nsm, st);
    default: return cudaErrorInvalidValue;
    }
}

/* pre-filter (thread per pair) + refine (warp per surviving pair); *surv_count must be zero on entry */
template <int WPL>
static cudaError_t launch_refine_t(const SweepParams &prm, const uint32_t *surv, const unsigned long long *surv_count, unsigned long long surv_cap,
                                   int nwarps, int nsm, cudaStream_t st)
{
    if (nwarps == 16) nwarps = PPP_REFINE_WARPS; /* the tuned setting asks for as many warps as the kernel was compiled for */
    while (nwarps > 1 && refine_smem_bytes(WPL, nwarps) > PPP_MAX_DYN_SMEM) --nwarps;
    const size_t smem = refine_smem_bytes(WPL, nwarps);
    cudaError_t e = cudaFuncSetAttribute(ppp_refine_kernel<WPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, PPP_MAX_DYN_SMEM);
    if (e != cudaSuccess) return e;
    int per_sm = 1;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ppp_refine_kernel<WPL>, nwarps * 32, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    ppp_refine_kernel<WPL><<<(unsigned)(nsm * per_sm), nwarps * 32, smem, st>>>(prm, surv, surv_count, surv_cap);
    return cudaGetLastError();
}

/* refine only: the survivor list was produced by ppp_launch_prefilter2 (ppp_prefilter.cu) */
extern "C" cudaError_t ppp_launch_refine(const SweepParams *prm, int wpl, const uint32_t *surv, const unsigned long long *surv_count,
                                         unsigned long long surv_cap, int nwarps, int nsm, cudaStream_t st)
{
    switch (wpl) {
    case 1: return launch_refine_t<1>(*prm, surv, surv_count, surv_cap, nwarps, nsm, st);
    case 2: return launch_refine_t<2>(*prm, surv, surv_count, surv_cap, nwarps, nsm, st);
    case 4: return launch_refine_t<4>(*prm, surv, surv_count, surv_cap, nwarps, nsm, st);
    case 8: return launch_refine_t<8>(*prm, surv, surv_count, surv_cap, nwarps, nsm, st);
    default: return cudaErrorInvalidValue;
    }
}

extern "C" cudaError_t ppp_launch_transpose_rows(const void *T, void *Tt, uint32_t nT, int wpl, uint32_t t_begin, uint32_t t_end, cudaStream_t st)
{
    if (t_end <= t_begin) return cudaSuccess;
    const uint32_t C = 8u * (uint32_t)wpl;
    dim3 grid((t_end - t_begin + 31) / 32, (C + 31) / 32), block(32, 8);
    ppp_transpose_kernel<<<grid, block, 0, st>>>(reinterpret_cast<const uint4 *>(T), reinterpret_cast<uint4 *>(Tt), nT, C, t_begin, t_end);
    return cudaGetLastError();
}
extern "C" cudaError_t ppp_launch_transpose(const void *T, void *Tt, uint32_t nT, int wpl, cudaStream_t st)
{
    return ppp_launch_transpose_rows(T, Tt, nT, wpl, 0u, nT, st);
}
