/*
 * prophylo_b200/csrc/ppp_prefilter.cu -- the integer screen of filtered PPP runs (top-k / threshold), sm_100a.
 *
 * Screens EVERY (query, target) pair of a target chunk against the query's staircase (ppp_topk.cu): a sweep candidate
 * of PhyloProf::Algorithm::ppp::get_match_score (/root/reference/lib/PhyloProf/Algorithm/ppp.pm:207-237) with running
 * counts (yes, number) can reach the query's threshold only if number <= nmax[yes].  Pairs with no such candidate are
 * finished here; the others go to the warp-per-pair refine kernel (ppp_sweep.cu) through the survivor list.
 *
 * Thread per target, a tile of up to 32 queries in shared memory, 8 queries at a time in registers:
 *   fast pass   per CH-word chunk of the row and per query: AND with the query's yes plane, carry-save adders fold the
 *               CH words so that 3 (CH = 4) or 4 (CH = 8) POPC give the chunk's yes count (POPC is the scarce pipe:
 *               16 lanes/clk/SM), one shared-memory lookup of the staircase at yes_after and one compare:
 *                   number_before + yes_in_chunk <= nmax[yes_after]
 *               (a candidate of rank y inside the chunk has number >= number_before + (y - yes_before), and
 *               nmax[y] - y is non-decreasing, so the chunk's last candidate is the easiest one).  The pass is
 *               branch-free: chunks that pass the screen are pushed on a per-thread queue in shared memory with
 *               predicated stores.
 *   drain       the queued (query, chunk) items are re-screened word by word and then candidate by candidate.  All
 *               lanes of a warp drain their own queues together, so the rare slow path runs with most lanes busy instead
 *               of one lane in 32 (a chunk passes the coarse screen for 0.2 .. 40 % of the pairs, depending on its
 *               position in the row, which would send nearly every warp down a divergent branch).
 * Queries whose presence plane is all ones (number == depth of the target, the common case) run in their own launch:
 * number_before comes from a per-thread prefix table of the target row, computed once per target block.
 * Queries with P == Y (the cross-score form: every genome the query knows is a "yes") run in a third one: number == yes
 * at every candidate, so a candidate passes iff yes <= nmax[yes], the slack nmax[y] - y is non-decreasing, and the pair has a
 * passing candidate iff its LAST one passes -- one lookup per pair at the row's total popc(T & Y), no queue, no drain.
 *
 * The staircase INDEX is kept doubled (a byte offset into the int16 tables: base + 2 * yes, advanced with one shift-add);
 * counts and staircase values are plain.
 */
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>

#include "../../include/ppp_b200.h"
#include "ppp_common.cuh"
#include "ppp_pair.cuh"

#define PF2_THREADS 512
#define PF2_QG 8    /* queries per register group */
#define PF2_QCAP 16 /* queue entries per thread; drained when a lane could overflow during the next chunk */
#ifndef PF2_HEAD
#define PF2_HEAD 1  /* leading chunks of a row screened word by word in the fast pass (number is still small there) */
#endif
#ifndef PF2_CH
#define PF2_CH 4
#endif
#define PF2_QT(MODE) ((MODE) ? 32 : 16) /* queries per shared-memory tile (MODE: 0 general, 1 all-ones P, 2 P == Y) */
#ifdef PF2_STATS
#define PF2_STAT(x) x
#else
#define PF2_STAT(x)
#endif

struct Pf2Params {
    const uint4 *Tt;          /* [W/4][ntt] chunk-major transposed target planes */
    const uint32_t *Q;        /* [nQ][2][W] */
    const QConst *qc;
    const uint16_t *nmax;     /* staircases */
    const uint32_t *nmax_off; /* [nQ] active offsets, PPP_NO_TABLE = the query accepts every pair */
    const uint32_t *qorder;   /* [nqk] query ids of this launch */
    uint32_t *surv;
    unsigned long long *surv_count;
    unsigned long long surv_cap;
    unsigned long long *stats; /* optional [8]: 1 pushes, 2 drains, 3 drain iterations, 4 word passes */
    unsigned long long *unit_counter; /* next unclaimed item of this launch (zero when the launch starts) */
    uint32_t ntt, t0, nTl, nqk;
    uint32_t stair_cap; /* int16 entries of shared memory reserved for the tile's staircases */
    uint32_t sb_blocks; /* target blocks per super-block of the item order (see the kernel) */
    uint32_t tile_q;    /* queries per tile, <= PF2_QT(MODE): fewer when the staircases of a full tile would not fit (very dense yes planes) */
};

__device__ __forceinline__ uint32_t lop3_xor3(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}
__device__ __forceinline__ uint32_t lop3_maj(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, 0xe8;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}

__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t v)
{
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

/* popcount of CH words through carry-save adders: CH = 4 -> 3 POPC, CH = 8 -> 4 POPC */
template <int CH>
__device__ __forceinline__ uint32_t popc_csa(const uint32_t (&b)[CH])
{
    if constexpr (CH == 4) {
        const uint32_t s = lop3_xor3(b[0], b[1], b[2]), c = lop3_maj(b[0], b[1], b[2]);
        return __popc(s) + __popc(b[3]) + 2u * __popc(c);
    } else {
        const uint32_t s1 = lop3_xor3(b[0], b[1], b[2]), c1 = lop3_maj(b[0], b[1], b[2]);
        const uint32_t s2 = lop3_xor3(b[3], b[4], b[5]), c2 = lop3_maj(b[3], b[4], b[5]);
        const uint32_t s3 = lop3_xor3(s1, s2, b[6]), c3 = lop3_maj(s1, s2, b[6]);
        const uint32_t t = lop3_xor3(c1, c2, c3), f = lop3_maj(c1, c2, c3);
        return __popc(s3) + __popc(b[7]) + 2u * __popc(t) + 4u * __popc(f);
    }
}

template <int CH>
__device__ __forceinline__ void load_chunk(const uint4 *__restrict__ Tt, uint32_t ntt, uint32_t t, int c, bool live, uint32_t (&tw)[CH])
{
#pragma unroll
    for (int i = 0; i < CH / 4; ++i) {
        const uint4 v = live ? __ldg(Tt + (size_t)(c * (CH / 4) + i) * ntt + t) : make_uint4(0, 0, 0, 0);
        tw[4 * i + 0] = v.x; tw[4 * i + 1] = v.y; tw[4 * i + 2] = v.z; tw[4 * i + 3] = v.w;
    }
}
template <int CH>
__device__ __forceinline__ void load_words(const uint32_t *__restrict__ src, uint32_t (&v)[CH])
{
#pragma unroll
    for (int i = 0; i < CH / 4; ++i) {
        const uint4 x = *reinterpret_cast<const uint4 *>(src + 4 * i);
        v[4 * i] = x.x; v[4 * i + 1] = x.y; v[4 * i + 2] = x.z; v[4 * i + 3] = x.w;
    }
}

/* queue entry: bits 0-15 doubled staircase index AFTER the chunk (or word), 16-20 chunk, 21-23 word, 24-28 compact
 * query slot, 29 word-level entry (pushed by the head of the fast pass).  Entries stay valid for the whole item (same
 * query tile, same target), so the queue is only drained when a lane runs out of room and at the end of the item. */
#define PF2_E_WORD 0x20000000u

template <int WPL, int MODE>
__global__ void __launch_bounds__(PF2_THREADS) ppp_prefilter2_kernel(const Pf2Params a)
{
    constexpr bool PALL = MODE == 1, PYY = MODE == 2;
    constexpr int W = 32 * WPL;
    constexpr int CH = (WPL == 8 || PF2_CH == 8) ? 8 : 4; /* words per chunk; at most 32 chunks per row */
    constexpr int NCH = W / CH;
    constexpr int HEAD = (PF2_HEAD < NCH) ? PF2_HEAD : NCH;
    constexpr int QW = MODE ? 1 : 2;       /* queue words per entry */
    constexpr int QT = PF2_QT(MODE);
    constexpr int NSLOT = QT + PF2_QG;     /* compact slots incl. padding of the last register group */
    constexpr int QSTRIDE = QW * PF2_THREADS;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw);
    uint32_t *s_qid = reinterpret_cast<uint32_t *>(smem_raw + 64);   /* [32] query id of compact slot i */
    uint32_t *s_sbase = s_qid + 32;                                  /* [40] doubled staircase index of slot i's entry 0 */
    uint32_t *s_misc = s_sbase + 40;                                 /* [0] nact, [1] inactive count */
    uint32_t *s_inact = s_misc + 8;                                  /* [32] query ids that accept every pair */
    uint32_t *s_pass = reinterpret_cast<uint32_t *>(smem_raw + 1024); /* [THREADS] per target: compact slots with a passing candidate */
    uint32_t *s_y = s_pass + PF2_THREADS;                            /* [NSLOT][W] yes planes of the compact slots */
    uint32_t *s_p = s_y + NSLOT * W;                                 /* [NSLOT][W] presence planes (general launch) */
    uint16_t *s_pref = reinterpret_cast<uint16_t *>(s_y + (MODE ? 1 : 2) * NSLOT * W); /* [NCH][THREADS] doubled number before chunk */
    uint32_t *s_queue = reinterpret_cast<uint32_t *>(s_pref + (PALL ? NCH * PF2_THREADS : 0)); /* [QCAP][QW][THREADS] (none for P == Y) */
    int16_t *s_stair = reinterpret_cast<int16_t *>(s_queue + (PYY ? 0 : PF2_QCAP) * QSTRIDE);  /* doubled staircases */
    const unsigned char *s_stair_b = reinterpret_cast<const unsigned char *>(s_stair);

    const int tid = threadIdx.x, lane = tid & 31;
    const uint32_t nTblocks = (a.nTl + PF2_THREADS - 1) / PF2_THREADS;
    const uint32_t nQtiles = (a.nqk + a.tile_q - 1) / a.tile_q;
    const uint32_t nItems = nQtiles * nTblocks; /* a launch covers at most 2^31 pairs = 2^18 items of 16 x 512 */
    /* Item order: super-blocks of sb_blocks target blocks (32 MB of planes); inside a super-block tile-major, i.e. every query
     * tile meets the super-block's targets before the next super-block starts.  A launch of 2^31 pairs spans the whole target
     * set (512 MB at config 3): in plain tile-major order every query tile streamed it from DRAM again (ncu: 81 GB per step);
     * a super-block stays in L2 while all tiles pass over it.
     * Items (query tile, block of 512 targets) are claimed from a global counter in runs of up to 16 consecutive
     * items, shrinking as the launch drains (guided self-scheduling: remaining / (2 x CTAs)).  Static spans left the CTA slots
     * 22 % idle on average (ncu: 26 of 32 warps resident per active cycle): item costs differ -- dense targets and dense yes
     * planes queue many more chunks -- and the CTA that finishes first leaves its SM half empty. */
    uint32_t *s_unit = reinterpret_cast<uint32_t *>(smem_raw + 16); /* [0] first item, [1] one past the last */
    unsigned int *unit_counter = reinterpret_cast<unsigned int *>(a.unit_counter); /* low word of the 64-bit slot */

    if (tid == 0) {
        mbar_init(bar, 1);
        fence_barrier_init();
    }
    __syncthreads();
    uint32_t parity = 0, cur_qt = 0xffffffffu;
    PF2_STAT(unsigned long long st_push = 0; unsigned long long st_drain = 0; unsigned long long st_iter = 0; unsigned long long st_word = 0;)

    for (;;) {
    __syncthreads(); /* every thread is done with the previous run of items (and has read s_unit) */
    if (tid == 0) {
        const uint32_t seen = *reinterpret_cast<volatile unsigned int *>(unit_counter);
        const uint32_t rem = seen < nItems ? nItems - seen : 0u;
        const uint32_t take = max(1u, min(16u, rem / (2u * gridDim.x)));
        const uint32_t first = atomicAdd(unit_counter, take);
        s_unit[0] = first;
        s_unit[1] = min(nItems, first + take);
    }
    __syncthreads();
    if (s_unit[0] >= nItems) break;
    for (uint32_t item = s_unit[0]; item < s_unit[1]; ++item) {
        const uint32_t per_sb = nQtiles * a.sb_blocks;
        const uint32_t sbi = item / per_sb, r = item - sbi * per_sb;
        const uint32_t nb = min(a.sb_blocks, nTblocks - sbi * a.sb_blocks); /* blocks of this super-block (the last one may be short) */
        const uint32_t qt = r / nb, tb = sbi * a.sb_blocks + (r - qt * nb);
        if (qt != cur_qt) {
            /* ---- stage the tile: compact the queries that have a staircase, copy their planes (TMA) and staircases */
            __syncthreads();
            const uint32_t q0 = qt * a.tile_q;
            const uint32_t nq = min(a.tile_q, a.nqk - q0);
            if (tid < 32) {
                uint32_t qid = 0, off = PPP_NO_TABLE, len = 0;
                if ((uint32_t)tid < nq) {
                    qid = a.qorder[q0 + tid];
                    off = a.nmax_off[qid];
                    len = a.qc[qid].ny + 1;
                }
                const bool act = off != PPP_NO_TABLE;
                const uint32_t am = __ballot_sync(FULL, act), im = __ballot_sync(FULL, (uint32_t)tid < nq && !act);
                const uint32_t slot = __popc(am & ((1u << tid) - 1u));
                uint32_t incl = act ? len : 0;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t v = __shfl_up_sync(FULL, incl, o);
                    if (lane >= o) incl += v;
                }
                const uint32_t nact = __popc(am);
                if (tid == 0) {
                    s_misc[0] = nact;
                    s_misc[1] = __popc(im);
                    fence_proxy_async();
                    mbar_expect_tx(bar, nact * (uint32_t)(QW * W * 4));
                }
                __syncwarp();
                if (act) {
                    s_qid[slot] = qid;
                    s_sbase[slot] = 2u * (1u + incl - len); /* entry 0 of the shared array is the dummy staircase */
                    if (MODE == 0) tma_load_1d(s_p + slot * W, a.Q + (size_t)qid * 2 * W, W * 4, bar);
                    tma_load_1d(s_y + slot * W, a.Q + (size_t)qid * 2 * W + W, W * 4, bar);
                } else if ((uint32_t)tid < nq) {
                    s_inact[__popc(im & ((1u << tid) - 1u))] = qid;
                }
                if ((uint32_t)tid >= nact) s_sbase[tid] = 0; /* padding slots use the dummy staircase */
                if (tid < 8) s_sbase[32 + tid] = 0;
            }
            __syncthreads();
            const uint32_t nact = s_misc[0];
            const uint32_t npad = (nact + PF2_QG - 1) / PF2_QG * PF2_QG;
            /* padding slots: all-zero planes */
            for (uint32_t i = nact * W + tid; i < npad * W; i += PF2_THREADS) {
                s_y[i] = 0;
                if (MODE == 0) s_p[i] = 0;
            }
            if (tid == 0) s_stair[0] = -1;
            /* staircases; entry 0 of each (no candidate has yes == 0) can never pass */
            for (uint32_t s = tid >> 5; s < nact; s += PF2_THREADS / 32) {
                const uint32_t qid = s_qid[s];
                const uint16_t *src = a.nmax + a.nmax_off[qid];
                const uint32_t len = a.qc[qid].ny + 1;
                int16_t *dst = s_stair + (s_sbase[s] >> 1);
                /* P == Y: the SLACK nmax[y] - y instead (number == yes at every candidate: it passes iff this is >= 0) */
                for (uint32_t i = lane; i < len; i += 32)
                    dst[i] = i ? (int16_t)((int)__ldg(src + i) - (PYY ? (int)i : 0)) : (int16_t)-1;
            }
            mbar_wait(bar, parity);
            parity ^= 1;
            cur_qt = qt;
            __syncthreads();
        }
        const uint32_t nact = s_misc[0], ninact = s_misc[1];
        const uint32_t npad = (nact + PF2_QG - 1) / PF2_QG * PF2_QG;
        const uint32_t tl = tb * PF2_THREADS + tid;
        const bool live = tl < a.nTl;
        const uint32_t t = a.t0 + (live ? tl : 0);
        s_pass[tid] = 0; /* by compact slot; written by whichever lane of the warp re-screens this target's entries */

        if (PALL && npad) { /* number (= depth) before every chunk of this thread's target row */
            uint32_t run = live ? 0u : 40000u; /* threads past the last target: a number no staircase admits, so nothing is queued */
#pragma unroll 4
            for (int c = 0; c < NCH; ++c) {
                uint32_t tw[CH];
                load_chunk<CH>(a.Tt, a.ntt, t, c, live, tw);
                s_pref[c * PF2_THREADS + tid] = (uint16_t)run;
                run += popc_csa<CH>(tw);
            }
        }

        /* this thread's queue: entry i at s_queue[i * QSTRIDE + tid] (+ THREADS for the second word); qa is the 32-bit
         * shared-memory address of the next free entry (a 64-bit generic pointer costs two instructions per bump) */
        const uint32_t qa0 = smem_u32(s_queue) + 4u * (uint32_t)tid;
        uint32_t qa = qa0;

        /* Slow path, all lanes of the warp together: re-screen the queued chunks word by word, then candidate by
         * candidate (number at the candidate <= nmax[its rank]).  The warp's entries are POOLED: lane i takes the
         * i-th, (i+32)-th ... entry of the concatenated queues whoever pushed it (queue lengths differ a lot between
         * dense and sparse targets), so a drain costs ceil(total / 32) rounds instead of the longest queue. */
        auto drain = [&]() {
#ifdef PF2_NODRAIN
            qa = qa0;
            return;
#endif
            PF2_STAT(st_drain += 1;)
            __syncwarp(); /* the other lanes' pushes are read below */
            const uint32_t cnt = (qa - qa0) / (4u * QSTRIDE);
            uint32_t incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t v = __shfl_up_sync(FULL, incl, o);
                if (lane >= o) incl += v;
            }
            const uint32_t excl = incl - cnt, total = __shfl_sync(FULL, incl, 31);
            const uint32_t wbase = (uint32_t)tid & ~31u; /* first thread of this warp */
            for (uint32_t base = 0; base < total; base += 32) {
                PF2_STAT(st_iter += 1;)
                const uint32_t g = base + lane;
                /* owner = the last lane whose exclusive prefix is <= g (lanes with empty queues share their successor's) */
                uint32_t own = 0;
#pragma unroll
                for (int step = 16; step; step >>= 1) {
                    const uint32_t cand = own + step;
                    const uint32_t ex = __shfl_sync(FULL, excl, cand & 31);
                    if (ex <= g) own = cand;
                }
                const uint32_t oex = __shfl_sync(FULL, excl, own);
                if (g < total) {
                    const uint32_t otid = wbase + own, kq = g - oex;
                    const uint32_t *qe = s_queue + (kq * QW) * PF2_THREADS + otid;
                    const uint32_t e = qe[0];
                    const uint32_t sl = (e >> 24) & 31u, c = (e >> 16) & 31u;
                    if (!((s_pass[otid] >> sl) & 1u)) {
                        const uint32_t otl = tb * PF2_THREADS + otid;
                        const bool olive = otl < a.nTl;
                        uint32_t tw[CH], aw[CH], bw[CH];
                        load_chunk<CH>(a.Tt, a.ntt, a.t0 + (olive ? otl : 0), (int)c, olive, tw);
                        load_words<CH>(s_y + sl * W + c * CH, bw);
                        if (PALL) {
#pragma unroll
                            for (int w = 0; w < CH; ++w) aw[w] = tw[w];
                        } else {
                            load_words<CH>(s_p + sl * W + c * CH, aw);
#pragma unroll
                            for (int w = 0; w < CH; ++w) aw[w] &= tw[w];
                        }
                        uint32_t cyw[CH], cnw[CH], cys = 0, cns = 0;
                        const uint32_t wsel = (e & PF2_E_WORD) ? ((e >> 21) & 7u) : (uint32_t)(CH - 1);
#pragma unroll
                        for (int w = 0; w < CH; ++w) {
                            bw[w] &= aw[w];
                            cyw[w] = __popc(bw[w]);
                            cnw[w] = __popc(aw[w]);
                            if ((uint32_t)w <= wsel) { /* counts from the chunk's start up to and including the entry's word */
                                cys += cyw[w];
                                cns += cnw[w];
                            }
                        }
                        /* state before the chunk: the entry carries the state after its chunk (or word); yb = staircase entry index */
                        uint32_t yb = ((e & 0xffffu) >> 1) - cys;
                        uint32_t nb = PALL ? (uint32_t)s_pref[c * PF2_THREADS + otid] : qe[PF2_THREADS] - cns;
                        const uint32_t wlo = (e & PF2_E_WORD) ? wsel : 0u;
                        bool ok = false;
#pragma unroll
                        for (int w = 0; w < CH; ++w) {
                            if ((uint32_t)w >= wlo && (uint32_t)w <= wsel && cyw[w] &&
                                (int)(nb + cyw[w]) <= (int)s_stair[yb + cyw[w]]) {
                                PF2_STAT(st_word += 1;)
                                uint32_t yy = yb, bits = bw[w];
                                while (bits) {
                                    const uint32_t bit = __ffs(bits) - 1;
                                    bits &= bits - 1;
                                    ++yy;
                                    ok |= (int)(nb + __popc(aw[w] & ((2u << bit) - 1u))) <= (int)s_stair[yy];
                                }
                            }
                            yb += cyw[w];
                            nb += cnw[w];
                        }
                        if (ok) atomicOr(&s_pass[otid], 1u << sl);
                    }
                }
            }
            __syncwarp();
            qa = qa0;
        };

        if constexpr (PYY) {
            /* P == Y: count popc(T & Y) over the whole row, one lookup at the end */
            uint32_t mypass = 0;
            for (uint32_t g0 = 0; g0 < npad; g0 += PF2_QG) {
                uint32_t ys[PF2_QG];
#pragma unroll
                for (int j = 0; j < PF2_QG; ++j) ys[j] = s_sbase[g0 + j];
                const uint32_t *yg = s_y + g0 * W;
                uint32_t tw[CH];
                load_chunk<CH>(a.Tt, a.ntt, t, 0, live, tw);
                const uint4 *tnext = a.Tt + t;
#pragma unroll 1
                for (int c = 0; c < NCH; ++c) {
                    uint32_t tn[CH];
                    if (c + 1 < NCH) tnext += (size_t)(CH / 4) * a.ntt;
#pragma unroll
                    for (int i = 0; i < CH / 4; ++i) {
                        const uint4 v = __ldg(tnext + (size_t)i * a.ntt);
                        tn[4 * i + 0] = v.x; tn[4 * i + 1] = v.y; tn[4 * i + 2] = v.z; tn[4 * i + 3] = v.w;
                    }
#pragma unroll
                    for (int j = 0; j < PF2_QG; ++j) {
                        uint32_t b[CH];
                        load_words<CH>(yg + j * W + c * CH, b);
#pragma unroll
                        for (int w = 0; w < CH; ++w) b[w] &= tw[w];
                        ys[j] += 2u * popc_csa<CH>(b);
                    }
#pragma unroll
                    for (int w = 0; w < CH; ++w) tw[w] = tn[w];
                }
#pragma unroll
                for (int j = 0; j < PF2_QG; ++j)
                    if ((int)*reinterpret_cast<const int16_t *>(s_stair_b + ys[j]) >= 0) mypass |= 1u << (g0 + j);
            }
            s_pass[tid] = mypass; /* threads past the last target are ignored below */
        } else {
        for (uint32_t g0 = 0; g0 < npad; g0 += PF2_QG) {
            uint32_t ys[PF2_QG];  /* doubled staircase index: base + 2 * yes */
            uint32_t nbq[PF2_QG]; /* number so far (general launch) */
#pragma unroll
            for (int j = 0; j < PF2_QG; ++j) {
                ys[j] = s_sbase[g0 + j];
                nbq[j] = live ? 0u : 40000u; /* see s_pref: threads past the last target never pass */
            }
            const uint32_t *yg = s_y + g0 * W, *pg = s_p + g0 * W;
            const uint32_t gtag = g0 << 24;

            uint32_t tw[CH];
            load_chunk<CH>(a.Tt, a.ntt, t, 0, live, tw);
            /* running pointers instead of per-chunk index arithmetic: the row's next chunk (threads past the last target read
             * target t0's row, which their impossible `number` keeps out of every queue) and the depth prefix of this thread */
            const uint4 *tnext = a.Tt + t;
            const uint16_t *pref = s_pref + tid;
#pragma unroll 1
            for (int c = 0; c < NCH; ++c) {
                uint32_t tn[CH]; /* prefetch the next chunk of the row */
                if (c + 1 < NCH) tnext += (size_t)(CH / 4) * a.ntt;
#pragma unroll
                for (int i = 0; i < CH / 4; ++i) {
                    const uint4 v = __ldg(tnext + (size_t)i * a.ntt);
                    tn[4 * i + 0] = v.x; tn[4 * i + 1] = v.y; tn[4 * i + 2] = v.z; tn[4 * i + 3] = v.w;
                }
                const uint32_t nb_all = PALL ? (uint32_t)*pref : 0u;
                pref += PF2_THREADS;
                const uint32_t ctag = ((uint32_t)c << 16) + gtag;
                if (c < HEAD) {
                    /* head of the row: word-level screen (number is small here, a 128-genome chunk would nearly
                     * always pass); immediate predicated pushes */
                    uint32_t nbw[CH];
                    if (PALL) {
                        uint32_t run = nb_all;
#pragma unroll
                        for (int w = 0; w < CH; ++w) {
                            nbw[w] = run;
                            run += __popc(tw[w]);
                        }
                    }
#pragma unroll
                    for (int j = 0; j < PF2_QG; ++j) {
                        uint32_t yw[CH], pw[CH];
                        load_words<CH>(yg + j * W + c * CH, yw);
                        if (!PALL) load_words<CH>(pg + j * W + c * CH, pw);
#pragma unroll
                        for (int w = 0; w < CH; ++w) {
                            const uint32_t aw = PALL ? tw[w] : (tw[w] & pw[w]);
                            const uint32_t cy = __popc(aw & yw[w]);
                            const uint32_t nb = PALL ? nbw[w] : nbq[j];
                            ys[j] += 2u * cy;
                            if (!PALL) nbq[j] += __popc(aw);
                            const int lim = (int)*reinterpret_cast<const int16_t *>(s_stair_b + ys[j]);
                            if ((int)(nb + cy) <= lim) {
                                sts_u32(qa, ys[j] | ctag | ((uint32_t)w << 21) | ((uint32_t)j << 24) | PF2_E_WORD);
                                if (!PALL) sts_u32(qa + 4u * PF2_THREADS, nbq[j]);
                                qa += 4u * QSTRIDE;
                                PF2_STAT(st_push += 1;)
                            }
                        }
                        if (__any_sync(FULL, qa > qa0 + 4u * (PF2_QCAP - CH) * QSTRIDE)) drain(); /* room for the next query's CH pushes */
                    }
                } else {
                    /* phase A: no shared-memory stores, so the eight queries' load/AND/POPC/lookup chains overlap */
                    uint32_t flags = 0;
#pragma unroll
                    for (int j = 0; j < PF2_QG; ++j) {
                        uint32_t b[CH];
                        load_words<CH>(yg + j * W + c * CH, b);
                        uint32_t nb;
                        if (PALL) {
#pragma unroll
                            for (int w = 0; w < CH; ++w) b[w] &= tw[w];
                            nb = nb_all;
                        } else {
                            uint32_t aw[CH];
                            load_words<CH>(pg + j * W + c * CH, aw);
#pragma unroll
                            for (int w = 0; w < CH; ++w) {
                                aw[w] &= tw[w];
                                b[w] &= aw[w];
                            }
                            nb = nbq[j];
                            nbq[j] = nb + popc_csa<CH>(aw);
                        }
                        const uint32_t cy = popc_csa<CH>(b);
                        ys[j] += 2u * cy; /* one shift-add: the index is a byte offset */
                        const int lim = (int)*reinterpret_cast<const int16_t *>(s_stair_b + ys[j]);
                        /* d < 0: the chunk cannot pass.  One 3-input add and one funnel shift that moves d's sign bit into the
                         * flag word (query j ends up at bit QG-1-j) instead of compare + select + or */
                        const int d = lim - (int)cy - (int)nb;
                        flags = __funnelshift_l((uint32_t)d, flags, 1);
                    }
                    /* phase B: predicated pushes */
                    if (flags != (1u << PF2_QG) - 1u) {
#pragma unroll
                        for (int j = 0; j < PF2_QG; ++j) {
                            if (!((flags >> (PF2_QG - 1 - j)) & 1u)) {
                                sts_u32(qa, ys[j] | ctag | ((uint32_t)j << 24));
                                if (!PALL) sts_u32(qa + 4u * PF2_THREADS, nbq[j]);
                                qa += 4u * QSTRIDE;
                                PF2_STAT(st_push += 1;)
                            }
                        }
                    }
                }
#pragma unroll
                for (int w = 0; w < CH; ++w) tw[w] = tn[w];
                if (__any_sync(FULL, qa > qa0 + 4u * (PF2_QCAP - PF2_QG) * QSTRIDE)) drain();
            }
        }
        drain();
        }

        /* ---- survivors: compact slots with a passing candidate + every query without a staircase (live threads) */
        __syncwarp();
        const uint32_t passbits = s_pass[tid];
        const uint32_t mycnt = live ? (__popc(passbits) + ninact) : 0;
        uint32_t incl = mycnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t v = __shfl_up_sync(FULL, incl, o);
            if (lane >= o) incl += v;
        }
        const uint32_t total = __shfl_sync(FULL, incl, 31);
        if (total) {
            unsigned long long base = 0;
            if (lane == 31) base = atomicAdd(a.surv_count, (unsigned long long)total);
            base = __shfl_sync(FULL, base, 31);
            unsigned long long slot = base + (incl - mycnt);
            if (live) {
                uint32_t pb = passbits;
                while (pb) {
                    const uint32_t s = __ffs(pb) - 1;
                    pb &= pb - 1;
                    if (slot < a.surv_cap) a.surv[slot] = (uint32_t)((uint64_t)s_qid[s] * a.nTl + tl);
                    ++slot;
                }
                for (uint32_t i = 0; i < ninact; ++i) {
                    if (slot < a.surv_cap) a.surv[slot] = (uint32_t)((uint64_t)s_inact[i] * a.nTl + tl);
                    ++slot;
                }
            }
        }
    }
    }
#ifdef PF2_STATS
    if (a.stats) {
        st_push = (unsigned long long)warp_sum_d((double)st_push);
        st_word = (unsigned long long)warp_sum_d((double)st_word);
        if (lane == 0) {
            atomicAdd(&a.stats[1], st_push);
            atomicAdd(&a.stats[2], st_drain);
            atomicAdd(&a.stats[3], st_iter);
            atomicAdd(&a.stats[4], st_word);
        }
    }
#endif
}

/* ---------------------------------------------------------------- host side */
static size_t pf2_smem_bytes(int wpl, int mode, uint32_t stair_cap)
{
    const size_t W = 32 * (size_t)wpl, CH = (wpl == 8 || PF2_CH == 8) ? 8 : 4, NCH = W / CH;
    const size_t nslot = PF2_QT(mode) + PF2_QG;
    size_t b = 1024 + PF2_THREADS * 4 + (mode ? 1 : 2) * nslot * W * 4;
    if (mode == 1) b += NCH * PF2_THREADS * 2;
    if (mode != 2) b += (size_t)PF2_QCAP * (mode ? 1 : 2) * PF2_THREADS * 4;
    b += ((size_t)stair_cap * 2 + 15) & ~(size_t)15;
    return b + 64;
}

/* largest staircase capacity (int16 entries) a tile may use; 0 = this layout cannot run the kernel at all */
extern "C" uint32_t ppp_prefilter2_tile_queries(int mode) { return PF2_QT(mode); }

extern "C" uint32_t ppp_prefilter2_max_stair(int wpl, int mode)
{
    const size_t fixed = pf2_smem_bytes(wpl, mode, 0);
    const size_t limit = 220 * 1024;
    if (fixed + 64 > limit) return 0;
    const size_t room = (limit - fixed) / 2;
    return (uint32_t)std::min<size_t>(room, 32767); /* the doubled index must fit 16 bits */
}

template <int WPL, int MODE>
static cudaError_t pf2_launch_t(const Pf2Params &p, int nsm, cudaStream_t st)
{
    const size_t smem = pf2_smem_bytes(WPL, MODE, p.stair_cap);
    cudaError_t e = cudaFuncSetAttribute(ppp_prefilter2_kernel<WPL, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, PPP_MAX_DYN_SMEM);
    if (e != cudaSuccess) return e;
    int per_sm = 1;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ppp_prefilter2_kernel<WPL, MODE>, PF2_THREADS, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    const uint64_t items = (uint64_t)((p.nqk + p.tile_q - 1) / p.tile_q) * ((p.nTl + PF2_THREADS - 1) / PF2_THREADS);
    uint64_t grid = std::min<uint64_t>((uint64_t)nsm * per_sm, items);
    if (grid < 1) grid = 1;
    ppp_prefilter2_kernel<WPL, MODE><<<(unsigned)grid, PF2_THREADS, smem, st>>>(p);
    return cudaGetLastError();
}

/* One launch over the queries listed in qorder[0..nqk): pall = 1 -> every listed query has an all-ones presence plane, 2 -> every
 * listed query has P == Y, 0 -> any presence plane.  tile_q consecutive entries of qorder form a tile (<= the kernel's 32 / 16);
 * stair_cap = the largest staircase footprint of a tile (int16 entries, dummy entry included).
 * *surv_count is NOT reset here (the all-ones and the general launch of a chunk append to the same list). */
extern "C" cudaError_t ppp_launch_prefilter2(const SweepParams *prm, int wpl, int pall, const void *Tt, uint32_t ntt, const uint32_t *qorder,
                                             uint32_t nqk, uint32_t tile_q, uint32_t stair_cap, uint32_t *surv, unsigned long long *surv_count,
                                             unsigned long long surv_cap, unsigned long long *stats, int nsm, cudaStream_t st)
{
    if (!nqk) return cudaSuccess;
    if (tile_q < 1 || tile_q > (uint32_t)PF2_QT(pall)) return cudaErrorInvalidValue;
    Pf2Params p;
    p.Tt = reinterpret_cast<const uint4 *>(Tt);
    p.Q = prm->Q;
    p.qc = prm->qc;
    p.nmax = prm->nmax;
    p.nmax_off = prm->nmax_off;
    p.qorder = qorder;
    p.surv = surv;
    p.surv_count = surv_count;
    p.surv_cap = surv_cap;
    p.stats = stats;
    p.unit_counter = prm->counters + (pall == 1 ? CNT_PF_UNIT_ALL : pall == 2 ? CNT_PF_UNIT_PYY : CNT_PF_UNIT_GEN);
    p.ntt = ntt;
    p.t0 = prm->t0;
    p.nTl = prm->t1 - prm->t0;
    p.nqk = nqk;
    p.stair_cap = stair_cap;
    p.tile_q = tile_q;
    p.sb_blocks = std::max<uint32_t>(1u, (32u << 20) / (PF2_THREADS * 128u * (uint32_t)wpl));
    switch (wpl * 4 + pall) {
    case 4: return pf2_launch_t<1, 0>(p, nsm, st);
    case 5: return pf2_launch_t<1, 1>(p, nsm, st);
    case 6: return pf2_launch_t<1, 2>(p, nsm, st);
    case 8: return pf2_launch_t<2, 0>(p, nsm, st);
    case 9: return pf2_launch_t<2, 1>(p, nsm, st);
    case 10: return pf2_launch_t<2, 2>(p, nsm, st);
    case 16: return pf2_launch_t<4, 0>(p, nsm, st);
    case 17: return pf2_launch_t<4, 1>(p, nsm, st);
    case 18: return pf2_launch_t<4, 2>(p, nsm, st);
    case 32: return pf2_launch_t<8, 0>(p, nsm, st);
    case 33: return pf2_launch_t<8, 1>(p, nsm, st);
    case 34: return pf2_launch_t<8, 2>(p, nsm, st);
    default: return cudaErrorInvalidValue;
    }
}
