/*
 * prophylo_b200/csrc/ppp_prefilter_tc.cu -- the chunk screen of filtered PPP runs on the 5th-generation tensor cores (sm_100a).
 *
 * Same question as ppp_prefilter.cu asks of EVERY (query, target) pair of a target chunk -- can any sweep candidate of
 * PhyloProf::Algorithm::ppp::get_match_score (/root/reference/lib/PhyloProf/Algorithm/ppp.pm:207-237) reach the query's
 * threshold, i.e. is `number <= nmax[yes]` anywhere along the row -- but the counts come from tcgen05.mma instead of POPC:
 * per 128-genome chunk of the rows,
 *        yes_after[t, q] = sum over genomes g up to the end of the chunk of  T[t, g] * Y[q, g]
 * is a {0,1} inner product, i.e. one K-block of a GEMM with M = 128 targets, N = 128 queries, K = 128 genomes, and
 * accumulating over the K-blocks IS the running prefix the staircase lookup needs.  Operands are the profile bits expanded
 * to unsigned bytes (target bit -> 0/1, query yes bit -> 0/2) in shared memory in the no-swizzle K-major core-matrix layout,
 * `tcgen05.mma.cta_group::1.kind::i8` accumulates exact int32 counts in tensor memory:
 *        P[t, q]  = &stair_q[0] + 2 * yes_after       (running, initialised to the shared-memory address of query q's
 *                                                      doubled staircase: the accumulator IS the lookup address)
 *        C[t, q]  = 2 * yes_in_chunk                   (restarted at every step)
 * and the epilogue warps read both back with tcgen05.ld and apply the chunk screen of ppp_prefilter.cu unchanged,
 *        number_before + yes_in_chunk <= nmax[yes_after]      (doubled: D2[t] + C <= lds.s16 [P]),
 * three instructions per (target, query, chunk): one shared-memory lookup, one 3-input add, one funnel shift that collects
 * the sign bits of 32 queries into a pass mask.  The first chunk of a row is screened word by word (K = 32 steps), as the
 * POPC kernel does.  Pairs with a passing chunk go to the warp-per-pair refine kernel (ppp_sweep.cu), whose own exact
 * candidate-level screen rejects the ones no candidate of which passes -- so the final results are identical to the POPC
 * path's; only the survivor list handed to refine is a superset.
 *
 * Queries whose presence plane is all ones only (number == depth of the target: D2 is a per-target prefix, which every
 * epilogue thread keeps for its own row from the packed target words); the general launch stays on the POPC kernel.
 *
 * Warp roles (416 threads, one CTA per SM): warps 0-7 epilogue (TMEM lane quarter = warp % 4, column half = warp / 4),
 * warps 8-11 producers (thread = one row of each operand tile: 16 packed bytes -> 128 operand bytes per chunk), warp 12
 * issues the MMAs (one elected lane).  Two target blocks of 128 are in flight per CTA and share every query tile, so that the
 * tensor core works on one block while the epilogue reads the other (accumulators: 2 x (P, C) x 128 columns = all 512 TMEM
 * columns).  Synchronisation is mbarriers only: full/empty per operand stage (producers <-> MMA), tfull/tfree per block
 * (MMA <-> epilogue), tcgen05.commit arrives for the tensor core.
 */
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>

#include "../../include/ppp_b200.h"
#include "ppp_common.cuh"
#include "ppp_pair.cuh"

#define TC_M 128
#define TC_N 128
#define TC_STAGES 2
#define TC_EPI_WARPS 8
#define TC_PROD_WARPS 4
#define TC_THREADS ((TC_EPI_WARPS + TC_PROD_WARPS + 1) * 32)
#define TC_TILE_BYTES (TC_M * 128)           /* one operand tile: 128 rows x 128 K-bytes */
#define TC_STAGE_BYTES (3 * TC_TILE_BYTES)   /* A0, A1, B */
#define TC_IDESC 0x08200020u                 /* kind::i8: D = S32 (2 << 4), A/B = unsigned 8 bit K-major, N = 128 (16 << 17), M = 128 (8 << 24) */

struct TcParams {
    const uint4 *Tt;          /* [W/4][ntt] chunk-major transposed target planes */
    const uint32_t *Q;        /* [nQ][2][W] */
    const QConst *qc;
    const uint16_t *nmax;
    const uint32_t *nmax_off;
    const uint32_t *qorder;   /* [nqk] query ids of this launch (all-ones presence planes) */
    uint32_t *surv;
    unsigned long long *surv_count;
    unsigned long long surv_cap;
    uint32_t ntt, t0, nTl, nqk;
    uint32_t stair_cap;       /* int16 entries of shared memory reserved for the tile's staircases */
};

/* ---------------------------------------------------------------- PTX wrappers */
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(TC_IDESC), "r"(accumulate)
        : "memory");
}
/* shared-memory matrix descriptor, no swizzle, K-major: 8-row x 16-byte core matrices of 128 contiguous bytes; LBO = distance
 * of the next core matrix along K (128 B), SBO = distance of the next 8-row group (1024 B); version 1 (sm_100) */
__device__ __forceinline__ uint64_t tc_smem_desc(uint32_t saddr)
{
    return (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)(128u >> 4) << 16) | ((uint64_t)(1024u >> 4) << 32) | ((uint64_t)1 << 46);
}
#define TC_LD32(taddr, r)                                                                                                                    \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, "  \
                 "%18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"                                              \
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),  \
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),     \
                   "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),     \
                   "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])                                                                        \
                 : "r"(taddr)                                                                                                               \
                 : "memory")
#define TC_ST32(taddr, r)                                                                                                                    \
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, " \
                 "%18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),                                  \
                 "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),   \
                 "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),    \
                 "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]),    \
                 "r"(r[31])                                                                                                                 \
                 : "memory")
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ int lds_s16(uint32_t addr)
{
    int v;
    asm volatile("ld.shared.s16 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

/* 16 profile bits (one half word) -> 16 operand bytes of value VAL each, stored as one core-matrix row */
template <uint32_t VAL>
__device__ __forceinline__ uint4 expand16(uint32_t bits)
{
    uint4 o;
    o.x = (((bits & 0xfu) * 0x00204081u) & 0x01010101u) * VAL;
    o.y = ((((bits >> 4) & 0xfu) * 0x00204081u) & 0x01010101u) * VAL;
    o.z = ((((bits >> 8) & 0xfu) * 0x00204081u) & 0x01010101u) * VAL;
    o.w = ((((bits >> 12) & 0xfu) * 0x00204081u) & 0x01010101u) * VAL;
    return o;
}
/* one row of an operand tile: the chunk's 4 words -> 8 core-matrix rows of 16 bytes (row r of the tile: 8-row group r / 8,
 * 128 B per core matrix along K, 16 B per row inside it) */
template <uint32_t VAL>
__device__ __forceinline__ void expand_row(unsigned char *tile, uint32_t r, const uint4 w)
{
    uint4 *dst = reinterpret_cast<uint4 *>(tile + (r >> 3) * 1024u + (r & 7u) * 16u);
    const uint32_t ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        dst[(2 * i) * 8] = expand16<VAL>(ws[i] & 0xffffu);      /* core matrix 2i     (+ 128 B each) */
        dst[(2 * i + 1) * 8] = expand16<VAL>(ws[i] >> 16);      /* core matrix 2i + 1 */
    }
}

template <int WPL>
__global__ void __launch_bounds__(TC_THREADS, 1) ppp_prefilter_tc_kernel(const TcParams a)
{
    constexpr int W = 32 * WPL;
    constexpr int NCH = W / 4;         /* 128-genome chunks per row */
    constexpr int NSTEP = NCH + 3;     /* chunk 0 is screened in four 32-genome steps */
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    uint64_t *bar_full = reinterpret_cast<uint64_t *>(smem_raw);   /* [STAGES] producers -> MMA */
    uint64_t *bar_empty = bar_full + TC_STAGES;                    /* [STAGES] MMA (commit) -> producers */
    uint64_t *bar_tfull = bar_empty + TC_STAGES;                   /* [2] MMA (commit) -> epilogue: accumulators of block b hold step s */
    uint64_t *bar_tfree = bar_tfull + 2;                           /* [2] epilogue -> MMA: accumulators of block b may be overwritten */
    uint32_t *s_tmem = reinterpret_cast<uint32_t *>(smem_raw + 128);
    uint32_t *s_qid = reinterpret_cast<uint32_t *>(smem_raw + 256);    /* [TC_N] query id of column j (padding: 0xffffffff) */
    uint32_t *s_sbase = s_qid + TC_N;                                   /* [TC_N] shared-memory ADDRESS of entry 0 of column j's doubled staircase */
    uint32_t *s_inact = s_sbase + TC_N;                                 /* [TC_N] queries of the tile that accept every pair (no staircase) */
    uint32_t *s_ninact = s_inact + TC_N;
    unsigned char *s_stage = smem_raw + 2048;
    int16_t *s_stair = reinterpret_cast<int16_t *>(s_stage + TC_STAGES * TC_STAGE_BYTES);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t nPairs = (a.nTl + 2 * TC_M - 1) / (2 * TC_M);
    const uint32_t nQtiles = (a.nqk + TC_N - 1) / TC_N;
    const uint64_t nItems = (uint64_t)nQtiles * nPairs;
    const uint64_t span = (nItems + gridDim.x - 1) / gridDim.x;
    const uint64_t i0 = (uint64_t)blockIdx.x * span, i1 = min(nItems, i0 + span);

    if (tid == 0) {
        for (int s = 0; s < TC_STAGES; ++s) {
            mbar_init(&bar_full[s], TC_PROD_WARPS * 32);
            mbar_init(&bar_empty[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&bar_tfull[b], 1);
            mbar_init(&bar_tfree[b], TC_EPI_WARPS * 32);
        }
        fence_barrier_init();
    }
    if (warp == TC_EPI_WARPS + TC_PROD_WARPS) { /* the MMA warp owns the tensor memory: all 512 columns */
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(s_tmem)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *s_tmem;

    /* barrier phases, tracked per role: a wait on parity p returns once the barrier has completed its phase p */
    uint32_t ph_full = 0, ph_empty = 0;          /* bit s = parity of the next wait on stage s */
    uint32_t ph_tfull = 0, ph_tfree = 0;         /* bit b */
    uint32_t chunk_seq = 0;                      /* running chunk counter of this CTA: stage = chunk_seq % STAGES */
    uint32_t cur_qt = 0xffffffffu;

    for (uint64_t item = i0; item < i1; ++item) {
        const uint32_t qt = (uint32_t)(item / nPairs), pr = (uint32_t)(item % nPairs);
        if (qt != cur_qt) {
            /* ---- stage the query tile: ids and doubled staircases of its columns (all warps; nothing is in flight here) */
            __syncthreads();
            const uint32_t q0 = qt * TC_N;
            const uint32_t nq = min((uint32_t)TC_N, a.nqk - q0);
            if (warp == 0) { /* offsets by a warp scan over four columns per lane */
                uint32_t len[4], off[4], qid[4], sum = 0;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint32_t col = 4 * lane + j;
                    qid[j] = col < nq ? a.qorder[q0 + col] : 0xffffffffu;
                    off[j] = qid[j] != 0xffffffffu ? a.nmax_off[qid[j]] : PPP_NO_TABLE;
                    len[j] = (qid[j] != 0xffffffffu && off[j] != PPP_NO_TABLE) ? a.qc[qid[j]].ny + 1 : 0;
                    sum += len[j];
                }
                uint32_t incl = sum;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t v = __shfl_up_sync(FULL, incl, o);
                    if (lane >= o) incl += v;
                }
                uint32_t run = 1 + incl - sum; /* entry 0 of the array is the dummy staircase */
                uint32_t nin = 0;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint32_t col = 4 * lane + j;
                    s_qid[col] = qid[j];
                    /* columns without a staircase (padding, or a query that accepts every pair: listed in s_inact and emitted
                     * for every live target below) point at the dummy entry, which never passes */
                    s_sbase[col] = smem_u32(s_stair) + 2u * (len[j] ? run : 0u);
                    run += len[j];
                    nin += (qid[j] != 0xffffffffu && !len[j]) ? 1u : 0u;
                }
                uint32_t iin = nin;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t v = __shfl_up_sync(FULL, iin, o);
                    if (lane >= o) iin += v;
                }
                uint32_t at = iin - nin;
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (qid[j] != 0xffffffffu && !len[j]) s_inact[at++] = qid[j];
                if (lane == 31) *s_ninact = iin;
            }
            __syncthreads();
            if (tid == 0) s_stair[0] = -2;
            for (uint32_t col = warp; col < nq; col += TC_THREADS / 32) {
                const uint32_t qid = s_qid[col];
                const uint32_t off = a.nmax_off[qid];
                if (off == PPP_NO_TABLE) continue;
                const uint32_t len = a.qc[qid].ny + 1;
                int16_t *dst = s_stair + ((s_sbase[col] - smem_u32(s_stair)) >> 1);
                const uint16_t *src = a.nmax + off;
                for (uint32_t i = lane; i < len; i += 32) dst[i] = i ? (int16_t)(2 * (int)__ldg(src + i)) : (int16_t)-2;
            }
            cur_qt = qt;
            __syncthreads();
        }
        const uint32_t tbase = pr * 2 * TC_M; /* first target (launch-local) of the pair of blocks */

        if (warp < TC_EPI_WARPS) {
            /* =============================================== EPILOGUE */
            const uint32_t quarter = warp & 3, half = warp >> 2;
            const uint32_t row = quarter * 32 + lane;
            const uint32_t lane_addr = (quarter * 32u) << 16;
            const uint32_t col0 = half * 64u;
            /* running accumulators start at the staircase addresses of their columns */
#pragma unroll 1
            for (int b = 0; b < 2; ++b) {
#pragma unroll 1
                for (int g = 0; g < 2; ++g) {
                    uint32_t v[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = s_sbase[col0 + 32 * g + j];
                    TC_ST32(tmem + lane_addr + (uint32_t)(b * 256) + col0 + 32u * g, v);
                }
            }
            tc_wait_st();
            tc_fence_before();
            mbar_arrive(&bar_tfree[0]);
            mbar_arrive(&bar_tfree[1]);
            uint32_t pass[2][2] = {{0, 0}, {0, 0}};
            /* doubled depth (= number, all-ones presence planes) before each step of this thread's row in either block, from
             * the packed target words themselves (immutable global data: no hand-over from the producers needed); the words
             * of the next chunk are fetched one step ahead */
            uint32_t trow[2];
            bool tlive[2];
#pragma unroll
            for (int b = 0; b < 2; ++b) {
                const uint32_t tl = tbase + b * TC_M + row;
                tlive[b] = tl < a.nTl;
                trow[b] = a.t0 + (tlive[b] ? tl : 0);
            }
            uint4 tw[2], twn[2];
            uint32_t run[2] = {0, 0};
#pragma unroll
            for (int b = 0; b < 2; ++b) tw[b] = tlive[b] ? __ldg(a.Tt + trow[b]) : make_uint4(0, 0, 0, 0);
#pragma unroll 1
            for (int s = 0; s < NSTEP; ++s) {
                const int c = s < 4 ? 0 : s - 3;
                if (s >= 3 && c + 1 < NCH) {
#pragma unroll
                    for (int b = 0; b < 2; ++b) twn[b] = tlive[b] ? __ldg(a.Tt + (size_t)(c + 1) * a.ntt + trow[b]) : make_uint4(0, 0, 0, 0);
                }
#pragma unroll
                for (int b = 0; b < 2; ++b) {
                    mbar_wait(&bar_tfull[b], (ph_tfull >> b) & 1u);
                    ph_tfull ^= 1u << b;
                    tc_fence_after();
                    const int kt = (int)run[b];
                    {
                        const uint32_t wsel = s == 0 ? tw[b].x : s == 1 ? tw[b].y : s == 2 ? tw[b].z : tw[b].w;
                        run[b] += s < 4 ? 2u * __popc(wsel) : 2u * (__popc(tw[b].x) + __popc(tw[b].y) + __popc(tw[b].z) + __popc(tw[b].w));
                    }
#pragma unroll
                    for (int g = 0; g < 2; ++g) {
                        uint32_t p[32], c32[32];
                        TC_LD32(tmem + lane_addr + (uint32_t)(b * 256) + col0 + 32u * g, p);
                        TC_LD32(tmem + lane_addr + (uint32_t)(b * 256 + 128) + col0 + 32u * g, c32);
                        tc_wait_ld();
                        if (g == 1 && s + 1 < NSTEP) { /* this step is in registers: the tensor core may go on with block b (after the
                                                         last step the next item's initialisation gives that signal) */
                            tc_fence_before();
                            mbar_arrive(&bar_tfree[b]);
                        }
                        uint32_t fl = 0;
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const int d = lds_s16(p[j]) - kt - (int)c32[j]; /* d < 0: no candidate of the chunk can pass */
                            fl = __funnelshift_l((uint32_t)d, fl, 1);       /* sign bits, column j at bit 31 - j */
                        }
                        pass[b][g] |= ~fl;
                    }
                }
                if (s >= 3) {
                    tw[0] = twn[0];
                    tw[1] = twn[1];
                }
            }
            /* ---- survivors of the two blocks: (query of column, target) for every set pass bit; queries without a staircase
             * accept every live target (emitted once, by the warps of column half 0) */
            const uint32_t ninact = half == 0 ? *s_ninact : 0u;
#pragma unroll 1
            for (int b = 0; b < 2; ++b) {
                const uint32_t tl = tbase + b * TC_M + row;
                const bool live = tl < a.nTl;
                uint32_t pb[2] = {live ? pass[b][0] : 0u, live ? pass[b][1] : 0u};
                const uint32_t mycnt = __popc(pb[0]) + __popc(pb[1]) + (live ? ninact : 0u);
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
#pragma unroll
                    for (int g = 0; g < 2; ++g) {
                        uint32_t m = pb[g];
                        while (m) {
                            const uint32_t bit = 31u - (uint32_t)__clz(m); /* column j sits at bit 31 - j */
                            m &= ~(1u << bit);
                            const uint32_t col = col0 + 32u * g + (31u - bit);
                            if (slot < a.surv_cap) a.surv[slot] = (uint32_t)((uint64_t)s_qid[col] * a.nTl + tl);
                            ++slot;
                        }
                    }
                    if (live)
                        for (uint32_t i = 0; i < ninact; ++i) {
                            if (slot < a.surv_cap) a.surv[slot] = (uint32_t)((uint64_t)s_inact[i] * a.nTl + tl);
                            ++slot;
                        }
                }
            }
        } else if (warp < TC_EPI_WARPS + TC_PROD_WARPS) {
            /* =============================================== PRODUCERS: thread = row r of A0, A1 and B */
            const uint32_t r = (uint32_t)tid - TC_EPI_WARPS * 32;
            const uint32_t tl0 = tbase + r, tl1 = tbase + TC_M + r;
            const bool live0 = tl0 < a.nTl, live1 = tl1 < a.nTl;
            const uint32_t qid = s_qid[r];
            const bool bq = qid != 0xffffffffu && a.nmax_off[qid] != PPP_NO_TABLE;
            const uint32_t *yrow = a.Q + (size_t)(bq ? qid : 0) * 2 * W + W;
            uint32_t cs = chunk_seq;
#pragma unroll 1
            for (int c = 0; c < NCH; ++c, ++cs) {
                const uint32_t st = cs % TC_STAGES;
                const uint4 w0 = live0 ? __ldg(a.Tt + (size_t)c * a.ntt + a.t0 + tl0) : make_uint4(0, 0, 0, 0);
                const uint4 w1 = live1 ? __ldg(a.Tt + (size_t)c * a.ntt + a.t0 + tl1) : make_uint4(0, 0, 0, 0);
                const uint4 wy = bq ? __ldg(reinterpret_cast<const uint4 *>(yrow) + c) : make_uint4(0, 0, 0, 0);
                if (cs >= TC_STAGES) { /* the stage's previous tenant has been consumed by the tensor core */
                    mbar_wait(&bar_empty[st], (ph_empty >> st) & 1u);
                    ph_empty ^= 1u << st;
                }
                unsigned char *stage = s_stage + st * TC_STAGE_BYTES;
                expand_row<1>(stage, r, w0);
                expand_row<1>(stage + TC_TILE_BYTES, r, w1);
                expand_row<2>(stage + 2 * TC_TILE_BYTES, r, wy);
                fence_proxy_async(); /* generic-proxy stores -> visible to the tensor core's async-proxy reads */
                mbar_arrive(&bar_full[st]);
            }
        } else if (lane == 0) {
            /* =============================================== MMA ISSUER (one thread) */
            uint32_t cs = chunk_seq;
#pragma unroll 1
            for (int c = 0; c < NCH; ++c, ++cs) {
                const uint32_t st = cs % TC_STAGES;
                mbar_wait(&bar_full[st], (ph_full >> st) & 1u);
                ph_full ^= 1u << st;
                tc_fence_after();
                const uint32_t sa = smem_u32(s_stage + st * TC_STAGE_BYTES);
                const uint64_t db = tc_smem_desc(sa + 2 * TC_TILE_BYTES);
                const int nsteps = c == 0 ? 4 : 1, kper = c == 0 ? 1 : 4;
                for (int s = 0; s < nsteps; ++s) {
#pragma unroll
                    for (int b = 0; b < 2; ++b) {
                        mbar_wait(&bar_tfree[b], (ph_tfree >> b) & 1u);
                        ph_tfree ^= 1u << b;
                        tc_fence_after();
                        const uint64_t da = tc_smem_desc(sa + b * TC_TILE_BYTES);
                        for (int k = 0; k < kper; ++k) {
                            const uint64_t koff = (uint64_t)(((s * kper + k) * 256) >> 4); /* 32 K-bytes = two core matrices = 256 B */
                            tc_mma_i8(tmem + (uint32_t)(b * 256), da + koff, db + koff, 1u);               /* P += */
                            tc_mma_i8(tmem + (uint32_t)(b * 256 + 128), da + koff, db + koff, k ? 1u : 0u); /* C  = / += */
                        }
                        tc_commit(&bar_tfull[b]);
                    }
                }
                tc_commit(&bar_empty[st]); /* arrives once every MMA that read this stage has completed */
            }
        }
        chunk_seq += NCH;
        /* the roles' barrier phases advance in lock step: per item every stage barrier completes NCH / STAGES phases, tfull[b]
         * NSTEP phases (one per commit) and tfree[b] NSTEP phases (the initialisation + NSTEP - 1 step releases); the empty-
         * barrier commits of an item's last TC_STAGES chunks are consumed by the producers of the next item */
    }

    tc_fence_before();
    __syncthreads();
    if (warp == TC_EPI_WARPS + TC_PROD_WARPS) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
    }
}

/* ---------------------------------------------------------------- host side */
static size_t tc_smem_bytes(int wpl, uint32_t stair_cap)
{
    const size_t nstep = 8 * (size_t)wpl + 3;
    (void)nstep;
    return 2048 + (size_t)TC_STAGES * TC_STAGE_BYTES + (((size_t)stair_cap * 2 + 15) & ~(size_t)15) + 64;
}

extern "C" uint32_t ppp_prefilter_tc_tile_queries(void) { return TC_N; }

/* largest staircase footprint (int16 entries, dummy entry included) a query tile may have; the doubled entries are looked up
 * through 32-bit shared-memory addresses, so only shared memory bounds it */
extern "C" uint32_t ppp_prefilter_tc_max_stair(int wpl)
{
    const size_t fixed = tc_smem_bytes(wpl, 0);
    const size_t limit = 226 * 1024;
    if (fixed + 64 > limit) return 0;
    return (uint32_t)std::min<size_t>((limit - fixed) / 2, 1u << 20);
}

template <int WPL>
static cudaError_t tc_launch_t(const TcParams &p, int nsm, cudaStream_t st)
{
    const size_t smem = tc_smem_bytes(WPL, p.stair_cap);
    cudaError_t e = cudaFuncSetAttribute(ppp_prefilter_tc_kernel<WPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const uint64_t items = (uint64_t)((p.nqk + TC_N - 1) / TC_N) * ((p.nTl + 2 * TC_M - 1) / (2 * TC_M));
    uint64_t grid = std::min<uint64_t>((uint64_t)nsm, items);
    if (grid < 1) grid = 1;
    ppp_prefilter_tc_kernel<WPL><<<(unsigned)grid, TC_THREADS, smem, st>>>(p);
    return cudaGetLastError();
}

/* One launch over the queries listed in qorder[0..nqk), all with an all-ones presence plane and (normally) an active
 * staircase; appends to the survivor list like ppp_launch_prefilter2.  Queries without a staircase are NOT handled here (the
 * caller routes launches that contain any to the POPC kernel). */
extern "C" cudaError_t ppp_launch_prefilter_tc(const SweepParams *prm, int wpl, const void *Tt, uint32_t ntt, const uint32_t *qorder, uint32_t nqk,
                                               uint32_t stair_cap, uint32_t *surv, unsigned long long *surv_count, unsigned long long surv_cap,
                                               int nsm, cudaStream_t st)
{
    if (!nqk) return cudaSuccess;
    TcParams p;
    p.Tt = reinterpret_cast<const uint4 *>(Tt);
    p.Q = prm->Q;
    p.qc = prm->qc;
    p.nmax = prm->nmax;
    p.nmax_off = prm->nmax_off;
    p.qorder = qorder;
    p.surv = surv;
    p.surv_count = surv_count;
    p.surv_cap = surv_cap;
    p.ntt = ntt;
    p.t0 = prm->t0;
    p.nTl = prm->t1 - prm->t0;
    p.nqk = nqk;
    p.stair_cap = stair_cap;
    switch (wpl) {
    case 1: return tc_launch_t<1>(p, nsm, st);
    case 2: return tc_launch_t<2>(p, nsm, st);
    case 4: return tc_launch_t<4>(p, nsm, st);
    case 8: return tc_launch_t<8>(p, nsm, st);
    default: return cudaErrorInvalidValue;
    }
}
