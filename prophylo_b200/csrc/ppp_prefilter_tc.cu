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
 * POPC kernel does.  A passing (target, query, step) is rare (under 1 % of the elements, but about one pair in ten has one
 * somewhere along its row), so it is pushed on a per-thread queue in shared memory and re-screened exactly -- word by word,
 * then candidate by candidate, from the packed words -- by the whole warp together (pooled drain, as in ppp_prefilter.cu);
 * only pairs with a passing CANDIDATE go to the warp-per-pair refine kernel (ppp_sweep.cu): the survivor list is exactly the
 * POPC path's.
 *
 * Two launches per target chunk, as in the POPC kernel.  PALL: queries whose presence plane is all ones (number == depth of
 * the target: D2 is a per-target prefix, which every epilogue thread keeps for its own row from the packed target words),
 * 128 queries per tile, accumulators P and C.  General presence planes: 64 queries per tile and three accumulators,
 *        Py = &stair_q[0] + 2 * yes_after,   Pn = 2 * number_after  (T . P),   Zc = 2 * popc(T & P & ~Y) in the chunk,
 * because number_before + yes_in_chunk = number_after - (number_in_chunk - yes_in_chunk): d = lds.s16 [Py] - Pn + Zc.
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
/* kind::i8 instruction descriptor: D = S32 (2 << 4), A/B = unsigned 8 bit K-major, N >> 3 at bit 17, M = 128 (8 << 24) */
#define TC_IDESC(N) (0x08000020u | (((uint32_t)(N) >> 3) << 17))
#define TC_NQ(PALL) ((PALL) ? 128 : 64)      /* queries per tile */
#define TC_WQ 512                            /* queue entries per epilogue warp: chunk-level passes of one (step, block) */

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
    uint32_t exact;           /* 1: chunk-level passes are re-screened exactly in the kernel (queue + pooled drain); 0: the kernel only
                                 writes, per (query tile, target), the mask of columns with a passing chunk (passmask) and the caller
                                 re-screens the flagged pairs (ppp_launch_rescreen_masks) */
    uint32_t *passmask;       /* exact == 0: [nQtiles][nTl][NQ / 32] column masks */
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
__device__ __forceinline__ void tc_mma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
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
#define TC_LD16(taddr, r)                                                                                                                    \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"       \
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),  \
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])                                              \
                 : "r"(taddr)                                                                                                               \
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

template <int WPL, bool PALL>
__global__ void __launch_bounds__(TC_THREADS, 1) ppp_prefilter_tc_kernel(const TcParams a)
{
    constexpr int W = 32 * WPL;
    constexpr int NCH = W / 4;         /* 128-genome chunks per row */
    constexpr int NSTEP = NCH + 3;     /* chunk 0 is screened in four 32-genome steps */
    constexpr int NQ = TC_NQ(PALL);    /* queries (accumulator columns) per tile */
    constexpr int NB = PALL ? 1 : 3;   /* B operand tiles per stage: 2Y | 2Y, 2P, 2(P & ~Y) */
    constexpr uint32_t B_TILE_BYTES = NQ * 128;
    constexpr uint32_t STAGE_BYTES = 2 * TC_TILE_BYTES + NB * B_TILE_BYTES;
    constexpr uint32_t IDESC = TC_IDESC(NQ);
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    uint64_t *bar_full = reinterpret_cast<uint64_t *>(smem_raw);   /* [STAGES] producers -> MMA */
    uint64_t *bar_empty = bar_full + TC_STAGES;                    /* [STAGES] MMA (commit) -> producers */
    uint64_t *bar_tfull = bar_empty + TC_STAGES;                   /* [2] MMA (commit) -> epilogue: accumulators of block b hold step s */
    uint64_t *bar_tfree = bar_tfull + 2;                           /* [2] epilogue -> MMA: accumulators of block b may be overwritten */
    uint32_t *s_tmem = reinterpret_cast<uint32_t *>(smem_raw + 128);
    uint32_t *s_qid = reinterpret_cast<uint32_t *>(smem_raw + 256);    /* [128] query id of column j (padding: 0xffffffff) */
    uint32_t *s_sbase = s_qid + 128;                                    /* [128] shared-memory ADDRESS of entry 0 of column j's doubled staircase */
    uint32_t *s_inact = s_sbase + 128;                                  /* [128] queries of the tile that accept every pair (no staircase) */
    uint32_t *s_ninact = s_inact + 128;
    uint32_t *s_passw = reinterpret_cast<uint32_t *>(smem_raw + 2048);  /* [2][4][TC_M] pass bits of row r of block b, 32 columns per word */
    uint2 *s_queue = reinterpret_cast<uint2 *>(smem_raw + 2048 + 2 * 4 * TC_M * 4); /* [EPI_WARPS][TC_WQ] */
    uint32_t *s_wqn = s_ninact + 4;                                                  /* [EPI_WARPS] entries queued */
    uint32_t *s_inactw = s_wqn + TC_EPI_WARPS;                                       /* [4] mask words of the columns without a staircase */
    unsigned char *s_stage = smem_raw + 2048 + 2 * 4 * TC_M * 4 + TC_EPI_WARPS * TC_WQ * 8;
    int16_t *s_stair = reinterpret_cast<int16_t *>(s_stage + TC_STAGES * STAGE_BYTES);
    const unsigned char *s_stair_b = reinterpret_cast<const unsigned char *>(s_stair);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t nPairs = (a.nTl + 2 * TC_M - 1) / (2 * TC_M);
    const uint32_t nQtiles = (a.nqk + NQ - 1) / NQ;
    const uint64_t nItems = (uint64_t)nQtiles * nPairs;
    const uint64_t span = (nItems + gridDim.x - 1) / gridDim.x;
    const uint64_t i0 = (uint64_t)blockIdx.x * span, i1 = min(nItems, i0 + span);

    if (tid < 4) s_inactw[tid] = 0;
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
    /* accumulators of block b (columns): PALL  P at 256 b, C at 256 b + 128;  general  Py at 256 b, Pn at 256 b + 64, Zc at 256 b + 128 */

    /* barrier phases, tracked per role: a wait on parity p returns once the barrier has completed its phase p */
    uint32_t ph_full = 0, ph_empty = 0;          /* bit s = parity of the next wait on stage s */
    uint32_t ph_tfull = 0, ph_tfree = 0;         /* bit b */
    uint32_t chunk_seq = 0;                      /* running chunk counter of this CTA: stage = chunk_seq % STAGES */
    uint32_t cur_qt = 0xffffffffu;

    for (uint64_t item = i0; item < i1; ++item) {
        const uint32_t qt = (uint32_t)(item / nPairs), pr = (uint32_t)(item % nPairs);
        if (qt != cur_qt) {
            /* ---- stage the query tile: ids and doubled staircases of its columns.  All warps: the roles meet here, and the
             * tensor core has finished with the previous tile (the epilogue waited for its last commit). */
            __syncthreads();
            const uint32_t q0 = qt * NQ;
            const uint32_t nq = min((uint32_t)NQ, a.nqk - q0);
            if (warp == 0) { /* offsets by a warp scan over four columns per lane */
                uint32_t len[4], qid[4], sum = 0;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint32_t col = 4 * lane + j;
                    qid[j] = col < nq ? a.qorder[q0 + col] : 0xffffffffu;
                    const uint32_t off = qid[j] != 0xffffffffu ? a.nmax_off[qid[j]] : PPP_NO_TABLE;
                    len[j] = (qid[j] != 0xffffffffu && off != PPP_NO_TABLE) ? a.qc[qid[j]].ny + 1 : 0;
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
                /* columns that accept every pair, as mask words (pair-level mode flags them for every live target) */
                uint32_t im = 0;
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (qid[j] != 0xffffffffu && !len[j]) im |= 1u << ((4 * lane + j) & 31);
                im |= __shfl_xor_sync(FULL, im, 1);
                im |= __shfl_xor_sync(FULL, im, 2);
                im |= __shfl_xor_sync(FULL, im, 4); /* lanes 8 w .. 8 w + 7 hold word w (columns 32 w .. 32 w + 31) */
                if ((lane & 7) == 0) s_inactw[lane >> 3] = im;
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
            const uint32_t col0 = half * (uint32_t)(NQ / 2);
            /* the running yes accumulators start at the staircase addresses of their columns */
#pragma unroll 1
            for (int b = 0; b < 2; ++b) {
#pragma unroll 1
                for (int g = 0; g < NQ / 64; ++g) {
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
            /* this thread's pass words (its row, its column half) */
            const uint32_t wfirst = PALL ? 2u * half : half, nwords = PALL ? 2u : 1u;
#pragma unroll
            for (int b = 0; b < 2; ++b)
                for (uint32_t w = 0; w < nwords; ++w) s_passw[(b * 4 + wfirst + w) * TC_M + row] = 0;
            /* the warp's queue of chunk-level passes of the current (step, block), re-screened exactly right after the step */
            uint2 *wq = s_queue + (uint32_t)warp * TC_WQ;
            uint32_t *wqn = s_wqn + warp;
            if (lane == 0) *wqn = 0;
            __syncwarp();
            const uint32_t stair_addr = smem_u32(s_stair);
            auto push = [&](uint32_t b, uint32_t col, uint32_t paddr, uint32_t na) {
                uint32_t *pw = &s_passw[(b * 4 + (col >> 5)) * TC_M + row];
                if ((*pw >> (col & 31)) & 1u) return; /* the pair is a survivor already */
                const uint32_t pos = atomicAdd(wqn, 1u);
                if (pos < TC_WQ) wq[pos] = make_uint2((paddr - stair_addr) | (na << 17), col | ((uint32_t)lane << 7));
                else atomicOr(pw, 1u << (col & 31)); /* queue full (a quarter of the warp's elements passed this step): accept unscreened */
            };
            /* packed words of this thread's target rows (both blocks) and of the query columns it serves to the drain
             * (column col0 + lane, and col0 + 32 + lane in the all-ones launch), for the CURRENT chunk; immutable global data */
            uint32_t trow[2];
            bool tlive[2];
#pragma unroll
            for (int b = 0; b < 2; ++b) {
                const uint32_t tl = tbase + b * TC_M + row;
                tlive[b] = tl < a.nTl;
                trow[b] = a.t0 + (tlive[b] ? tl : 0);
            }
            const uint32_t *qrow[2];
#pragma unroll
            for (int i = 0; i < (PALL ? 2 : 1); ++i) {
                const uint32_t qid = s_qid[col0 + 32u * i + lane];
                qrow[i] = a.Q + (size_t)(qid != 0xffffffffu ? qid : 0) * 2 * W;
            }
            uint4 tw[2], yq[2], pq = make_uint4(0, 0, 0, 0);
            auto load_chunk_words = [&](int c) {
#pragma unroll
                for (int b = 0; b < 2; ++b) tw[b] = tlive[b] ? __ldg(a.Tt + (size_t)c * a.ntt + trow[b]) : make_uint4(0, 0, 0, 0);
#pragma unroll
                for (int i = 0; i < (PALL ? 2 : 1); ++i) yq[i] = __ldg(reinterpret_cast<const uint4 *>(qrow[i] + W) + c);
                if (!PALL) pq = __ldg(reinterpret_cast<const uint4 *>(qrow[0]) + c);
            };
            yq[1] = make_uint4(0, 0, 0, 0);
            load_chunk_words(0);
            uint32_t run[2] = {0, 0}; /* all-ones launch: doubled depth (= number) of this thread's row so far */
            uint32_t accw[2][2] = {{0, 0}, {0, 0}}; /* pair-level mode: pass bits of this thread's columns, kept in registers */
            /* Exact re-screen of the queued (row, column) items of step s of block b by the whole warp: lane i takes the i-th,
             * (i + 32)-th ... entry whoever pushed it; the target words come from the owner lane and the query words from the
             * lane that serves the column (shuffles, no memory traffic); each item is re-counted word by word and then candidate by
             * candidate:  number at the candidate <= nmax[its rank]   (ppp_prefilter.cu's drain, ppp.pm:207-237) */
            auto drain = [&](int b, int s) {
                __syncwarp();
                const uint32_t n = min(*wqn, (uint32_t)TC_WQ);
                if (n == 0) return;
                const uint32_t wlo = s < 4 ? (uint32_t)s : 0u, whi = s < 4 ? (uint32_t)s : 3u;
                for (uint32_t base = 0; base < n; base += 32) {
                    const bool valid = base + lane < n;
                    const uint2 e = valid ? wq[base + lane] : make_uint2(0, 0);
                    const uint32_t col = e.y & 127u, own = (e.y >> 7) & 31u;
                    const uint32_t src = (col - col0) & 31u, second = (col - col0) >> 5;
                    uint32_t aw[4], bw[4];
                    aw[0] = __shfl_sync(FULL, tw[b].x, own); aw[1] = __shfl_sync(FULL, tw[b].y, own);
                    aw[2] = __shfl_sync(FULL, tw[b].z, own); aw[3] = __shfl_sync(FULL, tw[b].w, own);
                    {
                        const uint32_t y0 = __shfl_sync(FULL, yq[0].x, src), y1 = __shfl_sync(FULL, yq[0].y, src);
                        const uint32_t y2 = __shfl_sync(FULL, yq[0].z, src), y3 = __shfl_sync(FULL, yq[0].w, src);
                        bw[0] = y0; bw[1] = y1; bw[2] = y2; bw[3] = y3;
                        if (PALL) {
                            const uint32_t z0 = __shfl_sync(FULL, yq[1].x, src), z1 = __shfl_sync(FULL, yq[1].y, src);
                            const uint32_t z2 = __shfl_sync(FULL, yq[1].z, src), z3 = __shfl_sync(FULL, yq[1].w, src);
                            if (second) { bw[0] = z0; bw[1] = z1; bw[2] = z2; bw[3] = z3; }
                        } else {
                            aw[0] &= __shfl_sync(FULL, pq.x, src); aw[1] &= __shfl_sync(FULL, pq.y, src);
                            aw[2] &= __shfl_sync(FULL, pq.z, src); aw[3] &= __shfl_sync(FULL, pq.w, src);
                        }
                    }
                    if (valid) {
                        uint32_t cyw[4], cnw[4], cys = 0, cns = 0;
#pragma unroll
                        for (int w = 0; w < 4; ++w) {
                            bw[w] &= aw[w];
                            cyw[w] = 2u * __popc(bw[w]);
                            cnw[w] = 2u * __popc(aw[w]);
                            if ((uint32_t)w >= wlo && (uint32_t)w <= whi) {
                                cys += cyw[w];
                                cns += cnw[w];
                            }
                        }
                        uint32_t yb = (e.x & 0x1ffffu) - cys; /* byte offset of the staircase entry BEFORE the step */
                        uint32_t nb = (e.x >> 17) - cns;     /* doubled number before the step */
                        bool ok = false;
#pragma unroll
                        for (int w = 0; w < 4; ++w) {
                            if ((uint32_t)w >= wlo && (uint32_t)w <= whi) {
                                if (cyw[w] && (int)(nb + cyw[w]) <= (int)*reinterpret_cast<const int16_t *>(s_stair_b + yb + cyw[w])) {
                                    uint32_t yy = yb, bits = bw[w];
                                    while (bits) {
                                        const uint32_t bit = __ffs(bits) - 1;
                                        bits &= bits - 1;
                                        yy += 2;
                                        ok |= (int)(nb + 2u * __popc(aw[w] & ((2u << bit) - 1u))) <= (int)*reinterpret_cast<const int16_t *>(s_stair_b + yy);
                                    }
                                }
                                yb += cyw[w];
                                nb += cnw[w];
                            }
                        }
                        if (ok) atomicOr(&s_passw[((uint32_t)b * 4 + (col >> 5)) * TC_M + quarter * 32 + own], 1u << (col & 31));
                    }
                }
                __syncwarp();
                if (lane == 0) *wqn = 0;
                __syncwarp();
            };
#pragma unroll 1
            for (int s = 0; s < NSTEP; ++s) {
#pragma unroll
                for (int b = 0; b < 2; ++b) {
                    mbar_wait(&bar_tfull[b], (ph_tfull >> b) & 1u);
                    ph_tfull ^= 1u << b;
                    tc_fence_after();
                    const int kt = (int)run[b];
                    if (PALL) {
                        const uint32_t wsel = s == 0 ? tw[b].x : s == 1 ? tw[b].y : s == 2 ? tw[b].z : tw[b].w;
                        run[b] += s < 4 ? 2u * __popc(wsel) : 2u * (__popc(tw[b].x) + __popc(tw[b].y) + __popc(tw[b].z) + __popc(tw[b].w));
                    }
#pragma unroll
                    for (int g = 0; g < 2; ++g) {
                        uint32_t fl = 0;
                        if constexpr (PALL) {
                            uint32_t p[32], c32[32];
                            TC_LD32(tmem + lane_addr + (uint32_t)(b * 256) + col0 + 32u * g, p);
                            TC_LD32(tmem + lane_addr + (uint32_t)(b * 256 + 128) + col0 + 32u * g, c32);
                            tc_wait_ld();
                            if (g == 1 && s + 1 < NSTEP) { /* this step is in registers: the tensor core may go on with block b (after
                                                             the last step the next item's initialisation gives that signal) */
                                tc_fence_before();
                                mbar_arrive(&bar_tfree[b]);
                            }
#pragma unroll
                            for (int j = 0; j < 32; ++j) {
                                const int d = lds_s16(p[j]) - kt - (int)c32[j]; /* d < 0: no candidate of the chunk can pass */
                                fl = __funnelshift_l((uint32_t)d, fl, 1);       /* sign bits, column j at bit 31 - j */
                            }
                            if (!a.exact) {
                                accw[b][g] |= __brev(~fl); /* column j of the group at bit j */
                            } else if (~fl) { /* queue the passing columns for the exact screen */
#pragma unroll
                                for (int j = 0; j < 32; ++j)
                                    if (!((fl >> (31 - j)) & 1u)) push((uint32_t)b, col0 + 32u * g + j, p[j], run[b]);
                            }
                        } else {
                            uint32_t py[16], pn[16], zc[16];
                            TC_LD16(tmem + lane_addr + (uint32_t)(b * 256) + col0 + 16u * g, py);
                            TC_LD16(tmem + lane_addr + (uint32_t)(b * 256 + 64) + col0 + 16u * g, pn);
                            TC_LD16(tmem + lane_addr + (uint32_t)(b * 256 + 128) + col0 + 16u * g, zc);
                            tc_wait_ld();
                            if (g == 1 && s + 1 < NSTEP) {
                                tc_fence_before();
                                mbar_arrive(&bar_tfree[b]);
                            }
#pragma unroll
                            for (int j = 0; j < 16; ++j) {
                                /* number_before + yes_in_chunk = number_after - (number - yes) in the chunk */
                                const int d = lds_s16(py[j]) - (int)pn[j] + (int)zc[j];
                                fl = __funnelshift_l((uint32_t)d, fl, 1);
                            }
                            if (!a.exact) {
                                accw[b][0] |= (__brev(~fl) >> 16) << (16 * g); /* column j of the group at bit 16 g + j */
                            } else if (~fl & 0xffffu) {
#pragma unroll
                                for (int j = 0; j < 16; ++j)
                                    if (!((fl >> (15 - j)) & 1u)) push((uint32_t)b, col0 + 16u * g + j, py[j], pn[j]);
                            }
                        }
                    }
                    if (a.exact) drain(b, s);
                }
                if (s >= 3 && s + 1 < NSTEP) load_chunk_words(s - 2); /* the next step is a new chunk */
            }
            /* ---- survivors of the two blocks: (query of column, target) for every set pass bit; queries without a staircase
             * accept every live target (emitted once, by the warps of column half 0) */
            const uint32_t ninact = half == 0 ? *s_ninact : 0u;
            if (!a.exact) { /* pair-level mode: this row's column masks go to the re-screen kernel as they are */
#pragma unroll
                for (int b = 0; b < 2; ++b) {
                    const uint32_t tl = tbase + b * TC_M + row;
                    if (tl < a.nTl)
                        for (uint32_t w = 0; w < nwords; ++w)
                            a.passmask[((size_t)qt * a.nTl + tl) * (NQ / 32) + wfirst + w] = accw[b][w] | s_inactw[wfirst + w];
                }
            }
#pragma unroll 1
            for (int b = 0; a.exact && b < 2; ++b) {
                const uint32_t tl = tbase + b * TC_M + row;
                const bool live = tl < a.nTl;
                uint32_t pb[2] = {0, 0};
                for (uint32_t w = 0; w < nwords; ++w) pb[w] = live ? s_passw[(b * 4 + wfirst + w) * TC_M + row] : 0u;
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
                    for (int w = 0; w < 2; ++w) {
                        uint32_t m = pb[w];
                        while (m) {
                            const uint32_t bit = __ffs(m) - 1;
                            m &= m - 1;
                            const uint32_t col = 32u * (wfirst + w) + bit;
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
            /* =============================================== PRODUCERS: thread = row r of A0 and A1, and one or two B rows */
            const uint32_t r = (uint32_t)tid - TC_EPI_WARPS * 32;
            const uint32_t tl0 = tbase + r, tl1 = tbase + TC_M + r;
            const bool live0 = tl0 < a.nTl, live1 = tl1 < a.nTl;
            const uint32_t brow = PALL ? r : (r & 63u); /* the query column this thread expands */
            const uint32_t qid = s_qid[brow];
            const bool bq = qid != 0xffffffffu && a.nmax_off[qid] != PPP_NO_TABLE;
            const uint32_t *prow = a.Q + (size_t)(bq ? qid : 0) * 2 * W, *yrow = prow + W;
            uint32_t cs = chunk_seq;
#pragma unroll 1
            for (int c = 0; c < NCH; ++c, ++cs) {
                const uint32_t st = cs % TC_STAGES;
                const uint4 w0 = live0 ? __ldg(a.Tt + (size_t)c * a.ntt + a.t0 + tl0) : make_uint4(0, 0, 0, 0);
                const uint4 w1 = live1 ? __ldg(a.Tt + (size_t)c * a.ntt + a.t0 + tl1) : make_uint4(0, 0, 0, 0);
                const uint4 wy = bq ? __ldg(reinterpret_cast<const uint4 *>(yrow) + c) : make_uint4(0, 0, 0, 0);
                uint4 wp = make_uint4(0, 0, 0, 0);
                if (!PALL && bq) wp = __ldg(reinterpret_cast<const uint4 *>(prow) + c);
                if (cs >= TC_STAGES) { /* the stage's previous tenant has been consumed by the tensor core */
                    mbar_wait(&bar_empty[st], (ph_empty >> st) & 1u);
                    ph_empty ^= 1u << st;
                }
                unsigned char *stage = s_stage + st * STAGE_BYTES;
                expand_row<1>(stage, r, w0);
                expand_row<1>(stage + TC_TILE_BYTES, r, w1);
                unsigned char *bt = stage + 2 * TC_TILE_BYTES;
                if (PALL) {
                    expand_row<2>(bt, r, wy);
                } else if (r < 64) { /* 2 (Y & P) and 2 P */
                    expand_row<2>(bt, brow, make_uint4(wy.x & wp.x, wy.y & wp.y, wy.z & wp.z, wy.w & wp.w));
                    expand_row<2>(bt + B_TILE_BYTES, brow, wp);
                } else {             /* 2 (P & ~Y) */
                    expand_row<2>(bt + 2 * B_TILE_BYTES, brow, make_uint4(wp.x & ~wy.x, wp.y & ~wy.y, wp.z & ~wy.z, wp.w & ~wy.w));
                }
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
                const uint32_t sa = smem_u32(s_stage + st * STAGE_BYTES);
                const uint32_t sb = sa + 2 * TC_TILE_BYTES;
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
                            const uint32_t tb = tmem + (uint32_t)(b * 256);
                            if (PALL) {
                                tc_mma_i8(tb, da + koff, tc_smem_desc(sb) + koff, IDESC, 1u);               /* P += T . 2Y */
                                tc_mma_i8(tb + 128, da + koff, tc_smem_desc(sb) + koff, IDESC, k ? 1u : 0u); /* C  = / += */
                            } else {
                                tc_mma_i8(tb, da + koff, tc_smem_desc(sb) + koff, IDESC, 1u);                                        /* Py += T . 2(Y & P) */
                                tc_mma_i8(tb + 64, da + koff, tc_smem_desc(sb + B_TILE_BYTES) + koff, IDESC, (c | s | k) ? 1u : 0u); /* Pn (+)= T . 2P */
                                tc_mma_i8(tb + 128, da + koff, tc_smem_desc(sb + 2 * B_TILE_BYTES) + koff, IDESC, k ? 1u : 0u);      /* Zc = / += T . 2(P & ~Y) */
                            }
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
static size_t tc_smem_bytes(bool pall, uint32_t stair_cap)
{
    const size_t stage = 2 * (size_t)TC_TILE_BYTES + (pall ? 1 : 3) * (size_t)TC_NQ(pall) * 128;
    return 2048 + 2 * 4 * TC_M * 4 + (size_t)TC_EPI_WARPS * TC_WQ * 8 + (size_t)TC_STAGES * stage + (((size_t)stair_cap * 2 + 15) & ~(size_t)15) + 64;
}

extern "C" uint32_t ppp_prefilter_tc_tile_queries(int pall) { return TC_NQ(pall != 0); }

/* largest staircase footprint (int16 entries, dummy entry included) a query tile may have; the doubled entries are looked up
 * through 32-bit shared-memory addresses, so only shared memory bounds it */
extern "C" uint32_t ppp_prefilter_tc_max_stair(int wpl, int pall)
{
    (void)wpl;
    const size_t fixed = tc_smem_bytes(pall != 0, 0);
    const size_t limit = 226 * 1024;
    if (fixed + 64 > limit) return 0;
    return (uint32_t)std::min<size_t>((limit - fixed) / 2, 1u << 20);
}

template <int WPL, bool PALL>
static cudaError_t tc_launch_t(const TcParams &p, int nsm, cudaStream_t st)
{
    const size_t smem = tc_smem_bytes(PALL, p.stair_cap);
    cudaError_t e = cudaFuncSetAttribute(ppp_prefilter_tc_kernel<WPL, PALL>, cudaFuncAttributeMaxDynamicSharedMemorySize, PPP_MAX_DYN_SMEM);
    if (e != cudaSuccess) return e;
    const uint64_t items = (uint64_t)((p.nqk + TC_NQ(PALL) - 1) / TC_NQ(PALL)) * ((p.nTl + 2 * TC_M - 1) / (2 * TC_M));
    uint64_t grid = std::min<uint64_t>((uint64_t)nsm, items);
    if (grid < 1) grid = 1;
    ppp_prefilter_tc_kernel<WPL, PALL><<<(unsigned)grid, TC_THREADS, smem, st>>>(p);
    return cudaGetLastError();
}

/* One launch over the queries listed in qorder[0..nqk): pall != 0 -> every listed query has an all-ones presence plane.
 * Appends to the survivor list like ppp_launch_prefilter2 (*surv_count is not reset here). */
extern "C" cudaError_t ppp_launch_prefilter_tc(const SweepParams *prm, int wpl, int pall, const void *Tt, uint32_t ntt, const uint32_t *qorder,
                                               uint32_t nqk, uint32_t stair_cap, uint32_t *surv, unsigned long long *surv_count,
                                               unsigned long long surv_cap, int exact, uint32_t *passmask, int nsm, cudaStream_t st)
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
    p.exact = exact ? 1u : 0u;
    p.passmask = passmask;
    switch (wpl * 2 + (pall ? 1 : 0)) {
    case 2: return tc_launch_t<1, false>(p, nsm, st);
    case 3: return tc_launch_t<1, true>(p, nsm, st);
    case 4: return tc_launch_t<2, false>(p, nsm, st);
    case 5: return tc_launch_t<2, true>(p, nsm, st);
    case 8: return tc_launch_t<4, false>(p, nsm, st);
    case 9: return tc_launch_t<4, true>(p, nsm, st);
    case 16: return tc_launch_t<8, false>(p, nsm, st);
    case 17: return tc_launch_t<8, true>(p, nsm, st);
    default: return cudaErrorInvalidValue;
    }
}
