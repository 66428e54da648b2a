/*
 * prophylo_b200/csrc/ppp_ranked.cu -- warp-per-pair sweep for RANKED targets (ppp_set_targets_ranked).
 *
 * A ranked target is what BlastResult::get_profile / the value-sorted HMMERResult hash hand to the reference loop
 * (/root/reference/lib/PhyloProf/BlastResult.pm:192-237, lib/PhyloProf/Algorithm/ppp.pm:147-153): an ORDERED list of
 * taxa, best hit first, in an order of its own.  The host de-duplicates it (ppp.pm:189-192) and maps taxa to genome
 * indices of the query universe (0xFFFF = taxon foreign to every query); the raw 1-based position of each kept entry
 * travels along so that $best_depth counts the skipped duplicates exactly as the reference does (ppp.pm:169,231).
 *
 * In the target's own order the target plane is a run of ones and the query planes are gathered through the list:
 * for 32 list positions at a time every lane tests its position's genome bit in the query's P and Y planes (shared
 * memory) and two ballots produce the words  T&P  and  T&Y  in LIST order -- after which the pair is exactly the
 * bit-plane problem and evaluate_pair() is reused unchanged.  One warp owns one target (its list indices stay in
 * registers) and walks a 32-query tile staged by TMA.
 */
#include "ppp_pair.cuh"

template <int WPL>
__global__ void __launch_bounds__(256) ppp_sweep_ranked_kernel(const SweepParams prm)
{
    constexpr int NW = 32 * WPL; /* list words: up to 1024 * WPL positions */
    constexpr uint32_t PPP_QCAP = PPP_QCAP_OF(WPL);
    constexpr int GD = 32 * NW;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const uint32_t Wg = prm.qwords;

    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw);
    double *s_lf = reinterpret_cast<double *>(smem_raw + 128);
    constexpr int LF_BYTES = (((GD + 2) * 8) + 127) & ~127;
    uint32_t *s_q = reinterpret_cast<uint32_t *>(smem_raw + 128 + LF_BYTES); /* PPP_QTILE * 2 * Wg words */
    QConst *s_qc = reinterpret_cast<QConst *>(reinterpret_cast<unsigned char *>(s_q) + (size_t)PPP_QTILE * 2 * Wg * 4);
    uint32_t *s_off = reinterpret_cast<uint32_t *>(reinterpret_cast<unsigned char *>(s_qc) + PPP_QTILE * sizeof(QConst));
    uint32_t *s_queue = s_off + PPP_QTILE;
    uint32_t *my_queue = s_queue + warp * (3 * PPP_QCAP);
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
    const uint32_t nTl = prm.t1 - prm.t0;
    const uint32_t nQtiles = (prm.nQ + PPP_QTILE - 1) / PPP_QTILE;
    const uint32_t nSB = (nTl + PPP_SB - 1) / PPP_SB;
    const uint64_t nItems = (uint64_t)nQtiles * nSB;
    uint32_t *s_next = reinterpret_cast<uint32_t *>(smem_raw + 64);
    unsigned long long *s_item = reinterpret_cast<unsigned long long *>(smem_raw + 72);
    (void)nwarps;
    WarpCounters wc;
    uint32_t tile_parity = 0, cur_qt = 0xffffffffu;
    mbar_wait(&bars[0], 0);

    for (;;) { /* items from a global counter, targets of an item from a shared one (see ppp_sweep.cu) */
        __syncthreads();
        if (threadIdx.x == 0) *s_item = atomicAdd(&prm.counters[CNT_ITEM], 1ull);
        __syncthreads();
        const uint64_t item = *s_item;
        if (item >= nItems) break;
        const uint32_t qt = (uint32_t)(item / nSB), sb = (uint32_t)(item % nSB);
        const uint32_t q0 = qt * PPP_QTILE;
        const uint32_t nq = min((uint32_t)PPP_QTILE, prm.nQ - q0);
        if (threadIdx.x == 0) *s_next = 0;
        if (qt != cur_qt) {
            if (threadIdx.x == 0) {
                fence_proxy_async();
                const uint32_t b1 = nq * 2 * Wg * 4, b2 = nq * (uint32_t)sizeof(QConst);
                mbar_expect_tx(&bars[1], b1 + b2);
                tma_load_1d(s_q, prm.Q + (size_t)q0 * 2 * Wg, b1, &bars[1]);
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
      for (;;) {
        uint32_t pick = 0;
        if (lane == 0) pick = atomicAdd(s_next, 1u);
        pick = __shfl_sync(FULL, pick, 0);
        if (pick >= sb_count) break;
        const uint32_t tl = sb * PPP_SB + pick;
        const uint32_t t = prm.t0 + tl;
        const unsigned long long off0 = prm.roff[t];
        const uint32_t L = (uint32_t)(prm.roff[t + 1] - off0);
        /* this lane's list entries: position j*32 + lane for every list word j */
        uint16_t gi[NW];
#pragma unroll
        for (int j = 0; j < NW; ++j) {
            const uint32_t pos = (uint32_t)j * 32u + lane;
            gi[j] = (pos < L) ? __ldg(prm.ridx + off0 + pos) : (uint16_t)0xFFFF;
        }
        /* target plane in list order: ones for positions < L; depth prefix is the position itself */
        uint32_t tw[WPL];
#pragma unroll
        for (int w = 0; w < WPL; ++w) {
            const uint32_t base = (uint32_t)(lane * WPL + w) * 32u;
            tw[w] = (base + 32u <= L) ? 0xffffffffu : (base < L ? ((1u << (L - base)) - 1u) : 0u);
        }
        const uint32_t dpre = min(L, (uint32_t)(lane * WPL) * 32u), dtotal = L;

        PairOut mine;
        mine.score = 1.0; mine.yes = 0; mine.number = 0; mine.depth = 0; mine.margin = INFINITY;
        uint32_t mine_flags = 0;
        for (uint32_t qi = 0; qi < nq; ++qi) {
            const uint32_t *pq = s_q + (size_t)qi * 2 * Wg, *yq = pq + Wg;
            uint32_t tp[WPL], ty[WPL];
#pragma unroll
            for (int w = 0; w < WPL; ++w) tp[w] = ty[w] = 0;
#pragma unroll
            for (int j = 0; j < NW; ++j) {
                if ((uint32_t)j * 32u >= L) break; /* warp-uniform */
                const uint32_t g = gi[j];
                bool inp = false, iny = false;
                if (g != 0xFFFFu) {
                    const uint32_t sh = g & 31u, wi = g >> 5;
                    inp = (pq[wi] >> sh) & 1u;
                    iny = inp && ((yq[wi] >> sh) & 1u);
                }
                const uint32_t wp = __ballot_sync(FULL, inp), wy = __ballot_sync(FULL, iny);
                if (lane == j / WPL) {
                    tp[j % WPL] = wp;
                    ty[j % WPL] = wy;
                }
            }
            if (prm.dup) clip_yes_overflow<WPL>(ty, s_qc[qi].ny, lane);
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
        if ((uint32_t)lane < nq && !(mine_flags & PF_REJECTED)) {
            const uint32_t q = q0 + lane;
            if (mine_flags & PF_EXACT) {
                push_exact(prm, (uint32_t)((uint64_t)q * nTl + tl), (mine_flags & PF_UNDERFLOW) != 0);
                if (!prm.append) ppp_emit(prm, q, t, 1.0, 0, 0, 0, -1.f);
            } else {
                /* de-duplicated depth -> raw list position (ppp.pm:231 `$i + 1`) */
                const uint32_t depth = mine.depth ? (uint32_t)__ldg(prm.rraw + off0 + mine.depth - 1) : 0u;
                ppp_emit(prm, q, t, mine.score, mine.yes, mine.number, depth, mine.margin);
            }
        }
      } /* targets of the item */
    }
    flush_counters(prm, wc, lane);
}

static size_t ranked_smem_bytes(int wpl, int nwarps, uint32_t Wg)
{
    const size_t NW = 32 * (size_t)wpl, GD = 32 * NW;
    return 128 + ((((GD + 2) * 8) + 127) & ~(size_t)127) + (size_t)PPP_QTILE * 2 * Wg * 4 + PPP_QTILE * sizeof(QConst) + PPP_QTILE * 4 +
           (size_t)nwarps * 3 * PPP_QCAP_OF(wpl) * 4 + 64;
}

template <int WPL>
static cudaError_t launch_ranked_t(const SweepParams &prm, int nwarps, int nsm, cudaStream_t st)
{
    if (nwarps > 8) nwarps = 8;
    const size_t smem = ranked_smem_bytes(WPL, nwarps, prm.qwords);
    cudaError_t e = cudaFuncSetAttribute(ppp_sweep_ranked_kernel<WPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, PPP_MAX_DYN_SMEM);
    if (e != cudaSuccess) return e;
    int per_sm = 1;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ppp_sweep_ranked_kernel<WPL>, nwarps * 32, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    const uint32_t nTl = prm.t1 - prm.t0;
    const uint64_t items = (uint64_t)((prm.nQ + PPP_QTILE - 1) / PPP_QTILE) * ((nTl + PPP_SB - 1) / PPP_SB);
    uint64_t grid = (uint64_t)nsm * per_sm;
    if (grid > items) grid = items;
    if (grid < 1) grid = 1;
    ppp_sweep_ranked_kernel<WPL><<<(unsigned)grid, nwarps * 32, smem, st>>>(prm);
    return cudaGetLastError();
}

/* rwpl = list words per lane: 1 (lists up to 1024 entries) or 2 (up to 2048) */
extern "C" cudaError_t ppp_launch_sweep_ranked(const SweepParams *prm, int rwpl, int nwarps, int nsm, cudaStream_t st)
{
    switch (rwpl) {
    case 1: return launch_ranked_t<1>(*prm, nwarps, nsm, st);
    case 2: return launch_ranked_t<2>(*prm, nwarps, nsm, st);
    default: return cudaErrorInvalidValue;
    }
}
