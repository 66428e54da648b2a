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

#include "../../include/ppp_b200.h"
#include "ppp_common.cuh"

#define FULL 0xffffffffu

/* ---------------------------------------------------------------- TMA bulk copy + mbarrier (sm_90+/sm_100a) */
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!ok);
}
/* 1-D bulk copy global -> shared, completion on mbarrier; bytes % 16 == 0, both addresses 16-byte aligned */
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

/* ---------------------------------------------------------------- warp helpers */
__device__ __forceinline__ float warp_min_f(float v)
{
#pragma unroll
    for (int o = 16; o; o >>= 1) v = fminf(v, __shfl_xor_sync(FULL, v, o));
    return v;
}
__device__ __forceinline__ double warp_sum_d(double v)
{
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}

/* ---------------------------------------------------------------- tier A: float enclosure of ln P(X >= y | n, p)
 * Unified form: upper tail of Bin(n, p*) at k with m = n - k, ratio_i = (m - i + 1) r / (k + i), r = p* / (1 - p*):
 *   "up"    (y + 1 > (n+1) p):   p* = p, k = y,         lead = ln pmf_p(y; n),      ln f = lead + ln R
 *   "lower" (otherwise):         p* = q, k = n - y + 1, lead = ln pmf_p(y - 1; n),  ln f = ln(1 - exp(lead) R)
 * R = sum_{i>=0} prod_{j<=i} ratio_j.  After K terms the remainder t_K * ratio_{K+1} * R' is enclosed by
 *   R' <= 1 / (1 - ratio_{K+1})                       (ratios decrease)
 *   R' >= max(1, G'' z^2 / (1 + z^2))                 (Bahadur 1960; G'' = 1/(1 - ratio_{K+2}), z the standardised
 *                                                      distance of k + K + 1 from the mean n p*)
 * Division-free Horner keeps numerator and denominator products separately (k + K <= 8200 so k^8 < 2^127).
 */
__device__ __forceinline__ void tier_a(int y, int n, const QConst &qc, const double *__restrict__ lf, float &lo, float &hi)
{
    const int N_ = n - y;
    const bool up = (float)N_ * qc.thetaf < (float)(y + 1);
    const int k = up ? y : N_ + 1;
    const int m = n - k;
    const int kk = up ? y : y - 1;
    const float r = up ? qc.thetaf : qc.inv_thetaf;
    const double lead_d = ((lf[n] - lf[kk]) - lf[n - kk]) + ((double)kk * qc.lnp + (double)(n - kk) * qc.lnq);
    const float lead = (float)lead_d;

    float mf = (float)m, bf = (float)(k + 1);
    float tn = 1.f, D = 1.f, X = 1.f;
#pragma unroll
    for (int i = 0; i < PPP_K_TERMS; ++i) {
        const float a = fmaxf(mf, 0.f) * r;
        tn *= a;
        D *= bf;
        X = fmaf(X, bf, tn);
        mf -= 1.f;
        bf += 1.f;
    }
    const float invD = __fdividef(1.f, D);
    const float S = X * invD, tK = tn * invD;
    const float rho = __fdividef(fmaxf(mf, 0.f) * r, bf);              /* ratio_{K+1} */
    const float rho2 = __fdividef(fmaxf(mf - 1.f, 0.f) * r, bf + 1.f); /* ratio_{K+2} */
    float rem_lo = tK * rho;
    float rem_hi = (rho < 0.999f) ? __fdividef(rem_lo, 1.f - rho) : INFINITY;
    {
        const float pstar = up ? qc.pf : 1.f - qc.pf;
        const float d = bf - (float)n * pstar;
        if (d > 0.f && rho2 < 0.999f) {
            const float z2 = __fdividef(d * d, (float)n * qc.varf_unit);
            const float gb = __fdividef(z2, (1.f + z2) * (1.f - rho2));
            rem_lo *= fmaxf(1.f, gb * 0.9999f);
        }
    }
    const float Rlo = (S + rem_lo) * 0.9999f, Rhi = (S + rem_hi) * 1.0001f;
    if (up) {
        lo = lead + __logf(Rlo);
        hi = fminf(lead + __logf(Rhi), 0.f);
        lo -= 1e-4f + 2e-5f * fabsf(lo);
        hi = fminf(hi + 1e-4f + 2e-5f * fabsf(hi), 0.f);
    } else {
        const float e = __expf(lead);
        const float glo = e * Rlo * 0.999f, ghi = e * Rhi * 1.001f;
        hi = (glo < 0.5f) ? fminf(-glo, 0.f) : __logf(fmaxf(1.f - glo, 0.f)) + 1e-4f; /* ln(1-g) <= -g */
        hi = fminf(hi, 0.f);
        lo = (ghi < 0.9f) ? __logf(1.f - ghi) - 1e-4f : -INFINITY;
        if (ghi < 1e-3f) lo = -ghi * (1.f + ghi) - 1e-30f;
    }
}

/* ---------------------------------------------------------------- tier B: warp-cooperative FP64 evaluation
 * Sums the same ratio series to convergence, 32 terms per round: t_i = exp(lf[m]-lf[m-i] - (lf[k+i]-lf[k]) + i ln r).
 * Returns ln f; *f_out = f (computed as 1 - g on the lower side so that values near 1 keep their digits). */
__device__ __forceinline__ double tier_b(int y, int n, const QConst &qc, const double *__restrict__ lf, int lane, double *f_out)
{
    const int N_ = n - y;
    const bool up = (double)N_ * qc.theta < (double)(y + 1);
    const int k = up ? y : N_ + 1;
    const int m = n - k;
    const int kk = up ? y : y - 1;
    const double lnr = up ? qc.lntheta : -qc.lntheta;
    const double lead = ((lf[n] - lf[kk]) - lf[n - kk]) + ((double)kk * qc.lnp + (double)(n - kk) * qc.lnq);
    const double base = lf[m] + lf[k];
    double s = 0.0;
    for (int i0 = 0; i0 <= m; i0 += 32) {
        const int i = i0 + lane;
        double t = 0.0;
        if (i <= m) t = exp(((base - lf[m - i]) - lf[k + i]) + (double)i * lnr);
        s += t;
        if (!__any_sync(FULL, t > 1e-19)) break;
    }
    s = warp_sum_d(s);
    double lnf, f;
    if (up) {
        lnf = lead + log(s);
        f = exp(lnf);
    } else {
        const double g = exp(lead) * s;
        lnf = log1p(-g);
        f = 1.0 - g;
    }
    *f_out = f;
    return lnf;
}

/* ---------------------------------------------------------------- the sweep kernel */
struct PairOut {
    double score;
    uint32_t yes, number, depth;
    float margin;
};

template <int WPL>
__global__ void __launch_bounds__(512) ppp_sweep_kernel(const SweepParams prm)
{
    constexpr int W = 32 * WPL;  /* words per row */
    constexpr int GD = 32 * W;   /* padded genome count */
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int nwarps = blockDim.x >> 5;

    /* shared memory carve-up */
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw);                /* [0]: lf table, [1]: query tile */
    double *s_lf = reinterpret_cast<double *>(smem_raw + 128);              /* GD + 2 doubles */
    constexpr int LF_BYTES = (((GD + 2) * 8) + 127) & ~127;
    uint32_t *s_q = reinterpret_cast<uint32_t *>(smem_raw + 128 + LF_BYTES);    /* PPP_QTILE * 2 * W words */
    QConst *s_qc = reinterpret_cast<QConst *>(reinterpret_cast<unsigned char *>(s_q) + PPP_QTILE * 2 * W * 4);
    uint32_t *s_off = reinterpret_cast<uint32_t *>(reinterpret_cast<unsigned char *>(s_qc) + PPP_QTILE * sizeof(QConst)); /* [32] */
    uint32_t *s_queue = s_off + PPP_QTILE;
    uint32_t *my_queue = s_queue + warp * (2 * PPP_QCAP);                   /* [QCAP] (n | g<<16), then [QCAP] float lo */
    float *my_lo = reinterpret_cast<float *>(my_queue + PPP_QCAP);

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

    const uint32_t nTl = prm.t1 - prm.t0;                                  /* targets in this launch */
    const uint32_t nQtiles = (prm.nQ + PPP_QTILE - 1) / PPP_QTILE;
    const uint32_t nTblocks = (nTl + nwarps - 1) / nwarps;
    const uint64_t nItems = (uint64_t)nQtiles * nTblocks;

    unsigned long long cnt_cand = 0, cnt_acc = 0, cnt_viol = 0, cnt_chk = 0, cnt_full = 0;
    uint32_t tile_parity = 0;
    uint32_t cur_qt = 0xffffffffu;
    mbar_wait(&bars[0], 0);

    for (uint64_t item = blockIdx.x; item < nItems; item += gridDim.x) {
        const uint32_t qt = (uint32_t)(item / nTblocks);
        const uint32_t tb = (uint32_t)(item % nTblocks);
        const uint32_t q0 = qt * PPP_QTILE;
        const uint32_t nq = min((uint32_t)PPP_QTILE, prm.nQ - q0);
        if (qt != cur_qt) { /* block-uniform */
            __syncthreads(); /* everyone is done with the previous tile */
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
            if (prm.append) __syncthreads();
        }
        const uint32_t tl = tb * nwarps + warp;
        if (tl >= nTl) continue; /* warp-uniform */
        const uint32_t t = prm.t0 + tl;

        /* target row -> registers; depth offsets once per target */
        uint32_t tw[WPL];
        {
            const uint32_t *row = prm.T + (size_t)t * W + lane * WPL;
            if constexpr (WPL == 4) {
                const uint4 v = __ldg(reinterpret_cast<const uint4 *>(row));
                tw[0] = v.x; tw[1] = v.y; tw[2] = v.z; tw[3] = v.w;
            } else if constexpr (WPL == 8) {
                const uint4 v = __ldg(reinterpret_cast<const uint4 *>(row));
                const uint4 u = __ldg(reinterpret_cast<const uint4 *>(row) + 1);
                tw[0] = v.x; tw[1] = v.y; tw[2] = v.z; tw[3] = v.w;
                tw[4] = u.x; tw[5] = u.y; tw[6] = u.z; tw[7] = u.w;
            } else if constexpr (WPL == 2) {
                const uint2 v = __ldg(reinterpret_cast<const uint2 *>(row));
                tw[0] = v.x; tw[1] = v.y;
            } else {
                tw[0] = __ldg(row);
            }
        }
        uint32_t dlane = 0;
#pragma unroll
        for (int w = 0; w < WPL; ++w) dlane += __popc(tw[w]);
        uint32_t dincl = dlane;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t v = __shfl_up_sync(FULL, dincl, o);
            if (lane >= o) dincl += v;
        }
        const uint32_t dpre = dincl - dlane;
        const uint32_t dtotal = __shfl_sync(FULL, dincl, 31);

        PairOut mine;
        mine.score = 1.0; mine.yes = 0; mine.number = 0; mine.depth = 0; mine.margin = INFINITY;
        bool mine_exact = false, mine_rejected = false;

        for (uint32_t qi = 0; qi < nq; ++qi) {
            const QConst &qc = s_qc[qi];
            const uint32_t *pw_ = s_q + (size_t)qi * 2 * W + lane * WPL;
            const uint32_t *yw_ = pw_ + W;
            uint32_t tp[WPL], ty[WPL];
#pragma unroll
            for (int w = 0; w < WPL; ++w) {
                tp[w] = tw[w] & pw_[w];
                ty[w] = tp[w] & yw_[w];
            }
            PairOut res;
            res.score = 1.0; res.yes = 0; res.number = 0; res.depth = 0; res.margin = INFINITY;
            bool need_exact = false, rejected = false;

            if (qc.flags & PPP_QF_DOMAIN) {
                /* p outside [0,1]: bdtrc -> 0.0 at the first element of the list (ppp.pm:225-241) */
                if (dtotal) {
                    uint32_t firstword = 0xffffffffu;
#pragma unroll
                    for (int w = WPL - 1; w >= 0; --w) if (tw[w]) firstword = w;
                    const uint32_t has = __ballot_sync(FULL, firstword != 0xffffffffu);
                    const int src = __ffs(has) - 1;
                    uint32_t y1 = 0, n1 = 0;
                    if (lane == src) {
                        const uint32_t b = __ffs(tw[firstword]) - 1;
                        n1 = (tp[firstword] >> b) & 1u;
                        y1 = (ty[firstword] >> b) & 1u;
                    }
                    res.score = 0.0;
                    res.yes = __shfl_sync(FULL, y1, src);
                    res.number = __shfl_sync(FULL, n1, src);
                    res.depth = 1;
                    res.margin = 0.f;
                }
            } else if (!(qc.flags & PPP_QF_TRIVIAL)) {
                /* 1. counts + packed prefix scan (yes in low 16 bits, number in high 16 bits) */
                uint32_t cy = 0, cn = 0;
#pragma unroll
                for (int w = 0; w < WPL; ++w) {
                    cy += __popc(ty[w]);
                    cn += __popc(tp[w]);
                }
                const uint32_t packed = cy | (cn << 16);
                uint32_t incl = packed;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t v = __shfl_up_sync(FULL, incl, o);
                    if (lane >= o) incl += v;
                }
                const uint32_t excl = incl - packed;
                const uint32_t E = __shfl_sync(FULL, incl, 31) & 0xffffu;
                cnt_cand += (lane == 0) ? E : 0;

                /* filtered runs: the staircase nmax[yes] bounds the number a candidate may have and still pass the
                 * query's threshold (tail increases with number, decreases with yes); one lookup per lane decides
                 * whether ANY candidate of the lane's words can pass.  Most pairs end here. */
                if (prm.append) {
                    const uint32_t off = s_off[qi];
                    if (off != PPP_NO_TABLE) {
                        const uint16_t *stair = prm.nmax + off;
                        const bool lane_pass = cy > 0 && ((excl >> 16) + 1u) <= (uint32_t)__ldg(stair + (excl & 0xffffu) + cy);
                        rejected = !__any_sync(FULL, lane_pass);
                        if (!rejected) {
                            /* second look, exact per candidate: number <= nmax[yes] for at least one set bit of T&Y */
                            bool cand_pass = false;
                            if (lane_pass) {
                                uint32_t yrun = excl & 0xffffu, nbase = excl >> 16;
#pragma unroll
                                for (int w = 0; w < WPL; ++w) {
                                    uint32_t mbits = ty[w];
                                    while (mbits) {
                                        const uint32_t b = __ffs(mbits) - 1;
                                        mbits &= mbits - 1;
                                        ++yrun;
                                        const uint32_t nn = nbase + __popc(tp[w] & ((2u << b) - 1u));
                                        cand_pass |= nn <= (uint32_t)__ldg(stair + yrun);
                                    }
                                    nbase += __popc(tp[w]);
                                }
                            }
                            rejected = !__any_sync(FULL, cand_pass);
                        }
                    }
                }
                if (rejected) {
                } else if (E > 0 && (prm.force_exact || (qc.flags & PPP_QF_EXACT))) {
                    need_exact = true;
                } else if (E > 0) {
                    double best_lnf = INFINITY, best_f = 1.0, second_lnf = INFINITY;
                    uint32_t best_y = 0, best_n = 0, best_g = 0;
                    float Ustar = INFINITY;
                    /* candidates are processed in passes of PPP_QCAP ranks (one pass unless |T&Y| > 1024) */
                    for (uint32_t qbase = 0; qbase < E; qbase += PPP_QCAP) {
                        /* 2. enqueue: rank (= yes) is the queue slot */
                        {
                            uint32_t yrun = excl & 0xffffu, nbase = excl >> 16;
#pragma unroll
                            for (int w = 0; w < WPL; ++w) {
                                uint32_t mbits = ty[w];
                                while (mbits) {
                                    const uint32_t b = __ffs(mbits) - 1;
                                    mbits &= mbits - 1;
                                    const uint32_t nn = nbase + __popc(tp[w] & ((2u << b) - 1u));
                                    const uint32_t slot = yrun - qbase;
                                    if (slot < PPP_QCAP) my_queue[slot] = nn | ((uint32_t)((lane * WPL + w) * 32 + b) << 16);
                                    ++yrun;
                                }
                                nbase += __popc(tp[w]);
                            }
                        }
                        __syncwarp();
                        const uint32_t cnt = min((uint32_t)PPP_QCAP, E - qbase);
                        /* 3. tier A enclosures */
                        float umin = INFINITY;
                        for (uint32_t j = lane; j < cnt; j += 32) {
                            const uint32_t ent = my_queue[j];
                            float lo, hi;
                            tier_a((int)(qbase + j + 1), (int)(ent & 0xffffu), qc, s_lf, lo, hi);
                            my_lo[j] = lo;
                            umin = fminf(umin, hi);
                        }
                        Ustar = fminf(Ustar, warp_min_f(umin));
                        __syncwarp();
                        /* 4. survivors -> tier B (in rank order, so that the earliest minimum wins) */
                        const float cut = Ustar + 1e-5f + 1e-6f * fabsf(Ustar);
                        for (uint32_t j0 = 0; j0 < cnt; j0 += 32) {
                            const uint32_t j = j0 + lane;
                            const bool sv = (j < cnt) && !(my_lo[j] > cut); /* NaN-safe: unknown = survivor */
                            uint32_t ball = __ballot_sync(FULL, sv);
                            while (ball) {
                                const int src = __ffs(ball) - 1;
                                ball &= ball - 1;
                                const uint32_t ent = my_queue[j0 + src];
                                const int yy = (int)(qbase + j0 + src + 1), nn = (int)(ent & 0xffffu);
                                double f;
                                const double lnf = tier_b(yy, nn, qc, s_lf, lane, &f);
                                cnt_acc += (lane == 0);
                                if (lnf < best_lnf) {
                                    second_lnf = best_lnf;
                                    best_lnf = lnf; best_f = f;
                                    best_y = yy; best_n = nn; best_g = ent >> 16;
                                } else if (lnf < second_lnf) {
                                    second_lnf = lnf;
                                }
                            }
                        }
                        if (prm.check_bounds) { /* diagnostic: every enclosure must contain the accurate value */
                            for (uint32_t j = 0; j < cnt; ++j) {
                                const uint32_t ent = my_queue[j];
                                double f;
                                const double lnf = tier_b((int)(qbase + j + 1), (int)(ent & 0xffffu), qc, s_lf, lane, &f);
                                float lo, hi;
                                tier_a((int)(qbase + j + 1), (int)(ent & 0xffffu), qc, s_lf, lo, hi);
                                cnt_chk += (lane == 0);
                                if (!((double)lo <= lnf && lnf <= (double)hi)) cnt_viol += (lane == 0);
                            }
                        }
                        __syncwarp();
                    }
                    /* 5. decide */
                    const double margin = second_lnf - best_lnf;
                    const double tie_eps = 1e-7; /* >> Cephes' and tier B's relative error on f */
                    if (!(best_lnf < 0.0)) {
                        /* nothing provably below 1: leave it to the exact tier (saturation at 1 - MACHEP) */
                        need_exact = true;
                    } else if (margin < tie_eps || best_lnf < -700.0 || best_lnf > -1e-6) {
                        need_exact = true;
                    } else {
                        res.score = best_f;
                        res.yes = best_y;
                        res.number = best_n;
                        res.margin = (float)margin;
                        /* depth = popc(T[0..g]) : the lane that owns the word of g finishes it */
                        const int owner = (int)(best_g >> 5) / WPL;
                        uint32_t dd = 0;
                        if (lane == owner) {
                            dd = dpre;
                            const int wsel = (int)(best_g >> 5) % WPL;
#pragma unroll
                            for (int w = 0; w < WPL; ++w) {
                                if (w < wsel) dd += __popc(tw[w]);
                                if (w == wsel) dd += __popc(tw[w] & ((2u << (best_g & 31)) - 1u));
                            }
                        }
                        res.depth = __shfl_sync(FULL, dd, owner);
                    }
                }
            }
            if (lane == (int)qi) {
                mine = res;
                mine_exact = need_exact;
                mine_rejected = rejected;
            }
        }
        /* lane i holds the result of query q0 + i */
        if ((uint32_t)lane < nq && !mine_rejected) {
            const uint32_t q = q0 + lane;
            if (mine_exact) {
                const unsigned long long slot = atomicAdd(&prm.counters[CNT_EXACT], 1ull);
                if (slot < prm.exact_cap) prm.exact_list[slot] = (uint32_t)((uint64_t)q * nTl + tl);
                else atomicAdd(&prm.counters[CNT_OVERFLOW], 1ull);
                if (!prm.append) ppp_emit(prm, q, t, mine.score, mine.yes, mine.number, mine.depth, -1.f);
            } else {
                ppp_emit(prm, q, t, mine.score, mine.yes, mine.number, mine.depth, mine.margin);
            }
        }
        if (prm.append) cnt_full += __popc(__ballot_sync(FULL, (uint32_t)lane < nq && !mine_rejected));
    }
    /* flush per-warp statistics */
    cnt_cand = (unsigned long long)warp_sum_d((double)cnt_cand);
    if (lane == 0) {
        if (cnt_cand) atomicAdd(&prm.counters[CNT_CAND], cnt_cand);
        if (cnt_acc) atomicAdd(&prm.counters[CNT_ACCURATE], cnt_acc);
        if (cnt_viol) atomicAdd(&prm.counters[CNT_VIOL], cnt_viol);
        if (cnt_chk) atomicAdd(&prm.counters[CNT_CHECKS], cnt_chk);
        if (cnt_full) atomicAdd(&prm.counters[CNT_FULL], cnt_full);
    }
}

/* ---------------------------------------------------------------- launcher */
static size_t sweep_smem_bytes(int wpl, int nwarps)
{
    const size_t W = 32 * (size_t)wpl, GD = 32 * W;
    return 128 + ((((GD + 2) * 8) + 127) & ~(size_t)127) + PPP_QTILE * 2 * W * 4 + PPP_QTILE * sizeof(QConst) + PPP_QTILE * 4 + (size_t)nwarps * 2 * PPP_QCAP * 4 + 64;
}

template <int WPL>
static cudaError_t launch_t(const SweepParams &prm, int nwarps, int nsm, cudaStream_t st)
{
    const size_t smem = sweep_smem_bytes(WPL, nwarps);
    cudaError_t e = cudaFuncSetAttribute(ppp_sweep_kernel<WPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int per_sm = 1;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ppp_sweep_kernel<WPL>, nwarps * 32, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    const uint32_t nTl = prm.t1 - prm.t0;
    const uint64_t items = (uint64_t)((prm.nQ + PPP_QTILE - 1) / PPP_QTILE) * ((nTl + nwarps - 1) / nwarps);
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
