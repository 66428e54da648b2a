/*
 * prophylo_b200/csrc/ppp_pair.cuh -- device code shared by the warp-per-pair kernels (ppp_sweep.cu, ppp_ranked.cu):
 * TMA bulk-copy / mbarrier helpers, the FP64 series tier (tier B) and evaluate_pair(), the warp-wide evaluation of one
 * (query, target) pair = the body of PhyloProf::Algorithm::ppp::get_match_score (/root/reference/lib/PhyloProf/
 * Algorithm/ppp.pm:130-245) on bit words.  See ppp_sweep.cu for the algorithm overview.
 */
#pragma once
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

/* tier A (the float enclosure) lives in ppp_common.cuh: the staircase builder shares it. */

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
    const double r = up ? qc.theta : 1.0 / qc.theta;
    double s = 0.0;
    /* 128 terms a round, 4 consecutive terms per lane: the lane's first term from the log-factorial table (one exp),
     * the next three by the term ratio (m - i) r / (k + i + 1) in a division-free Horner form -- one exp and one
     * division per four terms instead of four exps */
    for (int i0 = 0; i0 <= m; i0 += 128) {
        const int i = i0 + 4 * lane;
        double t = 0.0, last = 0.0;
        if (i <= m) {
            const double t0 = exp(((base - lf[m - i]) - lf[k + i]) + (double)i * lnr);
            const double a1 = (double)max(m - i, 0) * r, b1 = (double)(k + i + 1);
            const double a2 = (double)max(m - i - 1, 0) * r, b2 = b1 + 1.0;
            const double a3 = (double)max(m - i - 2, 0) * r, b3 = b1 + 2.0;
            const double a12 = a1 * a2, b23 = b2 * b3;
            const double num = fma(b1, b23, fma(a1, b23, fma(a12, b3, a12 * a3))); /* b1b2b3 + a1b2b3 + a1a2b3 + a1a2a3 */
            const double inv = 1.0 / (b1 * b23);
            t = t0 * num * inv;
            last = t0 * (a12 * a3) * inv; /* the lane's smallest term (ratios decrease along the series) */
        }
        s += t;
        /* terms decrease monotonically beyond the mode, which lies at or before the series' start on both sides */
        if (!__any_sync(FULL, last > 1e-19)) break;
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

/* ---------------------------------------------------------------- one pair, one warp */
struct PairOut {
    double score;
    uint32_t yes, number, depth;
    float margin;
};
enum { PF_REJECTED = 1, PF_EXACT = 2, PF_UNDERFLOW = 4 };

struct WarpCounters {
    unsigned long long cand = 0, acc = 0, viol = 0, chk = 0, full = 0;
};

/* Evaluate one (query, target) pair with the whole warp.  Lane l owns words [l*WPL, (l+1)*WPL): tw = target words,
 * tp = tw & P, ty = tp & Y.  dpre / dtotal are the lane's exclusive depth prefix and the target's total popcount.
 * stair != nullptr: the query's pre-filter staircase (filtered runs); the pair is dropped (PF_REJECTED) unless some
 * candidate satisfies number <= stair[yes].  Returns flags; *out is valid unless rejected or flagged exact. */
template <int WPL>
__device__ __forceinline__ uint32_t evaluate_pair(const uint32_t (&tw)[WPL], const uint32_t (&tp)[WPL], const uint32_t (&ty)[WPL],
                                                  uint32_t dpre, uint32_t dtotal, const QConst &qc, const double *__restrict__ s_lf,
                                                  uint32_t *__restrict__ my_queue, uint32_t *__restrict__ my_y, float *__restrict__ my_lo,
                                                  const uint16_t *__restrict__ stair, const SweepParams &prm, int lane,
                                                  WarpCounters &wc, PairOut *out)
{
    constexpr uint32_t PPP_QCAP = PPP_QCAP_OF(WPL);
    PairOut res;
    res.score = 1.0; res.yes = 0; res.number = 0; res.depth = 0; res.margin = INFINITY;

    if (qc.flags & PPP_QF_DOMAIN) {
        /* p outside [0,1]: bdtrc -> 0.0 at the first element of the list (ppp.pm:225-241) */
        if (dtotal) {
            uint32_t first = 0, fp = 0, fy = 0;
            bool have = false;
#pragma unroll
            for (int w = WPL - 1; w >= 0; --w)
                if (tw[w]) { first = tw[w]; fp = tp[w]; fy = ty[w]; have = true; }
            const uint32_t has = __ballot_sync(FULL, have);
            const int src = __ffs(has) - 1;
            uint32_t y1 = 0, n1 = 0;
            if (lane == src) {
                const uint32_t b = __ffs(first) - 1;
                n1 = (fp >> b) & 1u;
                y1 = (fy >> b) & 1u;
            }
            res.score = 0.0;
            res.yes = __shfl_sync(FULL, y1, src);
            res.number = __shfl_sync(FULL, n1, src);
            res.depth = 1;
            res.margin = 0.f;
        }
        *out = res;
        return 0;
    }
    if (qc.flags & PPP_QF_TRIVIAL) {
        *out = res;
        return 0;
    }
    /* 1. counts + packed prefix scan (yes in the low 16 bits, number in the high 16 bits) */
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
    wc.cand += (lane == 0) ? E : 0;

    if (E == 0) {
        if (stair) return PF_REJECTED; /* score 1.0 cannot beat a threshold below 1/2 */
        *out = res;
        return 0;
    }
    if (prm.force_exact || (qc.flags & PPP_QF_EXACT)) {
        if (prm.bounds_only) { /* 1.0 is a valid upper bound of any score */
            *out = res;
            return 0;
        }
        return PF_EXACT;
    }
    if (prm.bounds_only) {
        /* threshold seeding needs only SOME valid upper bound of the pair's score: the minimum over a subset of the
         * candidates.  Each lane evaluates the LAST candidate inside its own words (rank = its inclusive yes count,
         * number counted up to that genome): one enclosure per lane instead of one per candidate. */
        float hi = INFINITY;
        if (cy) {
            uint32_t nn = excl >> 16, nlast = 0;
#pragma unroll
            for (int w = 0; w < WPL; ++w) {
                if (ty[w]) nlast = nn + __popc(tp[w] & ((2u << (31 - __clz(ty[w]))) - 1u));
                nn += __popc(tp[w]);
            }
            float lo;
            tier_a((int)(incl & 0xffffu), (int)nlast, qc, s_lf, lo, hi);
        }
        res.score = exp((double)warp_min_f(hi));
        *out = res;
        return 0;
    }

    /* 2. candidates.  Unfiltered: every set bit of T&Y.  Filtered: only those with number <= stair[yes] -- the
     * pair is only kept if its minimum is below the threshold, and then the minimum (and anything within the
     * staircase's margin of it, i.e. every possible near-tie) is among them. */
    uint32_t mycnt = 0;
    {
        uint32_t yrun = excl & 0xffffu, nbase = excl >> 16;
#pragma unroll
        for (int w = 0; w < WPL; ++w) {
            const uint32_t cyw = __popc(ty[w]);
            if (!stair) {
                mycnt += cyw;
            } else if (cyw && (nbase + 1u) <= (uint32_t)__ldg(stair + yrun + cyw)) { /* word-level screen, one lookup */
                uint32_t mbits = ty[w], yy = yrun;
                while (mbits) {
                    const uint32_t b = __ffs(mbits) - 1;
                    mbits &= mbits - 1;
                    ++yy;
                    const uint32_t nn = nbase + __popc(tp[w] & ((2u << b) - 1u));
                    mycnt += nn <= (uint32_t)__ldg(stair + yy);
                }
            }
            yrun += cyw;
            nbase += __popc(tp[w]);
        }
    }
    uint32_t cincl = mycnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(FULL, cincl, o);
        if (lane >= o) cincl += v;
    }
    const uint32_t cexcl = cincl - mycnt;
    const uint32_t EC = __shfl_sync(FULL, cincl, 31);
    if (stair && EC == 0) return PF_REJECTED; /* no candidate can reach the query's threshold */
    wc.full += (lane == 0);

    double best_lnf = INFINITY, best_f = 1.0, second_lnf = INFINITY;
    uint32_t best_y = 0, best_n = 0, best_g = 0;
    float Ustar = INFINITY;
    /* candidates are processed in passes of PPP_QCAP queue slots (one pass unless more than 512 candidates) */
    for (uint32_t qbase = 0; qbase < EC; qbase += PPP_QCAP) {
        {
            uint32_t yrun = excl & 0xffffu, nbase = excl >> 16, slot = cexcl - qbase;
#pragma unroll
            for (int w = 0; w < WPL; ++w) {
                uint32_t mbits = ty[w];
                const uint32_t cyw = __popc(mbits);
                if (stair && !(cyw && (nbase + 1u) <= (uint32_t)__ldg(stair + yrun + cyw))) mbits = 0; /* word-level screen */
                const uint32_t yword = yrun;
                yrun += cyw;
                uint32_t yy = yword;
                while (mbits) {
                    const uint32_t b = __ffs(mbits) - 1;
                    mbits &= mbits - 1;
                    ++yy;
                    const uint32_t nn = nbase + __popc(tp[w] & ((2u << b) - 1u));
                    if (!stair || nn <= (uint32_t)__ldg(stair + yy)) {
                        if (slot < PPP_QCAP) {
                            my_queue[slot] = nn | ((uint32_t)((lane * WPL + w) * 32 + b) << 16);
                            my_y[slot] = yy;
                        }
                        ++slot;
                    }
                }
                nbase += __popc(tp[w]);
            }
        }
        __syncwarp();
        const uint32_t cnt = min((uint32_t)PPP_QCAP, EC - qbase);
        /* 3. tier A enclosures */
        float umin = INFINITY;
        for (uint32_t j = lane; j < cnt; j += 32) {
            const uint32_t ent = my_queue[j];
            float lo, hi;
            tier_a((int)my_y[j], (int)(ent & 0xffffu), qc, s_lf, lo, hi);
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
                const int yy = (int)my_y[j0 + src], nn = (int)(ent & 0xffffu);
                double f;
                const double lnf = tier_b(yy, nn, qc, s_lf, lane, &f);
                wc.acc += (lane == 0);
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
                const double lnf = tier_b((int)my_y[j], (int)(ent & 0xffffu), qc, s_lf, lane, &f);
                float lo, hi;
                tier_a((int)my_y[j], (int)(ent & 0xffffu), qc, s_lf, lo, hi);
                wc.chk += (lane == 0);
                if (!((double)lo <= lnf && lnf <= (double)hi)) wc.viol += (lane == 0);
            }
        }
        __syncwarp();
    }
    /* 5. decide */
    const double margin = second_lnf - best_lnf;
    const double tie_eps = 1e-7; /* >> Cephes' and tier B's relative error on f */
    if (!(best_lnf < 0.0)) return PF_EXACT; /* nothing provably below 1: saturation at 1 - MACHEP is Cephes' call */
    if (margin < tie_eps || best_lnf < -700.0 || best_lnf > -1e-6)
        return PF_EXACT | (best_lnf < -700.0 ? PF_UNDERFLOW : 0);
    res.score = best_f;
    res.yes = best_y;
    res.number = best_n;
    res.margin = (float)margin;
    { /* depth = popc(T[0..g]): the lane that owns the word of g finishes it */
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
    *out = res;
    return 0;
}

/* -dup lists (ppp.pm:189-200, 222-224): the one re-counted taxon can make the running yes exceed |Y| = max_yes, at which
 * step the reference leaves the loop before evaluating.  That step is the LAST yes-increment of the list (every yes taxon
 * has been seen by then), so dropping the highest set bit of T&Y (list order: lane-major, word, bit) removes exactly it. */
template <int WPL>
__device__ __forceinline__ void clip_yes_overflow(uint32_t (&ty)[WPL], uint32_t ny, int lane)
{
    uint32_t c = 0;
#pragma unroll
    for (int w = 0; w < WPL; ++w) c += __popc(ty[w]);
    uint32_t tot = c;
#pragma unroll
    for (int o = 16; o; o >>= 1) tot += __shfl_xor_sync(FULL, tot, o);
    if (tot <= ny) return;
    const uint32_t holders = __ballot_sync(FULL, c != 0);
    if (lane == 31 - __clz(holders)) {
#pragma unroll
        for (int w = WPL - 1; w >= 0; --w) {
            if (ty[w]) {
                ty[w] &= ~(0x80000000u >> __clz(ty[w]));
                break;
            }
        }
    }
}

template <int WPL>
__device__ __forceinline__ void load_row(const uint32_t *__restrict__ row, uint32_t (&v)[WPL])
{
    if constexpr (WPL == 4) {
        const uint4 a = __ldg(reinterpret_cast<const uint4 *>(row));
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
    } else if constexpr (WPL == 8) {
        const uint4 a = __ldg(reinterpret_cast<const uint4 *>(row));
        const uint4 b = __ldg(reinterpret_cast<const uint4 *>(row) + 1);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
        v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else if constexpr (WPL == 2) {
        const uint2 a = __ldg(reinterpret_cast<const uint2 *>(row));
        v[0] = a.x; v[1] = a.y;
    } else {
        v[0] = __ldg(row);
    }
}

__device__ __forceinline__ void depth_prefix(uint32_t dlane, int lane, uint32_t &dpre, uint32_t &dtotal)
{
    uint32_t dincl = dlane;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(FULL, dincl, o);
        if (lane >= o) dincl += v;
    }
    dpre = dincl - dlane;
    dtotal = __shfl_sync(FULL, dincl, 31);
}

__device__ __forceinline__ void flush_counters(const SweepParams &prm, WarpCounters &wc, int lane)
{
    wc.cand = (unsigned long long)warp_sum_d((double)wc.cand);
    if (lane == 0) {
        if (wc.cand) atomicAdd(&prm.counters[CNT_CAND], wc.cand);
        if (wc.acc) atomicAdd(&prm.counters[CNT_ACCURATE], wc.acc);
        if (wc.viol) atomicAdd(&prm.counters[CNT_VIOL], wc.viol);
        if (wc.chk) atomicAdd(&prm.counters[CNT_CHECKS], wc.chk);
        if (wc.full) atomicAdd(&prm.counters[CNT_FULL], wc.full);
    }
}

__device__ __forceinline__ void push_exact(const SweepParams &prm, uint32_t local_id, bool underflow)
{
    const unsigned long long slot = atomicAdd(&prm.counters[CNT_EXACT], 1ull);
    if (slot < prm.exact_cap) prm.exact_list[slot] = local_id | (underflow ? 0x80000000u : 0u);
    else atomicAdd(&prm.counters[CNT_OVERFLOW], 1ull);
}

