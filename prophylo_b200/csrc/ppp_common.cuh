/*
 * prophylo_b200/csrc/ppp_common.cuh -- device/host shared layouts of the PPP scan.
 *
 * HBM layout (all little-endian, genome g = bit (g & 31) of word (g >> 5), canonical genome order):
 *   targets   T[nT][W]            uint32, W = 32 * WPL words per row (row = 1024*WPL bits, zero padded)
 *   queries   Q[nQ][2][W]         uint32, plane 0 = presence P, plane 1 = yes Y (Y subset of P); the two planes
 *                                 of a query are adjacent so that a tile of queries is ONE contiguous TMA bulk copy
 *   qconst    QConst[nQ]          per-query constants (64 B), computed on the host in long double
 *   lf        double[Gd + 2]      log-factorial table lf[k] = lgamma(k+1), k = 0..Gd+1
 *   hits      ppp_hit[...]        32-byte records (include/ppp_b200.h)
 */
#pragma once
#include <stdint.h>

#define PPP_QTILE 32      /* queries per shared-memory tile = lanes of a warp (lane i keeps query i's result) */
#define PPP_SB 128       /* targets per work item of the tile kernels (warps draw them dynamically) */
#define PPP_QCAP_MAX 512  /* candidates per warp queue pass (G > 1024) */
/* G <= 1024 (WPL == 1): smaller queues and 64 registers fit two CTAs per SM, which pays for latency-bound sweeps */
#define PPP_QCAP_OF(WPL) ((WPL) == 1 ? 256 : PPP_QCAP_MAX)
#define PPP_MINBLOCKS_OF(WPL) ((WPL) == 1 ? 2 : 1)
#define PPP_K_TERMS 8     /* series terms of the float enclosure tier */
/* Per-CTA dynamic shared-memory limit of sm_100.  Launchers opt every kernel in to THIS constant, not to the launch's own
 * size: the attribute is per function and device, and two contexts on one device (a device group that lists a GPU twice)
 * launch the same kernel with different sizes from different host threads -- one thread lowering the attribute under the
 * other's launch makes that launch fail with "invalid argument". */
#define PPP_MAX_DYN_SMEM (227 * 1024)

/* per-query flags */
#define PPP_QF_TRIVIAL 1u /* p == 0, p == 1 or NaN: bdtrc never improves on 1.0 -> (1,0,0,0)             */
#define PPP_QF_DOMAIN 2u  /* p < 0 or p > 1: Cephes "domain error" -> 0.0 at the first list element       */
#define PPP_QF_EXACT 4u   /* p so close to 0 or 1 that only the exact tier is trusted                     */

struct __attribute__((aligned(64))) QConst {
    double p;        /* verbatim */
    double lnp;      /* ln p */
    double lnq;      /* ln (1-p) */
    double lntheta;  /* ln (p / (1-p)) */
    double theta;    /* p / (1-p) */
    float thetaf;    /* float copies for the enclosure tier */
    float inv_thetaf;
    float pf;        /* p */
    float varf_unit; /* p (1-p) */
    uint32_t flags;
    uint32_t ny;     /* |Y| (Profile::get_yes, ppp.pm:136) */
    uint32_t npres;  /* |P| (Profile::get_total, Profile.pm:342-354) */
    uint32_t pad_;
};

/* pair flags carried in the top bits of the packed result while a pair waits for the exact tier */
struct SweepParams {
    const uint32_t *T;     /* [nT][W] */
    const uint32_t *Q;     /* [nQ][2][W] */
    const QConst *qc;      /* [nQ] */
    const double *lf;      /* [Gd+2] */
    void *hits;            /* ppp_hit[nQ*nT] (dense mode) */
    uint32_t *exact_list;  /* pair ids (q * nT + t) flagged for the exact tier */
    unsigned long long *counters; /* [0]=exact_count [1]=candidates [2]=accurate_evals [3]=bound_violations [4]=bound_checks */
    uint64_t exact_cap;
    uint32_t nT, nQ;
    uint32_t t0, t1;       /* target range of this launch */
    uint32_t G;            /* real genome count */
    uint32_t force_exact;
    uint32_t check_bounds;
    /* filtered runs (PPP_FILTER_TOPK / PPP_FILTER_THRESH_LOG10) */
    uint32_t append;           /* 0: dense hits[q*nT+t]; 1: per-query append hits[q*qcap + qcnt[q]++]; 2: global append;
                                  3: (bounds_only) dense matrix of provisional scores, double[nQ][t1 - t0] */
    uint32_t accept_mode;      /* 0: score < kth[q] (top-k, kth = 2.0 while the list is not full); 1: log10(score) < thresh;
                                  2: klo[q] < score <= kth[q] (re-scan of queries whose speculative threshold failed);
                                  3: score <= kth[q] (top-k launch screened against a SPECULATIVE threshold: ties at the
                                     threshold are kept, so the verification may accept a k-th score equal to it) */
    double thresh_log10;
    const double *kth;         /* [nQ] */
    const double *klo;         /* [nQ] accept_mode 2 */
    const uint16_t *nmax;      /* prefilter staircase: candidate (yes, number) can pass only if number <= nmax[off + yes] */
    const uint32_t *nmax_off;  /* [nQ] offset of query q's staircase, PPP_NO_TABLE = accept every pair */
    uint32_t *qcnt;            /* [nQ] append counters */
    uint32_t qcap;             /* per-query append capacity (>= targets per launch) */
    uint32_t qwords;           /* words per query plane row (universe width; = W for bit-plane targets) */
    uint32_t bounds_only;      /* 1: report exp(min upper enclosure) as a provisional score (threshold seeding); no tier B/C */
    uint32_t dup;              /* ranked lists packed for -dup (ppp.pm:189-200): a list may hold ONE repeated taxon, which can push
                                  yes past |Y| -- the reference stops there (ppp.pm:222-224), so that candidate is dropped */
    /* ranked targets (ppp_set_targets_ranked): per target an ordered, de-duplicated list of genome indices */
    const uint16_t *ridx;      /* genome index in the query universe, 0xFFFF = foreign taxon */
    const uint16_t *rraw;      /* 1-based raw list position of each entry (duplicates counted) */
    const unsigned long long *roff; /* [nT+1] entry offsets */
};

#define PPP_NO_TABLE 0xffffffffu

enum { CNT_EXACT = 0, CNT_CAND = 1, CNT_ACCURATE = 2, CNT_VIOL = 3, CNT_CHECKS = 4, CNT_OVERFLOW = 5, CNT_APPEND = 6,
       CNT_FULL = 7, CNT_N = 8,
       CNT_SURV = 8, /* survivor count of the pre-filter (slot CNT_N) */
       CNT_ITEM = 9, /* work-item counter of the tile kernels */
       CNT_SURV1 = 10, /* pairs with a passing CHUNK (tensor-core screen, pair-level mode): re-screened into CNT_SURV */
       CNT_ITEM2 = 11, /* work-item counter of the re-screen kernel */
       CNT_PF_UNIT_GEN = 12, CNT_PF_UNIT_ALL = 13, CNT_PF_UNIT_PYY = 14, /* work counters of the pre-filter's three launches of a chunk */
       CNT_SLOTS = 16 };

#ifdef __CUDACC__
#include "../../include/ppp_b200.h"
/* does a finished pair belong to the filtered output? */
__device__ __forceinline__ bool ppp_accept(const SweepParams &prm, uint32_t q, double score)
{
    if (prm.accept_mode == 0) return score < prm.kth[q];
    if (prm.accept_mode == 3) return score <= prm.kth[q];
    if (prm.accept_mode == 2) return score > prm.klo[q] && score <= prm.kth[q];
    return log10(score) < prm.thresh_log10;
}
/* write one result record: dense slot, per-query append or global append */
__device__ __forceinline__ void ppp_emit(const SweepParams &prm, uint32_t q, uint32_t t, double score, uint32_t yes,
                                         uint32_t number, uint32_t depth, float margin)
{
    ppp_hit *base = reinterpret_cast<ppp_hit *>(prm.hits);
    size_t idx;
    /* threshold seeding: a provisional score must stay STRICTLY above the real (Cephes-rounded) score of its pair,
     * also when that is exactly 0.0 (underflow), 1.0 (no hit) or a denormal -- acceptance is `score < kth` */
    if (prm.bounds_only) {
        score = score * (1.0 + 1e-6) + 1e-300;
        if (prm.append == 3) { /* seeding pass: only the bound itself is needed, as a dense [nQ][t1 - t0] matrix of doubles */
            reinterpret_cast<double *>(prm.hits)[(size_t)q * (prm.t1 - prm.t0) + (t - prm.t0)] = score;
            return;
        }
    }
    if (prm.append == 0) {
        idx = (size_t)q * prm.nT + t;
    } else {
        if (!ppp_accept(prm, q, score)) return;
        if (prm.append == 1) {
            const uint32_t slot = atomicAdd(&prm.qcnt[q], 1u);
            if (slot >= prm.qcap) {
                atomicAdd(&prm.counters[CNT_OVERFLOW], 1ull);
                return;
            }
            idx = (size_t)q * prm.qcap + slot;
        } else {
            idx = (size_t)atomicAdd(&prm.counters[CNT_APPEND], 1ull);
        }
    }
    const unsigned long long sb = (unsigned long long)__double_as_longlong(score);
    uint4 *dst = reinterpret_cast<uint4 *>(base + idx);
    dst[0] = make_uint4(q, t, (uint32_t)sb, (uint32_t)(sb >> 32));
    dst[1] = make_uint4(yes, number, depth, __float_as_uint(margin));
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
        /* ln(1-g) <= -g is only tight for small g (off by g^2 / 2: 0.045 at g = 0.3, the regime of unrelated pairs); __logf errs
         * by < 1e-6 absolute on [0.5, 1].  Measured on the dense G = 4096 probe: 3.42 -> 3.33 FP64-tier evaluations per pair --
         * near the mean the width of the enclosure comes from the series remainder after 8 terms, not from this bound. */
        hi = (glo < 0.01f) ? fminf(-glo, 0.f) : __logf(fmaxf(1.f - glo, 0.f)) + 1e-4f;
        hi = fminf(hi, 0.f);
        lo = (ghi < 0.9f) ? __logf(1.f - ghi) - 1e-4f : -INFINITY;
        if (ghi < 1e-3f) lo = -ghi * (1.f + ghi) - 1e-30f;
    }
}

#endif
