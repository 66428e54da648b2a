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
#define PPP_QCAP 1024     /* candidates per warp queue pass */
#define PPP_K_TERMS 8     /* series terms of the float enclosure tier */

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
    uint32_t append;           /* 0: dense hits[q*nT+t]; 1: per-query append hits[q*qcap + qcnt[q]++]; 2: global append */
    uint32_t accept_mode;      /* 0: score < kth[q] (top-k, kth = 2.0 while the list is not full); 1: log10(score) < thresh */
    double thresh_log10;
    const double *kth;         /* [nQ] */
    const uint16_t *nmax;      /* prefilter staircase: candidate (yes, number) can pass only if number <= nmax[off + yes] */
    const uint32_t *nmax_off;  /* [nQ] offset of query q's staircase, PPP_NO_TABLE = accept every pair */
    uint32_t *qcnt;            /* [nQ] append counters */
    uint32_t qcap;             /* per-query append capacity (>= targets per launch) */
    uint32_t pad2_;
};

#define PPP_NO_TABLE 0xffffffffu

enum { CNT_EXACT = 0, CNT_CAND = 1, CNT_ACCURATE = 2, CNT_VIOL = 3, CNT_CHECKS = 4, CNT_OVERFLOW = 5, CNT_APPEND = 6,
       CNT_FULL = 7, CNT_N = 8 };

#ifdef __CUDACC__
#include "../../include/ppp_b200.h"
/* does a finished pair belong to the filtered output? */
__device__ __forceinline__ bool ppp_accept(const SweepParams &prm, uint32_t q, double score)
{
    if (prm.accept_mode == 0) return score < prm.kth[q];
    return log10(score) < prm.thresh_log10;
}
/* write one result record: dense slot, per-query append or global append */
__device__ __forceinline__ void ppp_emit(const SweepParams &prm, uint32_t q, uint32_t t, double score, uint32_t yes,
                                         uint32_t number, uint32_t depth, float margin)
{
    ppp_hit *base = reinterpret_cast<ppp_hit *>(prm.hits);
    size_t idx;
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
#endif
