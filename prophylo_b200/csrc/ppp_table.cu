/*
 * prophylo_b200/csrc/ppp_table.cu -- EXACT TABLE mode for dense runs of few queries against many targets
 * (BASELINE config 2 / the bin/ppp.pl shape: one query profile vs up to a million targets).  Compile with -fmad=false.
 *
 * For a fixed query the reference's answer at a sweep step depends only on (yes, number):
 *     answer = bdtrc(yes - 1, number, p)                       lib/PhyloProf/Algorithm/ppp.pm:225
 * with 1 <= yes <= |Y| and yes <= number <= |P|.  When one query meets many targets the same (yes, number) is
 * evaluated again and again, so the whole table F[yes][number] is computed ONCE per query with the bit-exact device
 * restatement of Math::Cephes (cephes_exact.cuh) -- |Y| x |P| evaluations instead of sum-over-targets popc(T&Y) --
 * and every pair is then a walk over its candidates with one table lookup each:
 *     for every set bit of T & Y in genome order:  f = F[yes][number];  if (f < score) record      ppp.pm:228-237
 * One thread per target (coalesced chunk-major planes), scores bit-identical to the reference by construction.
 *
 * The walk visits only the steps where yes increments.  That equals the reference loop iff, for this query, the table
 * is non-negative and non-decreasing in number at fixed yes (then a step that only increments number can neither
 * improve on the running minimum nor trigger `last if answer <= 0` before the minimum is reached).  Both properties
 * are VERIFIED on the finished table (ppp_table_check_kernel); a query that fails the check is reported back and the
 * caller uses the general sweep for the run instead.
 */
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/ppp_b200.h"
#include "cephes_exact.cuh"
#include "ppp_common.cuh"

/* table of query q: tab[toff[q] + yes * (npres_q + 1) + number], yes in [0, ny], number in [0, npres] */
__global__ void ppp_table_build_kernel(const QConst *__restrict__ qc, const unsigned long long *__restrict__ toff, uint32_t nQ,
                                       double *__restrict__ tab)
{
    const unsigned long long total = toff[nQ];
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < total; i += stride) {
        /* locate the query by binary search over the offsets */
        uint32_t lo = 0, hi = nQ - 1;
        while (lo < hi) {
            const uint32_t mid = (lo + hi + 1) >> 1;
            if (toff[mid] <= i) lo = mid;
            else hi = mid - 1;
        }
        const QConst c = qc[lo];
        const unsigned long long r = i - toff[lo];
        const uint32_t cols = c.npres + 1;
        const uint32_t y = (uint32_t)(r / cols), n = (uint32_t)(r % cols);
        double f = 2.0; /* unreachable cells */
        if (y >= 1 && n >= y) f = dc_bdtrc((int)y - 1, (int)n, c.p);
        tab[i] = f;
    }
}

/* one warp per (query, yes) row: non-negative and non-decreasing in number? */
__global__ void ppp_table_check_kernel(const QConst *__restrict__ qc, const unsigned long long *__restrict__ toff, uint32_t nQ,
                                       const double *__restrict__ tab, uint32_t *__restrict__ bad)
{
    const int lane = threadIdx.x & 31;
    const unsigned long long warps = ((unsigned long long)gridDim.x * blockDim.x) >> 5;
    unsigned long long rows_before = 0;
    for (uint32_t q = 0; q < nQ; ++q) {
        const QConst c = qc[q];
        const uint32_t cols = c.npres + 1;
        for (unsigned long long w = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < c.ny; w += warps) {
            const uint32_t y = (uint32_t)w + 1;
            const double *row = tab + toff[q] + (unsigned long long)y * cols;
            bool ok = true;
            for (uint32_t n = y + lane; n <= c.npres; n += 32) {
                const double f = row[n];
                if (!(f >= 0.0)) ok = false;
                if (n > y && !(row[n - 1] <= f)) ok = false;
            }
            if (!__all_sync(0xffffffffu, ok) && lane == 0) atomicOr(&bad[q], 1u);
        }
        rows_before += c.ny;
    }
    (void)rows_before;
}

/* thread per target, queries in an outer loop; Tt = chunk-major planes [W/4][ntt] of uint4 */
template <int WPL>
__global__ void __launch_bounds__(256) ppp_table_sweep_kernel(const SweepParams prm, const uint4 *__restrict__ Tt, uint32_t ntt,
                                                              const unsigned long long *__restrict__ toff, const double *__restrict__ tab)
{
    constexpr int W = 32 * WPL, C = W / 4;
    __shared__ uint32_t s_p[W], s_y[W];
    const uint32_t nTl = prm.t1 - prm.t0;
    const uint32_t nTblocks = (nTl + blockDim.x - 1) / blockDim.x;
    const uint64_t nItems = (uint64_t)prm.nQ * nTblocks;
    uint32_t cur_q = 0xffffffffu;
    for (uint64_t item = blockIdx.x; item < nItems; item += gridDim.x) {
        const uint32_t q = (uint32_t)(item / nTblocks), tb = (uint32_t)(item % nTblocks);
        if (q != cur_q) { /* block-uniform */
            __syncthreads();
            for (int i = threadIdx.x; i < W; i += blockDim.x) {
                s_p[i] = prm.Q[(size_t)q * 2 * W + i];
                s_y[i] = prm.Q[(size_t)q * 2 * W + W + i];
            }
            __syncthreads();
            cur_q = q;
        }
        const uint32_t tl = tb * blockDim.x + threadIdx.x;
        if (tl >= nTl) continue;
        const uint32_t t = prm.t0 + tl;
        const QConst qcst = prm.qc[q];
        double score = 1.0;
        uint32_t best_yes = 0, best_number = 0, best_depth = 0;
        if (qcst.flags & PPP_QF_DOMAIN) {
            /* p outside [0,1]: 0.0 at the first list element (ppp.pm:225-241) */
            uint32_t depth_seen = 0;
            for (int c = 0; c < C && !depth_seen; ++c) {
                const uint4 tv = __ldg(Tt + (size_t)c * ntt + t);
                const uint32_t tw[4] = {tv.x, tv.y, tv.z, tv.w};
                for (int w = 0; w < 4 && !depth_seen; ++w)
                    if (tw[w]) {
                        const uint32_t b = __ffs(tw[w]) - 1;
                        const uint32_t pw = s_p[c * 4 + w], yw = s_y[c * 4 + w] & pw;
                        score = 0.0;
                        best_number = (pw >> b) & 1u;
                        best_yes = (yw >> b) & 1u;
                        best_depth = 1;
                        depth_seen = 1;
                    }
            }
        } else if (!(qcst.flags & PPP_QF_TRIVIAL)) {
            const uint32_t cols = qcst.npres + 1;
            const double *F = tab + toff[q];
            uint32_t yrun = 0, nrun = 0, drun = 0;
            for (int c = 0; c < C; ++c) {
                const uint4 tv = __ldg(Tt + (size_t)c * ntt + t);
                const uint32_t tw[4] = {tv.x, tv.y, tv.z, tv.w};
#pragma unroll
                for (int w = 0; w < 4; ++w) {
                    const uint32_t a = tw[w] & s_p[c * 4 + w];
                    uint32_t b = a & s_y[c * 4 + w];
                    while (b) {
                        const uint32_t bit = __ffs(b) - 1;
                        b &= b - 1;
                        ++yrun;
                        const uint32_t below = (2u << bit) - 1u;
                        const uint32_t nn = nrun + __popc(a & below);
                        const double f = __ldg(F + (size_t)yrun * cols + nn);
                        if (f < score) { /* strict: the first minimum wins (ppp.pm:228-237) */
                            score = f;
                            best_yes = yrun;
                            best_number = nn;
                            best_depth = drun + __popc(tw[w] & below);
                        }
                    }
                    nrun += __popc(a);
                    drun += __popc(tw[w]);
                }
            }
        }
        ppp_emit(prm, q, t, score, best_yes, best_number, best_depth, 0.f);
    }
}

extern "C" cudaError_t ppp_launch_table_build(const QConst *qc, const unsigned long long *toff, uint32_t nQ, double *tab, uint32_t *bad,
                                              int nsm, cudaStream_t st)
{
    ppp_table_build_kernel<<<nsm * 8, 128, 0, st>>>(qc, toff, nQ, tab);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    ppp_table_check_kernel<<<nsm * 4, 128, 0, st>>>(qc, toff, nQ, tab, bad);
    return cudaGetLastError();
}

extern "C" cudaError_t ppp_launch_table_sweep(const SweepParams *prm, int wpl, const void *Tt, uint32_t ntt, const unsigned long long *toff,
                                              const double *tab, int nsm, cudaStream_t st)
{
    const int grid = nsm * 8;
    const uint4 *t4 = reinterpret_cast<const uint4 *>(Tt);
    switch (wpl) {
    case 1: ppp_table_sweep_kernel<1><<<grid, 256, 0, st>>>(*prm, t4, ntt, toff, tab); break;
    case 2: ppp_table_sweep_kernel<2><<<grid, 256, 0, st>>>(*prm, t4, ntt, toff, tab); break;
    case 4: ppp_table_sweep_kernel<4><<<grid, 256, 0, st>>>(*prm, t4, ntt, toff, tab); break;
    case 8: ppp_table_sweep_kernel<8><<<grid, 256, 0, st>>>(*prm, t4, ntt, toff, tab); break;
    default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}
