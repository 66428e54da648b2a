/*
 * prophylo_b200/csrc/ppp_exact.cu -- the EXACT tier of the PPP scan.  Compile with -fmad=false.
 *
 * For every pair the sweep kernel could not decide beyond doubt, replay the reference loop
 *   PhyloProf::Algorithm::ppp::get_match_score    /root/reference/lib/PhyloProf/Algorithm/ppp.pm:169-242
 * literally -- every list element of the target, `answer = bdtrc(yes-1, number, p)` with the bit-exact device
 * restatement of Math::Cephes (cephes_exact.cuh), strict `<`, stop on `answer <= 0` -- so that cut-off index AND
 * score are the reference's, including Cephes' quantisation (exact 0.0 below MINLOG, 1 - MACHEP / 1.0 near one,
 * "domain error" -> 0.0).  One thread per flagged pair; flagged pairs are rare (near-ties, saturated or
 * underflowing sweeps), so this tier is sized for exactness, not throughput.
 */
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/ppp_b200.h"
#include "cephes_exact.cuh"
#include "ppp_common.cuh"

#define EXQ 1024 /* steps per queue pass */
#define FULLM 0xffffffffu

/* One WARP per flagged pair.  The steps of the reference loop that can change (yes, number) are the set bits of T & P
 * (an element outside P repeats the previous answer, which can neither improve on the running minimum nor move the
 * point where the loop stops); they are enumerated in list order into a shared-memory queue exactly as in the sweep
 * kernel, evaluated 32 at a time (one bdtrc per lane) and then folded in order with the reference's rule.
 *
 * Underflow mode (bit 31 of the list id; set when the sweep saw ln p < -700): the answer is decided among steps whose
 * tail is below e^-600 -- the global minimum is there and exact zeros only occur there -- so steps whose leading
 * binomial term alone exceeds e^-600 (tail >= leading term) are not evaluated. */
/* fold one pair exactly; tw/tp/ty are this lane's words of the target plane, T&P and T&P&Y (list order) */
template <int WPL>
__device__ __forceinline__ void exact_pair(const SweepParams &prm, const QConst &qc, const uint32_t (&tw)[WPL], const uint32_t (&tp)[WPL],
                                           const uint32_t (&ty)[WPL], unsigned long long *queue, int lane, bool underflow_mode,
                                           double &score, uint32_t &best_yes, uint32_t &best_number, uint32_t &best_depth)
{
    const double p = qc.p;
    const int max_yes = (int)qc.ny;
    uint32_t cy = 0, cn = 0, cd = 0;
#pragma unroll
    for (int w = 0; w < WPL; ++w) {
        cy += __popc(ty[w]);
        cn += __popc(tp[w]);
        cd += __popc(tw[w]);
    }
    uint32_t iy = cy, in_ = cn, id_ = cd;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t a = __shfl_up_sync(FULLM, iy, o), b = __shfl_up_sync(FULLM, in_, o), c = __shfl_up_sync(FULLM, id_, o);
        if (lane >= o) { iy += a; in_ += b; id_ += c; }
    }
    const uint32_t ypre = iy - cy, npre = in_ - cn, dpre = id_ - cd;
    const uint32_t S = __shfl_sync(FULLM, in_, 31); /* number of steps = popc(T & P) */

    score = 1.0; /* ppp.pm:132 */
    best_yes = best_number = best_depth = 0;
    bool stop = false;
    for (uint32_t qbase = 0; qbase < S && !stop; qbase += EXQ) {
        { /* enqueue steps [qbase, qbase + EXQ): slot = number - 1 */
            uint32_t yrun = ypre, nrun = npre, drun = dpre;
#pragma unroll
            for (int w = 0; w < WPL; ++w) {
                uint32_t mbits = tp[w];
                while (mbits) {
                    const uint32_t b = __ffs(mbits) - 1;
                    mbits &= mbits - 1;
                    const uint32_t below = (2u << b) - 1u;
                    ++nrun;
                    const uint32_t yy = yrun + __popc(ty[w] & below);
                    const uint32_t dd = drun + __popc(tw[w] & below);
                    const uint32_t slot = nrun - 1 - qbase;
                    if (slot < EXQ) queue[slot] = (unsigned long long)yy | ((unsigned long long)dd << 16) | ((unsigned long long)nrun << 32);
                }
                yrun += __popc(ty[w]);
                drun += __popc(tw[w]);
            }
        }
        __syncwarp();
        const uint32_t cnt = min((uint32_t)EXQ, S - qbase);
        for (uint32_t j0 = 0; j0 < cnt && !stop; j0 += 32) {
            const uint32_t j = j0 + lane;
            double answer = 2.0; /* 2.0 = "not evaluated" (can neither improve nor stop) */
            uint32_t yy = 0, nn = 0, dd = 0;
            if (j < cnt) {
                const unsigned long long e = queue[j];
                yy = (uint32_t)(e & 0xffffu);
                dd = (uint32_t)((e >> 16) & 0xffffu);
                nn = (uint32_t)(e >> 32);
                bool eval = true;
                if (underflow_mode) {
                    if (yy == 0) eval = false; /* bdtrc(-1, n, p) = 1.0 */
                    else {
                        const double lead = ((prm.lf[nn] - prm.lf[yy]) - prm.lf[nn - yy]) + ((double)yy * qc.lnp + (double)(nn - yy) * qc.lnq);
                        eval = !(lead > -600.0);
                    }
                }
                if (eval) answer = ((int)yy > max_yes) ? -1.0 : dc_bdtrc((int)yy - 1, (int)nn, p); /* ppp.pm:222-225 */
            }
            /* fold the 32 answers in list order (ppp.pm:228-241) */
            const uint32_t lim = min(32u, cnt - j0);
            for (uint32_t s = 0; s < lim; ++s) {
                const double a = __shfl_sync(FULLM, answer, s);
                if (a == 2.0) continue;
                if (a == -1.0) { stop = true; break; } /* yes > max_yes: `last` before bdtrc */
                if (a < score) {
                    score = a;
                    best_yes = __shfl_sync(FULLM, yy, s);
                    best_number = __shfl_sync(FULLM, nn, s);
                    best_depth = __shfl_sync(FULLM, dd, s);
                    continue;
                }
                if (a <= 0) { stop = true; break; }
            }
        }
        __syncwarp();
    }
}

template <int WPL>
__global__ void __launch_bounds__(128) ppp_exact_kernel(const SweepParams prm)
{
    constexpr int W = 32 * WPL;
    __shared__ unsigned long long s_queue[4][EXQ];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned long long n_list = min(prm.counters[CNT_EXACT], (unsigned long long)prm.exact_cap);
    const uint32_t nTl = prm.t1 - prm.t0;
    const unsigned long long nwarps = (unsigned long long)gridDim.x * 4;
    for (unsigned long long i = (unsigned long long)blockIdx.x * 4 + warp; i < n_list; i += nwarps) {
        const uint32_t raw = prm.exact_list[i];
        const uint32_t id = raw & 0x7fffffffu;
        const uint32_t q = id / nTl, t = prm.t0 + id % nTl;
        const uint32_t *Tw = prm.T + (size_t)t * W + lane * WPL;
        const uint32_t *Pw = prm.Q + (size_t)q * 2 * W + lane * WPL;
        const uint32_t *Yw = Pw + W;
        uint32_t tw[WPL], tp[WPL], ty[WPL];
#pragma unroll
        for (int w = 0; w < WPL; ++w) {
            tw[w] = Tw[w];
            tp[w] = tw[w] & Pw[w];
            ty[w] = tp[w] & Yw[w];
        }
        double score;
        uint32_t by, bn, bd;
        exact_pair<WPL>(prm, prm.qc[q], tw, tp, ty, s_queue[warp], lane, (raw >> 31) != 0, score, by, bn, bd);
        if (lane == 0) ppp_emit(prm, q, t, score, by, bn, bd, 0.f);
    }
}

/* ranked targets: gather T&P / T&Y in list order with ballots (see ppp_ranked.cu), then the same fold */
template <int WPL>
__global__ void __launch_bounds__(128) ppp_exact_ranked_kernel(const SweepParams prm)
{
    constexpr int NW = 32 * WPL;
    __shared__ unsigned long long s_queue[4][EXQ];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned long long n_list = min(prm.counters[CNT_EXACT], (unsigned long long)prm.exact_cap);
    const uint32_t nTl = prm.t1 - prm.t0, Wg = prm.qwords;
    const unsigned long long nwarps = (unsigned long long)gridDim.x * 4;
    for (unsigned long long i = (unsigned long long)blockIdx.x * 4 + warp; i < n_list; i += nwarps) {
        const uint32_t raw = prm.exact_list[i];
        const uint32_t id = raw & 0x7fffffffu;
        const uint32_t q = id / nTl, t = prm.t0 + id % nTl;
        const unsigned long long off0 = prm.roff[t];
        const uint32_t L = (uint32_t)(prm.roff[t + 1] - off0);
        const uint32_t *pq = prm.Q + (size_t)q * 2 * Wg, *yq = pq + Wg;
        uint32_t tw[WPL], tp[WPL], ty[WPL];
#pragma unroll
        for (int w = 0; w < WPL; ++w) {
            const uint32_t base = (uint32_t)(lane * WPL + w) * 32u;
            tw[w] = (base + 32u <= L) ? 0xffffffffu : (base < L ? ((1u << (L - base)) - 1u) : 0u);
            tp[w] = ty[w] = 0;
        }
#pragma unroll
        for (int j = 0; j < NW; ++j) {
            if ((uint32_t)j * 32u >= L) break;
            const uint32_t pos = (uint32_t)j * 32u + lane;
            const uint32_t g = (pos < L) ? (uint32_t)prm.ridx[off0 + pos] : 0xFFFFu;
            bool inp = false, iny = false;
            if (g != 0xFFFFu) {
                inp = (pq[g >> 5] >> (g & 31u)) & 1u;
                iny = inp && ((yq[g >> 5] >> (g & 31u)) & 1u);
            }
            const uint32_t wp = __ballot_sync(FULLM, inp), wy = __ballot_sync(FULLM, iny);
            if (lane == j / WPL) {
                tp[j % WPL] = wp;
                ty[j % WPL] = wy;
            }
        }
        double score;
        uint32_t by, bn, bd;
        exact_pair<WPL>(prm, prm.qc[q], tw, tp, ty, s_queue[warp], lane, (raw >> 31) != 0, score, by, bn, bd);
        if (lane == 0) ppp_emit(prm, q, t, score, by, bn, bd ? (uint32_t)prm.rraw[off0 + bd - 1] : 0u, 0.f);
    }
}

extern "C" cudaError_t ppp_launch_exact(const SweepParams *prm, int W, int nsm, cudaStream_t st)
{
    const int grid = nsm * 8;
    switch (W / 32) {
    case 1: ppp_exact_kernel<1><<<grid, 128, 0, st>>>(*prm); break;
    case 2: ppp_exact_kernel<2><<<grid, 128, 0, st>>>(*prm); break;
    case 4: ppp_exact_kernel<4><<<grid, 128, 0, st>>>(*prm); break;
    case 8: ppp_exact_kernel<8><<<grid, 128, 0, st>>>(*prm); break;
    default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

extern "C" cudaError_t ppp_launch_exact_ranked(const SweepParams *prm, int rwpl, int nsm, cudaStream_t st)
{
    const int grid = nsm * 8;
    switch (rwpl) {
    case 1: ppp_exact_ranked_kernel<1><<<grid, 128, 0, st>>>(*prm); break;
    case 2: ppp_exact_ranked_kernel<2><<<grid, 128, 0, st>>>(*prm); break;
    default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

__global__ void ppp_debug_bdtrc_kernel(const int32_t *k, const int32_t *n, const double *p, double *out, size_t m)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < m; i += (size_t)gridDim.x * blockDim.x)
        out[i] = dc_bdtrc(k[i], n[i], p[i]);
}

extern "C" cudaError_t ppp_launch_debug_bdtrc(const int32_t *k, const int32_t *n, const double *p, double *out, size_t m,
                                              cudaStream_t st)
{
    ppp_debug_bdtrc_kernel<<<296, 128, 0, st>>>(k, n, p, out, m);
    return cudaGetLastError();
}
