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

__global__ void ppp_exact_kernel(const SweepParams prm, int W)
{
    const unsigned long long n_list = min(prm.counters[CNT_EXACT], (unsigned long long)prm.exact_cap);
    const uint32_t nTl = prm.t1 - prm.t0;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n_list;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const uint32_t id = prm.exact_list[i];
        const uint32_t q = id / nTl, t = prm.t0 + id % nTl;
        const uint32_t *Tw = prm.T + (size_t)t * W;
        const uint32_t *Pw = prm.Q + (size_t)q * 2 * W;
        const uint32_t *Yw = Pw + W;
        const double p = prm.qc[q].p;
        const int max_yes = (int)prm.qc[q].ny;
        const bool p_in_range = (p >= 0.0) && (p <= 1.0); /* false for NaN as well: then nothing can be skipped */

        double score = 1.0; /* ppp.pm:132 */
        int yes = 0, number = 0, depth = 0;
        int best_yes = 0, best_number = 0, best_depth = 0;
        bool stop = false;
        for (int w = 0; w < W && !stop; ++w) {
            uint32_t m = Tw[w];
            const uint32_t pw = Pw[w], yw = Yw[w] & pw;
            while (m) {
                const uint32_t b = __ffs(m) - 1;
                m &= m - 1;
                ++depth; /* $i + 1 (ppp.pm:169,231) */
                const int in_p = (pw >> b) & 1u, in_y = (yw >> b) & 1u;
                /* an element outside P repeats the previous (yes, number): for p in [0,1] its answer equals the
                 * previous one (or 1.0 while yes == 0), so it can neither improve nor change where the loop ends */
                if (!in_p && p_in_range) continue;
                number += in_p; /* ppp.pm:207-210 */
                yes += in_y;    /* ppp.pm:212-220 */
                if (yes > max_yes) { stop = true; break; } /* ppp.pm:222-224 */
                const double answer = dc_bdtrc(yes - 1, number, p); /* ppp.pm:225 */
                if (answer < score) {                                /* ppp.pm:228-237 */
                    best_depth = depth;
                    score = answer;
                    best_yes = yes;
                    best_number = number;
                    continue;
                }
                if (answer <= 0) { stop = true; break; } /* ppp.pm:239-241 */
            }
        }
        ppp_emit(prm, q, t, score, (uint32_t)best_yes, (uint32_t)best_number, (uint32_t)best_depth, 0.f);
    }
}

extern "C" cudaError_t ppp_launch_exact(const SweepParams *prm, int W, int nsm, cudaStream_t st)
{
    ppp_exact_kernel<<<nsm * 4, 64, 0, st>>>(*prm, W);
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
