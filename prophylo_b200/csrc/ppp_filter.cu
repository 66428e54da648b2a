/*
 * prophylo_b200/csrc/ppp_filter.cu -- device threshold filter over dense hit records.
 *
 * Keeps the pairs a ppp_cutoff-style consumer can ever print: bin/ppp_cutoff.pl:122,154 only looks at rows whose
 * printed "%.3f" log10 score is < 0, i.e. log10(score) < -0.0005 (PPP_FILTER_THRESH_LOG10 with that threshold).
 * Warp-aggregated append: one atomicAdd per warp.
 */
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/ppp_b200.h"

__global__ void ppp_filter_kernel(const ppp_hit *__restrict__ hits, uint64_t npairs, double thresh_log10, int keep_all,
                                  ppp_hit *__restrict__ out, uint64_t cap, unsigned long long *count)
{
    const int lane = threadIdx.x & 31;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t start = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    /* all lanes of a warp run the same number of iterations (ballot needs the full warp) */
    const uint64_t iters = (npairs + stride - 1) / stride;
    for (uint64_t it = 0; it < iters; ++it) {
        const uint64_t i = start + it * stride;
        bool keep = false;
        ppp_hit h;
        if (i < npairs) {
            const uint4 *src = reinterpret_cast<const uint4 *>(hits + i);
            const uint4 a = src[0], b = src[1];
            h.q = a.x; h.t = a.y;
            h.score = __longlong_as_double((long long)(((unsigned long long)a.w << 32) | a.z));
            h.yes = b.x; h.number = b.y; h.depth = b.z; h.margin = __uint_as_float(b.w);
            keep = keep_all || (log10(h.score) < thresh_log10);
        }
        const uint32_t ball = __ballot_sync(0xffffffffu, keep);
        if (!ball) continue;
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(count, (unsigned long long)__popc(ball));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (keep) {
            const unsigned long long slot = base + __popc(ball & ((1u << lane) - 1u));
            if (slot < cap) out[slot] = h;
        }
    }
}

extern "C" cudaError_t ppp_launch_filter(const void *hits, uint64_t npairs, double thresh_log10, int keep_all, void *out,
                                         uint64_t cap, unsigned long long *count, int nsm, cudaStream_t st)
{
    ppp_filter_kernel<<<nsm * 8, 256, 0, st>>>(reinterpret_cast<const ppp_hit *>(hits), npairs, thresh_log10, keep_all,
                                               reinterpret_cast<ppp_hit *>(out), cap, count);
    return cudaGetLastError();
}
