/*
 * prophylo_b200/csrc/ppp_debug.cu -- diagnostic entry behind ppp_debug_enclosure(): the fast tiers of the sweep evaluated at
 * caller-chosen (yes, number, p) points, so that the parity tests can check them against the exact tier and the oracle
 * directly instead of only through whole pairs.
 *
 * The staircase prune of filtered runs (ppp_topk.cu: tail_may_be_le) is sound only if tier A's lower bound never exceeds the
 * true ln P(X >= yes | number, p) = ln Math::Cephes::bdtrc(yes-1, number, p) (/root/reference/lib/PhyloProf/Algorithm/
 * ppp.pm:225); a `lo` above the truth would silently drop a true top-k hit.  One warp per point: lane 0 evaluates the float
 * enclosure (tier_a, ppp_common.cuh), the warp the FP64 series (tier_b, ppp_pair.cuh).
 */
#include <cuda_runtime.h>
#include <stdint.h>

#include "ppp_common.cuh"
#include "ppp_pair.cuh"

__global__ void ppp_debug_enclosure_kernel(const QConst *__restrict__ qc, const double *__restrict__ lf, const int32_t *__restrict__ y,
                                           const int32_t *__restrict__ n, float *__restrict__ lo, float *__restrict__ hi,
                                           double *__restrict__ lnf, size_t m)
{
    const int lane = threadIdx.x & 31;
    const size_t warps = ((size_t)gridDim.x * blockDim.x) >> 5;
    for (size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < m; i += warps) {
        const QConst c = qc[i];
        float l = 0.f, h = 0.f;
        if (lane == 0) tier_a(y[i], n[i], c, lf, l, h);
        double f;
        const double v = tier_b(y[i], n[i], c, lf, lane, &f);
        if (lane == 0) {
            lo[i] = l;
            hi[i] = h;
            lnf[i] = v;
        }
    }
}

extern "C" cudaError_t ppp_launch_debug_enclosure(const QConst *qc, const double *lf, const int32_t *y, const int32_t *n, float *lo, float *hi,
                                                  double *lnf, size_t m, int nsm, cudaStream_t st)
{
    ppp_debug_enclosure_kernel<<<nsm * 8, 256, 0, st>>>(qc, lf, y, n, lo, hi, lnf, m);
    return cudaGetLastError();
}

/* ---------------------------------------------------------------- pipe-peak microbenchmarks (ppp_pipe_peak)
 * The all-pairs screen is bound by SM issue pipes, not by HBM, so its roofline denominator is a
 * pipe rate.  SURVEY.md section 8(d) asks for that rate to be MEASURED on the box under the run's clocks rather than
 * taken from a table: each kernel below keeps 8 independent dependent-chains per thread of one instruction kind
 * (volatile asm: ptxas can neither fold nor reorder them away), 64 warps per SM, and runs for a few milliseconds.
 *   0  popc.b32      (XU pipe: the screen's scarce instruction)
 *   1  lop3.b32      (ALU)
 *   2  add.u32       (IADD3, ALU)
 *   3  fma.rn.f64    (FP64 pipe: the series tier)
 */
#define MB_CHAINS 8
#define MB_UNROLL 8

template <int KIND>
__global__ void __launch_bounds__(1024) ppp_pipe_peak_kernel(uint32_t iters, uint32_t seed, uint32_t *sink)
{
    uint32_t x[MB_CHAINS];
    double d[MB_CHAINS];
    const uint32_t a = seed | 1u, b = seed * 2654435761u + threadIdx.x;
    const double da = 0.999999 + 1e-9 * (double)(seed & 7), db = 1e-3;
#pragma unroll
    for (int c = 0; c < MB_CHAINS; ++c) {
        x[c] = b + 977u * c + blockIdx.x;
        d[c] = 1.0 + 1e-3 * c + 1e-6 * threadIdx.x;
    }
    for (uint32_t i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < MB_UNROLL; ++u) {
#pragma unroll
            for (int c = 0; c < MB_CHAINS; ++c) {
                if (KIND == 0) asm volatile("popc.b32 %0, %0;" : "+r"(x[c]));
                if (KIND == 1) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[c]) : "r"(a), "r"(b));
                if (KIND == 2) asm volatile("add.u32 %0, %0, %1;" : "+r"(x[c]) : "r"(a));
                if (KIND == 3) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[c]) : "d"(da), "d"(db));
            }
        }
    }
    uint32_t acc = 0;
#pragma unroll
    for (int c = 0; c < MB_CHAINS; ++c) acc += x[c] + (uint32_t)__double2int_rn(d[c]);
    if (acc == 0xdeadbeefu) sink[0] = acc; /* keeps the chains live */
}

extern "C" cudaError_t ppp_launch_pipe_peak(int kind, uint32_t iters, uint32_t *sink, int nsm, cudaStream_t st, double *lane_ops)
{
    const dim3 grid((unsigned)(2 * nsm)), block(1024);
    switch (kind) {
    case 0: ppp_pipe_peak_kernel<0><<<grid, block, 0, st>>>(iters, 12345u, sink); break;
    case 1: ppp_pipe_peak_kernel<1><<<grid, block, 0, st>>>(iters, 12345u, sink); break;
    case 2: ppp_pipe_peak_kernel<2><<<grid, block, 0, st>>>(iters, 12345u, sink); break;
    case 3: ppp_pipe_peak_kernel<3><<<grid, block, 0, st>>>(iters, 12345u, sink); break;
    default: return cudaErrorInvalidValue;
    }
    *lane_ops = (double)grid.x * 1024.0 * (double)iters * MB_UNROLL * MB_CHAINS;
    return cudaGetLastError();
}
