// Developer probe: throughput of the legacy warp-level integer MMA (mma.sync m16n8k32 s8, SASS IMMA.16832) on sm_100a,
// the instruction an AND+POPC contraction of 0/1 bytes would use outside tcgen05.  nvcc -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int CHAINS>
__global__ void __launch_bounds__(512) imma_kernel(int iters, int *sink)
{
    int c[CHAINS][4];
    uint32_t a0 = threadIdx.x, a1 = threadIdx.x * 3, a2 = threadIdx.x * 5, a3 = threadIdx.x * 7, b0 = threadIdx.x * 11, b1 = threadIdx.x * 13;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i)
            asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+r"(c[i][0]), "+r"(c[i][1]), "+r"(c[i][2]), "+r"(c[i][3])
                         : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
    int s = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    if (s == 0x7fffffff) sink[0] = s;
}
int main()
{
    int *sink;
    cudaMalloc(&sink, 4);
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int iters = 4096;
    for (int warps = 4; warps <= 16; warps *= 2) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            imma_kernel<8><<<p.multiProcessorCount * 2, warps * 32>>>(iters, sink);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
        }
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        const double n = (double)p.multiProcessorCount * 2 * warps * iters * 8;
        printf("warps/CTA %2d (x2 CTAs/SM): %.3f ms, %.1f G IMMA.16832/s = %.2f per clk per SM at %.0f MHz, %.1f T MAC/s\n", warps, ms, n / ms / 1e6,
               n / (ms * 1e-3) / p.multiProcessorCount / (p.clockRate * 1e3), p.clockRate / 1e3, n * 4096 / ms / 1e9);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
