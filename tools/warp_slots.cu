// Which hardware warp slot (%warpid; scheduler = slot % 4) do the warps of two co-resident 320-thread CTAs get?
// Prints, for a few SMs, the slots of each CTA's warps and the warps-per-scheduler histogram of the SM.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(320, 2) probe(int *out, int iters)
{
    extern __shared__ unsigned char smem[];
    unsigned smid, wid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    asm volatile("mov.u32 %0, %%warpid;" : "=r"(wid));
    // keep the CTA alive until every CTA of the grid has started, so that two are co-resident on each SM
    volatile int *flag = out;
    if (threadIdx.x == 0) atomicAdd(out, 1);
    while (*flag < (int)gridDim.x) { }
    if ((threadIdx.x & 31) == 0) {
        int *rec = out + 1 + (blockIdx.x * 10 + (threadIdx.x >> 5)) * 2;
        rec[0] = (int)smid; rec[1] = (int)wid;
    }
    smem[threadIdx.x] = (unsigned char)iters;
}
int main()
{
    int dev = 0; cudaSetDevice(dev);
    cudaDeviceProp pr; cudaGetDeviceProperties(&pr, dev);
    const int grid = pr.multiProcessorCount * 2, smem_bytes = 100 * 1024;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    int *d; cudaMalloc(&d, (1 + grid * 20) * sizeof(int)); cudaMemset(d, 0, (1 + grid * 20) * sizeof(int));
    probe<<<grid, 320, smem_bytes>>>(d, 1);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("kernel failed\n"); return 1; }
    int *h = new int[1 + grid * 20]; cudaMemcpy(h, d, (1 + grid * 20) * sizeof(int), cudaMemcpyDeviceToHost);
    int worst = 0, best = 99, hist_all[4] = {0, 0, 0, 0};
    for (int sm = 0; sm < pr.multiProcessorCount; ++sm) {
        int hist[4] = {0, 0, 0, 0}, nc = 0;
        for (int b = 0; b < grid; ++b) {
            if (h[1 + b * 20] != sm) continue;
            ++nc;
            if (sm < 4) { printf("sm %d cta %d slots:", sm, b); for (int w = 0; w < 10; ++w) printf(" %d", h[1 + (b * 10 + w) * 2 + 1]); printf("\n"); }
            for (int w = 0; w < 10; ++w) hist[h[1 + (b * 10 + w) * 2 + 1] & 3]++;
        }
        int mx = 0; for (int k = 0; k < 4; ++k) { if (hist[k] > mx) mx = hist[k]; hist_all[k] += hist[k]; }
        if (sm < 4) printf("sm %d: %d CTAs, warps per scheduler %d %d %d %d\n", sm, nc, hist[0], hist[1], hist[2], hist[3]);
        if (nc == 2) { if (mx > worst) worst = mx; if (mx < best) best = mx; }
    }
    printf("max warps on one scheduler over the SMs with two CTAs: best %d worst %d; all SMs: %d %d %d %d\n", best, worst, hist_all[0], hist_all[1], hist_all[2], hist_all[3]);
    return 0;
}
