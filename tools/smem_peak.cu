// smem_peak.cu — measures the shared-memory pipe of this GPU: bytes per clock per SM that conflict-free LDS.128 and
// STS.128 sustain, the denominator of bench.py's roofline for the on-chip decoders (which are bound by shared-memory
// wavefronts, not by HBM).  Prints one JSON line.   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o smem_peak smem_peak.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

constexpr int kThreads = 512, kIters = 4096, kUnroll = 8;

template <bool STORE>
__global__ void __launch_bounds__(kThreads) smem_loop(float4 *sink, long long *cycles)
{
    extern __shared__ float4 cell[];
    const int tid = threadIdx.x;
    for (int i = tid; i < 4096; i += kThreads) cell[i] = make_float4((float)i, 1.f, 2.f, 3.f);
    __syncthreads();
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    // a quarter-warp touches 8 consecutive 16-byte cells = all 32 banks once: conflict-free 128-bit accesses
    int idx = tid;
    const long long t0 = clock64();
    for (int it = 0; it < kIters; ++it) {
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const int a = (idx + u * kThreads) & 4095;
            if (STORE) {
                cell[a] = acc;
                acc.x += 1.f;
            } else {
                const float4 v = cell[a];
                acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            }
        }
        idx = (idx + 37 * 8) & 4095;
    }
    const long long t1 = clock64();
    if (tid == 0) cycles[blockIdx.x] = t1 - t0;
    if (acc.x == -1.f) sink[blockIdx.x * kThreads + tid] = acc;      // keeps the loop alive
    if (STORE && cell[tid & 4095].x == -7.f) sink[0] = cell[tid];
}

template <bool STORE> double run(int sms, int ctas_per_sm, double *gbps, double *mhz)
{
    const int grid = sms * ctas_per_sm;
    float4 *sink; long long *cyc;
    cudaMalloc(&sink, sizeof(float4) * grid * kThreads);
    cudaMalloc(&cyc, sizeof(long long) * grid);
    cudaFuncSetAttribute(smem_loop<STORE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    smem_loop<STORE><<<grid, kThreads, 65536>>>(sink, cyc);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    smem_loop<STORE><<<grid, kThreads, 65536>>>(sink, cyc);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    long long *h = (long long *)malloc(sizeof(long long) * grid);
    cudaMemcpy(h, cyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
    double mean = 0;
    for (int i = 0; i < grid; ++i) mean += (double)h[i];
    mean /= grid;
    const double bytes_per_cta = (double)kThreads * kIters * kUnroll * 16.0;
    *gbps = bytes_per_cta * grid / (ms * 1e-3) / 1e9;
    free(h); cudaFree(sink); cudaFree(cyc);
    // the CTAs of the single wave run side by side for the whole launch: cycles one of them counted / launch time = SM clock
    *mhz = mean / (ms * 1e-3) / 1e6;                                  // a LOWER bound of the clock (launch ramp and tail are in ms)
    return 0.0;
}

int main()
{
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, 0) != cudaSuccess) { printf("{\"error\": \"no CUDA device\"}\n"); return 1; }
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    double g_ld = 0, g_st = 0, mhz_ld = 0, mhz_st = 0;
    const double ld = run<false>(p.multiProcessorCount, 2, &g_ld, &mhz_ld);
    const double st = run<true>(p.multiProcessorCount, 2, &g_st, &mhz_st);
    (void)ld; (void)st; (void)mhz_st;
    const double per_clk = 1e9 / (p.multiProcessorCount * (clk_khz / 1e3) * 1e6);      // GB/s -> bytes per clock per SM at the maximum SM clock
    printf("{\"gpu\": \"%s\", \"sm_count\": %d, \"sm_clock_mhz_max\": %.0f, \"sm_clock_mhz_lower_bound\": %.0f, "
           "\"lds128_bytes_per_clk_per_sm\": %.2f, \"sts128_bytes_per_clk_per_sm\": %.2f, \"lds128_GBps\": %.1f, \"sts128_GBps\": %.1f, "
           "\"how\": \"2 CTAs x 512 threads per SM, conflict-free 128-bit accesses, CUDA events around the launch; bytes per clock = GB/s / (SMs x maximum SM clock)\"}\n",
           p.name, p.multiProcessorCount, clk_khz / 1e3, mhz_ld, g_ld * per_clk, g_st * per_clk, g_ld, g_st);
    return 0;
}
