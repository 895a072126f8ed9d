// micro-benchmark: FP64 mma.sync m8n8k4 vs DFMA throughput on the current GPU
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_dmma(double * out, int iters)
{
    double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-4;
    double c0[2] = {0, 0}, c1[2] = {0, 0}, c2[2] = {0, 0}, c3[2] = {0, 0};
    for (int i = 0; i < iters; ++i)
    {
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0[0]), "+d"(c0[1]) : "d"(a), "d"(b));
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c1[0]), "+d"(c1[1]) : "d"(a), "d"(b));
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c2[0]), "+d"(c2[1]) : "d"(a), "d"(b));
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c3[0]), "+d"(c3[1]) : "d"(a), "d"(b));
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = c0[0] + c0[1] + c1[0] + c1[1] + c2[0] + c2[1] + c3[0] + c3[1];
}
__global__ void k_dfma(double * out, int iters)
{
    double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-4;
    double c[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < iters; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) c[j] = fma(a, b, c[j]);
    double s = 0; for (int j = 0; j < 8; ++j) s += c[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// latency: one warp, a dependent chain of DMMAs (ILP accumulators interleaved)
template <int ILP>
__global__ void k_dmma_lat(long long * out, double * sink, int iters)
{
    double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-4;
    double c[ILP][2];
    for (int j = 0; j < ILP; ++j) { c[j][0] = 0; c[j][1] = 0; }
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i)
#pragma unroll
        for (int j = 0; j < ILP; ++j)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[j][0]), "+d"(c[j][1]) : "d"(a), "d"(b));
    long long t1 = clock64();
    double s = 0; for (int j = 0; j < ILP; ++j) s += c[j][0] + c[j][1];
    sink[threadIdx.x] = s;
    if (threadIdx.x == 0) out[0] = t1 - t0;
}
__global__ void k_dfma_lat(long long * out, double * sink, int iters)
{
    double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-4, c = 0;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) c = fma(a, b, c);
    long long t1 = clock64();
    sink[threadIdx.x] = c;
    if (threadIdx.x == 0) out[0] = t1 - t0;
}
int main()
{
    {
        long long * d; double * sink; cudaMalloc(&d, 8); cudaMalloc(&sink, 32 * 8); long long h;
        const int iters = 4096;
        k_dmma_lat<1><<<1, 32>>>(d, sink, iters); cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost); printf("DMMA dependent chain: %.1f cycles per DMMA\n", (double)h / iters);
        k_dmma_lat<2><<<1, 32>>>(d, sink, iters); cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost); printf("DMMA 2 chains: %.1f cycles per DMMA\n", (double)h / iters / 2);
        k_dmma_lat<4><<<1, 32>>>(d, sink, iters); cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost); printf("DMMA 4 chains: %.1f cycles per DMMA\n", (double)h / iters / 4);
        k_dmma_lat<8><<<1, 32>>>(d, sink, iters); cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost); printf("DMMA 8 chains: %.1f cycles per DMMA\n", (double)h / iters / 8);
        k_dfma_lat<<<1, 32>>>(d, sink, iters); cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost); printf("DFMA dependent chain: %.1f cycles per DFMA\n", (double)h / iters);
    }
    double * out; cudaMalloc(&out, 148 * 8 * 1024 * sizeof(double));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int warps = 4; warps <= 32; warps *= 2)
    {
        const int iters = 20000; float ms;
        k_dmma<<<148, warps * 32>>>(out, 100); cudaDeviceSynchronize();
        cudaEventRecord(e0); k_dmma<<<148, warps * 32>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        double fma_mma = 148.0 * warps * iters * 4 * 256;
        printf("DMMA  %2d warps/SM: %.2f TFMA/s (%.1f FMA/clk/SM @1.9GHz)\n", warps, fma_mma / ms / 1e9, fma_mma / (ms * 1e-3) / 148 / 1.9e9);
        k_dfma<<<148, warps * 32>>>(out, 100); cudaDeviceSynchronize();
        cudaEventRecord(e0); k_dfma<<<148, warps * 32>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        double fma_f = 148.0 * warps * 32 * iters * 8.0;
        printf("DFMA  %2d warps/SM: %.2f TFMA/s (%.1f FMA/clk/SM @1.9GHz)\n", warps, fma_f / ms / 1e9, fma_f / (ms * 1e-3) / 148 / 1.9e9);
    }
    return 0;
}
