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
int main()
{
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
