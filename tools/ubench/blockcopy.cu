// Calibration of the sweep kernels' memory pattern: how fast can this GPU move element blocks (a few KB each, at scattered rows) when the
// kernel does nothing else?  One warp per block, lanes on consecutive 8-byte (or 16-byte) words, every load of a block issued before its stores.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o blockcopy blockcopy.cu && ./blockcopy
// Prints GB/s (read + written bytes) for the cfg2 shape (10 496 blocks of 256 doubles -> 256) and the cfg5 shape (16 172 blocks of 729 -> 486),
// identity and random row order, one-shot and grid-stride CTAs, back-to-back launches over rotating buffers larger than L2.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <numeric>
#include <random>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("cuda error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

template <int VEC>
__global__ void __launch_bounds__(128) block_copy(const double * __restrict__ src, double * __restrict__ dst, const int * __restrict__ rows_in,
                                                  const int * __restrict__ rows_out, int n, int s_in, int s_out)
{
    const int lane = threadIdx.x & 31;
    for (int b = blockIdx.x * 4 + (threadIdx.x >> 5); b < n; b += gridDim.x * 4)
    {
        const double * __restrict__ x = src + (size_t)rows_in[b] * s_in;
        double * __restrict__ y = dst + (size_t)rows_out[b] * s_out;
        if (VEC == 2)
        {
            // 16 bytes per lane (needs even sizes)
            double2 v[12];
            const int nv = s_out / 2;
#pragma unroll
            for (int i = 0; i < 12; ++i) if (lane + 32 * i < nv) v[i] = __ldg(reinterpret_cast<const double2 *>(x) + lane + 32 * i);
#pragma unroll
            for (int i = 0; i < 12; ++i) if (lane + 32 * i < nv) reinterpret_cast<double2 *>(y)[lane + 32 * i] = v[i];
        }
        else
        {
            double v[24];
#pragma unroll
            for (int i = 0; i < 24; ++i) if (lane + 32 * i < s_in) v[i] = __ldg(x + lane + 32 * i);
#pragma unroll
            for (int i = 0; i < 24; ++i) if (lane + 32 * i < s_out) y[lane + 32 * i] = v[i];
        }
    }
}

static void run(const char * name, int n, int s_in, int s_out, bool random, int grid, int vec)
{
    const int nbuf = std::max(2, (int)(600e6 / (8.0 * n * (s_in + s_out))) + 1);
    std::vector<double *> src(nbuf), dst(nbuf);
    for (int i = 0; i < nbuf; ++i) { CK(cudaMalloc(&src[i], (size_t)n * s_in * 8)); CK(cudaMalloc(&dst[i], (size_t)n * s_out * 8)); CK(cudaMemset(src[i], 0, (size_t)n * s_in * 8)); }
    std::vector<int> ri(n), ro(n);
    std::iota(ri.begin(), ri.end(), 0); std::iota(ro.begin(), ro.end(), 0);
    if (random) { std::mt19937 g(1); std::shuffle(ri.begin(), ri.end(), g); std::shuffle(ro.begin(), ro.end(), g); }
    int * d_ri, * d_ro;
    CK(cudaMalloc(&d_ri, n * 4)); CK(cudaMalloc(&d_ro, n * 4));
    CK(cudaMemcpy(d_ri, ri.data(), n * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(d_ro, ro.data(), n * 4, cudaMemcpyHostToDevice));
    const int g = grid > 0 ? grid : (n + 3) / 4;
    auto launch = [&](int i)
    {
        if (vec == 2) block_copy<2><<<g, 128>>>(src[i % nbuf], dst[i % nbuf], d_ri, d_ro, n, s_in, s_out);
        else block_copy<1><<<g, 128>>>(src[i % nbuf], dst[i % nbuf], d_ri, d_ro, n, s_in, s_out);
    };
    for (int i = 0; i < 2 * nbuf; ++i) launch(i);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int reps = 10 * nbuf;
    CK(cudaEventRecord(e0));
    for (int i = 0; i < reps; ++i) launch(i);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    const double us = ms * 1e3 / reps, bytes = 8.0 * n * (s_in + (vec == 2 ? s_out : s_out));
    printf("%-34s %s rows, grid %5d, %2d B/lane: %7.2f us  %7.1f GB/s\n", name, random ? "random  " : "identity", g, vec * 8, us, bytes / us / 1e3);
    for (int i = 0; i < nbuf; ++i) { cudaFree(src[i]); cudaFree(dst[i]); }
    cudaFree(d_ri); cudaFree(d_ro);
}

int main()
{
    for (int random = 0; random < 2; ++random)
    {
        for (int grid : { 0, 148 * 8, 148 * 4 })
        {
            run("cfg2 block 256 -> 256 (10496)", 10496, 256, 256, random, grid, 1);
            run("cfg2 block 256 -> 256 (10496)", 10496, 256, 256, random, grid, 2);
            run("cfg5 block 729 -> 486 (16172)", 16172, 729, 486, random, grid, 1);
            run("cfg5 block 486 -> 486 (16172)", 16172, 486, 486, random, grid, 2);
        }
    }
    run("large block 256 -> 256 (500000)", 500000, 256, 256, 1, 0, 2);
    return 0;
}
