// K1, column form (variant 8): FastMultiplyLU::transform_1D (reference source/FastMultiplyLU.cpp:436-512) with one THREAD per
// column of the (outer x inner) plane and the KT outputs of that column in registers.
//
//   dst[e][o][q][i] = coef * sum_{f in R_t(e), L/U ok} sum_k src[f][o][k][i] * B[pair(f,e)][k][q]   (+ old value)
//
// Why this form (ncu of sweep_tc_kernel, profiles/r02_sweep_kernels.md): the tensor-core kernels execute ~40 warp instructions per
// FP64 MMA -- staging loops, run decode, bulk-copy issue, fragment addressing, unit decode, barriers.  The arithmetic is ~2 flop/B,
// far below the FP64 ridge, and DFMA has the same FP64 rate as DMMA on this part, so the tensor core buys nothing once its
// bookkeeping outweighs the 8x instruction saving in the inner product.  Here nothing is staged: a WARP owns a unit = (target
// element, group of 32*NC columns); lane l holds columns c0 + j*32 + l (j < NC), so every source load of the warp is NC*KF coalesced
// 256-byte rows straight from global memory / L1, the operator block of an entry is one broadcast line, and an entry costs NC*KF
// loads + KF*KT/2 operator loads + NC*KF*KT DFMAs.  Units are ordered fibre by fibre, so the re-reads of a source by the other
// targets of its fibre hit L1 (short fibres) or L2 (long ones).  The source of the next entry is fetched while the current one is
// multiplied.  The few targets with long entry lists (coarse elements of long fibres) are HEAVY units: one CTA each, the entry list cut
// into four contiguous parts, partial sums added in warp order through shared memory (deterministic).  No shared-memory staging,
// no block barrier on the normal path, no layout restrictions: any (KF, KT), any inner, odd or even; coef, accumulate, destination
// maps and accumulate-from are applied in the epilogue.
//
// Measured (B200, profiles/r02_sweep_kernels.md): faster than the lean tensor-core kernel where blocks are large and K is small (the 6-D
// shapes: <3,2> on 729-double blocks 57 vs 82 us, <3,3> 63 vs 89 us), slower on 256-double <4,4> blocks (long entry lists of the L sweeps)
// and on blocks below ~200 doubles; the auto mode of the context picks per shape (csrc/capi.cu: launch_sweep).  Variants of this kernel with
// a software pipeline over units, L1 prefetch by asynchronous copies and persistent CTAs were measured and were not faster (same file).
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

#include "kernels.cuh"

namespace amdg {

static const int COL_THREADS = 128;
static const int COL_NW = COL_THREADS / 32;

template <int KF, int NC>
__device__ __forceinline__ void col_load(double (&x)[NC][KF], const double * __restrict__ row, const int (&offf)[NC], int inner)
{
#pragma unroll
    for (int j = 0; j < NC; ++j)
    {
        const double * __restrict__ p = row + offf[j];
#pragma unroll
        for (int k = 0; k < KF; ++k) x[j][k] = __ldg(p + (int64_t)k * inner);
    }
}

// entries [p0, p0 + n) of one target into the accumulators of this lane's columns
template <int KF, int KT, int NC>
__device__ __forceinline__ void col_entries(double (&acc)[NC][KT], const int2 * __restrict__ ent, int p0, int n, const double * __restrict__ src,
                                            int64_t s_from, const double * __restrict__ blocks, const int (&offf)[NC], int inner, int lane, unsigned junk)
{
    constexpr unsigned FULL = 0xffffffffu;
    constexpr bool VEC_B = ((KF * KT) & 1) == 0;
    for (int base = 0; base < n; base += 32)
    {
        const int cnt = min(32, n - base);
        int2 my = make_int2(0, 0);
        if (lane < cnt)
        {
            my = __ldg(ent + p0 + base + lane);                     // lane r holds entry r: (source element row, 1D pair id)
#ifdef AMDG_COL_TOUCH
            // operator blocks of all entries of the batch towards L1: an asynchronous copy into a junk word (no destination register, nothing waits)
#pragma unroll
            for (int b = 0; b < KF * KT * 8; b += 32)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(junk), "l"(reinterpret_cast<const char *>(blocks) + (int64_t)my.y * (KF * KT * 8) + b) : "memory");
#endif
        }
        double xn[NC][KF];
        col_load<KF, NC>(xn, src + (int64_t)__shfl_sync(FULL, my.x, 0) * s_from, offf, inner);
        for (int r = 0; r < cnt; ++r)
        {
            double xc[NC][KF];
#pragma unroll
            for (int j = 0; j < NC; ++j)
#pragma unroll
                for (int k = 0; k < KF; ++k) xc[j][k] = xn[j][k];
            const int pair = __shfl_sync(FULL, my.y, r);
            const int e_next = __shfl_sync(FULL, my.x, min(r + 1, cnt - 1));
            if (r + 1 < cnt) col_load<KF, NC>(xn, src + (int64_t)e_next * s_from, offf, inner);      // in flight during the products below
            const double * __restrict__ B = blocks + (int64_t)pair * (KF * KT);
            double bv[KF * KT];
            if (VEC_B)
            {
#pragma unroll
                for (int v = 0; v < KF * KT / 2; ++v)
                {
                    const double2 t2 = __ldg(reinterpret_cast<const double2 *>(B) + v);
                    bv[2 * v] = t2.x; bv[2 * v + 1] = t2.y;
                }
            }
            else
            {
#pragma unroll
                for (int v = 0; v < KF * KT; ++v) bv[v] = __ldg(B + v);
            }
#pragma unroll
            for (int k = 0; k < KF; ++k)
#pragma unroll
                for (int q = 0; q < KT; ++q)
#pragma unroll
                    for (int j = 0; j < NC; ++j) acc[j][q] = fma(xc[j][k], bv[k * KT + q], acc[j][q]);
        }
    }
}

// DUAL: some job of the launch has a second destination (SweepJob::dst2).  A template parameter because the kernel's register allocation is fragile:
// the extra pointer and branch in the epilogue of the plain form cost 10 % of the cfg5 stage (17.1 -> 18.85 ms, gpurun call AK)
template <int KF, int KT, int NC, bool DUAL>
__global__ void __launch_bounds__(COL_THREADS) sweep_col_kernel(const ColArgs a)
{
    __shared__ double s_red[(COL_NW - 1) * NC * KT * 32];
    __shared__ double s_junk[COL_THREADS];
    asm volatile("griddepcontrol.launch_dependents;");
    const int jb = blockIdx.y, comp = blockIdx.z;
    const SweepJob J = a.job[jb];
    const int inner = a.inner;
    const int W = J.outer * inner;
    const int64_t s_from = (int64_t)W * KF, s_to = (int64_t)W * KT;
    const double * __restrict__ src = J.src + (int64_t)comp * a.n_elem * s_from;
    double * __restrict__ dst = J.dst + (int64_t)comp * a.n_elem * s_to;
    const long long * __restrict__ dmap = J.dst_map;
    const double * __restrict__ accf = J.acc_from ? J.acc_from + (int64_t)comp * a.n_elem * s_to : nullptr;
    const long long * __restrict__ dmap2 = DUAL ? J.dst2_map : nullptr;            // second copy of the output (one component per job), see SweepJob::dst2
    const bool accumulate = J.accumulate != 0;
    const double coef = J.coef;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned junk = (unsigned)__cvta_generic_to_shared(&s_junk[threadIdx.x]);
    const bool heavy = (int)blockIdx.x < a.n_heavy;
    // units of this warp: a heavy CTA runs one unit with all its warps, a normal CTA `upc` consecutive units, warp w the units w, w + NW, ...
    int u, u_end, u_step;
    if (heavy) { u = blockIdx.x; u_end = u + 1; u_step = 1; }
    else
    {
        const int first = a.n_heavy + ((int)blockIdx.x - a.n_heavy) * a.upc;
        u = first + warp; u_end = min(first + a.upc, a.n_unit); u_step = COL_NW;
    }
    int g_cur = -1;
    int offf[NC], offt[NC];
    bool first_unit = true;
    for (; u < u_end; u += u_step)
    {
        const int4 U = __ldg(reinterpret_cast<const int4 *>(a.units) + u);      // (target element row, first entry, entries, column group)
        if (U.w != g_cur)
        {
            g_cur = U.w;
#pragma unroll
            for (int j = 0; j < NC; ++j)
            {
                const int c = U.w * (32 * NC) + j * 32 + lane;
                const bool on = c < W;
                const int cc = on ? c : 0;
                const int o = cc / inner, i = cc - o * inner;
                offf[j] = o * KF * inner + i;
                offt[j] = on ? o * KT * inner + i : -1;
            }
        }
        double * y = dst + (dmap ? __ldg(dmap + U.x) : (long long)U.x * s_to);
        const double * yr = accf ? accf + (int64_t)U.x * s_to : y;
        if (first_unit)
        {
            // everything above reads host-written tables only; the coefficient arrays may still be written by the previous kernel of the stream
            asm volatile("griddepcontrol.wait;" ::: "memory");
            first_unit = false;
        }
        if (accumulate && (!heavy || warp == 0))
        {
            // an accumulating sweep reads its destination: start those lines on their way now, the entry loop hides the trip
#pragma unroll
            for (int j = 0; j < NC; ++j)
                if (offt[j] >= 0)
                {
#pragma unroll
                    for (int q = 0; q < KT; ++q) asm volatile("prefetch.global.L1 [%0];" :: "l"(yr + offt[j] + (int64_t)q * inner));
                }
        }
        double acc[NC][KT];
#pragma unroll
        for (int j = 0; j < NC; ++j)
#pragma unroll
            for (int q = 0; q < KT; ++q) acc[j][q] = 0.0;
        if (!heavy) col_entries<KF, KT, NC>(acc, a.ent, U.y, U.z, src, s_from, a.blocks, offf, inner, lane, junk);
        else
        {
            const int chunk = (U.z + COL_NW - 1) / COL_NW;
            const int q0 = min(U.z, warp * chunk), q1 = min(U.z, q0 + chunk);
            if (q1 > q0) col_entries<KF, KT, NC>(acc, a.ent, U.y + q0, q1 - q0, src, s_from, a.blocks, offf, inner, lane, junk);
            if (warp > 0)
            {
#pragma unroll
                for (int j = 0; j < NC; ++j)
#pragma unroll
                    for (int q = 0; q < KT; ++q) s_red[(((warp - 1) * NC + j) * KT + q) * 32 + lane] = acc[j][q];
            }
            __syncthreads();
            if (warp > 0) break;
#pragma unroll
            for (int w = 1; w < COL_NW; ++w)
#pragma unroll
                for (int j = 0; j < NC; ++j)
#pragma unroll
                    for (int q = 0; q < KT; ++q) acc[j][q] += s_red[(((w - 1) * NC + j) * KT + q) * 32 + lane];
        }
        // epilogue: all old values of an accumulating sweep are fetched before the first store
        double old[NC][KT];
#pragma unroll
        for (int j = 0; j < NC; ++j)
#pragma unroll
            for (int q = 0; q < KT; ++q) old[j][q] = (accumulate && offt[j] >= 0) ? yr[offt[j] + (int64_t)q * inner] : 0.0;
        double * y2 = (DUAL && dmap2) ? J.dst2 + __ldg(dmap2 + U.x) : nullptr;
#pragma unroll
        for (int j = 0; j < NC; ++j)
            if (offt[j] >= 0)
            {
#pragma unroll
                for (int q = 0; q < KT; ++q)
                {
                    const double v = fma(coef, acc[j][q], old[j][q]);
                    y[offt[j] + (int64_t)q * inner] = v;
                    if (DUAL && y2) y2[offt[j] + (int64_t)q * inner] = v;
                }
            }
    }
#ifdef AMDG_COL_TOUCH
    asm volatile("cp.async.wait_all;" ::: "memory");
#endif
}

#ifdef AMDG_COL_PROBE
// SASS inspection builds (nvcc -DAMDG_COL_PROBE -cubin): two instantiations only
template __global__ void sweep_col_kernel<4, 4, 2, false>(const ColArgs);
template __global__ void sweep_col_kernel<3, 2, 4, false>(const ColArgs);
template __global__ void sweep_col_kernel<3, 2, 4, true>(const ColArgs);
#else
template <int KF, int KT, int NC>
static cudaError_t launch_col_t(const ColArgs & a, cudaStream_t st)
{
    static const bool pdl = !(std::getenv("AMDG_TC_PDL") && std::atoi(std::getenv("AMDG_TC_PDL")) == 0);
    const int n_normal = a.n_unit - a.n_heavy;
    const int n_cta = a.n_heavy + (n_normal + a.upc - 1) / a.upc;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)std::max(n_cta, 1), (unsigned)a.n_job, (unsigned)a.n_comp);
    cfg.blockDim = dim3(COL_THREADS, 1, 1);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
    bool dual = false;
    for (int i = 0; i < a.n_job; ++i) dual = dual || a.job[i].dst2 != nullptr;
    if constexpr (KF <= 4 && KT <= 4)
    {
        if (dual) return cudaLaunchKernelEx(&cfg, sweep_col_kernel<KF, KT, NC, true>, a);
    }
    else if (dual) return cudaErrorInvalidValue;                 // capi.cu routes such launches through the scatter fall-back
    return cudaLaunchKernelEx(&cfg, sweep_col_kernel<KF, KT, NC, false>, a);
}

int col_max_nc(int kf, int kt) { return (kf <= 4 && kt <= 4) ? 4 : 2; }

template <int KF, int KT>
static cudaError_t launch_col_nc(const ColArgs & a, int nc, cudaStream_t st)
{
    if (nc == 1) return launch_col_t<KF, KT, 1>(a, st);
    if (nc == 2) return launch_col_t<KF, KT, 2>(a, st);
    if constexpr (KF <= 4 && KT <= 4)
    {
        if (nc == 3) return launch_col_t<KF, KT, 3>(a, st);
        if (nc == 4) return launch_col_t<KF, KT, 4>(a, st);
    }
    return cudaErrorInvalidValue;
}

#define AMDG_DISPATCH_KT_COL(KF_)                                                                  \
    switch (kt) {                                                                                  \
        case 1: return launch_col_nc<KF_, 1>(a, nc, st); case 2: return launch_col_nc<KF_, 2>(a, nc, st);  \
        case 3: return launch_col_nc<KF_, 3>(a, nc, st); case 4: return launch_col_nc<KF_, 4>(a, nc, st);  \
        case 5: return launch_col_nc<KF_, 5>(a, nc, st); case 6: return launch_col_nc<KF_, 6>(a, nc, st);  \
        default: return cudaErrorInvalidValue; }

cudaError_t launch_sweep_col(const ColArgs & a, int kf, int kt, int nc, cudaStream_t st)
{
    switch (kf)
    {
        case 1: AMDG_DISPATCH_KT_COL(1) case 2: AMDG_DISPATCH_KT_COL(2) case 3: AMDG_DISPATCH_KT_COL(3)
        case 4: AMDG_DISPATCH_KT_COL(4) case 5: AMDG_DISPATCH_KT_COL(5) case 6: AMDG_DISPATCH_KT_COL(6)
        default: return cudaErrorInvalidValue;
    }
}
#endif

}  // namespace amdg
