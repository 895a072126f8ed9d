// K1, column form (variant 8): FastMultiplyLU::transform_1D (reference source/FastMultiplyLU.cpp:436-512) with one THREAD per
// column of the (outer x inner) plane and the KT outputs of that column in registers.
//
//   dst[e][o][q][i] = coef * sum_{f in R_t(e), L/U ok} sum_k src[f][o][k][i] * B[pair(f,e)][k][q]   (+ old value)
//
// Why this form (ncu of sweep_tc_kernel, profiles/r02_sweep_kernels.md): the tensor-core kernels execute ~40 warp instructions per
// FP64 MMA -- staging loops, run decode, bulk-copy issue, fragment addressing, unit decode, barriers -- and sit at 43-54 % issue
// utilisation with DRAM at 13 %.  The arithmetic is ~2 flop/B, far below the FP64 ridge, and DFMA has the same FP64 rate as DMMA
// on this part, so the tensor core buys nothing once its bookkeeping outweighs the 8x instruction saving in the inner product.
// Here nothing is staged: a WARP owns a unit = (target element, group of 32*NC columns); lane l holds columns c0 + j*32 + l
// (j < NC), so every source load of the warp is NC*KF coalesced 256-byte rows straight from global memory / L1, the operator block
// of an entry is one broadcast line, and an entry costs NC*KF loads + KF*KT/2 operator loads + NC*KF*KT DFMAs.  Units are ordered
// fibre by fibre, so the re-reads of a source by the other targets of its fibre hit L1 (short fibres) or L2 (long ones).  The sources
// of the next DEPTH entries are in flight while the current one is multiplied.  The few targets with long entry lists (coarse elements
// of long fibres) are HEAVY units: one CTA per (target, 32 columns), one column per lane, the entry list cut into four contiguous
// parts, partial sums added in warp order through shared memory (deterministic).  No shared-memory staging, no block barrier on the
// normal path, no layout restrictions: any (KF, KT), any inner, odd or even; coef, accumulate, destination maps and accumulate-from
// are applied in the epilogue.
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

#include "kernels.cuh"

namespace amdg {

static const int COL_THREADS = 128;
static const int COL_NW = COL_THREADS / 32;
#ifndef AMDG_COL_MINB
#define AMDG_COL_MINB 4
#endif

// L1 prefetch.  prefetch.global.L1 does not allocate in L1 on this part (ncu: the L1 hit rate of the loads that follow does not move, the L2 hit
// rate does); an asynchronous 8-byte copy with .ca into a junk word of shared memory does: it has no destination register, nothing ever waits
// for it, and the sector it reads is in L1 when the real load arrives.
__device__ __forceinline__ void col_prefetch(const void * p, unsigned junk)
{
#ifdef AMDG_COL_PTX_PREFETCH
    asm volatile("prefetch.global.L1 [%0];" :: "l"(p));
#else
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(junk), "l"((const void *)((unsigned long long)p & ~7ull)) : "memory");
#endif
}

// entries of one target into the accumulators of this lane's columns.  `my` = entry of this lane (lane r holds entry r of the list: source
// element row, 1D pair id), cnt <= 32 entries; px = address of the lane's (column, k) in element row 0 (a 64-bit pointer each, so that the
// address of a load is ONE multiply-add: px + row * row_bytes).  DEPTH source rows are in flight in registers.
template <int KF, int KT, int NC, int DEPTH>
__device__ __forceinline__ void col_entries(double (&acc)[NC][KT], int2 my, int cnt, unsigned row_bytes, const double * __restrict__ blocks,
                                            const char * const (&px)[NC * KF])
{
    constexpr unsigned FULL = 0xffffffffu;
    constexpr bool VEC_B = ((KF * KT) & 1) == 0;
    auto ld = [&](int v, int e) -> double
    {
        return __ldg(reinterpret_cast<const double *>(px[v] + (unsigned long long)(unsigned)e * row_bytes));
    };
    double x[DEPTH][NC * KF];
#pragma unroll
    for (int d = 0; d < DEPTH; ++d)
    {
        const int e = __shfl_sync(FULL, my.x, min(d, cnt - 1));
        if (d < cnt)
        {
#pragma unroll
            for (int v = 0; v < NC * KF; ++v) x[d][v] = ld(v, e);
        }
    }
    for (int r0 = 0; r0 < cnt; r0 += DEPTH)
    {
#pragma unroll
        for (int d = 0; d < DEPTH; ++d)
        {
            const int r = r0 + d;
            if (r < cnt)
            {
                double xc[NC * KF];
#pragma unroll
                for (int v = 0; v < NC * KF; ++v) xc[v] = x[d][v];
                const int pair = __shfl_sync(FULL, my.y, r);
                const int e_next = __shfl_sync(FULL, my.x, min(r + DEPTH, cnt - 1));
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
                if (r + DEPTH < cnt)
                {
#pragma unroll
                    for (int v = 0; v < NC * KF; ++v) x[d][v] = ld(v, e_next);                     // in flight during the next DEPTH products
                }
#pragma unroll
                for (int k = 0; k < KF; ++k)
#pragma unroll
                    for (int q = 0; q < KT; ++q)
#pragma unroll
                        for (int j = 0; j < NC; ++j) acc[j][q] = fma(xc[j * KF + k], bv[k * KT + q], acc[j][q]);
            }
        }
    }
}

// offsets of the lane's columns of group g (NCX columns per lane): source pointers (column, k) in element row 0, destination offsets (column, q);
// -1 = column beyond the plane
template <int KF, int KT, int NCX>
__device__ __forceinline__ void col_offsets(const char * (&px)[NCX * KF], int (&offy)[NCX * KT], const double * __restrict__ src, int g, int W, int inner, int lane)
{
#pragma unroll
    for (int j = 0; j < NCX; ++j)
    {
        const int c = g * (32 * NCX) + j * 32 + lane;
        const bool on = c < W;
        const int cc = on ? c : 0;
        const int o = cc / inner, i = cc - o * inner;
#pragma unroll
        for (int k = 0; k < KF; ++k) px[j * KF + k] = reinterpret_cast<const char *>(src + ((o * KF + k) * inner + i));
#pragma unroll
        for (int q = 0; q < KT; ++q) offy[j * KT + q] = on ? (o * KT + q) * inner + i : -1;
    }
}

template <int KT, int NCX>
__device__ __forceinline__ void col_store(const double (&acc)[NCX][KT], const int (&offy)[NCX * KT], double * __restrict__ y, const double * __restrict__ yr,
                                          double coef, bool accumulate)
{
    // all old values of an accumulating sweep are fetched before the first store
    double old[NCX * KT];
#pragma unroll
    for (int v = 0; v < NCX * KT; ++v) old[v] = (accumulate && offy[v] >= 0) ? yr[offy[v]] : 0.0;
#pragma unroll
    for (int j = 0; j < NCX; ++j)
#pragma unroll
        for (int q = 0; q < KT; ++q)
            if (offy[j * KT + q] >= 0) y[offy[j * KT + q]] = fma(coef, acc[j][q], old[j * KT + q]);
}

// Normal CTAs run a software pipeline over the units of each warp, three stages deep, so that no load of the compute stage waits for DRAM:
//   stage A (unit u + 2 NW): the unit record;  (unit u + NW): its entry list (lane r <- entry r);
//   stage B (unit u + NW)  : prefetch into L1 -- the operator blocks of all its entries (one instruction, lane r <- block of entry r), every
//                            32-byte sector of its source rows (lane-per-sector, offsets from the host's table of the column group), and the old
//                            values of an accumulating sweep;
//   stage C (unit u)       : the products, sources and operator blocks now coming from L1, and the store.
template <int KF, int KT, int NC>
__global__ void __launch_bounds__(COL_THREADS, AMDG_COL_MINB) sweep_col_kernel(const ColArgs a)
{
#ifndef AMDG_COL_DEPTH
    constexpr int DEPTH = 1;                                        // source rows in flight in registers per warp (they come from L1)
#else
    constexpr int DEPTH = AMDG_COL_DEPTH;
#endif
    constexpr int DEPTH_H = KF <= 4 ? 4 : 2;
#ifndef AMDG_COL_PD
#define AMDG_COL_PD 2
#endif
    constexpr int PD = AMDG_COL_PD;                                 // units between the prefetch and the products
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int OPB = KF * KT * 8;                                // bytes of an operator block
    __shared__ double s_red[(COL_NW - 1) * KT * 32];
    __shared__ double s_junk[COL_THREADS];
    asm volatile("griddepcontrol.launch_dependents;");
    const int jb = blockIdx.y, comp = blockIdx.z;
    const SweepJob J = a.job[jb];
    const int inner = a.inner;
    const int W = J.outer * inner;
    const int s_from = W * KF, s_to = W * KT;
    const unsigned row_bytes = (unsigned)s_from * 8u;
    const double * __restrict__ src = J.src + (int64_t)comp * a.n_elem * s_from;
    double * __restrict__ dst = J.dst + (int64_t)comp * a.n_elem * s_to;
    const long long * __restrict__ dmap = J.dst_map;
    const double * __restrict__ accf = J.acc_from ? J.acc_from + (int64_t)comp * a.n_elem * s_to : nullptr;
    const bool accumulate = J.accumulate != 0;
    const double coef = J.coef;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const char * __restrict__ blocks_b = reinterpret_cast<const char *>(a.blocks);
    const unsigned junk = (unsigned)__cvta_generic_to_shared(&s_junk[threadIdx.x]);

    if ((int)blockIdx.x < a.n_heavy)
    {
        // heavy unit: (target, 32 columns), one column per lane, the entry list in COL_NW contiguous parts
        const int4 U = __ldg(reinterpret_cast<const int4 *>(a.units) + blockIdx.x);
        const char * px[KF]; int offy[KT];
        col_offsets<KF, KT, 1>(px, offy, src, U.w, W, inner, lane);
        double * y = dst + (dmap ? __ldg(dmap + U.x) : (long long)U.x * s_to);
        const double * yr = accf ? accf + (int64_t)U.x * s_to : y;
        const int chunk = (U.z + COL_NW - 1) / COL_NW;
        const int q0 = min(U.z, warp * chunk), q1 = min(U.z, q0 + chunk);
        asm volatile("griddepcontrol.wait;" ::: "memory");
        double acc[1][KT];
#pragma unroll
        for (int q = 0; q < KT; ++q) acc[0][q] = 0.0;
        for (int base = q0; base < q1; base += 32)
        {
            const int cnt = min(32, q1 - base);
            int2 my = make_int2(0, 0);
            if (lane < cnt)
            {
                my = __ldg(a.ent + U.y + base + lane);
#pragma unroll
                for (int b = 0; b < OPB; b += 32) col_prefetch(blocks_b + (int64_t)my.y * OPB + b, junk);
            }
            col_entries<KF, KT, 1, DEPTH_H>(acc, my, cnt, row_bytes, a.blocks, px);
        }
        if (warp > 0)
        {
#pragma unroll
            for (int q = 0; q < KT; ++q) s_red[((warp - 1) * KT + q) * 32 + lane] = acc[0][q];
        }
        __syncthreads();
        if (warp > 0) return;
#pragma unroll
        for (int w = 1; w < COL_NW; ++w)
#pragma unroll
            for (int q = 0; q < KT; ++q) acc[0][q] += s_red[((w - 1) * KT + q) * 32 + lane];
        col_store<KT, 1>(acc, offy, y, yr, coef, accumulate);
        return;
    }

    // normal CTA c: units [cta_ptr[c], cta_ptr[c+1]), warp w the units w, w + NW, ...
    const int cta = (int)blockIdx.x - a.n_heavy;
    const int u_first = __ldg(a.cta_ptr + cta), u_end = __ldg(a.cta_ptr + cta + 1);
    const int4 none = make_int4(-1, 0, 0, 0);
    auto load_unit = [&](int uu) -> int4 { return uu < u_end ? __ldg(reinterpret_cast<const int4 *>(a.units) + uu) : none; };
    auto load_ent = [&](const int4 & U) -> int2 { return (U.x >= 0 && lane < U.z) ? __ldg(a.ent + U.y + lane) : make_int2(0, 0); };
    int u = u_first + warp;
    // unit queue: Uq[i] = unit u + i NW (record), mq[i] = its entries; prefetches run PD units ahead of the products
    int4 Uq[PD + 2]; int2 mq[PD + 1];
#pragma unroll
    for (int i = 0; i < PD + 2; ++i) Uq[i] = load_unit(u + i * COL_NW);
#pragma unroll
    for (int i = 0; i < PD + 1; ++i) mq[i] = load_ent(Uq[i]);
    int g_cur = -1, g_pf = -1;
    const char * px[NC * KF]; int offy[NC * KT];
    int pfo[COL_PFR];                                               // prefetch offsets (bytes from the row start) of this lane, -1 = none
    auto issue_prefetch = [&](const int4 & U, int2 my)
    {
        if (U.w != g_pf)
        {
            g_pf = U.w;
            const int * __restrict__ t = a.pf + (int64_t)U.w * (32 * COL_PFR);
#pragma unroll
            for (int i = 0; i < COL_PFR; ++i) pfo[i] = __ldg(t + i * 32 + lane);
        }
        if (lane < U.z)
        {
#pragma unroll
            for (int b = 0; b < OPB; b += 32) col_prefetch(blocks_b + (int64_t)my.y * OPB + b, junk);
        }
        const char * __restrict__ sb = reinterpret_cast<const char *>(src);
        for (int r = 0; r < U.z; ++r)
        {
            const int e = __shfl_sync(FULL, my.x, r);
            const char * rowp = sb + (unsigned long long)(unsigned)e * row_bytes;
#pragma unroll
            for (int i = 0; i < COL_PFR; ++i)
                if (pfo[i] >= 0) col_prefetch(rowp + (unsigned)pfo[i], junk);
        }
        if (accumulate && U.w == g_cur)
        {
            const double * yr = accf ? accf + (int64_t)U.x * s_to : dst + (dmap ? __ldg(dmap + U.x) : (long long)U.x * s_to);
#pragma unroll
            for (int v = 0; v < NC * KT; ++v)
                if (offy[v] >= 0) col_prefetch(yr + offy[v], junk);
        }
    };
    // everything above reads host-written tables only; the coefficient arrays may still be written by the previous kernel of the stream
    asm volatile("griddepcontrol.wait;" ::: "memory");
#pragma unroll
    for (int i = 0; i < PD; ++i)
        if (Uq[i].x >= 0) issue_prefetch(Uq[i], mq[i]);
    while (Uq[0].x >= 0)
    {
        const int4 Unew = load_unit(u + (PD + 2) * COL_NW);
        const int2 mnew = load_ent(Uq[PD + 1]);
        const int4 Ua = Uq[0]; const int2 ma = mq[0];
        if (Ua.w != g_cur) { g_cur = Ua.w; col_offsets<KF, KT, NC>(px, offy, src, Ua.w, W, inner, lane); }
        if (Uq[PD].x >= 0) issue_prefetch(Uq[PD], mq[PD]);
        double * y = dst + (dmap ? __ldg(dmap + Ua.x) : (long long)Ua.x * s_to);
        const double * yr = accf ? accf + (int64_t)Ua.x * s_to : y;
        double acc[NC][KT];
#pragma unroll
        for (int j = 0; j < NC; ++j)
#pragma unroll
            for (int q = 0; q < KT; ++q) acc[j][q] = 0.0;
        if (Ua.z > 0) col_entries<KF, KT, NC, DEPTH>(acc, ma, Ua.z, row_bytes, a.blocks, px);
        col_store<KT, NC>(acc, offy, y, yr, coef, accumulate);
#pragma unroll
        for (int i = 0; i < PD + 1; ++i) Uq[i] = Uq[i + 1];
        Uq[PD + 1] = Unew;
#pragma unroll
        for (int i = 0; i < PD; ++i) mq[i] = mq[i + 1];
        mq[PD] = mnew;
        u += COL_NW;
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
}

#ifdef AMDG_COL_PROBE
// SASS inspection builds (nvcc -DAMDG_COL_PROBE -cubin): two instantiations only
template __global__ void sweep_col_kernel<4, 4, 2>(const ColArgs);
template __global__ void sweep_col_kernel<3, 2, 4>(const ColArgs);
#else
template <int KF, int KT, int NC>
static cudaError_t launch_col_t(const ColArgs & a, cudaStream_t st)
{
    static const bool pdl = !(std::getenv("AMDG_TC_PDL") && std::atoi(std::getenv("AMDG_TC_PDL")) == 0);
    const int n_cta = a.n_heavy + a.n_cta;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)std::max(n_cta, 1), (unsigned)a.n_job, (unsigned)a.n_comp);
    cfg.blockDim = dim3(COL_THREADS, 1, 1);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, sweep_col_kernel<KF, KT, NC>, a);
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
