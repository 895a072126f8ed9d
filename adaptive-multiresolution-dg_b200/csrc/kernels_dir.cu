// K1, register-direct tensor-core form: FastMultiplyLU::transform_1D (reference source/FastMultiplyLU.cpp:436-512) with no
// shared-memory staging and no block barrier.
//
// Why a third form: ncu of sweep_tc_kernel (profiles/r01_sweep_tc_ncu.md) shows ~45 warp instructions per FP64 MMA, 0.7 eligible
// warps per scheduler and DRAM at 9-17 %: one-shot CTAs spend their life in staging loops, index decode and barriers.  Most
// elements of the named grids sit on short fibres (cfg2: 71 % on fibres of <= 8 elements, cfg5: 86 %), whose whole operator is a
// handful of 8x4 tiles.  Here a WARP owns a unit (dir_items.hpp): it loads each source's B fragments (4 source indices x 8
// columns, 8 bytes per lane) straight from global memory into registers, eight loads in flight per warp, feeds them to the row
// tiles named by the source's bit mask (A fragments stream from L1) and stores the C fragments from registers.  All index
// arithmetic of a column tile is two table look-ups (tab_b, tab_c: built once per block shape on the host), so the kernel is
// the same code for every (KF, KT).  coef is folded into the B fragments and the old value of an accumulating sweep initialises
// the accumulator, so the epilogue is a plain store -- through an optional per-element destination map (SweepJob::dst_map), which
// is how the multi-GPU path stores straight into the peer that owns the element in the next layout.
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

#include "kernels.cuh"

namespace amdg {

#ifndef AMDG_DIR_MIN_CTAS
#define AMDG_DIR_MIN_CTAS 5
#endif
static const int DIR_THREADS = 128;

__device__ __forceinline__ void dir_dmma(double (&c)[2], double a, double b)
{
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

struct DirU { int pool_ofs, fib_ofs, nfib, m, ct0, nct, n_src, prog, variant, n_rt, rt[4]; };

// NRT row tiles x G column tiles per group; NA > 1: one tile, entries spread over NA accumulators (added in a fixed order)
template <int NRT, int G, int NA>
__device__ __forceinline__ void dir_run(const DirArgs & a, const DirU & U, const double * __restrict__ src, double * __restrict__ dst,
                                        const long long * __restrict__ dst_map, const double * __restrict__ acc_from, int64_t s_from, int64_t s_to,
                                        double coef, bool accumulate, bool vec, int lane)
{
    constexpr int SB = NA > 1 ? 8 : 8 / G;                         // sources per batch: eight B fragments in flight
    constexpr unsigned FULL = 0xffffffffu;
    const double * __restrict__ Ag = a.a_tab[U.prog];
    const int * __restrict__ pool = a.pool + U.pool_ofs;
    const int g_lane = (lane >> 2) >> a.tg_shift;                  // target slot of this lane's C row
    const int kk = lane & 3;
    const int dkp = (min(4 + kk, a.kf - 1) - min(kk, a.kf - 1)) * a.inner;     // second k-part (KF > 4)
    const bool scale = coef != 1.0;
    int code0 = 0, msk0 = 0;
    if (lane < U.n_src) { code0 = __ldg(pool + lane); msk0 = __ldg(pool + U.n_src + lane); }
    // everything above reads host-written tables only; the coefficient arrays may still be written by the previous kernel
    asm volatile("griddepcontrol.wait;" ::: "memory");
    for (int b = 0; b < U.nfib; ++b)
    {
        const int fo = U.fib_ofs + b * U.m;
        long long yoff[NRT]; bool ton[NRT]; int erow[NRT];
#pragma unroll
        for (int r = 0; r < NRT; ++r)
        {
            const int tl = U.rt[r] * a.tg + g_lane;
            ton[r] = r < U.n_rt && tl < U.m;
            const int e = ton[r] ? __ldg(a.elem_pool + fo + tl) : 0;
            erow[r] = e;
            yoff[r] = dst_map ? __ldg(dst_map + e) : (long long)e * s_to;
        }
        const int srow0 = lane < U.n_src ? __ldg(a.elem_pool + fo + (code0 >> 1)) : 0;
        const int ct_end = U.ct0 + U.nct;
        for (int ctg = U.ct0; ctg < ct_end; ctg += G)
        {
            const int gv = min(G, ct_end - ctg);
            int bo[G];
#pragma unroll
            for (int j = 0; j < G; ++j) bo[j] = __ldg(a.tab_b + (ctg + j) * 32 + lane);
            // the destination offsets are looked up where they are used (L1 hits) instead of living in registers across the source loop
            auto cofs = [&](int j) -> int2 { int2 v = __ldg(a.tab_c + (ctg + j) * 32 + lane); if (j >= gv) v.x = -1; return v; };
            double acc[NRT][G][NA][2];
#pragma unroll
            for (int r = 0; r < NRT; ++r)
#pragma unroll
                for (int j = 0; j < G; ++j)
#pragma unroll
                    for (int q = 0; q < NA; ++q) { acc[r][j][q][0] = 0.0; acc[r][j][q][1] = 0.0; }
            if (accumulate)
            {
#pragma unroll
                for (int r = 0; r < NRT; ++r)
                {
                    if (!ton[r]) continue;
                    const double * y = acc_from ? acc_from + (long long)erow[r] * s_to : dst + yoff[r];
#pragma unroll
                    for (int j = 0; j < G; ++j)
                    {
                        const int2 co = cofs(j);
                        if (co.x < 0) continue;
                        if (vec) { const double2 v = *reinterpret_cast<const double2 *>(y + co.x); acc[r][j][0][0] = v.x; acc[r][j][0][1] = v.y; }
                        else { acc[r][j][0][0] = y[co.x]; if (co.y >= 0) acc[r][j][0][1] = y[co.y]; }
                    }
                }
            }
            int abase = 0;
            for (int s0 = 0; s0 < U.n_src; s0 += 32)
            {
                int code = code0, msk = msk0, srow = srow0;
                if (s0 > 0)
                {
                    const int sl = s0 + lane;
                    code = 0; msk = 0; srow = 0;
                    if (sl < U.n_src) { code = __ldg(pool + sl); msk = __ldg(pool + U.n_src + sl); srow = __ldg(a.elem_pool + fo + (code >> 1)); }
                }
                const int ns = min(32, U.n_src - s0);
                for (int i0 = 0; i0 < ns; i0 += SB)
                {
                    double bv[SB][G]; int mk[SB];
#pragma unroll
                    for (int ii = 0; ii < SB; ++ii)
                    {
                        const int i = i0 + ii;
                        const int row = __shfl_sync(FULL, srow, i);
                        const int cd = __shfl_sync(FULL, code, i);
                        mk[ii] = __shfl_sync(FULL, msk, i);
                        if (mk[ii])
                        {
                            const double * __restrict__ bp = src + (int64_t)row * s_from + ((cd & 1) ? dkp : 0);
#pragma unroll
                            for (int j = 0; j < G; ++j) if (j < gv) { bv[ii][j] = __ldg(bp + bo[j]); if (scale) bv[ii][j] *= coef; }
                        }
                    }
#pragma unroll
                    for (int ii = 0; ii < SB; ++ii)
                    {
                        if (!mk[ii]) continue;
#pragma unroll
                        for (int r = 0; r < NRT; ++r)
                        {
                            if (!((mk[ii] >> r) & 1)) continue;
                            const double av = __ldg(Ag + (int64_t)(abase + __popc(mk[ii] & ((1 << r) - 1))) * 32 + lane);
#pragma unroll
                            for (int j = 0; j < G; ++j) if (j < gv) dir_dmma(acc[r][j][NA > 1 ? (ii & (NA - 1)) : 0], av, bv[ii][j]);
                        }
                        abase += __popc(mk[ii]);
                    }
                }
            }
#pragma unroll
            for (int r = 0; r < NRT; ++r)
            {
                if (!ton[r]) continue;
                double * y = dst + yoff[r];
#pragma unroll
                for (int j = 0; j < G; ++j)
                {
                    const int2 co = cofs(j);
                    if (co.x < 0) continue;
                    double v0 = acc[r][j][0][0], v1 = acc[r][j][0][1];
                    if (NA == 4)
                    {
                        v0 = (acc[r][j][0][0] + acc[r][j][1][0]) + (acc[r][j][2][0] + acc[r][j][3][0]);
                        v1 = (acc[r][j][0][1] + acc[r][j][1][1]) + (acc[r][j][2][1] + acc[r][j][3][1]);
                    }
                    if (vec) *reinterpret_cast<double2 *>(y + co.x) = make_double2(v0, v1);
                    else { y[co.x] = v0; if (co.y >= 0) y[co.y] = v1; }
                }
            }
        }
    }
}

__global__ void __launch_bounds__(DIR_THREADS, AMDG_DIR_MIN_CTAS) sweep_dir_kernel(const DirArgs a)
{
    // programmatic dependent launch: the next kernel's CTAs may start their (table-only) prologue under this grid's tail
    asm volatile("griddepcontrol.launch_dependents;");
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int u = blockIdx.x * (DIR_THREADS / 32) + warp;
    if (u >= a.n_unit) return;
    int hv = 0;
    if (lane < 14) hv = __ldg(reinterpret_cast<const int *>(a.units + u) + lane);
    DirU U;
    U.pool_ofs = __shfl_sync(0xffffffffu, hv, 0); U.fib_ofs = __shfl_sync(0xffffffffu, hv, 1); U.nfib = __shfl_sync(0xffffffffu, hv, 2);
    U.m = __shfl_sync(0xffffffffu, hv, 3); U.ct0 = __shfl_sync(0xffffffffu, hv, 4); U.nct = __shfl_sync(0xffffffffu, hv, 5);
    U.n_src = __shfl_sync(0xffffffffu, hv, 6); U.prog = __shfl_sync(0xffffffffu, hv, 7); U.variant = __shfl_sync(0xffffffffu, hv, 8);
    U.n_rt = __shfl_sync(0xffffffffu, hv, 9);
#pragma unroll
    for (int r = 0; r < 4; ++r) U.rt[r] = __shfl_sync(0xffffffffu, hv, 10 + r);
    const int jb = blockIdx.y, comp = blockIdx.z;
    const int W = a.job[jb].outer * a.inner;
    const int64_t s_from = (int64_t)W * a.kf, s_to = (int64_t)W * a.kt;
    const double * __restrict__ src = a.job[jb].src + (int64_t)comp * a.n_elem * s_from;
    double * __restrict__ dst = a.job[jb].dst + (int64_t)comp * a.n_elem * s_to;
    const long long * __restrict__ dmap = a.job[jb].dst_map;
    const double * __restrict__ accf = a.job[jb].acc_from ? a.job[jb].acc_from + (int64_t)comp * a.n_elem * s_to : nullptr;
    const double coef = a.job[jb].coef;
    const bool accumulate = a.job[jb].accumulate != 0;
    const bool vec = (a.vec_ok >> jb) & 1;
    switch (U.variant)
    {
        case 0: dir_run<1, 8, 1>(a, U, src, dst, dmap, accf, s_from, s_to, coef, accumulate, vec, lane); break;
        case 1: dir_run<2, 4, 1>(a, U, src, dst, dmap, accf, s_from, s_to, coef, accumulate, vec, lane); break;
        case 2: dir_run<4, 2, 1>(a, U, src, dst, dmap, accf, s_from, s_to, coef, accumulate, vec, lane); break;
        default: dir_run<1, 1, 4>(a, U, src, dst, dmap, accf, s_from, s_to, coef, accumulate, vec, lane); break;
    }
}

cudaError_t launch_sweep_dir(const DirArgs & a, cudaStream_t st)
{
    static const bool pdl = !(std::getenv("AMDG_TC_PDL") && std::atoi(std::getenv("AMDG_TC_PDL")) == 0);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)((a.n_unit + DIR_THREADS / 32 - 1) / (DIR_THREADS / 32)), (unsigned)a.n_job, (unsigned)a.n_comp);
    cfg.blockDim = dim3(DIR_THREADS, 1, 1);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, sweep_dir_kernel, a);
}

}  // namespace amdg
