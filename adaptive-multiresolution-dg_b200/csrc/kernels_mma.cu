// K1 sweep kernel, whole-fibre tensor-core form (kernel variant 4).  Split from kernels.cu so that the translation units build in parallel.
#include <algorithm>
#include <cstdlib>
#include "kernels.cuh"
#include "cp_async.cuh"
#include "../../include/amdg.h"

namespace amdg {

// -------------------------------------------------------------------------------------------------------------
// K1, tensor-core form.  All fibres of one shape share one block-sparse matrix (mma_items.hpp); a CTA stages the
// source coefficients of a few fibres of a shape (a rectangle of columns) in shared memory with 16-byte async
// copies and then every warp walks row tiles of the shape's tile program: per tile entry one A fragment (operator
// values, L1/L2) feeds up to eight FP64 m8n8k4 MMAs, one per 8-column tile, whose B fragments come from shared
// memory.  ~3 instructions per 256 FMAs instead of ~12 in the list kernels, no per-row list walking, and long rows
// (coarse targets of long fibres) are just longer MMA chains.
// -------------------------------------------------------------------------------------------------------------
#ifndef AMDG_MMA_THREADS
#define AMDG_MMA_THREADS 256
#endif
static const int MMA_THREADS = AMDG_MMA_THREADS;
static const int MMA_SMEM_DOUBLES = 9 * 1024;          // 72 KiB per CTA: three CTAs per SM

int mma_smem_capacity_doubles() { return MMA_SMEM_DOUBLES; }

__device__ __forceinline__ void dmma8x8x4(double (&c)[2], double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

// A fragments x B fragments of the entries [p0, p1) of one row tile.  xb already holds the lane's column / source-index
// offsets; the eight 8-column tiles of a 64-column group are at compile-time offsets j*8*SC from it.
template <int KF, int NKP, int NT, int SC>
__device__ __forceinline__ void mma_rows(double (&acc)[8][2], const double * A, const int * s_ent, int p0, int p1,
                                         const double * xb, int rowsize, int sk, int kl, int lane, bool a_global)
{
    auto row_of = [&](int es) -> const double *
    {
        if (NKP == 1) return xb + es * rowsize;                         // the source-index offset kl*sk is folded into xb
        const int f = es / NKP;
        return xb + f * rowsize + min((es - f * NKP) * 4 + kl, KF - 1) * sk;
    };
    if (!a_global)
    {
        // operator values staged in shared memory
#pragma unroll 2
        for (int p = p0; p < p1; ++p)
        {
            const double av = A[(int64_t)p * 32 + lane];
            const double * xr = row_of(s_ent[p]);
#pragma unroll
            for (int j = 0; j < NT; ++j) dmma8x8x4(acc[j], av, xr[j * 8 * SC]);
        }
        return;
    }
    // operator values streamed from L2: PF fragments in flight per lane (rolling), bodies unconditional
    constexpr int PF = 8;
    double av[PF];
    const double * __restrict__ Ap = A + (int64_t)p0 * 32 + lane;
    const int n = p1 - p0;
#pragma unroll
    for (int v = 0; v < PF; ++v) av[v] = __ldg(Ap + (int64_t)min(v, n - 1) * 32);
    int p = 0;
    for (; p + PF <= n; p += PF)
    {
#pragma unroll
        for (int v = 0; v < PF; ++v)
        {
            const double cur = av[v];
            av[v] = __ldg(Ap + (int64_t)min(p + PF + v, n - 1) * 32);
            const double * xr = row_of(s_ent[p0 + p + v]);
#pragma unroll
            for (int j = 0; j < NT; ++j) dmma8x8x4(acc[j], cur, xr[j * 8 * SC]);
        }
    }
#pragma unroll
    for (int v = 0; v < PF; ++v)
    {
        if (p + v < n)
        {
            const double * xr = row_of(s_ent[p0 + p + v]);
#pragma unroll
            for (int j = 0; j < NT; ++j) dmma8x8x4(acc[j], av[v], xr[j * 8 * SC]);
        }
    }
}

// Shared-memory layout of a staged source row (one element of a fibre, the item's column rectangle):
//   INNER1 == false: X[k][col], col = o_local*ni + i_local compact, pitch pk = 4 (mod 8) doubles between source indices k
//                    -> conflict-free B fragments (lane = (col%8)*4 + k), tile j of a 64-column group at +8j;
//   INNER1 == true (sweep along the last dimension, inner == 1): X[col][k] -- the element's own memory order, so a row is
//                    one contiguous copy and the B fragment of lane (col, k) sits at col*KF + k.
template <int KF, int KT, bool INNER1>
__global__ void __launch_bounds__(MMA_THREADS, 768 / MMA_THREADS) sweep_mma_kernel(const MmaArgs a)
{
    extern __shared__ __align__(16) double Xs[];
    constexpr int KTP = KT <= 1 ? 1 : (KT <= 2 ? 2 : (KT <= 4 ? 4 : 8));
    constexpr int TG = 8 / KTP;
    constexpr int NKP = (KF + 3) / 4;
    constexpr int SC = INNER1 ? KF : 1;
    constexpr int NW = MMA_THREADS / 32;
#define MMA_STAMP(i) do { if (a.dbg && threadIdx.x == 0) a.dbg[(int64_t)blockIdx.x * 8 + (i)] = clock64(); } while (0)
    MMA_STAMP(0);
    asm volatile("griddepcontrol.launch_dependents;");          // programmatic dependent launch, as in sweep_tc_kernel
    const MmaItem it = a.items[blockIdx.x];
    if (a.dbg && threadIdx.x == 0)
    {
        unsigned long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        a.dbg[(int64_t)blockIdx.x * 8 + 5] = (long long)gt;
        a.dbg[(int64_t)blockIdx.x * 8 + 6] = it.m; a.dbg[(int64_t)blockIdx.x * 8 + 7] = it.nfib * 1000000 + it.n_ent * 100 + it.no * it.ni;
    }
    MMA_STAMP(1);
    const int jb = blockIdx.y / a.n_comp, comp = blockIdx.y % a.n_comp;
    const SweepJob J = a.job[jb];
    const int inner = INNER1 ? 1 : a.inner;
    const int W = J.outer * inner;
    const int64_t s_from = (int64_t)W * KF, s_to = (int64_t)W * KT;
    const double * __restrict__ src = J.src + (int64_t)comp * a.n_elem * s_from;
    double * __restrict__ dst = J.dst + (int64_t)comp * a.n_elem * s_to;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int m = it.m;
    const int ncols = it.no * it.ni;
    const int pk = it.pk;
    const int sk = INNER1 ? 1 : pk;
    const int rowsize = INNER1 ? ncols * KF : KF * pk;
    const int nrow = it.nfib * m;
    // shared memory: X[nrow][rowsize] | staged operator values (small pieces only) | piece ints: rt_ptr[n_rt+1] rt_id[n_rt] ent_src[n_ent] | elem[nrow]
    double * s_A = Xs + (((int64_t)nrow * rowsize + 1) & ~(int64_t)1);
    int * s_prog = reinterpret_cast<int *>(s_A + (it.stage_a ? it.n_ent * 32 : 0));
    const int n_prog_ints = 2 * it.n_rt + 1 + it.n_ent;
    int * s_elem = s_prog + n_prog_ints;
    const double * __restrict__ Ag = a.a_tab[it.prog];
    const double * A = it.stage_a ? s_A : Ag;

    // ---- stage (everything in flight at once): the program, the element rows, the operator values, the source rows
    {
        const int * __restrict__ ep = a.elem_pool + it.elem_ofs;
        // a row is nrun runs of runlen contiguous doubles: run r = (o_local, k) at global offset r*inner, shared offset k*pk + o_local*ni
        const int nrun = INNER1 ? 1 : it.no * KF;
        const int runlen = INNER1 ? ncols * KF : it.ni;
        const int64_t col_base = INNER1 ? (int64_t)it.o0 * KF : (int64_t)it.o0 * KF * inner + it.i0;
        const bool vec = ((runlen & 1) == 0) && ((col_base & 1) == 0) && (INNER1 || (inner & 1) == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0) && ((s_from & 1) == 0);
        const int cpr = vec ? (runlen >> 1) : runlen;                 // copies per run
        const int per_row = nrun * cpr;
        const unsigned cpr_magic = cpr <= 1 ? 0u : 0xffffffffu / (unsigned)cpr + 1u;
        for (int c = tid; c < n_prog_ints; c += MMA_THREADS) cp_async4(s_prog + c, a.prog_pool + it.prog_ofs + c);
        for (int c = tid; c < nrow; c += MMA_THREADS) cp_async4(s_elem + c, ep + c);
        if (it.stage_a) for (int c = tid; c < it.n_ent * 16; c += MMA_THREADS) cp_async16(s_A + 2 * c, Ag + 2 * c);
        auto copy_offsets = [&](int c, int & so, int & dof)
        {
            const int r = INNER1 ? 0 : (cpr <= 1 ? c : (int)__umulhi((unsigned)c, cpr_magic));
            const int w = (c - r * cpr) * (vec ? 2 : 1);
            const int o_l = r / KF, k = r - o_l * KF;
            so = r * inner + w;
            dof = INNER1 ? w : k * pk + o_l * it.ni + w;
        };
        // from here on the coefficient arrays of earlier kernels are read: wait for them (everything above touches work lists only)
        asm volatile("griddepcontrol.wait;" ::: "memory");
        // Rows of this warp: row = warp + NW*r.  Lane r fetches the element row of row r (one load for 32 rows, handed out
        // by shuffles).  The copy pattern inside a row is the same for every row: each lane computes its copy offsets once.
        if (per_row >= 32)
        {
            constexpr int MAXC = 8;
            int so[MAXC], dof[MAXC];
            const int ncp = (per_row + 31) >> 5;
#pragma unroll
            for (int i = 0; i < MAXC; ++i)
            {
                const int c = lane + 32 * i;
                so[i] = -1; dof[i] = 0;
                if (i < ncp && c < per_row) copy_offsets(c, so[i], dof[i]);
            }
            for (int rbase = 0; warp + NW * rbase < nrow; rbase += 32)
            {
                const int myrow = warp + NW * (rbase + lane);
                const int e_lane = myrow < nrow ? __ldg(ep + myrow) : 0;
                const int nr = min(32, (nrow - warp - NW * rbase + NW - 1) / NW);
                for (int r = 0; r < nr; ++r)
                {
                    const int e = __shfl_sync(0xffffffffu, e_lane, r);
                    const double * __restrict__ g = src + (int64_t)e * s_from + col_base;
                    double * xr = Xs + (int64_t)(warp + NW * (rbase + r)) * rowsize;
#pragma unroll
                    for (int i = 0; i < MAXC; ++i)
                        if (so[i] >= 0) { if (vec) cp_async16(xr + dof[i], g + so[i]); else cp_async8(xr + dof[i], g + so[i]); }
                    for (int c = lane + 32 * MAXC; c < per_row; c += 32)          // rows longer than 32*MAXC copies (rare)
                    {
                        int s2, d2; copy_offsets(c, s2, d2);
                        if (vec) cp_async16(xr + d2, g + s2); else cp_async8(xr + d2, g + s2);
                    }
                }
            }
        }
        else
        {
            // short rows: several rows per warp pass; lane = (row in pass, copy)
            const int rpp = 32 / per_row;
            const int sub = lane / per_row, c = lane - sub * per_row;
            const bool lane_on = sub < rpp;
            int so, dof; copy_offsets(c, so, dof);
            for (int rbase = 0; warp + NW * rbase < nrow; rbase += 32)
            {
                const int myrow = warp + NW * (rbase + lane);
                const int e_lane = myrow < nrow ? __ldg(ep + myrow) : 0;
                const int nr = min(32, (nrow - warp - NW * rbase + NW - 1) / NW);
                for (int r0 = 0; r0 < nr; r0 += rpp)
                {
                    const int r = r0 + sub;
                    const bool on = lane_on && r < nr;
                    const int e = __shfl_sync(0xffffffffu, e_lane, on ? r : 0);
                    if (!on) continue;
                    const double * __restrict__ g = src + (int64_t)e * s_from + col_base;
                    double * xr = Xs + (int64_t)(warp + NW * (rbase + r)) * rowsize;
                    if (vec) cp_async16(xr + dof, g + so); else cp_async8(xr + dof, g + so);
                }
            }
        }
        cp_async_commit();
    }
    MMA_STAMP(2);

    // ---- fragment coordinates: B fragment (source index lane%4, column lane/4), C fragment (row lane/4, columns (lane%4)*2, +1)
    const int kl = lane & 3, cb = lane >> 2, cc2 = (lane & 3) * 2;
    const int rr = lane >> 2;                                      // C fragment row -> (target g, output q)
    const int cg_ = rr / KTP, cq = rr - cg_ * KTP;
    const int * s_rt_ptr = s_prog, * s_rt_order = s_prog + it.n_rt + 1, * s_ent = s_prog + 2 * it.n_rt + 1;
    // pairs of output columns go out as one 16-byte store when they are adjacent and aligned in the destination block
    const bool vecst = !INNER1 && ((it.ni & 1) == 0) && ((it.i0 & 1) == 0) && ((inner & 1) == 0) && ((s_to & 1) == 0) && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0);
    const int q_off = INNER1 ? cq : cq * inner;
    for (int cg0 = 0; cg0 < ncols; cg0 += 64)
    {
        // destination offsets of the lane's first column of every 8-column tile; the second column is +1, +KT (INNER1) or wraps
        int soff[8]; unsigned vmask = 0, wrap = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j)
        {
            const int c2 = cg0 + j * 8 + cc2;
            const int c3 = c2 < ncols ? c2 : 0;
            if (INNER1) soff[j] = (it.o0 + c3) * KT;
            else
            {
                const int o2 = it.ni == 1 ? c3 : (int)__umulhi((unsigned)c3, it.ni_magic), i2 = c3 - o2 * it.ni;
                soff[j] = (it.o0 + o2) * KT * inner + it.i0 + i2;
                if (i2 + 1 >= it.ni) wrap |= 1u << j;
            }
            if (c2 < ncols) vmask |= 1u << (2 * j);
            if (c2 + 1 < ncols) vmask |= 2u << (2 * j);
        }
        const int step_wrap = INNER1 ? KT : KT * inner - it.ni + 1;
        const int ntile = min(8, (ncols - cg0 + 7) >> 3);
        if (cg0 == 0) { cp_async_wait_all(); __syncthreads(); MMA_STAMP(3); }
        const int col_b = min(cg0 + cb, ncols - 1);                // B column of tile 0 (tiles beyond the rectangle are never stored)
        const double * xcol = Xs + col_b * SC + (NKP == 1 ? min(kl, KF - 1) * sk : 0);

        const int n_units = it.n_rt * it.nfib;
        int ri = warp / it.nfib, b = warp - ri * it.nfib;
        for (int u = warp; u < n_units; u += NW)
        {
            const int rt = s_rt_order[ri];                               // row tile id (targets rt*TG ..)
            const int p0 = s_rt_ptr[ri], p1 = s_rt_ptr[ri + 1];           // entries are stored in position order
            const double * xb = xcol + (int64_t)b * m * rowsize;
            double acc[8][2];
#pragma unroll
            for (int j = 0; j < 8; ++j) { acc[j][0] = 0.0; acc[j][1] = 0.0; }
            // NT = number of 8-column tiles rounded up to 1/2/4/8: the MMAs of a loop body are unconditional (tiles
            // beyond the rectangle read memory of the next rows / the staged tables and are never stored)
            if (p1 > p0)
            {
                const bool ag = !it.stage_a;
                if (ntile > 4) mma_rows<KF, NKP, 8, SC>(acc, A, s_ent, p0, p1, xb, rowsize, sk, kl, lane, ag);
                else if (ntile > 2) mma_rows<KF, NKP, 4, SC>(acc, A, s_ent, p0, p1, xb, rowsize, sk, kl, lane, ag);
                else if (ntile > 1) mma_rows<KF, NKP, 2, SC>(acc, A, s_ent, p0, p1, xb, rowsize, sk, kl, lane, ag);
                else mma_rows<KF, NKP, 1, SC>(acc, A, s_ent, p0, p1, xb, rowsize, sk, kl, lane, ag);
            }
            // epilogue: C fragment row rr = (target cg_, output cq), columns cc2, cc2+1 of every tile
            const int e_loc = rt * TG + cg_;
            if (e_loc < m && cq < KT)
            {
                const int e = s_elem[b * m + e_loc];
                double * y = dst + (int64_t)e * s_to + q_off;
                if (vecst)
                {
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                    {
                        if (j >= ntile || !((vmask >> (2 * j)) & 1u)) continue;
                        double2 * yp = reinterpret_cast<double2 *>(y + soff[j]);
                        double2 v = make_double2(J.coef * acc[j][0], J.coef * acc[j][1]);
                        if (J.accumulate) { const double2 o = *yp; v.x += o.x; v.y += o.y; }
                        *yp = v;
                    }
                }
                else
                {
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                    {
                        if (j >= ntile) continue;
#pragma unroll
                        for (int h = 0; h < 2; ++h)
                        {
                            if (!((vmask >> (2 * j + h)) & 1u)) continue;
                            double * yp = y + soff[j] + (h == 0 ? 0 : (INNER1 || ((wrap >> j) & 1u) ? step_wrap : 1));
                            double v = J.coef * acc[j][h];
                            if (J.accumulate) v += *yp;
                            *yp = v;
                        }
                    }
                }
            }
            b += NW; while (b >= it.nfib) { b -= it.nfib; ++ri; }
        }
    }
    MMA_STAMP(4);
}

template <int KF, int KT, bool INNER1>
static cudaError_t launch_mma_t(const MmaArgs & a, int smem_doubles, cudaStream_t st)
{
    static PerDeviceOnce configured;
    if (!configured.done())
    {
        cudaError_t e = cudaFuncSetAttribute(sweep_mma_kernel<KF, KT, INNER1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(MMA_SMEM_DOUBLES * sizeof(double)));
        if (e != cudaSuccess) return e;
        configured.mark();
    }
    // programmatic stream serialization: the head of this grid overlaps the tail of the previous kernel (the kernel waits itself
    // before it touches coefficient data); matters most for the small sweeps this form serves in auto mode
    static const bool pdl = !(std::getenv("AMDG_TC_PDL") && std::atoi(std::getenv("AMDG_TC_PDL")) == 0);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)a.n_item, (unsigned)(a.n_job * a.n_comp), 1);
    cfg.blockDim = dim3(MMA_THREADS, 1, 1);
    cfg.dynamicSmemBytes = (size_t)std::min(smem_doubles, MMA_SMEM_DOUBLES) * sizeof(double);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, sweep_mma_kernel<KF, KT, INNER1>, a);
}

#define AMDG_DISPATCH_KT_M(KF_)                                                                   \
    switch (kt) {                                                                                 \
        case 1: return a.inner == 1 ? launch_mma_t<KF_, 1, true>(a, smem_doubles, st) : launch_mma_t<KF_, 1, false>(a, smem_doubles, st); \
        case 2: return a.inner == 1 ? launch_mma_t<KF_, 2, true>(a, smem_doubles, st) : launch_mma_t<KF_, 2, false>(a, smem_doubles, st); \
        case 3: return a.inner == 1 ? launch_mma_t<KF_, 3, true>(a, smem_doubles, st) : launch_mma_t<KF_, 3, false>(a, smem_doubles, st); \
        case 4: return a.inner == 1 ? launch_mma_t<KF_, 4, true>(a, smem_doubles, st) : launch_mma_t<KF_, 4, false>(a, smem_doubles, st); \
        case 5: return a.inner == 1 ? launch_mma_t<KF_, 5, true>(a, smem_doubles, st) : launch_mma_t<KF_, 5, false>(a, smem_doubles, st); \
        case 6: return a.inner == 1 ? launch_mma_t<KF_, 6, true>(a, smem_doubles, st) : launch_mma_t<KF_, 6, false>(a, smem_doubles, st); \
        default: return cudaErrorInvalidValue; }

cudaError_t launch_sweep_mma(const MmaArgs & a, int kf, int kt, int smem_doubles, cudaStream_t st)
{
    switch (kf)
    {
        case 1: AMDG_DISPATCH_KT_M(1) case 2: AMDG_DISPATCH_KT_M(2) case 3: AMDG_DISPATCH_KT_M(3)
        case 4: AMDG_DISPATCH_KT_M(4) case 5: AMDG_DISPATCH_KT_M(5) case 6: AMDG_DISPATCH_KT_M(6)
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace amdg
