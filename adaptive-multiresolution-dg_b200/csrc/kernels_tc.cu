// K1, lean tensor-core form (default): the same work items, tile programs and operator fragments as sweep_mma_kernel
// (kernels.cu, mma_items.hpp), executed by small CTAs with a short prologue.
//
// Why a second form: on the named grids most fibres are short (cfg2: 58 % of the elements sit on fibres of 1-4
// elements), so an item is a handful of element blocks and a warp owns one or two row tiles of it.  ncu of
// sweep_mma_kernel (round 1) showed ~1 360 warp instructions per warp of which ~100 are MMA work: the kernel is its own
// prologue, at 24 resident warps per SM.  Here:
//   * 128 threads per CTA, 80 registers (ncu: profiles/r02_roofline_kernels_ncu.md), <= 36 KiB of shared memory -> 6 CTAs
//     (24 warps) per SM (registers and shared memory both cap there), so that the load phase of some CTAs overlaps the
//     compute/store phase of others (one-shot CTAs, no pipeline state);
//   * units are (row tile, fibre, 32-column group): 4 accumulator tiles instead of 8, one code path for the MMA
//     loop (operator fragments from shared memory or straight from L1/L2), unit decode by two multiplications;
//   * copy offsets of a row are computed once per lane (4 slots) and the rest of a long row incrementally.
// Numerics are identical to sweep_mma_kernel (same fragment order, same summation order per output).
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

#include "kernels.cuh"

namespace amdg {

static const int TC_THREADS = 128;
#ifndef AMDG_TC_MIN_CTAS
#define AMDG_TC_MIN_CTAS 6
#endif
static const int TC_MIN_CTAS = AMDG_TC_MIN_CTAS;
static const int TC_SMEM_DOUBLES = 4608;            // upper bound (36 KiB, six CTAs per SM); typical lists need <= 27 KiB: eight CTAs per SM

int tc_smem_capacity_doubles() { return TC_SMEM_DOUBLES; }

__device__ __forceinline__ void tc_cp16(void * smem, const void * gmem)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void tc_cp8(void * smem, const void * gmem)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void tc_cp4(void * smem, const void * gmem)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
// bulk (TMA) copy of one contiguous run into shared memory, completion counted in bytes on an mbarrier
__device__ __forceinline__ void tc_bulk(void * smem, const void * gmem, unsigned bytes, unsigned bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void tc_dmma(double (&c)[2], double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

// entries [p0, p1) of one row tile: one A fragment per entry feeds `nt` MMAs (8-column tiles at +8*SC doubles).
// Operator fragments that are not staged come from L2 (hundreds of cycles): they are fetched one group of four entries
// ahead of the MMAs that use them.
template <int KF, int NKP, int SC, bool A_SHARED>
__device__ __forceinline__ void tc_rows(double (&acc)[4][2], const double * A, const int * s_ent, int p0, int p1,
                                        const double * xb, int rowsize, int sk, int kl, int lane, int nt)
{
    auto body = [&](int p, double av)
    {
        const int es = s_ent[p];
        const double * xr;
        if (NKP == 1) xr = xb + es * rowsize;                            // the source-index offset is folded into xb
        else { const int f = es / NKP; xr = xb + f * rowsize + min((es - f * NKP) * 4 + kl, KF - 1) * sk; }
        if (nt == 4)
        {
#pragma unroll
            for (int j = 0; j < 4; ++j) tc_dmma(acc[j], av, xr[j * 8 * SC]);
        }
        else
        {
#pragma unroll
            for (int j = 0; j < 3; ++j) if (j < nt) tc_dmma(acc[j], av, xr[j * 8 * SC]);
        }
    };
    if (A_SHARED)
    {
#pragma unroll 2
        for (int p = p0; p < p1; ++p) body(p, A[p * 32 + lane]);
        return;
    }
    const double * __restrict__ Ap = A + (int64_t)p0 * 32 + lane;
    const int n = p1 - p0;
    double cur[4], nxt[4];
#pragma unroll
    for (int v = 0; v < 4; ++v) cur[v] = __ldg(Ap + min(v, n - 1) * 32);
    for (int p = 0; p < n; p += 4)
    {
#pragma unroll
        for (int v = 0; v < 4; ++v) nxt[v] = __ldg(Ap + min(p + 4 + v, n - 1) * 32);
#pragma unroll
        for (int v = 0; v < 4; ++v) if (p + v < n) body(p0 + p + v, cur[v]);
#pragma unroll
        for (int v = 0; v < 4; ++v) cur[v] = nxt[v];
    }
}

// one 8-column tile only (narrow rectangles: the long fibres): the entries go round-robin into the four accumulators,
// so that four MMAs are in flight instead of one dependent chain; the caller adds them up in a fixed order
template <int KF, int NKP, bool A_SHARED>
__device__ __forceinline__ void tc_rows_narrow(double (&acc)[4][2], const double * A, const int * s_ent, int p0, int p1,
                                               const double * xb, int rowsize, int sk, int kl, int lane)
{
    auto bfrag = [&](int p) -> double
    {
        const int es = s_ent[p];
        if (NKP == 1) return xb[es * rowsize];
        const int f = es / NKP; return xb[f * rowsize + min((es - f * NKP) * 4 + kl, KF - 1) * sk];
    };
    if (A_SHARED)
    {
        int p = p0;
        for (; p + 4 <= p1; p += 4)
        {
#pragma unroll
            for (int j = 0; j < 4; ++j) tc_dmma(acc[j], A[(p + j) * 32 + lane], bfrag(p + j));
        }
        for (int j = 0; p < p1; ++p, ++j)
        {
            const double av = A[p * 32 + lane], bv = bfrag(p);
            if (j == 0) tc_dmma(acc[0], av, bv); else if (j == 1) tc_dmma(acc[1], av, bv); else tc_dmma(acc[2], av, bv);
        }
        return;
    }
    // operator fragments from L2: eight entries in flight (two groups of four)
    const double * __restrict__ Ap = A + (int64_t)p0 * 32 + lane;
    const int n = p1 - p0;
    double cur[4], nxt[4];
#pragma unroll
    for (int v = 0; v < 4; ++v) cur[v] = __ldg(Ap + min(v, n - 1) * 32);
    for (int p = 0; p < n; p += 4)
    {
#pragma unroll
        for (int v = 0; v < 4; ++v) nxt[v] = __ldg(Ap + min(p + 4 + v, n - 1) * 32);
#pragma unroll
        for (int v = 0; v < 4; ++v) if (p + v < n) tc_dmma(acc[v], cur[v], bfrag(p0 + p + v));
#pragma unroll
        for (int v = 0; v < 4; ++v) cur[v] = nxt[v];
    }
}

// Shared-memory layout of a staged source row as in sweep_mma_kernel:
//   INNER1 == false: X[k][col], col = o_local*ni + i_local, pitch pk = 4 (mod 8) doubles between source indices k;
//   INNER1 == true : X[col][k], the element's own memory order (one contiguous copy per row).
// MODE 0: general; 1: sweep along the last dimension (inner == 1); 2: as 0, and the host has checked that every job overwrites its
// destination and that all stores are aligned 16-byte pairs (the common case) -- the general store path is compiled out
template <int KF, int KT, int MODE>
__global__ void __launch_bounds__(TC_THREADS, TC_MIN_CTAS) sweep_tc_kernel(const MmaArgs a)
{
    constexpr bool INNER1 = MODE == 1;
    constexpr bool FAST = MODE == 2;
    extern __shared__ __align__(16) double Xs[];
    constexpr int KTP = KT <= 1 ? 1 : (KT <= 2 ? 2 : (KT <= 4 ? 4 : 8));
    constexpr int TG = 8 / KTP;
    constexpr int NKP = (KF + 3) / 4;
    constexpr int SC = INNER1 ? KF : 1;
    constexpr int NW = TC_THREADS / 32;
#define TC_STAMP(i) do { if (a.dbg && threadIdx.x == 0) a.dbg[(int64_t)blockIdx.x * 8 + (i)] = clock64(); } while (0)
    TC_STAMP(0);
    // programmatic dependent launch: the next kernel of the stream may start its CTAs as soon as every CTA of this grid has got
    // here (they then run their own prologue and park at griddepcontrol.wait below until this grid has completed)
    asm volatile("griddepcontrol.launch_dependents;");
    const MmaItem it = a.items[blockIdx.x];
    if (a.dbg && threadIdx.x == 0)
    {
        unsigned long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        a.dbg[(int64_t)blockIdx.x * 8 + 5] = (long long)gt;
        a.dbg[(int64_t)blockIdx.x * 8 + 6] = it.m; a.dbg[(int64_t)blockIdx.x * 8 + 7] = it.nfib * 1000000 + it.n_ent * 100 + it.no * it.ni;
    }
    TC_STAMP(1);
    const int jb = blockIdx.y, comp = blockIdx.z;
    const SweepJob J = a.job[jb];
    const int inner = INNER1 ? 1 : a.inner;
    const int W = J.outer * inner;
    const int64_t s_from = (int64_t)W * KF, s_to = (int64_t)W * KT;
    const double * __restrict__ src = J.src + (int64_t)comp * a.n_elem * s_from;
    double * __restrict__ dst = J.dst + (int64_t)comp * a.n_elem * s_to;
    // optional destination map (element row -> offset of its block relative to dst, e.g. in peer memory) and accumulate-from array (kernels.cuh)
    const long long * __restrict__ dmap = J.dst_map;
    const double * __restrict__ accf = J.acc_from ? J.acc_from + (int64_t)comp * a.n_elem * s_to : nullptr;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int m = it.m;
    const int ncols = it.no * it.ni;
    const int pk = it.pk;
    const int sk = INNER1 ? 1 : pk;
    const int rowsize = INNER1 ? ncols * KF : KF * pk;
    const int nrow = it.nfib * it.nsrc;                            // staged rows; the element list below covers whole fibres (targets)
    const int nelem = it.nfib * m;
    double * s_A = Xs + ((nrow * rowsize + 1) & ~1);
    int * s_prog = reinterpret_cast<int *>(s_A + (it.stage_a ? it.n_ent * 32 : 0));
    const int n_prog_ints = 2 * it.n_rt + 1 + it.n_ent;
    int * s_elem = s_prog + n_prog_ints;
    const double * __restrict__ Ag = a.a_tab[it.prog];

    // ---- stage: the program, the element rows, (small) operator fragments and the source rows, all in flight at once
    __shared__ __align__(8) unsigned long long s_bar;
    bool use_bulk = false;
    {
        const int * __restrict__ ep = a.elem_pool + it.src_ofs;        // element rows of the staged rows
        // a row is nrun runs of runlen contiguous doubles: run r = (o_local, k) at global offset r*inner, shared offset k*pk + o_local*ni
        const int nrun = INNER1 ? 1 : it.no * KF;
        const int runlen = INNER1 ? ncols * KF : it.ni;
        const int64_t col_base = INNER1 ? (int64_t)it.o0 * KF : (int64_t)it.o0 * KF * inner + it.i0;
        const bool vec = ((runlen & 1) == 0) && ((col_base & 1) == 0) && (INNER1 || (inner & 1) == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0) && ((s_from & 1) == 0);
        // runs of >= 128 aligned bytes go as one bulk copy each (one instruction per run instead of one 16-byte copy per lane)
        use_bulk = vec && runlen >= 16;
        const int cpr = vec ? (runlen >> 1) : runlen;                 // copies per run
        const int per_row = nrun * cpr;
        const unsigned cpr_magic = (cpr <= 1 || use_bulk) ? 0u : 0xffffffffu / (unsigned)cpr + 1u;
        for (int c = tid; c < n_prog_ints; c += TC_THREADS) tc_cp4(s_prog + c, a.prog_pool + it.prog_ofs + c);
        for (int c = tid; c < nelem; c += TC_THREADS) tc_cp4(s_elem + c, a.elem_pool + it.elem_ofs + c);
        if (it.stage_a) for (int c = tid; c < it.n_ent * 16; c += TC_THREADS) tc_cp16(s_A + 2 * c, Ag + 2 * c);
        auto copy_offsets = [&](int c, int & so, int & dof)
        {
            const int r = INNER1 ? 0 : (cpr <= 1 ? c : (int)__umulhi((unsigned)c, cpr_magic));
            const int w = (c - r * cpr) * (vec ? 2 : 1);
            const int o_l = r / KF, k = r - o_l * KF;
            so = r * inner + w;
            dof = INNER1 ? w : k * pk + o_l * it.ni + w;
        };
        // everything above reads only the work lists (written once by the host); from here on the coefficient arrays of earlier
        // kernels are touched: wait for them (returns at once when the launch carries no programmatic dependency)
        asm volatile("griddepcontrol.wait;" ::: "memory");
        if (use_bulk)
        {
            const unsigned bar = (unsigned)__cvta_generic_to_shared(&s_bar);
            if (tid == 0)
            {
                asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar) : "memory");
                asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            }
            __syncthreads();
            const int n_copy = nrow * nrun;
            const unsigned run_bytes = (unsigned)runlen * 8u;
            if (tid == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"((unsigned)n_copy * run_bytes) : "memory");
            for (int c = tid; c < n_copy; c += TC_THREADS)
            {
                const int row = nrun == 1 ? c : (int)__umulhi((unsigned)c, it.nrun_magic), r = c - row * nrun;
                const int o_l = r / KF, k = r - o_l * KF;
                const int e = __ldg(ep + row);
                tc_bulk(Xs + row * rowsize + (INNER1 ? 0 : k * pk + o_l * it.ni), src + (int64_t)e * s_from + col_base + (int64_t)r * inner, run_bytes, bar);
            }
        }
        else if (!INNER1 && !vec && ncols >= 32)
        {
            // odd block edges (no 16-byte alignment): 8-byte copies, one column (o, i) per lane and pass with its KF source indices in
            // a row -- two multiplications per column instead of a run decode per copy
            for (int rbase = 0; warp + NW * rbase < nrow; rbase += 32)
            {
                const int myrow = warp + NW * (rbase + lane);
                const int e_lane = myrow < nrow ? __ldg(ep + myrow) : 0;
                const int nr = min(32, (nrow - warp - NW * rbase + NW - 1) / NW);
                for (int r = 0; r < nr; ++r)
                {
                    const int e = __shfl_sync(0xffffffffu, e_lane, r);
                    const double * __restrict__ g = src + (int64_t)e * s_from + col_base;
                    double * xr = Xs + (warp + NW * (rbase + r)) * rowsize;
                    for (int col = lane; col < ncols; col += 32)
                    {
                        const int o_l = it.ni == 1 ? col : (int)__umulhi((unsigned)col, it.ni_magic), i_l = col - o_l * it.ni;
                        const double * __restrict__ gs = g + o_l * KF * inner + i_l;
                        double * xd = xr + col;
#pragma unroll
                        for (int k = 0; k < KF; ++k) tc_cp8(xd + k * pk, gs + k * inner);
                    }
                }
            }
        }
        else if (per_row >= 32)
        {
            // rows of this warp: warp + NW*r; lane r fetches the element row of row r, handed out by shuffles; the copy pattern
            // inside a row is the same for every row, so each lane computes its first MAXC copy offsets once
            constexpr int MAXC = 4;
            int so[MAXC], dof[MAXC];
#pragma unroll
            for (int i = 0; i < MAXC; ++i)
            {
                const int c = lane + 32 * i;
                so[i] = -1; dof[i] = 0;
                if (c < per_row) copy_offsets(c, so[i], dof[i]);
            }
#define TC_ROW_LOOP(CP)                                                                                                   \
            for (int rbase = 0; warp + NW * rbase < nrow; rbase += 32)                                                    \
            {                                                                                                             \
                const int myrow = warp + NW * (rbase + lane);                                                             \
                const int e_lane = myrow < nrow ? __ldg(ep + myrow) : 0;                                                  \
                const int nr = min(32, (nrow - warp - NW * rbase + NW - 1) / NW);                                         \
                for (int r = 0; r < nr; ++r)                                                                              \
                {                                                                                                         \
                    const int e = __shfl_sync(0xffffffffu, e_lane, r);                                                    \
                    const double * __restrict__ g = src + (int64_t)e * s_from + col_base;                                 \
                    double * xr = Xs + (warp + NW * (rbase + r)) * rowsize;                                               \
                    _Pragma("unroll")                                                                                     \
                    for (int i = 0; i < MAXC; ++i) if (so[i] >= 0) CP(xr + dof[i], g + so[i]);                            \
                    for (int c = lane + 32 * MAXC; c < per_row; c += 32)                                                  \
                    {                                                                                                     \
                        int s2, d2; copy_offsets(c, s2, d2);                                                              \
                        CP(xr + d2, g + s2);                                                                              \
                    }                                                                                                     \
                }                                                                                                         \
            }
            if (vec) { TC_ROW_LOOP(tc_cp16) } else { TC_ROW_LOOP(tc_cp8) }
#undef TC_ROW_LOOP
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
                    double * xr = Xs + (warp + NW * (rbase + r)) * rowsize;
                    if (vec) tc_cp16(xr + dof, g + so); else tc_cp8(xr + dof, g + so);
                }
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    TC_STAMP(2);

    // ---- fragment coordinates: B fragment (source index lane%4, column lane/4), C fragment (row lane/4, columns (lane%4)*2, +1)
    const int kl = lane & 3, cb = lane >> 2, cc2 = (lane & 3) * 2;
    const int rr = lane >> 2;                                      // C fragment row -> (target g, output q)
    const int cg_ = rr / KTP, cq = rr - cg_ * KTP;
    const int * s_rt_ptr = s_prog, * s_rt_order = s_prog + it.n_rt + 1, * s_ent = s_prog + 2 * it.n_rt + 1;
    const bool vecst = !INNER1 && ((it.ni & 1) == 0) && ((it.i0 & 1) == 0) && ((inner & 1) == 0) && ((s_to & 1) == 0) && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) &&
                       (!accf || (reinterpret_cast<uintptr_t>(accf) & 15) == 0);
    const int q_off = (INNER1 ? it.o0 * KT + cq : it.o0 * KT * inner + cq * inner + it.i0);
    const int step_wrap = INNER1 ? KT : KT * inner - it.ni + 1;
    const int n_rf = it.n_rt * it.nfib;
    const int n_units = ((ncols + 31) >> 5) * n_rf;
    const bool row_on = cq < KT;
    const bool fast_store = FAST || (vecst && !J.accumulate);
    const double coef = J.coef;
    const double * A = it.stage_a ? s_A : Ag;
    const int fib_stride = it.nsrc * rowsize;
    const double * xlane = Xs + (NKP == 1 ? min(kl, KF - 1) * sk : 0);

    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    if (use_bulk)
    {
        const unsigned bar = (unsigned)__cvta_generic_to_shared(&s_bar);
        unsigned done = 0;
        while (!done)
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(bar) : "memory");
    }
    TC_STAMP(3);

    if (it.ksplit)
    {
        // coarse targets of a long fibre (one 8-column tile, nothing staged): the entry list of a row tile is cut into NW contiguous
        // parts, one per warp; operator and source fragments stream from L2, four entries ahead of the MMAs; partial sums go through
        // shared memory and warp 0 adds them in warp order and stores
        double * s_red = reinterpret_cast<double *>((reinterpret_cast<uintptr_t>(s_elem + nelem) + 7) & ~(uintptr_t)7);
        int off0, second0, boff;                                     // store offsets of the lane's C columns; source offset of its B column
        {
            const int c3 = cc2 < ncols ? cc2 : 0, cB = min(cb, ncols - 1);
            if (INNER1) { off0 = c3 * KT; second0 = KT; boff = (it.o0 + cB) * KF; }
            else
            {
                const int o2 = it.ni == 1 ? c3 : (int)__umulhi((unsigned)c3, it.ni_magic), i2 = c3 - o2 * it.ni;
                off0 = o2 * KT * inner + i2;
                second0 = i2 + 1 < it.ni ? 1 : step_wrap;
                const int oB = it.ni == 1 ? cB : (int)__umulhi((unsigned)cB, it.ni_magic), iB = cB - oB * it.ni;
                boff = (it.o0 + oB) * KF * inner + it.i0 + iB;
            }
        }
        const int kstride = INNER1 ? 1 : inner;
        for (int u = 0; u < n_rf; ++u)
        {
            const int ri = it.nfib == 1 ? u : (int)__umulhi((unsigned)u, it.nfib_magic), b = u - ri * it.nfib;
            const int rt = s_rt_order[ri];
            const int p0 = s_rt_ptr[ri], p1 = s_rt_ptr[ri + 1];
            const int chunk = (((p1 - p0 + NW - 1) / NW) + 3) & ~3;
            const int q0 = min(p1, p0 + warp * chunk), q1 = min(p1, q0 + chunk);
            const int * sel = s_elem + b * m;
            double acc[4][2];
#pragma unroll
            for (int j = 0; j < 4; ++j) { acc[j][0] = 0.0; acc[j][1] = 0.0; }
            const int n = q1 - q0;
            if (n > 0)
            {
                auto bload = [&](int p) -> double
                {
                    const int es = s_ent[p];
                    const int f = NKP == 1 ? es : es / NKP;
                    const int k = NKP == 1 ? min(kl, KF - 1) : min((es - f * NKP) * 4 + kl, KF - 1);
                    return __ldg(src + (int64_t)sel[f] * s_from + boff + k * kstride);
                };
                const double * __restrict__ Ap = Ag + (int64_t)q0 * 32 + lane;
                double ca[4], cbv[4], na[4], nb[4];
#pragma unroll
                for (int v = 0; v < 4; ++v) { const int p = min(v, n - 1); ca[v] = __ldg(Ap + p * 32); cbv[v] = bload(q0 + p); }
                for (int p = 0; p < n; p += 4)
                {
#pragma unroll
                    for (int v = 0; v < 4; ++v) { const int pp = min(p + 4 + v, n - 1); na[v] = __ldg(Ap + pp * 32); nb[v] = bload(q0 + pp); }
#pragma unroll
                    for (int v = 0; v < 4; ++v) if (p + v < n) tc_dmma(acc[v], ca[v], cbv[v]);
#pragma unroll
                    for (int v = 0; v < 4; ++v) { ca[v] = na[v]; cbv[v] = nb[v]; }
                }
            }
            const double r0 = (acc[0][0] + acc[1][0]) + (acc[2][0] + acc[3][0]);
            const double r1 = (acc[0][1] + acc[1][1]) + (acc[2][1] + acc[3][1]);
            if (warp > 0) { s_red[(warp * 32 + lane) * 2] = r0; s_red[(warp * 32 + lane) * 2 + 1] = r1; }
            __syncthreads();
            if (warp == 0)
            {
                double v0 = r0, v1 = r1;
#pragma unroll
                for (int w = 1; w < NW; ++w) { v0 += s_red[(w * 32 + lane) * 2]; v1 += s_red[(w * 32 + lane) * 2 + 1]; }
                const int e_loc = rt * TG + cg_;
                if (e_loc < m && row_on && cc2 < ncols)
                {
                    const int e_t = sel[e_loc];
                    double * y = dst + (dmap ? __ldg(dmap + e_t) : (long long)e_t * s_to) + q_off;
                    const double * yr = accf ? accf + (int64_t)e_t * s_to + q_off : y;
                    v0 *= coef; v1 *= coef;
                    if (J.accumulate) v0 += yr[off0];
                    y[off0] = v0;
                    if (cc2 + 1 < ncols)
                    {
                        if (J.accumulate) v1 += yr[off0 + second0];
                        y[off0 + second0] = v1;
                    }
                }
            }
            __syncthreads();
        }
        TC_STAMP(4);
        return;
    }

    // per 32-column group (changes rarely along a warp's units): tile count, the lane's B column and its store offsets
    int cg_cur = -1, nt = 0, off[4], second[4];
    unsigned vmask = 0;                                            // bit j: first column of tile j is stored; bit 4+j: the second one too
    const double * xcol = xlane;
    for (int u = warp; u < n_units; u += NW)
    {
        const int cgi = n_rf == 1 ? u : (int)__umulhi((unsigned)u, it.unit_magic);           // 32-column group
        const int rem = u - cgi * n_rf;
        const int ri = it.nfib == 1 ? rem : (int)__umulhi((unsigned)rem, it.nfib_magic), b = rem - ri * it.nfib;
        if (cgi != cg_cur)
        {
            cg_cur = cgi;
            const int cg0 = cgi << 5;
            nt = min(4, (ncols - cg0 + 7) >> 3);
            xcol = xlane + min(cg0 + cb, ncols - 1) * SC;          // B column of tile 0 (columns beyond the rectangle are never stored)
            vmask = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j)
            {
                const int c2 = cg0 + j * 8 + cc2;
                const int c3 = c2 < ncols ? c2 : 0;
                if (INNER1) { off[j] = c3 * KT; second[j] = KT; }
                else
                {
                    const int o2 = it.ni == 1 ? c3 : (int)__umulhi((unsigned)c3, it.ni_magic), i2 = c3 - o2 * it.ni;
                    off[j] = o2 * KT * inner + i2;
                    second[j] = i2 + 1 < it.ni ? 1 : step_wrap;
                }
                if (j < nt && c2 < ncols) vmask |= 1u << j;
                if (j < nt && c2 + 1 < ncols) vmask |= 16u << j;
            }
        }
        const int rt = s_rt_order[ri];                               // row tile id (targets rt*TG ..)
        const int p0 = s_rt_ptr[ri], p1 = s_rt_ptr[ri + 1];           // entries are stored in position order
        const double * xb = xcol + b * fib_stride;
        const int e_loc = rt * TG + cg_;
        const bool tgt_on = e_loc < m && row_on;
        const int e_t = tgt_on ? s_elem[b * m + e_loc] : 0;
        double * y = dst + ((!FAST && dmap) ? __ldg(dmap + e_t) : (long long)e_t * s_to) + q_off;
        const double * yr = (!FAST && accf) ? accf + (int64_t)e_t * s_to + q_off : y;
        if (!FAST && J.accumulate && tgt_on)
        {
            // an accumulating sweep reads its destination: start those lines towards L1 now, the MMA loop hides the trip
#pragma unroll
            for (int j = 0; j < 4; ++j)
            {
                if ((vmask >> j) & 1u) asm volatile("prefetch.global.L1 [%0];" :: "l"(yr + off[j]));
                if (!vecst && ((vmask >> j) & 16u)) asm volatile("prefetch.global.L1 [%0];" :: "l"(yr + off[j] + second[j]));
            }
        }
        double acc[4][2];
#pragma unroll
        for (int j = 0; j < 4; ++j) { acc[j][0] = 0.0; acc[j][1] = 0.0; }
        if (nt == 1 && p1 - p0 >= 8)
        {
            if (it.stage_a) tc_rows_narrow<KF, NKP, true>(acc, A, s_ent, p0, p1, xb, rowsize, sk, kl, lane);
            else tc_rows_narrow<KF, NKP, false>(acc, A, s_ent, p0, p1, xb, rowsize, sk, kl, lane);
            acc[0][0] = (acc[0][0] + acc[1][0]) + (acc[2][0] + acc[3][0]);
            acc[0][1] = (acc[0][1] + acc[1][1]) + (acc[2][1] + acc[3][1]);
        }
        else if (it.stage_a) tc_rows<KF, NKP, SC, true>(acc, A, s_ent, p0, p1, xb, rowsize, sk, kl, lane, nt);
        else tc_rows<KF, NKP, SC, false>(acc, A, s_ent, p0, p1, xb, rowsize, sk, kl, lane, nt);
        // epilogue: C fragment row rr = (target cg_, output cq), columns cc2, cc2+1 of every tile
        if (tgt_on)
        {
            if (FAST || fast_store)
            {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if ((vmask >> j) & 1u) *reinterpret_cast<double2 *>(y + off[j]) = make_double2(coef * acc[j][0], coef * acc[j][1]);
            }
            else if (!FAST)
            {
                // general path.  All old values of an accumulating sweep are fetched before the first store (the stores may alias the
                // loads for the compiler, which would serialise one L2 round trip per tile); aligned pairs go as 16-byte accesses
                const bool accu = J.accumulate != 0;
                if (vecst)
                {
                    double2 old[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                    {
                        old[j] = make_double2(0.0, 0.0);
                        if (accu && ((vmask >> j) & 1u)) old[j] = *reinterpret_cast<const double2 *>(yr + off[j]);
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if ((vmask >> j) & 1u) *reinterpret_cast<double2 *>(y + off[j]) = make_double2(coef * acc[j][0] + old[j].x, coef * acc[j][1] + old[j].y);
                }
                else
                {
                    double old[4][2];
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                    {
                        old[j][0] = 0.0; old[j][1] = 0.0;
                        if (accu && ((vmask >> j) & 1u)) old[j][0] = yr[off[j]];
                        if (accu && ((vmask >> j) & 16u)) old[j][1] = yr[off[j] + second[j]];
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                    {
                        if ((vmask >> j) & 1u) y[off[j]] = coef * acc[j][0] + old[j][0];
                        if ((vmask >> j) & 16u) y[off[j] + second[j]] = coef * acc[j][1] + old[j][1];
                    }
                }
            }
        }
    }
    TC_STAMP(4);
}

template <int KF, int KT, int MODE>
static cudaError_t launch_tc_t(const MmaArgs & a, int smem_doubles, cudaStream_t st)
{
    static PerDeviceOnce configured;
    if (!configured.done())
    {
        cudaError_t e = cudaFuncSetAttribute(sweep_tc_kernel<KF, KT, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(TC_SMEM_DOUBLES * sizeof(double)));
        if (e != cudaSuccess) return e;
        configured.mark();
    }
    if (smem_doubles > TC_SMEM_DOUBLES) return cudaErrorInvalidValue;
    // launched with programmatic stream serialization: the head of this grid (item headers, element rows, index arithmetic) overlaps
    // the tail of the previous kernel; the kernel orders its data accesses itself (griddepcontrol.wait)
    static const bool pdl = !(std::getenv("AMDG_TC_PDL") && std::atoi(std::getenv("AMDG_TC_PDL")) == 0);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)a.n_item, (unsigned)a.n_job, (unsigned)a.n_comp);
    cfg.blockDim = dim3(TC_THREADS, 1, 1);
    cfg.dynamicSmemBytes = (size_t)smem_doubles * sizeof(double);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, sweep_tc_kernel<KF, KT, MODE>, a);
}

template <int KF, int KT>
static cudaError_t launch_tc_m(const MmaArgs & a, int mode, int smem_doubles, cudaStream_t st)
{
    if (mode == 1) return launch_tc_t<KF, KT, 1>(a, smem_doubles, st);
    if (mode == 2) return launch_tc_t<KF, KT, 2>(a, smem_doubles, st);
    return launch_tc_t<KF, KT, 0>(a, smem_doubles, st);
}

#define AMDG_DISPATCH_KT_TC(KF_)                                                                  \
    switch (kt) {                                                                                 \
        case 1: return launch_tc_m<KF_, 1>(a, mode, smem_doubles, st);                            \
        case 2: return launch_tc_m<KF_, 2>(a, mode, smem_doubles, st);                            \
        case 3: return launch_tc_m<KF_, 3>(a, mode, smem_doubles, st);                            \
        case 4: return launch_tc_m<KF_, 4>(a, mode, smem_doubles, st);                            \
        case 5: return launch_tc_m<KF_, 5>(a, mode, smem_doubles, st);                            \
        case 6: return launch_tc_m<KF_, 6>(a, mode, smem_doubles, st);                            \
        default: return cudaErrorInvalidValue; }

cudaError_t launch_sweep_tc(const MmaArgs & a, int kf, int kt, int smem_doubles, cudaStream_t st)
{
    // mode 2 (plain aligned 16-byte stores only): every rectangle of the lean lists has even ni and i0 when inner is even
    int mode = a.inner == 1 ? 1 : 0;
    if (mode == 0 && (a.inner & 1) == 0)
    {
        bool fast = true;
        for (int i = 0; i < a.n_job && fast; ++i)
        {
            const int64_t s_to = (int64_t)a.job[i].outer * a.inner * kt;
            fast = !a.job[i].accumulate && !a.job[i].dst_map && (s_to & 1) == 0 && (reinterpret_cast<uintptr_t>(a.job[i].dst) & 15) == 0;
        }
        if (fast) mode = 2;
    }
    switch (kf)
    {
        case 1: AMDG_DISPATCH_KT_TC(1) case 2: AMDG_DISPATCH_KT_TC(2) case 3: AMDG_DISPATCH_KT_TC(3)
        case 4: AMDG_DISPATCH_KT_TC(4) case 5: AMDG_DISPATCH_KT_TC(5) case 6: AMDG_DISPATCH_KT_TC(6)
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace amdg
