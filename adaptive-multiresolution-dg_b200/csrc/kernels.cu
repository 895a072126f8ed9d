// FP64 CUDA kernels for sm_100a.  See kernels.cuh for the map to the reference functions.
#include <algorithm>
#include "kernels.cuh"
#include "cp_async.cuh"
#include "../../include/amdg.h"

namespace amdg {


// -------------------------------------------------------------------------------------------------------------
// K1, gather form.  One thread owns one (target slot, column) and the KT outputs of that column; it walks the
// target's neighbour list (U sources first, then L sources) and, per source, reads the KF source entries of
// the same column and the (KF x KT) operator block of the 1D pair.  Reads of a source block are served by
// L1/L2 (a block is re-read by every related target); the fibre-staged kernel below removes that re-read.
//   dst[e][o][q][i] = coef * sum_f sum_k src[f][o][k][i] * B[pair(f,e)][k][q]   (+ dst if accumulate)
// reference loop: source/FastMultiplyLU.cpp:476-506
// -------------------------------------------------------------------------------------------------------------
template <int KF, int KT>
__global__ void __launch_bounds__(128) sweep_gather_kernel(const SweepArgs a)
{
    const int jb = blockIdx.y / a.n_comp, comp = blockIdx.y % a.n_comp;
    const SweepJob J = a.job[jb];
    const int inner = a.inner;
    const int W = J.outer * inner;
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t slot = g / W;
    if (slot >= a.n_elem) return;
    const int c = (int)(g - slot * W);
    const int o = c / inner, i = c - o * inner;
    const int64_t s_from = (int64_t)W * KF, s_to = (int64_t)W * KT;
    const double * __restrict__ src = J.src + (int64_t)comp * a.n_elem * s_from;
    double * __restrict__ dst = J.dst + (int64_t)comp * a.n_elem * s_to;

    const int e = a.slot_elem[slot];
    const int fbase = a.slot_fbase[slot];
    int64_t n0 = a.nbr_ptr[slot], n1 = a.nbr_ptr[slot + 1];
    const int split = a.nbr_split[slot];
    if (a.lu == AMDG_LU_U) n1 = n0 + split;
    else if (a.lu == AMDG_LU_L) n0 = n0 + split;

    double acc[KT];
#pragma unroll
    for (int q = 0; q < KT; ++q) acc[q] = 0.0;

    const int64_t col_off = (int64_t)o * KF * inner + i;
    // neighbours in groups of four: the three dependent loads of an entry (neighbour record -> element row -> source values) are issued for the
    // whole group before the first product, so a target with n entries waits for ~3 n / 4 round trips instead of 3 n (the adaptive mode of the
    // context sweeps small short-lived grids with this kernel: a launch there is one such dependency chain).  Products stay in entry order.
    constexpr int G = 4;
    for (int64_t p = n0; p < n1; p += G)
    {
        NbrDev nb[G];
        int f[G];
        double xv[G][KF];
#pragma unroll
        for (int j = 0; j < G; ++j) nb[j] = a.nbr[min(p + j, n1 - 1)];
#pragma unroll
        for (int j = 0; j < G; ++j) f[j] = a.slot_elem[fbase + nb[j].local];
#pragma unroll
        for (int j = 0; j < G; ++j)
        {
            const double * __restrict__ x = src + (int64_t)f[j] * s_from + col_off;
#pragma unroll
            for (int k = 0; k < KF; ++k) xv[j][k] = __ldg(x + (int64_t)k * inner);
        }
#pragma unroll
        for (int j = 0; j < G; ++j)
        {
            if (p + j >= n1) break;
            const double * __restrict__ B = a.blocks + (int64_t)nb[j].pair * (KF * KT);
#pragma unroll
            for (int k = 0; k < KF; ++k)
            {
#pragma unroll
                for (int q = 0; q < KT; ++q) acc[q] = fma(xv[j][k], __ldg(B + k * KT + q), acc[q]);
            }
        }
    }
    double * y = dst + (int64_t)e * s_to + (int64_t)o * KT * inner + i;
#pragma unroll
    for (int q = 0; q < KT; ++q)
    {
        double v = J.coef * acc[q];
        if (J.accumulate) v += y[(int64_t)q * inner];
        y[(int64_t)q * inner] = v;
    }
}

template <int KF, int KT>
static cudaError_t launch_gather_t(const SweepArgs & a, cudaStream_t st)
{
    // all jobs of one launch share `outer` in this variant (checked by the caller): grid.x from job 0
    const int64_t W = (int64_t)a.job[0].outer * a.inner;
    const int64_t total = a.n_elem * W;
    dim3 grid((unsigned)((total + 127) / 128), (unsigned)(a.n_job * a.n_comp));
    sweep_gather_kernel<KF, KT><<<grid, 128, 0, st>>>(a);
    return cudaGetLastError();
}

#define AMDG_DISPATCH_KT(KF_)                                                       \
    switch (kt) {                                                                   \
        case 1: return FN<KF_, 1>(a, st); case 2: return FN<KF_, 2>(a, st);          \
        case 3: return FN<KF_, 3>(a, st); case 4: return FN<KF_, 4>(a, st);          \
        case 5: return FN<KF_, 5>(a, st); case 6: return FN<KF_, 6>(a, st);          \
        default: return cudaErrorInvalidValue; }

bool sweep_shape_supported(int kf, int kt) { return kf >= 1 && kf <= 6 && kt >= 1 && kt <= 6; }

cudaError_t launch_sweep_gather(const SweepArgs & a, int kf, int kt, cudaStream_t st)
{
#define FN launch_gather_t
    switch (kf)
    {
        case 1: AMDG_DISPATCH_KT(1) case 2: AMDG_DISPATCH_KT(2) case 3: AMDG_DISPATCH_KT(3)
        case 4: AMDG_DISPATCH_KT(4) case 5: AMDG_DISPATCH_KT(5) case 6: AMDG_DISPATCH_KT(6)
        default: return cudaErrorInvalidValue;
    }
#undef FN
}

// -------------------------------------------------------------------------------------------------------------
// K1, fibre-staged form.  A CTA owns one work item = a run of whole fibres along dimension t (or, for a fibre too
// long for shared memory, a column range of it).  Phase 1 copies the item's source entries into shared memory
// X[row][k][column] (every source block is read from HBM exactly once per sweep); phase 2 gives every thread
// one target row and up to CT columns: it walks the row's neighbour list, takes the KF source entries per column
// from shared memory and the (KF x KT) operator block from L1/L2 (shared by all lanes of the row), and keeps
// CT*KT accumulators in registers; the epilogue applies coef / accumulate and stores.
// The thread block is 256 threads viewed as (2^lcx column lanes) x (256 >> lcx row lanes), per item.
// -------------------------------------------------------------------------------------------------------------
#ifndef AMDG_FIBRE_THREADS
#define AMDG_FIBRE_THREADS 256
#endif
static const int FIBRE_THREADS = AMDG_FIBRE_THREADS;

static const int FIBRE_SMEM_DOUBLES = 12 * 1024;     // upper bound (96 KiB); the context picks the launch size

int fibre_smem_capacity_doubles() { return FIBRE_SMEM_DOUBLES; }
int fibre_threads() { return FIBRE_THREADS; }

template <int KF, int KT, int CT>
__global__ void __launch_bounds__(FIBRE_THREADS) sweep_fibre_kernel(const FibreSweepArgs a)
{
    extern __shared__ double X[];
#define AMDG_STAMP(i) do { if (a.dbg && threadIdx.x == 0) a.dbg[(int64_t)blockIdx.x * 8 + (i)] = clock64(); } while (0)
    AMDG_STAMP(0);
    const FibreItem it = a.items[blockIdx.x];
    const int jb = blockIdx.y / a.n_comp, comp = blockIdx.y % a.n_comp;
    const SweepJob J = a.job[jb];
    const int inner = a.inner;
    const int W = J.outer * inner;
    const int64_t s_from = (int64_t)W * KF, s_to = (int64_t)W * KT;
    const double * __restrict__ src = J.src + (int64_t)comp * a.n_elem * s_from;
    double * __restrict__ dst = J.dst + (int64_t)comp * a.n_elem * s_to;
    const int P = it.pitch;
    const int cx = 1 << it.lcx;
    const int tid = threadIdx.x;
    if (a.dbg && tid == 0) { a.dbg[(int64_t)blockIdx.x * 8 + 6] = it.packed; a.dbg[(int64_t)blockIdx.x * 8 + 7] = it.packed ? it.nsrc : it.nslot; unsigned smid; asm("mov.u32 %0, %%smid;" : "=r"(smid)); a.dbg[(int64_t)blockIdx.x * 8 + 5] = smid; }
    AMDG_STAMP(1);
    const int tx = tid & (cx - 1);
    const int ncol = min(it.ncol, W - it.col0);

    int cc[CT]; int off_from[CT], off_to[CT]; bool ok[CT];
#pragma unroll
    for (int r = 0; r < CT; ++r)
    {
        cc[r] = tx + r * cx;
        ok[r] = cc[r] < ncol;
        const int col = it.col0 + (ok[r] ? cc[r] : 0);
        const int o = col / inner, i = col - o * inner;
        off_from[r] = o * KF * inner + i;
        off_to[r] = o * KT * inner + i;
    }

    if (it.packed)
    {
        // ---- packed item: source rows, the operator blocks of the item's distinct pairs and the item-local
        // neighbour lists all go to shared memory (cp.async, everything in flight at once); phase 2 touches global
        // memory only for the stores.
        const int * __restrict__ srcs = a.pool_slots + it.src_ofs;       // element rows of the sources
        const int * __restrict__ tgts = a.pool_slots + it.tgt_ofs;       // element rows of the targets
        double * Bs = X + (int64_t)it.nsrc * KF * P;
        int * rowptr = reinterpret_cast<int *>(Bs + (int64_t)it.npair * (KF * KT));      // [ntgt+1]
        int * rsplit = rowptr + it.ntgt + 1;                                               // [ntgt]
        int * telem = rsplit + it.ntgt;                                                    // [ntgt]
        int * ent = telem + it.ntgt;                                                       // [nnz][2]
        const int ty = tid >> it.lcx, ny = FIBRE_THREADS >> it.lcx;
        // all index loads first (sources and pair ids), so their latencies overlap; then every copy is issued
        int e[4]; int pr_id[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) e[u] = (ty + u * ny < it.nsrc) ? __ldg(srcs + ty + u * ny) : -1;
        const int nb_copy = it.npair * (KF * KT);
#pragma unroll
        for (int u = 0; u < 4; ++u) { const int idx = tid + u * FIBRE_THREADS; pr_id[u] = idx < nb_copy ? __ldg(a.pool_pairs + it.pair_ofs + idx / (KF * KT)) : -1; }
#pragma unroll
        for (int u = 0; u < 4; ++u)
        {
            if (e[u] < 0) continue;
            const double * __restrict__ g = src + (int64_t)e[u] * s_from;
            double * xr = X + (int64_t)(ty + u * ny) * KF * P;
#pragma unroll
            for (int r = 0; r < CT; ++r)
            {
                if (!ok[r]) continue;
#pragma unroll
                for (int k = 0; k < KF; ++k) cp_async8(xr + k * P + cc[r], g + off_from[r] + (int64_t)k * inner);
            }
        }
        for (int j0 = ty + 4 * ny; j0 < it.nsrc; j0 += ny)
        {
            const double * __restrict__ g = src + (int64_t)__ldg(srcs + j0) * s_from;
            double * xr = X + (int64_t)j0 * KF * P;
#pragma unroll
            for (int r = 0; r < CT; ++r)
            {
                if (!ok[r]) continue;
#pragma unroll
                for (int k = 0; k < KF; ++k) cp_async8(xr + k * P + cc[r], g + off_from[r] + (int64_t)k * inner);
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
        {
            const int idx = tid + u * FIBRE_THREADS;
            if (pr_id[u] >= 0) cp_async8(Bs + idx, a.blocks + (int64_t)pr_id[u] * (KF * KT) + idx % (KF * KT));
        }
        for (int idx = tid + 4 * FIBRE_THREADS; idx < nb_copy; idx += FIBRE_THREADS)
        {
            const int pr = idx / (KF * KT), r = idx - pr * (KF * KT);
            cp_async8(Bs + idx, a.blocks + (int64_t)__ldg(a.pool_pairs + it.pair_ofs + pr) * (KF * KT) + r);
        }
        cp_async_commit();
        AMDG_STAMP(2);
        for (int j = tid; j <= it.ntgt; j += FIBRE_THREADS) rowptr[j] = __ldg(a.pool_rowptr + it.row_ofs + j);
        for (int j = tid; j < it.ntgt; j += FIBRE_THREADS) { rsplit[j] = __ldg(a.pool_rsplit + it.row_ofs + j); telem[j] = __ldg(tgts + j); }
        const int nnz = __ldg(a.pool_rowptr + it.row_ofs + it.ntgt);
        for (int i = tid; i < nnz; i += FIBRE_THREADS)
        {
            const NbrDev nb = a.pool_ent[it.ent_ofs + i];
            ent[2 * i] = nb.local; ent[2 * i + 1] = nb.pair;
        }
        cp_async_wait_all();
        __syncthreads();
        AMDG_STAMP(3);

        for (int j = ty; j < it.ntgt; j += ny)
        {
            int n0 = rowptr[j], n1 = rowptr[j + 1];
            if (a.lu == AMDG_LU_U) n1 = n0 + rsplit[j];
            else if (a.lu == AMDG_LU_L) n0 = n0 + rsplit[j];
            double acc[CT][KT];
#pragma unroll
            for (int r = 0; r < CT; ++r)
#pragma unroll
                for (int q = 0; q < KT; ++q) acc[r][q] = 0.0;
            for (int p = n0; p < n1; ++p)
            {
                const double * xr = X + (int64_t)ent[2 * p] * KF * P;
                const double * B = Bs + ent[2 * p + 1] * (KF * KT);
#pragma unroll
                for (int k = 0; k < KF; ++k)
                {
                    double bk[KT];
#pragma unroll
                    for (int q = 0; q < KT; ++q) bk[q] = B[k * KT + q];
#pragma unroll
                    for (int r = 0; r < CT; ++r)
                    {
                        const double xv = ok[r] ? xr[k * P + cc[r]] : 0.0;
#pragma unroll
                        for (int q = 0; q < KT; ++q) acc[r][q] = fma(xv, bk[q], acc[r][q]);
                    }
                }
            }
            double * y = dst + (int64_t)telem[j] * s_to;
#pragma unroll
            for (int r = 0; r < CT; ++r)
            {
                if (!ok[r]) continue;
#pragma unroll
                for (int q = 0; q < KT; ++q)
                {
                    double v = J.coef * acc[r][q];
                    double * yp = y + off_to[r] + (int64_t)q * inner;
                    if (J.accumulate) v += *yp;
                    *yp = v;
                }
            }
        }
        AMDG_STAMP(4);
        return;
    }

    // ---- streamed item: stage the whole fibre (column range) ...
    {
        const int ty = tid >> it.lcx, ny = FIBRE_THREADS >> it.lcx;
        for (int j = ty; j < it.nslot; j += ny)
        {
            const int e = a.slot_elem[it.slot0 + j];
            const double * __restrict__ g = src + (int64_t)e * s_from;
            double * xr = X + (int64_t)j * KF * P;
#pragma unroll
            for (int r = 0; r < CT; ++r)
            {
                if (!ok[r]) continue;
#pragma unroll
                for (int k = 0; k < KF; ++k) xr[k * P + cc[r]] = __ldg(g + off_from[r] + (int64_t)k * inner);
            }
        }
    }

    // ... then one warp per target row (the first ntgt rows of the fibre); the 32 lanes are
    // (32/cx neighbour slices) x (cx column lanes); neighbour entries are fetched 32 at a time (coalesced) and
    // broadcast with shuffles, operator blocks come from L2; partial sums of the slices are reduced with shuffles.
    __syncthreads();
    {
        const int lane = tid & 31, warp = tid >> 5;
        constexpr int BWP = KF * KT + 2;                                              // slab row pitch (doubles)
        double * Bw = X + (int64_t)it.nslot * KF * P + (int64_t)warp * 32 * BWP;      // per-warp slab behind the staged rows
        const int cxw = min(cx, 32);
        const int ns = 32 / cxw;
        const int sl = lane / cxw;
        for (int j = warp; j < it.ntgt; j += FIBRE_THREADS / 32)
        {
            const int slot = it.slot0 + j;
            const int frow = 0;                      // a streamed item stages exactly one fibre
            int64_t n0 = a.nbr_ptr[slot], n1 = a.nbr_ptr[slot + 1];
            const int split = a.nbr_split[slot];
            if (a.lu == AMDG_LU_U) n1 = n0 + split;
            else if (a.lu == AMDG_LU_L) n0 = n0 + split;
            double acc[CT][KT];
#pragma unroll
            for (int r = 0; r < CT; ++r)
#pragma unroll
                for (int q = 0; q < KT; ++q) acc[r][q] = 0.0;
            for (int64_t base = n0; base < n1; base += 32)
            {
                const int cnt = (int)min((int64_t)32, n1 - base);
                NbrDev mine; mine.local = 0; mine.pair = 0;
                // every lane fetches one neighbour entry and that entry's operator block (32 blocks in flight per warp)
                // into the warp's shared-memory slab; the slices then read their blocks from there
                __syncwarp();
                if (lane < cnt)
                {
                    mine = a.nbr[base + lane];
                    const double * __restrict__ Bg = a.blocks + (int64_t)mine.pair * (KF * KT);
#pragma unroll
                    for (int r = 0; r < KF * KT; ++r) Bw[lane * BWP + r] = __ldg(Bg + r);
                }
                __syncwarp();
                for (int i0 = 0; i0 < cnt; i0 += ns)
                {
                    const int i = i0 + sl;
                    const bool valid = i < cnt;
                    const int local = __shfl_sync(0xffffffffu, mine.local, valid ? i : 0);
                    if (!valid) continue;
                    const double * xr = X + (int64_t)(frow + local) * KF * P;
                    const double * B = Bw + i * BWP;
#pragma unroll
                    for (int k = 0; k < KF; ++k)
                    {
                        double bk[KT];
#pragma unroll
                        for (int q = 0; q < KT; ++q) bk[q] = B[k * KT + q];
#pragma unroll
                        for (int r = 0; r < CT; ++r)
                        {
                            const double xv = ok[r] ? xr[k * P + cc[r]] : 0.0;
#pragma unroll
                            for (int q = 0; q < KT; ++q) acc[r][q] = fma(xv, bk[q], acc[r][q]);
                        }
                    }
                }
            }
            for (int off = cxw; off < 32; off <<= 1)
            {
#pragma unroll
                for (int r = 0; r < CT; ++r)
#pragma unroll
                    for (int q = 0; q < KT; ++q) acc[r][q] += __shfl_xor_sync(0xffffffffu, acc[r][q], off);
            }
            if (sl == 0)
            {
                double * y = dst + (int64_t)a.slot_elem[slot] * s_to;
#pragma unroll
                for (int r = 0; r < CT; ++r)
                {
                    if (!ok[r]) continue;
#pragma unroll
                    for (int q = 0; q < KT; ++q)
                    {
                        double v = J.coef * acc[r][q];
                        double * yp = y + off_to[r] + (int64_t)q * inner;
                        if (J.accumulate) v += *yp;
                        *yp = v;
                    }
                }
            }
        }
    }
}

template <int KF, int KT, int CT>
static cudaError_t launch_fibre_t(const FibreSweepArgs & a, cudaStream_t st)
{
    static PerDeviceOnce configured;
    const size_t smem = (size_t)std::min(std::max(a.smem_doubles, 256), FIBRE_SMEM_DOUBLES) * sizeof(double);
    if (!configured.done())
    {
        cudaError_t e = cudaFuncSetAttribute(sweep_fibre_kernel<KF, KT, CT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(FIBRE_SMEM_DOUBLES * sizeof(double)));
        if (e != cudaSuccess) return e;
        configured.mark();
    }
    dim3 grid((unsigned)a.n_item, (unsigned)(a.n_job * a.n_comp));
    sweep_fibre_kernel<KF, KT, CT><<<grid, FIBRE_THREADS, smem, st>>>(a);
    return cudaGetLastError();
}

template <int KF, int KT>
static cudaError_t launch_fibre_ct(const FibreSweepArgs & a, int ct, cudaStream_t st)
{
    if (ct <= 1) return launch_fibre_t<KF, KT, 1>(a, st);
    if (ct == 2) return launch_fibre_t<KF, KT, 2>(a, st);
    return launch_fibre_t<KF, KT, 4>(a, st);
}

#define AMDG_DISPATCH_KT_F(KF_)                                                                   \
    switch (kt) {                                                                                 \
        case 1: return launch_fibre_ct<KF_, 1>(a, ct, st); case 2: return launch_fibre_ct<KF_, 2>(a, ct, st); \
        case 3: return launch_fibre_ct<KF_, 3>(a, ct, st); case 4: return launch_fibre_ct<KF_, 4>(a, ct, st); \
        case 5: return launch_fibre_ct<KF_, 5>(a, ct, st); case 6: return launch_fibre_ct<KF_, 6>(a, ct, st); \
        default: return cudaErrorInvalidValue; }

cudaError_t launch_sweep_fibre(const FibreSweepArgs & a, int kf, int kt, int ct, cudaStream_t st)
{
    switch (kf)
    {
        case 1: AMDG_DISPATCH_KT_F(1) case 2: AMDG_DISPATCH_KT_F(2) case 3: AMDG_DISPATCH_KT_F(3)
        case 4: AMDG_DISPATCH_KT_F(4) case 5: AMDG_DISPATCH_KT_F(5) case 6: AMDG_DISPATCH_KT_F(6)
        default: return cudaErrorInvalidValue;
    }
}

// -------------------------------------------------------------------------------------------------------------
// K2 point-wise flux (source/Interplation.cpp:256-295; FluxFunction, source/Interplation.cpp, include/Interpolation.h:396-437)
// -------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double flux_eval(int id, const double * prm, double u, const double * x, int dim)
{
    switch (id)
    {
        case AMDG_FLUX_LINEAR: return prm[0] * u;
        case AMDG_FLUX_BURGERS: return u * u / 2.;
        case AMDG_FLUX_SIN: return sin(u);
        case AMDG_FLUX_COS: return cos(u);
        case AMDG_FLUX_BUCKLEY_X: return u * u / (u * u + (1. - u) * (1. - u));
        case AMDG_FLUX_BUCKLEY_Y: return (u * u * (1. - 5. * (1. - u) * (1. - u))) / (u * u + (1. - u) * (1. - u));
        case AMDG_FLUX_VLASOV_SMOOTH_E:
        {
            // generalised interp_Vlasov_2D2V (source/Interplation.cpp:4508-4580): component t = (int)prm[0];
            // t < dim/2: v_t * f ; t >= dim/2: E_t(x) * f with the prescribed smooth field of the oracle harness
            const int t = (int)prm[0], hd = dim / 2;
            double c;
            if (t < hd) c = x[hd + t];
            else { c = 0.; for (int s = 0; s < hd; ++s) c += sin(2. * 3.1415926535897932384626433832795 * (x[s] + 0.125 * (t - hd + 1))); }
            return c * u;
        }
    }
    return 0.;
}

__global__ void __launch_bounds__(256) pointwise_kernel(const PointwiseArgs a)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < a.n_points; p += stride)
    {
        const double u = a.up[p];
        double x[8];
        if (a.pts) { for (int t = 0; t < a.dim; ++t) x[t] = a.pts[p * a.dim + t]; }
        for (int c = 0; c < a.n_flux; ++c) a.fp[(int64_t)c * a.n_points + p] = flux_eval(a.flux_id[c], a.params[c], u, x, a.dim);
    }
}

cudaError_t launch_pointwise(const PointwiseArgs & a, cudaStream_t st)
{
    const int64_t nb = (a.n_points + 255) / 256;
    pointwise_kernel<<<(unsigned)(nb < 148 * 16 ? (nb > 0 ? nb : 1) : 148 * 16), 256, 0, st>>>(a);
    return cudaGetLastError();
}

// point-wise expressions: LagrInterpolation::eval_fp_Lag with all VEC_NUM unknowns (source/Interplation.cpp:256-295), eval_coe_u_Lag with a
// coefficient of position (:648-698) and the Vlasov bodies with the field values of DGSolution::copy_up_intp_to_f (:4451-4497, 4523-4573;
// source/DGSolution.cpp:1024-1065).  The reference takes a std::function; here every output is a small stack program over
//   VAR v (point value of unknown v), X t (coordinate t of the interpolation point, from the 1D point table and the element's 1D orders),
//   OTHER j (point value of field j at the same local point of the field element this element maps to), CONST, + - * / and a few functions.
__global__ void __launch_bounds__(256) pointwise_expr_kernel(const PwExprArgs a)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < a.n_points; p += stride)
    {
        const int64_t e = p / a.block; const int loc = (int)(p - e * a.block);
        for (int c = 0; c < a.n_out; ++c)
        {
            double st[PW_STACK]; int sp = 0;
            for (int i = a.out_ptr[c]; i < a.out_ptr[c + 1]; ++i)
            {
                const int op = a.op[i], arg = a.arg[i];
                switch (op)
                {
                    case PW_VAR: st[sp++] = a.up[arg][p]; break;
                    case PW_X: { const int pt = (loc / a.stride[arg]) % a.edge; st[sp++] = __ldg(a.pts1d + __ldg(a.ord1d + e * a.dim + arg) * a.edge + pt); break; }
                    case PW_OTHER: { const int64_t eo = a.other_map ? __ldg(a.other_map + e) : e; st[sp++] = a.other[arg][eo * a.block + loc]; break; }
                    case PW_CONST: st[sp++] = a.consts[arg]; break;
                    case PW_ADD: --sp; st[sp - 1] += st[sp]; break;
                    case PW_SUB: --sp; st[sp - 1] -= st[sp]; break;
                    case PW_MUL: --sp; st[sp - 1] *= st[sp]; break;
                    case PW_DIV: --sp; st[sp - 1] /= st[sp]; break;
                    case PW_NEG: st[sp - 1] = -st[sp - 1]; break;
                    case PW_SIN: st[sp - 1] = sin(st[sp - 1]); break;
                    case PW_COS: st[sp - 1] = cos(st[sp - 1]); break;
                    case PW_SQR: st[sp - 1] *= st[sp - 1]; break;
                    case PW_EXP: st[sp - 1] = exp(st[sp - 1]); break;
                    case PW_SQRT: st[sp - 1] = sqrt(st[sp - 1]); break;
                    case PW_ABS: st[sp - 1] = fabs(st[sp - 1]); break;
                    case PW_POW: --sp; st[sp - 1] = pow(st[sp - 1], st[sp]); break;
                    case PW_TANH: st[sp - 1] = tanh(st[sp - 1]); break;
                    case PW_MIN: --sp; st[sp - 1] = fmin(st[sp - 1], st[sp]); break;
                    case PW_MAX: --sp; st[sp - 1] = fmax(st[sp - 1], st[sp]); break;
                    default: break;
                }
            }
            a.out[c][p] = st[0];
        }
    }
}

cudaError_t launch_pointwise_expr(const PwExprArgs & a, cudaStream_t st)
{
    const int64_t nb = (a.n_points + 255) / 256;
    pointwise_expr_kernel<<<(unsigned)(nb < 148 * 16 ? (nb > 0 ? nb : 1) : 148 * 16), 256, 0, st>>>(a);
    return cudaGetLastError();
}

// rows of a local array to mapped destinations (peer memory): the part of a layout switch whose data no sweep produces at the right moment
__global__ void __launch_bounds__(256) scatter_rows_kernel(const double * __restrict__ src, int64_t n_rows, int width, double * __restrict__ dst,
                                                           const long long * __restrict__ map)
{
    const int64_t total = n_rows * width, stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += stride)
    {
        const int64_t e = g / width; const int i = (int)(g - e * width);
        dst[__ldg(map + e) + i] = src[g];
    }
}

cudaError_t launch_scatter_rows(const double * src, int64_t n_rows, int width, double * dst, const long long * map, cudaStream_t st)
{
    const int64_t nb = (n_rows * width + 255) / 256;
    scatter_rows_kernel<<<(unsigned)(nb < 148 * 16 ? (nb > 0 ? nb : 1) : 148 * 16), 256, 0, st>>>(src, n_rows, width, dst, map);
    return cudaGetLastError();
}

// element rows of an array into a compact array and back (the sub-grid of the *_coarse_grid transforms, capi.cu: amdg_apply_tensor_coarse):
// gather: dst[e][i] = src[rows[e]][i];  scatter-add: dst[rows[e]][i] += src[e][i]
__global__ void __launch_bounds__(256) rows_gather_kernel(const double * __restrict__ src, const int * __restrict__ rows, int64_t n_rows, int width, double * __restrict__ dst)
{
    const int64_t total = n_rows * width, stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += stride)
    {
        const int64_t e = g / width; const int i = (int)(g - e * width);
        dst[g] = src[(int64_t)__ldg(rows + e) * width + i];
    }
}
__global__ void __launch_bounds__(256) rows_scatter_add_kernel(const double * __restrict__ src, const int * __restrict__ rows, int64_t n_rows, int width, double * __restrict__ dst)
{
    const int64_t total = n_rows * width, stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += stride)
    {
        const int64_t e = g / width; const int i = (int)(g - e * width);
        dst[(int64_t)__ldg(rows + e) * width + i] += src[g];
    }
}
cudaError_t launch_rows_gather(const double * src, const int * rows, int64_t n_rows, int width, double * dst, cudaStream_t st)
{
    const int64_t nb = (n_rows * width + 255) / 256;
    rows_gather_kernel<<<(unsigned)(nb < 148 * 16 ? (nb > 0 ? nb : 1) : 148 * 16), 256, 0, st>>>(src, rows, n_rows, width, dst);
    return cudaGetLastError();
}
cudaError_t launch_rows_scatter_add(const double * src, const int * rows, int64_t n_rows, int width, double * dst, cudaStream_t st)
{
    const int64_t nb = (n_rows * width + 255) / 256;
    rows_scatter_add_kernel<<<(unsigned)(nb < 148 * 16 ? (nb > 0 ? nb : 1) : 148 * 16), 256, 0, st>>>(src, rows, n_rows, width, dst);
    return cudaGetLastError();
}

// Cross-GPU barrier over peer-mapped flags (one process per GPU, flags in IPC-shared device memory).  Thread r < world: publish this rank's
// new epoch in rank r's flag array (system-scope release: everything earlier kernels of this stream stored to peers is visible first), then wait
// until rank r's epoch has arrived here (acquire).  The epoch lives on the device so that the kernel can be replayed from a CUDA graph.  A peer
// that never arrives makes the kernel give up after timeout_cycles and raise *error instead of hanging the device.
__global__ void peer_barrier_kernel(const PeerBarrierArgs a)
{
    __shared__ unsigned s_epoch;
    if (threadIdx.x == 0) { s_epoch = *a.epoch + 1; }
    __syncthreads();
    const unsigned epoch = s_epoch;
    const int r = threadIdx.x;
    if (r < a.world)
    {
        __threadfence_system();
        unsigned * remote = a.flags_of[r] + a.rank;
        asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(remote), "r"(epoch) : "memory");
        const unsigned * mine = a.flags_of[a.rank] + r;
        const long long t0 = clock64();
        unsigned v = 0;
        while (true)
        {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
            if ((int)(v - epoch) >= 0) break;
            if (clock64() - t0 > a.timeout_cycles) { atomicExch(a.error, 1u); break; }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) { *a.epoch = epoch; }
}

cudaError_t launch_peer_barrier(const PeerBarrierArgs & a, cudaStream_t st)
{
    peer_barrier_kernel<<<1, 32, 0, st>>>(a);
    return cudaGetLastError();
}

// Hermite (PMAX=3) point-wise flux in 2D, scalar: HermInterpolation::eval_fp_Her_2D (source/Interplation.cpp:2045-2288).
// Local index along a dim: 0,1 = value at point 0,1; 2,3 = derivative at point 0,1 (deg_pt_deri_1d).  Block (a,b):
//   value/value: f(u);  deriv/value: f'(u) u_x;  value/deriv: f'(u) u_y;  deriv/deriv: f''(u) u_x u_y + f'(u) u_xy
__device__ __forceinline__ void flux_derivs(int id, const double * prm, double u, double & f, double & f1, double & f2)
{
    switch (id)
    {
        case AMDG_FLUX_LINEAR: f = prm[0] * u; f1 = prm[0]; f2 = 0.; break;
        case AMDG_FLUX_BURGERS: f = u * u / 2.; f1 = u; f2 = 1.; break;
        case AMDG_FLUX_SIN: f = sin(u); f1 = cos(u); f2 = -sin(u); break;
        case AMDG_FLUX_COS: f = cos(u); f1 = -sin(u); f2 = -cos(u); break;
        default: f = f1 = f2 = 0.;
    }
}

__global__ void __launch_bounds__(256) pointwise_herm2d_kernel(const PointwiseArgs a)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < a.n_points; p += stride)
    {
        const int64_t e = p >> 4; const int loc = (int)(p & 15);
        const int ia = loc >> 2, ib = loc & 3;
        const double * up = a.up + e * 16;
        const double u = up[(ia & 1) * 4 + (ib & 1)];                    // value slot of the same point
        const double w = up[loc];
        for (int c = 0; c < a.n_flux; ++c)
        {
            double f, f1, f2; flux_derivs(a.flux_id[c], a.params[c], u, f, f1, f2);
            double v;
            if (ia < 2 && ib < 2) v = f;
            else if (ia >= 2 && ib >= 2) v = f2 * up[ia * 4 + (ib - 2)] * up[(ia - 2) * 4 + ib] + f1 * w;
            else v = f1 * w;
            a.fp[(int64_t)c * a.n_points + p] = v;
        }
    }
}

cudaError_t launch_pointwise_herm2d(const PointwiseArgs & a, cudaStream_t st)
{
    const int64_t nb = (a.n_points + 255) / 256;
    pointwise_herm2d_kernel<<<(unsigned)(nb < 148 * 16 ? (nb > 0 ? nb : 1) : 148 * 16), 256, 0, st>>>(a);
    return cudaGetLastError();
}

// interpolation point coordinates of all element points: pts[e][p][t] = pts1d[ord1d[e][t]*edge + p_t]
__global__ void __launch_bounds__(256) point_coords_kernel(const double * __restrict__ pts1d, const int * __restrict__ ord1d,
                                                           int64_t n_elem, int dim, int edge, int block, double * __restrict__ pts)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t total = n_elem * block;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += stride)
    {
        const int64_t e = g / block; int r = (int)(g - e * block);
        for (int t = dim - 1; t >= 0; --t)
        {
            const int p = r % edge; r /= edge;
            pts[g * dim + t] = pts1d[ord1d[e * dim + t] * edge + p];
        }
    }
}

cudaError_t launch_point_coords(const double * pts1d, const int * ord1d, int64_t n_elem, int dim, int edge, double * pts, cudaStream_t st)
{
    int block = 1; for (int t = 0; t < dim; ++t) block *= edge;
    const int64_t nb = (n_elem * block + 255) / 256;
    point_coords_kernel<<<(unsigned)(nb < 148 * 16 ? (nb > 0 ? nb : 1) : 148 * 16), 256, 0, st>>>(pts1d, ord1d, n_elem, dim, edge, block, pts);
    return cudaGetLastError();
}

// -------------------------------------------------------------------------------------------------------------
// K4 explicit RK stage (source/ODESolver.cpp:209-301): u <- c0*u_tn + c1*(u_base + c2*dt*rhs)
// -------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) rk_stage_kernel(double c_tn, double c_u, double c_rhs, const double * __restrict__ u_tn,
                                                       double * __restrict__ u, const double * __restrict__ rhs, int64_t n)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += stride)
    {
        // same association as the reference expressions, e.g. 3/4*u_tn + 1/4*(u + dt*rhs)
        double v = c_rhs * rhs[p];
        if (c_u != 0.0) v = c_u * (u[p] + v);
        u[p] = (c_tn == 1.0 ? u_tn[p] : c_tn * u_tn[p]) + v;
    }
}

cudaError_t launch_rk_stage(int scheme, int stage, double dt, const double * u_tn, double * u, const double * rhs, int64_t n, cudaStream_t st)
{
    // u = c_tn*u_tn + [c_u != 0 ? c_u*(u + c_rhs*rhs) : c_rhs*rhs]
    double c_tn = 1., c_u = 0., c_rhs = dt;
    if (scheme == AMDG_RK_EULER) { if (stage != 0) return cudaErrorInvalidValue; }
    else if (scheme == AMDG_RK_RK2SSP) { if (stage == 1) { c_tn = 0.5; c_u = 0.5; } else if (stage != 0) return cudaErrorInvalidValue; }
    else if (scheme == AMDG_RK_RK2MID) { if (stage == 0) c_rhs = 0.5 * dt; else if (stage != 1) return cudaErrorInvalidValue; }
    else if (scheme == AMDG_RK_RK3SSP)
    {
        if (stage == 1) { c_tn = 3. / 4.; c_u = 1. / 4.; }
        else if (stage == 2) { c_tn = 1. / 3.; c_u = 2. / 3.; }
        else if (stage != 0) return cudaErrorInvalidValue;
    }
    else if (scheme == AMDG_RK_RK3HEUN)            // RK3HeunLinear::step_stage, source/ODESolver.cpp:314-330
    {
        if (stage == 0) c_rhs = 1. / 3. * dt; else if (stage == 1) c_rhs = 1. / 2. * dt; else if (stage != 2) return cudaErrorInvalidValue;
    }
    else return cudaErrorInvalidValue;
    const int64_t nb = (n + 255) / 256;
    rk_stage_kernel<<<(unsigned)(nb < 148 * 16 ? (nb > 0 ? nb : 1) : 148 * 16), 256, 0, st>>>(c_tn, c_u, c_rhs, u_tn, u, rhs, n);
    return cudaGetLastError();
}

// RK4ODE2nd::step_stage (source/ODESolver.cpp:578-615) for u_tt = L u written as (u, v = u_t): k_u = dt v, k_v = dt rhs are
// kept per stage ([4][n] each); stages 0..2 set (u, v) = (u_tn, v_tn) + c k, stage 3 the classical 1/6 (k1 + 2 k2 + 2 k3 + k4)
__global__ void __launch_bounds__(256) rk4_ode2nd_stage_kernel(int stage, double dt, const double * __restrict__ u_tn, const double * __restrict__ v_tn,
                                                               double * __restrict__ u, double * __restrict__ v, const double * __restrict__ rhs,
                                                               double * __restrict__ ku, double * __restrict__ kv, int64_t n)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += stride)
    {
        const double a = dt * v[p], b = dt * rhs[p];
        if (stage < 3)
        {
            ku[(int64_t)stage * n + p] = a; kv[(int64_t)stage * n + p] = b;
            if (stage < 2) { u[p] = u_tn[p] + a / 2.; v[p] = v_tn[p] + b / 2.; }
            else { u[p] = u_tn[p] + a; v[p] = v_tn[p] + b; }
        }
        else
        {
            u[p] = u_tn[p] + 1. / 6. * (ku[p] + 2 * ku[n + p] + 2 * ku[2 * n + p] + a);
            v[p] = v_tn[p] + 1. / 6. * (kv[p] + 2 * kv[n + p] + 2 * kv[2 * n + p] + b);
        }
    }
}

cudaError_t launch_rk4_ode2nd_stage(int stage, double dt, const double * u_tn, const double * v_tn, double * u, double * v, const double * rhs,
                                    double * ku, double * kv, int64_t n, cudaStream_t st)
{
    if (stage < 0 || stage > 3) return cudaErrorInvalidValue;
    const int64_t nb = (n + 255) / 256;
    rk4_ode2nd_stage_kernel<<<(unsigned)(nb < 148 * 16 ? (nb > 0 ? nb : 1) : 148 * 16), 256, 0, st>>>(stage, dt, u_tn, v_tn, u, v, rhs, ku, kv, n);
    return cudaGetLastError();
}

__global__ void __launch_bounds__(256) axpby_kernel(int64_t n, double alpha, const double * __restrict__ x, double beta, double * __restrict__ y)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += stride)
        y[p] = (beta == 0.0) ? alpha * x[p] : alpha * x[p] + beta * y[p];
}

// y = sum_i c_i x_i (+ beta y): the right-hand sides of several tensor applications, penalty terms and pushed partial sums joined in one pass
__global__ void __launch_bounds__(256) lincomb_kernel(const LincombArgs a)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < a.n; p += stride)
    {
        double v = a.beta == 0.0 ? 0.0 : a.beta * a.y[p];
        for (int i = 0; i < a.k; ++i) v += a.c[i] * a.x[i][p];
        a.y[p] = v;
    }
}

cudaError_t launch_lincomb(const LincombArgs & a, cudaStream_t st)
{
    const int64_t nb = (a.n + 255) / 256;
    lincomb_kernel<<<(unsigned)(nb < 148 * 16 ? (nb > 0 ? nb : 1) : 148 * 16), 256, 0, st>>>(a);
    return cudaGetLastError();
}

// Velocity moments of f accumulated into the right-hand side of a field solution (DGAdapt::compute_moment_1D2V / _2D2V, reference
// source/DGAdapt.cpp:243-338): for every field element (level 0 in the n_v trailing dimensions) that has a partner in f (map >= 0) and every index
// x of the leading dimensions,  rhs_E[e][x, 0..0] += weight * sum_{dv} prod_i c(order_i, dv_i) f[map[e]][x, dv],  dv_i in 0..order_i,
// c(0,0) = 1, c(1,0) = 1/2, c(1,1) = 1/(2 sqrt 3) (moments of the level-0 Alpert basis on [0,1]; higher levels have vanishing moments).
__global__ void __launch_bounds__(256) moment_kernel(const MomentArgs a)
{
    const int64_t n = a.n_field * a.x_block;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += stride)
    {
        const int64_t e = p / a.x_block; const int x = (int)(p - e * a.x_block);
        const int fe = a.map[e];
        if (fe < 0) continue;
        const double * __restrict__ f = a.f + ((int64_t)fe * a.x_block + x) * a.v_block;
        double s = 0.0;
        for (int combo = 0; combo < a.n_combo; ++combo) s += a.coef[combo] * f[a.offset[combo]];
        a.rhs[((int64_t)e * a.x_block + x) * a.v_block] += a.weight * s;
    }
}

cudaError_t launch_moment(const MomentArgs & a, cudaStream_t st)
{
    const int64_t nb = (a.n_field * a.x_block + 255) / 256;
    moment_kernel<<<(unsigned)(nb < 148 * 16 ? (nb > 0 ? nb : 1) : 148 * 16), 256, 0, st>>>(a);
    return cudaGetLastError();
}

// Refinement / coarsening indicator of DGAdapt (reference source/DGAdapt.cpp:1018-1030): norm[e] = sum over the indicator variables of the l2
// norm of the element's Alpert coefficients.  One warp per element row, lanes over the block, fixed-order shuffle reduction (deterministic).
__global__ void __launch_bounds__(256) indicator_norm_kernel(const IndicatorArgs a)
{
    const int lane = threadIdx.x & 31;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t e = (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5); e < a.n_elem; e += warps)
    {
        double norm = 0.0;
        for (int v = 0; v < a.n_var; ++v)
        {
            const double * __restrict__ u = a.u[v] + e * a.block;
            double s = 0.0;
            for (int i = lane; i < a.block; i += 32) { const double x = u[i]; s += x * x; }
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            norm += sqrt(s);
        }
        if (lane == 0) a.norm[e] = norm;
    }
}

cudaError_t launch_indicator_norm(const IndicatorArgs & a, cudaStream_t st)
{
    const int64_t nb = (a.n_elem + 7) / 8;
    indicator_norm_kernel<<<(unsigned)(nb < 148 * 16 ? (nb > 0 ? nb : 1) : 148 * 16), 256, 0, st>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_axpby(int64_t n, double alpha, const double * x, double beta, double * y, cudaStream_t st)
{
    const int64_t nb = (n + 255) / 256;
    axpby_kernel<<<(unsigned)(nb < 148 * 16 ? (nb > 0 ? nb : 1) : 148 * 16), 256, 0, st>>>(n, alpha, x, beta, y);
    return cudaGetLastError();
}

}  // namespace amdg
