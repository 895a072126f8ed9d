// FP64 CUDA kernels for sm_100a.  See kernels.cuh for the map to the reference functions.
#include <algorithm>
#include "kernels.cuh"
#include "../../include/amdg.h"

namespace amdg {

static const int PIPE_HDR_INTS = 16;     // == PIPE_HDR of pipe_items.hpp

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
    for (int64_t p = n0; p < n1; ++p)
    {
        const NbrDev nb = a.nbr[p];
        const int f = a.slot_elem[fbase + nb.local];
        const double * __restrict__ x = src + (int64_t)f * s_from + col_off;
        const double * __restrict__ B = a.blocks + (int64_t)nb.pair * (KF * KT);
#pragma unroll
        for (int k = 0; k < KF; ++k)
        {
            const double xv = __ldg(x + (int64_t)k * inner);
#pragma unroll
            for (int q = 0; q < KT; ++q) acc[q] = fma(xv, __ldg(B + k * KT + q), acc[q]);
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
__device__ __forceinline__ void cp_async8(void * smem, const void * gmem)
{
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async4(void * smem, const void * gmem)
{
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

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
    static bool configured = false;
    const size_t smem = (size_t)std::min(std::max(a.smem_doubles, 256), FIBRE_SMEM_DOUBLES) * sizeof(double);
    if (!configured)
    {
        cudaError_t e = cudaFuncSetAttribute(sweep_fibre_kernel<KF, KT, CT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(FIBRE_SMEM_DOUBLES * sizeof(double)));
        if (e != cudaSuccess) return e;
        configured = true;
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
// K1, pipelined persistent form (default).  The work list (pipe_items.hpp) is a sequence of self-contained records;
// a CTA walks its share of the list (cost-sorted, round-robin) with a 4-deep software pipeline of async copies:
//     iteration n:   table entry of item n+3  ->  record of item n+2  ->  source rows + operator blocks of item n+1
//                    (all cp.async, one commit group)            ||   compute item n from shared memory
// so the HBM latency of the data, of the index records and of the table entries is hidden behind the FMA work of
// earlier items, and each source block is read from HBM once per sweep.  Shared memory: 2 data stages + 3 record
// stages.  Thread = (row lane, column lane) with CT adjacent columns (16-byte shared loads) x KT accumulators.
// Rows above the cut of a long fibre are produced as partial sums per subtree and added up by whichever CTA
// finishes that (fibre, column chunk) last (arrival counter + __threadfence), deterministically.
// -------------------------------------------------------------------------------------------------------------
static const int PIPE_THREADS = 512;                  // one CTA per SM, 16 warps
static const int PIPE_SMEM_BUDGET = 208 * 1024;       // bytes per CTA

int pipe_threads() { return PIPE_THREADS; }
int pipe_smem_budget_bytes() { return PIPE_SMEM_BUDGET; }

__device__ __forceinline__ void cp_async16(void * smem, const void * gmem)
{
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gmem) : "memory");
}

// per-(job, component) view resolved once per kernel into shared memory
struct PipeView { const double * src; double * dst; double coef; int W; int accumulate; };

template <int KF, int KT, int CT>
struct PipeLanes
{
    int cc[CT], off_from[CT], off_to[CT]; bool ok[CT]; bool pair16;
    int ty, ny, P, nsrc, ntgt, npair;
    __device__ __forceinline__ void init(const int * m, int W, int inner, unsigned inner_magic, int tid, bool aligned16)
    {
        nsrc = m[0]; ntgt = m[1]; npair = m[2];
        const int col0 = m[4], lcx = m[6]; P = m[7];
        const int ncol = min(m[5], W - col0);
        const int cx = 1 << lcx;
        const int tx = tid & (cx - 1); ty = tid >> lcx; ny = PIPE_THREADS >> lcx;
#pragma unroll
        for (int r = 0; r < CT; ++r)
        {
            cc[r] = tx * CT + r;
            ok[r] = cc[r] < ncol;
            const int col = col0 + (ok[r] ? cc[r] : 0);
            const int o = inner == 1 ? col : (int)__umulhi((unsigned)col, inner_magic), i = col - o * inner;      // col / inner, exact for col*inner < 2^32
            off_from[r] = o * KF * inner + i;
            off_to[r] = o * KT * inner + i;
        }
        pair16 = false;
        if (CT >= 2) pair16 = aligned16 && ((inner & 1) == 0) && ((col0 & 1) == 0) && ok[CT - 1];
    }
};

template <int KF, int KT, int CT>
__global__ void __launch_bounds__(PIPE_THREADS, 1) sweep_pipe_kernel(const PipeArgs a)
{
    extern __shared__ __align__(16) double smem_pipe[];
    double * data = smem_pipe;                                                        // [2][data_doubles]
    int * meta = reinterpret_cast<int *>(data + 2 * (int64_t)a.data_doubles);         // [3][meta_ints]
    int2 * tabr = reinterpret_cast<int2 *>(meta + 3 * a.meta_ints);                   // [4]
    PipeView * views = reinterpret_cast<PipeView *>(tabr + 4);                         // [gy]
    __shared__ int s_last;
    const int tid = threadIdx.x;
    const int gy = a.n_job * a.n_comp;
    const int inner = a.inner;
    const unsigned inner_magic = a.inner_magic;
    if (tid < gy)
    {
        const SweepJob J = a.job[tid / a.n_comp];
        const int W = J.outer * inner;
        PipeView v;
        v.src = J.src + (int64_t)(tid % a.n_comp) * a.n_elem * ((int64_t)W * KF);
        v.dst = J.dst + (int64_t)(tid % a.n_comp) * a.n_elem * ((int64_t)W * KT);
        v.coef = J.coef; v.W = W; v.accumulate = J.accumulate;
        views[tid] = v;
    }
    // work index w = blockIdx.x + n*gridDim.x -> (item, by) kept incrementally for n, n+1, n+2, n+3
    const int qd = gridDim.x / gy, rd = gridDim.x % gy;
    int it_i[4], it_b[4];
    it_i[0] = blockIdx.x / gy; it_b[0] = blockIdx.x % gy;
#pragma unroll
    for (int u = 1; u < 4; ++u) { it_i[u] = it_i[u - 1] + qd; it_b[u] = it_b[u - 1] + rd; if (it_b[u] >= gy) { it_b[u] -= gy; ++it_i[u]; } }
    // before the loop the four slots hold items n=0..3; iteration n (starting at -3) needs items n+3, n+2, n+1, n
    long long t_wait = 0, t_issue = 0, t_comp = 0, t_fin = 0; int n_iter = 0;
    PipeLanes<KF, KT, CT> Lc, Ln;          // lanes of the item being computed / of the item whose data is being fetched
    for (int n = -3;; ++n)
    {
        const long long c0 = a.dbg ? clock64() : 0;
        cp_async_wait_all();
        __syncthreads();
        const long long c1 = a.dbg ? clock64() : 0;
        t_wait += c1 - c0;
        // slot of item k (k >= 0) in the incremental table: ring of 4 advanced below
        // ---- issue: table entry of item n+3, record of item n+2, data of item n+1
        {
            const int i3 = it_i[(n + 3) & 3];
            if (i3 < a.n_item && tid == 0) cp_async8(&tabr[(n + 3) & 3], &a.tab[i3]);
            if (n + 2 >= 0 && it_i[(n + 2) & 3] < a.n_item)
            {
                const int2 te = tabr[(n + 2) & 3];
                int * md = meta + ((n + 2) % 3) * a.meta_ints;
                for (int c = tid; c < (te.y >> 2); c += PIPE_THREADS) cp_async16(md + 4 * c, a.rec + te.x + 4 * c);
            }
            if (n + 1 >= 0 && it_i[(n + 1) & 3] < a.n_item)
            {
                const PipeView V = views[it_b[(n + 1) & 3]];
                const int64_t s_from = (int64_t)V.W * KF;
                const int * m = meta + ((n + 1) % 3) * a.meta_ints;
                Ln.init(m, V.W, inner, inner_magic, tid, ((reinterpret_cast<uintptr_t>(V.src) & 15) == 0) && ((s_from & 1) == 0));
                double * X = data + ((n + 1) & 1) * (int64_t)a.data_doubles;
                const int * m_src = m + PIPE_HDR_INTS;
                const int rowstride = KF * Ln.P;
                for (int row = Ln.ty; row < Ln.nsrc; row += Ln.ny)
                {
                    const double * __restrict__ g = V.src + (int64_t)m_src[row] * s_from;
                    double * xr = X + row * rowstride;
                    if (Ln.pair16)
                    {
#pragma unroll
                        for (int k = 0; k < KF; ++k)
#pragma unroll
                            for (int r = 0; r < CT; r += 2) cp_async16(xr + k * Ln.P + Ln.cc[r], g + Ln.off_from[r] + k * inner);
                    }
                    else
                    {
#pragma unroll
                        for (int r = 0; r < CT; ++r)
                        {
                            if (!Ln.ok[r]) continue;
#pragma unroll
                            for (int k = 0; k < KF; ++k) cp_async8(xr + k * Ln.P + Ln.cc[r], g + Ln.off_from[r] + k * inner);
                        }
                    }
                }
                double * Bs = X + ((Ln.nsrc * rowstride + 1) & ~1);
                const int * m_pair = m + PIPE_HDR_INTS + Ln.nsrc + 2 * Ln.ntgt + 1;
                if constexpr ((KF * KT) % 2 == 0)
                {
                    constexpr int CH = (KF * KT) / 2;          // 16-byte chunks per operator block
                    const int nch = Ln.npair * CH;
                    for (int idx = tid; idx < nch; idx += PIPE_THREADS)
                    {
                        const int pr = idx / CH, r = idx - pr * CH;
                        cp_async16(Bs + 2 * idx, a.blocks + (int64_t)m_pair[pr] * (KF * KT) + 2 * r);
                    }
                }
                else
                {
                    const int nb_copy = Ln.npair * (KF * KT);
                    for (int idx = tid; idx < nb_copy; idx += PIPE_THREADS)
                    {
                        const int pr = idx / (KF * KT), r = idx - pr * (KF * KT);
                        cp_async8(Bs + idx, a.blocks + (int64_t)m_pair[pr] * (KF * KT) + r);
                    }
                }
            }
            cp_async_commit();
        }
        const long long c2 = a.dbg ? clock64() : 0;
        t_issue += c2 - c1;
        if (n < 0) { Lc = Ln; continue; }
        const int item = it_i[n & 3], by = it_b[n & 3];
        if (item >= a.n_item) break;
        ++n_iter;

        // ---- compute item n (its lanes were set up when its data was issued, one iteration ago)
        const PipeView V = views[by];
        const int64_t s_to = (int64_t)V.W * KT;
        double * __restrict__ dst = V.dst;
        const int * m = meta + (n % 3) * a.meta_ints;
        const PipeLanes<KF, KT, CT> & L = Lc;
        const int rowstride = KF * L.P;
        const double * X = data + (n & 1) * (int64_t)a.data_doubles;
        const double * Bs = X + ((L.nsrc * rowstride + 1) & ~1);
        const int * m_dest = m + PIPE_HDR_INTS + L.nsrc;
        const int * m_rowptr = m_dest + L.ntgt;
        const int * m_ent = m + m[9];
        const int final_idx = m[8];
        for (int j = L.ty; j < L.ntgt && L.ok[0]; j += L.ny)
        {
            const int n0 = m_rowptr[j], n1 = m_rowptr[j + 1];
            double acc[CT][KT];
#pragma unroll
            for (int r = 0; r < CT; ++r)
#pragma unroll
                for (int q = 0; q < KT; ++q) acc[r][q] = 0.0;
            for (int p = n0; p < n1; ++p)
            {
                const int2 en = *reinterpret_cast<const int2 *>(m_ent + 2 * p);
                const double * xr = X + en.x * rowstride;
                const double * B = Bs + en.y * (KF * KT);
#pragma unroll
                for (int k = 0; k < KF; ++k)
                {
                    double bk[KT];
#pragma unroll
                    for (int q = 0; q < KT; ++q) bk[q] = B[k * KT + q];
                    double xv[CT];
                    if (CT >= 2)
                    {
#pragma unroll
                        for (int r = 0; r < CT; r += 2)
                        {
                            const double2 v = *reinterpret_cast<const double2 *>(xr + k * L.P + L.cc[r]);
                            xv[r] = v.x; xv[r + 1] = v.y;
                        }
                    }
                    else xv[0] = xr[k * L.P + L.cc[0]];
#pragma unroll
                    for (int r = 0; r < CT; ++r)
#pragma unroll
                        for (int q = 0; q < KT; ++q) acc[r][q] = fma(xv[r], bk[q], acc[r][q]);
                }
            }
            const int dest = m_dest[j];
            if (dest >= 0)
            {
                double * y = dst + (int64_t)dest * s_to;
#pragma unroll
                for (int r = 0; r < CT; ++r)
                {
                    if (!L.ok[r]) continue;
#pragma unroll
                    for (int q = 0; q < KT; ++q)
                    {
                        double v = V.coef * acc[r][q];
                        double * yp = y + L.off_to[r] + q * inner;
                        if (V.accumulate) v += *yp;
                        *yp = v;
                    }
                }
            }
            else
            {
                double * y = a.partial + ((int64_t)by * a.n_slot + (-(dest + 1))) * s_to;
#pragma unroll
                for (int r = 0; r < CT; ++r)
                {
                    if (!L.ok[r]) continue;
#pragma unroll
                    for (int q = 0; q < KT; ++q) __stcg(y + L.off_to[r] + q * inner, acc[r][q]);
                }
            }
        }
        const long long c3 = a.dbg ? clock64() : 0;
        t_comp += c3 - c2;
        if (final_idx >= 0)
        {
            // arrival: the last item of this (fibre, column chunk, job) adds the partial sums of the top rows
            const int * fq = a.fin + a.fin_ofs[final_idx];
            __threadfence();
            __syncthreads();
            if (tid == 0) s_last = (atomicAdd(&a.counters[(int64_t)final_idx * gy + by], 1) == fq[0] - 1);
            __syncthreads();
            if (s_last)
            {
                __threadfence();
                const int ntop = fq[1], fc0 = fq[2], fnc = min(fq[3], V.W - fq[2]);
                const int warp = tid >> 5, lane = tid & 31;
                const int * rp = fq + 4;
                const double * pbase = a.partial + (int64_t)by * a.n_slot * s_to;
                for (int row = 0; row < ntop; ++row)
                {
                    const int elem = rp[0], nsl = rp[1];
                    if ((row & (PIPE_THREADS / 32 - 1)) == warp)
                    {
                        double * y = dst + (int64_t)elem * s_to;
                        for (int o = lane; o < fnc * KT; o += 32)
                        {
                            const int c = o / KT, q = o - c * KT;
                            const int col = fc0 + c;
                            const int oo = inner == 1 ? col : (int)__umulhi((unsigned)col, inner_magic), ii = col - oo * inner;
                            const int64_t off = (int64_t)oo * KT * inner + (int64_t)q * inner + ii;
                            double sum = 0.0;
                            int s2 = 0;
                            for (; s2 + 4 <= nsl; s2 += 4)
                            {
                                const double v0 = __ldcg(pbase + (int64_t)rp[2 + s2] * s_to + off), v1 = __ldcg(pbase + (int64_t)rp[3 + s2] * s_to + off);
                                const double v2 = __ldcg(pbase + (int64_t)rp[4 + s2] * s_to + off), v3 = __ldcg(pbase + (int64_t)rp[5 + s2] * s_to + off);
                                sum += v0; sum += v1; sum += v2; sum += v3;
                            }
                            for (; s2 < nsl; ++s2) sum += __ldcg(pbase + (int64_t)rp[2 + s2] * s_to + off);
                            double v = V.coef * sum;
                            if (V.accumulate) v += y[off];
                            y[off] = v;
                        }
                    }
                    rp += 2 + nsl;
                }
                if (tid == 0) a.counters[(int64_t)final_idx * gy + by] = 0;
            }
            t_fin += (a.dbg ? clock64() : 0) - c3;
        }
        // advance: slot n&3 now takes item n+4
        {
            const int u = n & 3, prev = (n + 3) & 3;
            it_i[u] = it_i[prev] + qd; it_b[u] = it_b[prev] + rd; if (it_b[u] >= gy) { it_b[u] -= gy; ++it_i[u]; }
        }
        Lc = Ln;
    }
    if (a.dbg && tid == 0)
    {
        long long * q = a.dbg + (int64_t)blockIdx.x * 8;
        q[0] = t_wait; q[1] = t_issue; q[2] = t_comp; q[3] = t_fin; q[4] = n_iter; q[5] = 1;
    }
}

template <int KF, int KT, int CT>
static cudaError_t launch_pipe_t(const PipeArgs & a, int n_sm, cudaStream_t st)
{
    const size_t smem = (size_t)2 * a.data_doubles * sizeof(double) + (size_t)3 * a.meta_ints * sizeof(int) + 4 * sizeof(int2) + 64 * sizeof(PipeView);
    static int per_sm = -1; static size_t smem_cfg = 0;
    if (per_sm < 0 || smem > smem_cfg)
    {
        cudaError_t e = cudaFuncSetAttribute(sweep_pipe_kernel<KF, KT, CT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
        if (e != cudaSuccess) return e;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sweep_pipe_kernel<KF, KT, CT>, PIPE_THREADS, smem);
        if (e != cudaSuccess) return e;
        if (per_sm < 1) per_sm = 1;
        smem_cfg = smem;
    }
    const int64_t n_work = (int64_t)a.n_item * a.n_job * a.n_comp;
    const int64_t grid = std::min<int64_t>(n_work, (int64_t)n_sm * per_sm);
    sweep_pipe_kernel<KF, KT, CT><<<(unsigned)std::max<int64_t>(grid, 1), PIPE_THREADS, smem, st>>>(a);
    return cudaGetLastError();
}

template <int KF, int KT>
static cudaError_t launch_pipe_ct(const PipeArgs & a, int ct, int n_sm, cudaStream_t st)
{
    if (ct <= 1) return launch_pipe_t<KF, KT, 1>(a, n_sm, st);
    if (ct == 2) return launch_pipe_t<KF, KT, 2>(a, n_sm, st);
    return launch_pipe_t<KF, KT, 4>(a, n_sm, st);
}

#define AMDG_DISPATCH_KT_P(KF_)                                                                   \
    switch (kt) {                                                                                 \
        case 1: return launch_pipe_ct<KF_, 1>(a, ct, n_sm, st); case 2: return launch_pipe_ct<KF_, 2>(a, ct, n_sm, st); \
        case 3: return launch_pipe_ct<KF_, 3>(a, ct, n_sm, st); case 4: return launch_pipe_ct<KF_, 4>(a, ct, n_sm, st); \
        case 5: return launch_pipe_ct<KF_, 5>(a, ct, n_sm, st); case 6: return launch_pipe_ct<KF_, 6>(a, ct, n_sm, st); \
        default: return cudaErrorInvalidValue; }

cudaError_t launch_sweep_pipe(const PipeArgs & a, int kf, int kt, int ct, int n_sm, cudaStream_t st)
{
    switch (kf)
    {
        case 1: AMDG_DISPATCH_KT_P(1) case 2: AMDG_DISPATCH_KT_P(2) case 3: AMDG_DISPATCH_KT_P(3)
        case 4: AMDG_DISPATCH_KT_P(4) case 5: AMDG_DISPATCH_KT_P(5) case 6: AMDG_DISPATCH_KT_P(6)
        default: return cudaErrorInvalidValue;
    }
}

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
    static bool configured = false;
    if (!configured)
    {
        cudaError_t e = cudaFuncSetAttribute(sweep_mma_kernel<KF, KT, INNER1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(MMA_SMEM_DOUBLES * sizeof(double)));
        if (e != cudaSuccess) return e;
        configured = true;
    }
    dim3 grid((unsigned)a.n_item, (unsigned)(a.n_job * a.n_comp));
    sweep_mma_kernel<KF, KT, INNER1><<<grid, MMA_THREADS, (size_t)std::min(smem_doubles, MMA_SMEM_DOUBLES) * sizeof(double), st>>>(a);
    return cudaGetLastError();
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

cudaError_t launch_axpby(int64_t n, double alpha, const double * x, double beta, double * y, cudaStream_t st)
{
    const int64_t nb = (n + 255) / 256;
    axpby_kernel<<<(unsigned)(nb < 148 * 16 ? (nb > 0 ? nb : 1) : 148 * 16), 256, 0, st>>>(n, alpha, x, beta, y);
    return cudaGetLastError();
}

}  // namespace amdg
