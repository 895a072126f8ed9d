// K1 sweep kernel, pipelined persistent list form (kernel variant 3).  Split from kernels.cu so that the translation units build in parallel.
#include <algorithm>
#include "kernels.cuh"
#include "cp_async.cuh"
#include "../../include/amdg.h"

namespace amdg {

static const int PIPE_HDR_INTS = 16;     // == PIPE_HDR of pipe_items.hpp

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

}  // namespace amdg
