// K1, warp-specialised streaming form: FastMultiplyLU::transform_1D (reference source/FastMultiplyLU.cpp:436-512) as a persistent
// producer / consumer pipeline.
//
// Why: ncu of the one-shot kernels (profiles/r01_sweep_tc_ncu.md, profiles/r02_dir_ncu.md) shows DRAM at 2-17 %, 0.1-0.7 eligible
// warps per scheduler and every stall on the long scoreboard: a CTA (or warp) that loads, waits, computes and stores is bounded by
// the chain of memory latencies, not by bandwidth, and 30-45 warp instructions are spent per FP64 MMA on staging and index
// arithmetic.  Here one CTA per SM stays resident: warp 8 is the PRODUCER, it streams the source rows of the next items into a
// ring of four 32 KiB shared-memory stages -- one bulk (TMA) copy per contiguous run, completion counted in bytes on an mbarrier;
// 8-byte async copies where odd block sizes rule bulk copies out -- so that up to 96 KiB per SM are in flight whatever the
// consumers do.  Warps 0-7 are CONSUMERS: they wait for a stage, walk the entry lists of their (fibre, row tile, 64-column group)
// sub-units with B fragments from shared memory (the element's own memory order, offsets from a per-rectangle table), A fragments
// from L1, eight independent accumulator tiles per warp, and store from registers (coef, accumulate, destination map and
// "accumulate-from" as in kernels_dir.cu).  Row tiles whose source list cannot be staged (coarse targets of long fibres) are HEAVY
// items: one row tile x one column tile, the eight warps split the entry list, sources stream from L2 sixteen at a time, partial
// sums are added through shared memory in warp order (deterministic).
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

#include "kernels.cuh"

namespace amdg {

static const int WS_NC = 8;                       // consumer warps
static const int WS_THREADS = (WS_NC + 1) * 32;
static const int WS_STAGES = 4;
static const int WS_STAGE_DOUBLES = 4096;         // 32 KiB
static const int WS_RED_DOUBLES = WS_NC * 64;
static const int WS_HDR_ITEMS = 96;               // item headers of the CTA kept in shared memory (more: read from global memory)
static const int WS_ROWS_INTS = 4096;             // resolved element rows of the CTA's items kept in shared memory

int ws_stage_doubles() { return WS_STAGE_DOUBLES; }

__device__ __forceinline__ void ws_dmma(double (&c)[2], double a, double b)
{
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}
__device__ __forceinline__ unsigned ws_sptr(const void * p) { return (unsigned)__cvta_generic_to_shared(p); }
// shared-memory load by 32-bit shared address (volatile: stays behind the mbarrier wait that publishes the stage)
__device__ __forceinline__ double ws_lds(unsigned addr) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr)); return v; }
// bounded wait: a broken pipeline traps (the launch fails) instead of hanging the device
__device__ __forceinline__ void ws_wait(unsigned bar, unsigned parity)
{
    unsigned done = 0;
    const long long t0 = clock64();
    while (true)
    {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (done) break;
        if (clock64() - t0 > 4000000000ll) asm volatile("trap;");
    }
}

struct WsHdr { int prog, pool_ofs, fib_ofs, nfib, m, n_rt, n_src, n_ent, tab, nct, src_origin, dst_origin, nrun, run_len, gstride, slot, kstride, heavy, vec, rows_ofs; };
__device__ __forceinline__ WsHdr ws_header(const int * it, int lane)
{
    int hv = 0;
    if (lane < 20) hv = it[lane];
    WsHdr h;
#define WS_F(name, i) h.name = __shfl_sync(0xffffffffu, hv, i)
    WS_F(prog, 0); WS_F(pool_ofs, 1); WS_F(fib_ofs, 2); WS_F(nfib, 3); WS_F(m, 4); WS_F(n_rt, 5); WS_F(n_src, 6); WS_F(n_ent, 7); WS_F(tab, 8); WS_F(nct, 9);
    WS_F(src_origin, 10); WS_F(dst_origin, 11); WS_F(nrun, 12); WS_F(run_len, 13); WS_F(gstride, 14); WS_F(slot, 15); WS_F(kstride, 16); WS_F(heavy, 17); WS_F(vec, 18); WS_F(rows_ofs, 19);
#undef WS_F
    return h;
}

__global__ void __launch_bounds__(WS_THREADS, 1) sweep_ws_kernel(const WsArgs a)
{
    extern __shared__ __align__(128) double ws_smem[];
    __shared__ __align__(8) unsigned long long bar_full[WS_STAGES], bar_empty[WS_STAGES];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    asm volatile("griddepcontrol.launch_dependents;");
    if (tid == 0)
    {
        for (int s = 0; s < WS_STAGES; ++s)
        {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(ws_sptr(&bar_full[s])), "r"(32) : "memory");
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(ws_sptr(&bar_empty[s])), "r"(WS_NC) : "memory");
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int jb = blockIdx.y, comp = blockIdx.z;
    const int W = a.job[jb].outer * a.inner;
    const int64_t s_from = (int64_t)W * a.kf, s_to = (int64_t)W * a.kt;
    const double * __restrict__ src = a.job[jb].src + (int64_t)comp * a.n_elem * s_from;
    const int it0 = __ldg(a.cta_ptr + blockIdx.x), it1 = __ldg(a.cta_ptr + blockIdx.x + 1);
    // the CTA's item headers and resolved element rows move to shared memory once (one round trip to L2 for all of them, under the tail of the
    // previous kernel): neither the producer nor the consumers wait for work-list loads per item afterwards
    int * s_hdr = reinterpret_cast<int *>(ws_smem + WS_STAGES * WS_STAGE_DOUBLES + WS_RED_DOUBLES);
    int * s_rows = s_hdr + WS_HDR_ITEMS * 20;
    const int r0 = __ldg(a.rows_ptr + blockIdx.x), r1 = __ldg(a.rows_ptr + blockIdx.x + 1);
    const int n_items = it1 - it0, n_rows = r1 - r0;
    const bool hdr_sm = n_items <= WS_HDR_ITEMS, rows_sm = n_rows <= WS_ROWS_INTS;
    if (hdr_sm) for (int i = tid; i < n_items * 20; i += WS_THREADS) s_hdr[i] = __ldg(reinterpret_cast<const int *>(a.items + it0) + i);
    if (rows_sm) for (int i = tid; i < n_rows; i += WS_THREADS) s_rows[i] = __ldg(a.rows + r0 + i);
    const int * hdrs = hdr_sm ? s_hdr : reinterpret_cast<const int *>(a.items + it0);
    const int * rows = rows_sm ? s_rows : a.rows + r0;
    __syncthreads();
    // the work lists are host-written; from here on the coefficient arrays of earlier kernels are read and written
    asm volatile("griddepcontrol.wait;" ::: "memory");

    if (warp == WS_NC)
    {
        // ---------------------------------------------------------------- producer
        const bool bulk = (a.bulk_jobs >> jb) & 1u;
        int stage = 0; unsigned phase = 0;
        for (int it = 0; it < n_items; ++it)
        {
            const WsHdr h = ws_header(hdrs + it * 20, lane);
            if (h.heavy) continue;
            ws_wait(ws_sptr(&bar_empty[stage]), phase ^ 1u);
            const int * srow = rows + h.rows_ofs;
            const int n_slots = h.nfib * h.n_src;
            double * sbase = ws_smem + stage * WS_STAGE_DOUBLES;
            const unsigned bar = ws_sptr(&bar_full[stage]);
            if (bulk)
            {
                int mine = 0; for (int sl = lane; sl < n_slots; sl += 32) ++mine;
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"((unsigned)(mine * h.slot * 8)) : "memory");
                const unsigned run_bytes = (unsigned)h.run_len * 8u;
                for (int sl = lane; sl < n_slots; sl += 32)
                {
                    const double * g = src + (int64_t)srow[sl] * s_from + h.src_origin;
                    const unsigned d = ws_sptr(sbase + (int64_t)sl * h.slot);
                    for (int r = 0; r < h.nrun; ++r)
                        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                     :: "r"(d + (unsigned)r * run_bytes), "l"(g + (int64_t)r * h.gstride), "r"(run_bytes), "r"(bar) : "memory");
                }
            }
            else
            {
                for (int sl = 0; sl < n_slots; ++sl)
                {
                    const double * g = src + (int64_t)srow[sl] * s_from + h.src_origin;
                    double * d = sbase + (int64_t)sl * h.slot;
                    for (int r = 0; r < h.nrun; ++r)
                        for (int x = lane; x < h.run_len; x += 32)
                            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(ws_sptr(d + r * h.run_len + x)), "l"(g + (int64_t)r * h.gstride + x) : "memory");
                }
                asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" :: "r"(bar) : "memory");
            }
            if (++stage == WS_STAGES) { stage = 0; phase ^= 1u; }
        }
        return;
    }

    // -------------------------------------------------------------------- consumers
    double * __restrict__ dst = a.job[jb].dst + (int64_t)comp * a.n_elem * s_to;
    const long long * __restrict__ dmap = a.job[jb].dst_map;
    const double * __restrict__ accf = a.job[jb].acc_from ? a.job[jb].acc_from + (int64_t)comp * a.n_elem * s_to : nullptr;
    const double coef = a.job[jb].coef;
    const bool accumulate = a.job[jb].accumulate != 0;
    const bool vec_job = (a.vec_jobs >> jb) & 1u;
    const int g_lane = (lane >> 2) >> a.tg_shift;
    const int kk = lane & 3;
    const int dk = min(4 + kk, a.kf - 1) - min(kk, a.kf - 1);        // second k-part (KF > 4), in units of the k stride
    double * red = ws_smem + WS_STAGES * WS_STAGE_DOUBLES;
    int stage = 0; unsigned phase = 0;
    for (int it = 0; it < n_items; ++it)
    {
        const WsHdr h = ws_header(hdrs + it * 20, lane);
        const int * __restrict__ rt_ptr = a.pool + h.pool_ofs;
        const int * __restrict__ rt_id = rt_ptr + h.n_rt + 1;
        const int * __restrict__ ent = rt_id + h.n_rt;
        const double * __restrict__ Ag = a.a_tab[h.prog];
        const bool vec = vec_job && h.vec;
        const int dkp = dk * h.kstride;
        if (!h.heavy)
        {
            const int ngroups = (h.nct + 7) >> 3;
            const int n_sub = h.nfib * h.n_rt * ngroups;
            ws_wait(ws_sptr(&bar_full[stage]), phase);
            const unsigned sbase = ws_sptr(ws_smem) + (unsigned)(stage * WS_STAGE_DOUBLES) * 8u;
            const unsigned slot8 = (unsigned)h.slot * 8u, dkp8 = (unsigned)dkp * 8u;
            for (int su = warp; su < n_sub; su += WS_NC)
            {
                const int g = su % ngroups, rem = su / ngroups;
                const int ri = rem % h.n_rt, b = rem / h.n_rt;
                const int p0 = __ldg(rt_ptr + ri), p1 = __ldg(rt_ptr + ri + 1);
                const int e_t = rows[h.rows_ofs + h.nfib * h.n_src + (b * h.n_rt + ri) * a.tg + g_lane];
                const bool ton = e_t >= 0;
                const int e = ton ? e_t : 0;
                const long long yoff = (dmap ? __ldg(dmap + e) : (long long)e * s_to) + h.dst_origin;
                const int tile0 = h.tab + g * 8;
                const int gv = min(8, h.nct - g * 8);
                int bo[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) bo[j] = __ldg(a.tab_b + (tile0 + j) * 32 + lane);
                // old values of an accumulating sweep: requested now, consumed after the entry loop
                double old[8][2];
                if (accumulate)
                {
                    const double * y = accf ? accf + (long long)e * s_to + h.dst_origin : dst + yoff;
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                    {
                        old[j][0] = 0.0; old[j][1] = 0.0;
                        if (!ton || j >= gv) continue;
                        const int2 co = __ldg(a.tab_c + (tile0 + j) * 32 + lane);
                        if (co.x < 0) continue;
                        if (vec) { const double2 v = *reinterpret_cast<const double2 *>(y + co.x); old[j][0] = v.x; old[j][1] = v.y; }
                        else { old[j][0] = y[co.x]; if (co.y >= 0) old[j][1] = y[co.y]; }
                    }
                }
                double acc[8][2];
#pragma unroll
                for (int j = 0; j < 8; ++j) { acc[j][0] = 0.0; acc[j][1] = 0.0; }
                unsigned ba[8];
                {
                    const unsigned fbase = sbase + (unsigned)(b * h.n_src * h.slot) * 8u;
#pragma unroll
                    for (int j = 0; j < 8; ++j) ba[j] = fbase + (unsigned)bo[j] * 8u;
                }
                if (p0 < p1)
                {
                    // entry p+1 (its source code and operator fragment) is fetched while entry p runs; the fetch past the last entry of a row
                    // tile reads the next tile's entry or the padding the host leaves behind both arrays
                    const int * __restrict__ ep = ent + p0;
                    const double * __restrict__ ap = Ag + (int64_t)p0 * 32 + lane;
                    int code = __ldg(ep); double av = __ldg(ap);
                    if (gv == 8)
                    {
                        for (int p = p0; p < p1; ++p)
                        {
                            ++ep; ap += 32;
                            const int code_n = __ldg(ep); const double av_n = __ldg(ap);
                            const unsigned off = (unsigned)(code >> 1) * slot8 + ((code & 1) ? dkp8 : 0u);
                            double bv[8];
#pragma unroll
                            for (int j = 0; j < 8; ++j) bv[j] = ws_lds(ba[j] + off);
#pragma unroll
                            for (int j = 0; j < 8; ++j) ws_dmma(acc[j], av, bv[j]);
                            code = code_n; av = av_n;
                        }
                    }
                    else
                    {
                        for (int p = p0; p < p1; ++p)
                        {
                            ++ep; ap += 32;
                            const int code_n = __ldg(ep); const double av_n = __ldg(ap);
                            const unsigned off = (unsigned)(code >> 1) * slot8 + ((code & 1) ? dkp8 : 0u);
#pragma unroll
                            for (int j = 0; j < 7; ++j) if (j < gv) ws_dmma(acc[j], av, ws_lds(ba[j] + off));
                            code = code_n; av = av_n;
                        }
                    }
                }
                if (ton)
                {
                    double * y = dst + yoff;
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                    {
                        if (j >= gv) continue;
                        const int2 co = __ldg(a.tab_c + (tile0 + j) * 32 + lane);
                        if (co.x < 0) continue;
                        double v0 = coef * acc[j][0], v1 = coef * acc[j][1];
                        if (accumulate) { v0 += old[j][0]; v1 += old[j][1]; }
                        if (vec) *reinterpret_cast<double2 *>(y + co.x) = make_double2(v0, v1);
                        else { y[co.x] = v0; if (co.y >= 0) y[co.y] = v1; }
                    }
                }
            }
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(ws_sptr(&bar_empty[stage])) : "memory");
            if (++stage == WS_STAGES) { stage = 0; phase ^= 1u; }
        }
        else
        {
            // heavy item: one row tile, one column tile (tables of the whole plane: B offsets are global), the warps split the entries
            const int p0 = __ldg(rt_ptr), p1 = __ldg(rt_ptr + 1);
            const int n = p1 - p0;
            const int chunk = (((n + WS_NC - 1) / WS_NC) + 3) & ~3;
            const int q0 = min(p1, p0 + warp * chunk), q1 = min(p1, q0 + chunk);
            const int bo = __ldg(a.tab_b + h.tab * 32 + lane);
            const int e_t = rows[h.rows_ofs + h.n_ent + g_lane];
            const bool ton = e_t >= 0;
            for (int b = 0; b < h.nfib; ++b)              // (heavy items carry one fibre)
            {
                double acc[4][2];
#pragma unroll
                for (int j = 0; j < 4; ++j) { acc[j][0] = 0.0; acc[j][1] = 0.0; }
                for (int qb = q0; qb < q1; qb += 32)
                {
                    // lane l holds entry qb + l: its code and the element row of its source
                    const int ql = qb + lane;
                    int code = 0, row = 0;
                    if (ql < q1) { code = __ldg(ent + ql); row = rows[h.rows_ofs + ql - p0]; }
                    const int nb = min(32, q1 - qb);
                    for (int i0 = 0; i0 < nb; i0 += 16)
                    {
                        double bv[16], av[16];
#pragma unroll
                        for (int ii = 0; ii < 16; ++ii)
                        {
                            const int i = i0 + ii;
                            const int r_i = __shfl_sync(0xffffffffu, row, i & 31), c_i = __shfl_sync(0xffffffffu, code, i & 31);
                            bv[ii] = 0.0; av[ii] = 0.0;
                            if (i < nb)
                            {
                                bv[ii] = __ldg(src + (int64_t)r_i * s_from + h.src_origin + bo + ((c_i & 1) ? dkp : 0));
                                av[ii] = __ldg(Ag + (int64_t)(qb + i) * 32 + lane);
                            }
                        }
#pragma unroll
                        for (int ii = 0; ii < 16; ++ii) if (i0 + ii < nb) ws_dmma(acc[ii & 3], av[ii], bv[ii]);
                    }
                }
                const double r0 = (acc[0][0] + acc[1][0]) + (acc[2][0] + acc[3][0]);
                const double r1 = (acc[0][1] + acc[1][1]) + (acc[2][1] + acc[3][1]);
                red[(warp * 32 + lane) * 2] = r0; red[(warp * 32 + lane) * 2 + 1] = r1;
                asm volatile("bar.sync 1, %0;" :: "r"(WS_NC * 32) : "memory");
                if (warp == 0 && ton)
                {
                    double v0 = 0.0, v1 = 0.0;
#pragma unroll
                    for (int w = 0; w < WS_NC; ++w) { v0 += red[(w * 32 + lane) * 2]; v1 += red[(w * 32 + lane) * 2 + 1]; }
                    const int e = e_t;
                    const long long yoff = (dmap ? __ldg(dmap + e) : (long long)e * s_to) + h.dst_origin;
                    const int2 co = __ldg(a.tab_c + h.tab * 32 + lane);
                    if (co.x >= 0)
                    {
                        v0 *= coef; v1 *= coef;
                        double * y = dst + yoff;
                        if (accumulate)
                        {
                            const double * yo = accf ? accf + (long long)e * s_to + h.dst_origin : y;
                            v0 += yo[co.x]; if (co.y >= 0) v1 += yo[co.y];
                        }
                        y[co.x] = v0; if (co.y >= 0) y[co.y] = v1;
                    }
                }
                asm volatile("bar.sync 1, %0;" :: "r"(WS_NC * 32) : "memory");
            }
        }
    }
}

cudaError_t launch_sweep_ws(const WsArgs & a, int n_cta, cudaStream_t st)
{
    const size_t smem = (size_t)(WS_STAGES * WS_STAGE_DOUBLES + WS_RED_DOUBLES) * sizeof(double) + (size_t)(WS_HDR_ITEMS * 20 + WS_ROWS_INTS) * sizeof(int);
    // the opt-in is per device: set it on every launch (cheap), not once per process
    cudaError_t e = cudaFuncSetAttribute(sweep_ws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    static const bool pdl = !(std::getenv("AMDG_TC_PDL") && std::atoi(std::getenv("AMDG_TC_PDL")) == 0);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)n_cta, (unsigned)a.n_job, (unsigned)a.n_comp);
    cfg.blockDim = dim3(WS_THREADS, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, sweep_ws_kernel, a);
}

}  // namespace amdg
