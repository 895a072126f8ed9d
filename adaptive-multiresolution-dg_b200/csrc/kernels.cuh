// Device kernels of the fast sparse-grid transform path (FP64, sm_100a).
//
// K1 sweep1d  : FastMultiplyLU::transform_1D (reference source/FastMultiplyLU.cpp:436-512) plus the
//               resize/zero/copy/accumulate passes around it (:666-738) folded into the epilogue.
// K2 pointwise: LagrInterpolation::eval_fp_Lag (source/Interplation.cpp:256-295) and the Vlasov products.
// K4 rk_stage : ExplicitRK::step_stage (source/ODESolver.cpp:209-301).
// Hierarchisation (K3) is a K1 sweep with the stencil operator built by amdg_op_register_hier.
#pragma once
#include <atomic>
#include <cstdint>
#include <cuda_runtime.h>

namespace amdg {

// Function attributes (the dynamic shared-memory opt-in) are per device: every launch wrapper keeps one of these per kernel
// instantiation and sets the attribute the first time it launches on a device (the C ABI makes the context's device current first).
struct PerDeviceOnce
{
    std::atomic<unsigned long long> mask{0};
    static unsigned long long bit() { int d = 0; cudaGetDevice(&d); return 1ull << (d & 63); }
    bool done() const { return (mask.load(std::memory_order_acquire) & bit()) != 0; }
    void mark() { mask.fetch_or(bit(), std::memory_order_release); }
};

struct NbrDev { int local; int pair; };

// one sweep along one dimension, for a batch of `n_job` (src, dst) pairs that share operator, relation,
// L/U part and the fibre tables but may differ in block shape (outer) -- see the shared-prefix schedule.
struct SweepJob
{
    const double * src;
    double * dst;
    int outer;          // prod of block edges of dims < t
    int accumulate;     // dst += result
    double coef;
    // optional (lean and column kernels): offset in doubles, relative to dst, of the destination block of every element row; null = row * S_to.
    // The fibre-partitioned multi-GPU path points these at peer memory: the last sweep before a layout switch stores every element block
    // straight into the buffer of the rank that owns it in the next layout.
    const long long * dst_map = nullptr;
    // optional (lean and column kernels, with accumulate): the old values are read from acc_from (plain row * S_to layout) instead of the
    // destination -- "remote = local partial sum + sweep", the last accumulating sweep before a layout switch
    const double * acc_from = nullptr;
    // optional (column kernel; the other kernels are followed by a row scatter, capi.cu: launch_sweep): a SECOND copy of every output block at
    // dst2 + dst2_map[row] -- a buffer that is consumed locally in this layout AND in the other layout after the next switch is stored to both
    // places from the same epilogue instead of being moved by a separate pass
    double * dst2 = nullptr;
    const long long * dst2_map = nullptr;
};

static const int MAX_JOBS = 32;

struct SweepArgs
{
    // grid tables (dimension t, relation kind)
    const int * slot_elem;      // [n]
    const int * slot_fbase;     // [n] first slot of the slot's fibre
    const int64_t * nbr_ptr;    // [n+1]
    const int * nbr_split;      // [n]
    const NbrDev * nbr;
    const double * blocks;      // [n_pairs][KF][KT]
    int64_t n_elem;
    int inner;                  // prod of block edges of dims > t (same for every job)
    int lu;                     // AMDG_LU_*
    int n_comp;                 // components per job, consecutive n_elem*S apart
    int n_job;
    SweepJob job[MAX_JOBS];
};

// fibre-staged variant: one CTA stages the source entries of a run of whole fibres (or of a column range of one
// long fibre) in shared memory, so every source block is read from HBM once per sweep.
struct FibreItem
{
    int slot0, nslot;   // streamed item: the fibre's slots [slot0, slot0+nslot) are staged
    int col0, ncol;     // column range of the (outer x inner) plane staged by this item
    int lcx;            // log2 of the column-lane count
    int pitch;          // shared-memory row pitch in doubles
    int npair;          // packed item: distinct 1D pairs whose operator blocks are staged (0 = streamed item)
    int pair_ofs;       //   first entry in the `pairs` pool
    int nsrc, src_ofs;  // packed item: rows to stage = slots pool[src_ofs .. +nsrc)
    int ntgt, tgt_ofs;  // target rows: packed = slots pool[tgt_ofs .. +ntgt); streamed = the first ntgt rows of the fibre
    int ent_ofs;        // packed item: first neighbour entry in the `ent` pool; row pointers at rowptr pool[tgt_ofs + item index ..]
    int row_ofs;        //   first entry in the rowptr / rsplit pools (ntgt+1 / ntgt entries)
    int packed;         // 1 = packed item, 0 = streamed item
    int pad1;
};

struct FibreSweepArgs
{
    const int * slot_elem;
    const int * slot_fbase;
    const int64_t * nbr_ptr;
    const int * nbr_split;
    const NbrDev * nbr;
    // pools of the packed items (built per work list)
    const int * pool_slots;     // source / target slot lists
    const int * pool_pairs;     // distinct pair ids
    const int * pool_rowptr;    // per target row: first entry (relative to the item's ent_ofs)
    const int * pool_rsplit;    // per target row: number of leading U entries
    const NbrDev * pool_ent;    // (item-local source row, item-local pair)
    const double * blocks;
    const FibreItem * items;
    int n_item;
    int64_t n_elem;
    int inner;
    int lu;
    int n_comp;
    int n_job;
    int smem_doubles;
    long long * dbg;            // optional: per CTA 8 clock64 stamps (profiling builds)
    SweepJob job[MAX_JOBS];
};

// pipelined persistent sweep kernel (work list built by pipe_items.hpp)
struct PipeArgs
{
    const int * rec;            // record pool
    const int2 * tab;           // per item: (offset, length) into rec
    int n_item;
    const int * fin;            // final tables (long fibres)
    const int * fin_ofs;
    int * counters;             // [n_final][gy] arrival counters (zero between launches)
    double * partial;           // [gy][n_slot][S_to] partial sums of the rows above the cut
    int n_slot;
    const double * blocks;
    int64_t n_elem;
    int inner;
    unsigned inner_magic;       // ceil(2^32 / inner): col / inner == umulhi(col, magic) for col*inner < 2^32
    int n_comp;
    int n_job;
    int data_doubles;           // shared-memory data stage size (doubles, even)
    int meta_ints;              // shared-memory record stage size (ints, multiple of 4)
    long long * dbg;
    SweepJob job[MAX_JOBS];
};

// tensor-core sweep kernel (tile programs built by mma_items.hpp)
struct __align__(16) MmaItem
{
    int prog;           // tile program (shape): index into a_tab
    int elem_ofs, nfib; // element rows of this item's fibres: elem_pool[elem_ofs + b*m + f]
    int o0, no;         // column rectangle: outer indices [o0, o0+no) ...
    int i0, ni;         //   ... x inner indices [i0, i0+ni)
    int pk;             // shared-memory pitch between consecutive source indices k (doubles)
    int m, n_rt;        // fibre length, row tiles of the program piece
    int prog_ofs;       // piece ints in prog_pool: rt_ptr[n_rt+1] | rt_id[n_rt] | ent_src[n_ent]  (row tiles in position order)
    int n_ent;
    unsigned ni_magic;  // ceil(2^32 / ni): c / ni == umulhi(c, magic) for the small column counts used here
    int stage_a;        // 1: the operator values of the piece are staged in shared memory too
    unsigned nfib_magic;    // ceil(2^32 / nfib) and ceil(2^32 / (n_rt * nfib)): unit decode of sweep_tc_kernel
    unsigned unit_magic;
    // sweep_tc_kernel stages only the rows the piece reads: nsrc rows per fibre, element rows at elem_pool[src_ofs + b*nsrc + s]
    // (ent_src of the piece indexes these rows); sweep_mma_kernel stages whole fibres (nsrc == m, src_ofs == elem_ofs)
    int nsrc, src_ofs;
    int ksplit;             // 1: the few coarse targets of a long fibre (they read most of it): nothing is staged (nsrc == 0, ent_src holds
                            //    fibre-local source indices, sources stream from L2), at most 8 columns, and every row tile is walked by all warps
                            //    of the CTA: entries split four ways, partial sums added in warp order through shared memory
    unsigned nrun_magic;    // ceil(2^32 / (no * KF)): run decode of the bulk-copy staging
};
struct MmaArgs
{
    const MmaItem * items; int n_item;
    const int * prog_pool;          // tile programs
    const double * const * a_tab;   // per program: operator values in fragment order [entry][32]
    const int * elem_pool;          // element rows of the fibres of every item
    long long * dbg;                // optional per-CTA clock stamps
    int64_t n_elem;
    int inner;
    int n_comp;
    int n_job;
    SweepJob job[MAX_JOBS];
};

// column-form sweep kernel (kernels_col.cu): one thread per column, no staging
struct __align__(16) ColUnit { int tgt; int ent0; int n_ent; int g; };     // target element row, first entry, entries, column group (32*NC columns)
struct ColArgs
{
    const ColUnit * units; int n_unit;      // heavy units first (one CTA each), then the normal ones (upc per CTA), fibre by fibre
    int n_heavy, upc;
    const int2 * ent;                       // per (dimension, relation): (source element row, canonical 1D pair id), slot by slot, "U" sources first
    const double * blocks;                  // operator blocks [pair][KF][KT]
    int64_t n_elem;
    int inner;
    int n_comp, n_job;
    SweepJob job[MAX_JOBS];
};
cudaError_t launch_sweep_col(const ColArgs & a, int kf, int kt, int nc, cudaStream_t st);                       // kernels_col.cu
int col_max_nc(int kf, int kt);

// point-wise expressions (amdg_pointwise_expr): a small stack program per output, evaluated at every interpolation point
enum { PW_END = 0, PW_VAR = 1, PW_X = 2, PW_OTHER = 3, PW_CONST = 4, PW_ADD = 5, PW_SUB = 6, PW_MUL = 7, PW_DIV = 8, PW_NEG = 9, PW_SIN = 10, PW_COS = 11,
       PW_SQR = 12, PW_EXP = 13, PW_SQRT = 14, PW_ABS = 15, PW_POW = 16, PW_TANH = 17, PW_MIN = 18, PW_MAX = 19 };
static const int PW_MAX_OPS = 192, PW_MAX_CONST = 32, PW_MAX_IO = 8, PW_STACK = 8;
struct PwExprArgs
{
    const double * up[PW_MAX_IO];       // [n_points] per variable
    const double * other[PW_MAX_IO];    // [n_other_elem][block] per field (DGSolution::copy_up_intp_to_f gathers these per element)
    double * out[PW_MAX_IO];            // [n_points] per output
    const int * other_map;              // element row of the field grid for every element of this grid, or null (same rows)
    const double * pts1d;               // [T*edge] interpolation point coordinates (LagrBasis::intep_pt)
    const int * ord1d;                  // [n_elem][dim]
    int64_t n_points;
    int block, edge, dim, n_out;
    int stride[8];                      // edge^(dim-1-t)
    int out_ptr[PW_MAX_IO + 1];         // program of output c: ops [out_ptr[c], out_ptr[c+1])
    short op[PW_MAX_OPS], arg[PW_MAX_OPS];
    double consts[PW_MAX_CONST];
};

struct PointwiseArgs
{
    const double * up;      // [n_points]
    double * fp;            // [n_flux][n_points]
    const double * pts;     // [n_points][dim] or null
    int64_t n_points;
    int n_flux, dim;
    int flux_id[8];
    double params[8][4];
};

cudaError_t launch_sweep_gather(const SweepArgs & a, int kf, int kt, cudaStream_t st);
cudaError_t launch_sweep_fibre(const FibreSweepArgs & a, int kf, int kt, int ct, cudaStream_t st);
int fibre_smem_capacity_doubles();
int fibre_threads();
cudaError_t launch_sweep_pipe(const PipeArgs & a, int kf, int kt, int ct, int n_sm, cudaStream_t st);
int pipe_threads();
int pipe_smem_budget_bytes();
cudaError_t launch_sweep_mma(const MmaArgs & a, int kf, int kt, int smem_doubles, cudaStream_t st);
int mma_smem_capacity_doubles();
cudaError_t launch_sweep_tc(const MmaArgs & a, int kf, int kt, int smem_doubles, cudaStream_t st);     // kernels_tc.cu
int tc_smem_capacity_doubles();
cudaError_t launch_pointwise(const PointwiseArgs & a, cudaStream_t st);
cudaError_t launch_pointwise_expr(const PwExprArgs & a, cudaStream_t st);
// rows of a local array to mapped destinations (peer memory): dst[map[e] + i] = src[e*width + i]
cudaError_t launch_scatter_rows(const double * src, int64_t n_rows, int width, double * dst, const long long * map, cudaStream_t st);
// cross-GPU barrier through peer-mapped flag arrays: flags_of[r] = rank r's flag array [world] (device pointers valid on this device)
struct PeerBarrierArgs { unsigned * flags_of[16]; unsigned * epoch; unsigned * error; int world, rank; long long timeout_cycles; };
cudaError_t launch_peer_barrier(const PeerBarrierArgs & a, cudaStream_t st);
cudaError_t launch_pointwise_herm2d(const PointwiseArgs & a, cudaStream_t st);
cudaError_t launch_rk_stage(int scheme, int stage, double dt, const double * u_tn, double * u, const double * rhs, int64_t n, cudaStream_t st);
cudaError_t launch_rk4_ode2nd_stage(int stage, double dt, const double * u_tn, const double * v_tn, double * u, double * v, const double * rhs,
                                    double * ku, double * kv, int64_t n, cudaStream_t st);
cudaError_t launch_axpby(int64_t n, double alpha, const double * x, double beta, double * y, cudaStream_t st);
struct MomentArgs { const double * f; double * rhs; const int * map; int64_t n_field; int x_block, v_block, n_combo; int offset[16]; double coef[16]; double weight; };
cudaError_t launch_moment(const MomentArgs & a, cudaStream_t st);
cudaError_t launch_rows_gather(const double * src, const int * rows, int64_t n_rows, int width, double * dst, cudaStream_t st);
cudaError_t launch_rows_scatter_add(const double * src, const int * rows, int64_t n_rows, int width, double * dst, cudaStream_t st);
struct IndicatorArgs { const double * u[16]; double * norm; int64_t n_elem; int block, n_var; };
cudaError_t launch_indicator_norm(const IndicatorArgs & a, cudaStream_t st);
struct LincombArgs { const double * x[16]; double c[16]; double * y; double beta; int64_t n; int k; };
cudaError_t launch_lincomb(const LincombArgs & a, cudaStream_t st);
cudaError_t launch_point_coords(const double * pts1d, const int * ord1d, int64_t n_elem, int dim, int edge, double * pts, cudaStream_t st);
bool sweep_shape_supported(int kf, int kt);

}  // namespace amdg
