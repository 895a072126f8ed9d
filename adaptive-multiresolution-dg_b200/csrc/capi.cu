// C ABI (include/amdg.h) over the host tables (grid.hpp) and the sm_100a kernels (kernels.cu).
// No torch types, no CPU compute path: a context without a device can only build and export tables.
#include <chrono>
#include <cstdio>
#include <cstring>
#include <array>
#include <map>
#include <memory>
#include <string>
#include <tuple>
#include <vector>
#include <cuda_runtime.h>

#include "../../include/amdg.h"
#include "grid.hpp"
#include "pipe_items.hpp"
#include "mma_items.hpp"
#include "kernels.cuh"

using namespace amdg;

static thread_local std::string g_err;
static int fail(int code, const std::string & msg) { g_err = msg; return code; }
static thread_local int g_arena_mb_override = -1;   // sub-contexts (coarse views) take a small metadata arena and no L2 access-policy window (they run on the parent's stream)
#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail(AMDG_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); } while (0)

struct Op
{
    int kf = 0, kt = 0;
    bool hier = false;
    std::vector<double> blocks;     // host copy [n_pairs][kf][kt]
    double * d_blocks = nullptr;
};

struct DevDim
{
    int * slot_elem = nullptr; int * slot_fbase = nullptr;
    int64_t * nbr_ptr[2] = { nullptr, nullptr }; int * nbr_split[2] = { nullptr, nullptr }; NbrDev * nbr[2] = { nullptr, nullptr };
    int64_t * fibre_ptr = nullptr;
};

// plan of one fibre shape for the lean tensor-core kernel (get_mma_lean): pieces of its tile program with their source rows and
// column rectangles; depends on the shape and the block shape only, so it is kept across grid changes
struct LeanRect { int o0, no, i0, ni, pk; };
struct LeanPiece { ShapeProg prog; std::vector<int> src; std::vector<LeanRect> rects; bool stage_a = false, ksplit = false, whole = false; long long hash = 0; };

struct amdg_ctx
{
    int dim = 0, nmax = 0, edge_alpt = 0, edge_intp = 0, device = -1;
    int sched = AMDG_SCHED_SHARED, kernel_variant = 0;
    cudaStream_t stream = nullptr; bool own_stream = false;
    Pairs1D pairs;
    Grid grid; bool have_grid = false;
    // adaptive runs (DGAdapt::refine / coarsen every step) change the grid every few dozen sweeps: building per-grid work lists for the fast kernels then
    // costs more than the sweeps they serve.  A grid that replaces one which lived fewer than adaptive_life sweeps puts the context in adaptive mode:
    // sweeps below tc_min_doubles run the gather kernel, which needs nothing beyond the tables amdg_grid_set uploads, until the grid has lived that long.
    int64_t sweeps_on_grid = 0, adaptive_life = 512; bool adaptive_mode = false;
    int adaptive_max_fibre = 128;      // ... and only along dimensions whose longest fibre is short: a gather thread walks its target's whole entry list (measured:
                                       // 12-14 us per sweep on fibres of a few dozen elements, ~200 us on the 512-element fibres of the full 2-D NMAX = 9 grid)
    Grid grid_spare;                                 // the previous grid's tables: amdg_grid_set builds into their storage (no fresh pages)
    std::vector<NbrCache> nbr_caches;                // neighbour lists per fibre shape, kept across grid changes (grid.hpp)
    // *_coarse_grid transforms (amdg_apply_tensor_coarse): the elements with sum of levels <= mesh_nmax as a grid of their own
    struct CoarseView { amdg_ctx * sub = nullptr; int * d_rows = nullptr; int64_t n = 0; std::vector<int> op_map; double * buf[2] = { nullptr, nullptr }; int64_t cap[2] = { 0, 0 }; };
    std::map<int, CoarseView> coarse;
    char * meta_stage = nullptr; size_t meta_stage_cap = 0;   // pinned staging of the grid tables: one copy per amdg_grid_set
    std::vector<DevDim> ddims;
    int * d_ord1d = nullptr;
    std::vector<std::unique_ptr<Op>> ops;
    std::vector<double *> scratch; std::vector<int64_t> scratch_cap;
    double * h2d = nullptr; int64_t h2d_cap = 0;     // device staging for the host-buffer entry points
    double * d2h = nullptr; int64_t d2h_cap = 0;
    int64_t launches = 0;
    // fibre-staged kernel: work lists per (dim, columns W, source edge)
    struct ItemList
    {
        FibreItem * d_items = nullptr; int * d_slots = nullptr; int * d_pairs = nullptr; int * d_rowptr = nullptr; int * d_rsplit = nullptr; NbrDev * d_ent = nullptr;
        int n = 0; int ct = 1; int smem = 0; bool ok = false;
    };
    std::map<std::tuple<int, int, int, int, int, int>, ItemList> items;      // key: (dim t, columns W, kf, kt, relation, parallel class)
    int smem_doubles = 8192, item_target = 148 * 4;
    // pipelined kernel: work lists per (dim t, columns W, kf, kt, relation*4+lu, parallel class)
    struct PipeList
    {
        int * d_rec = nullptr; int2 * d_tab = nullptr; int * d_fin = nullptr; int * d_fin_ofs = nullptr; int * d_counters = nullptr;
        int n_item = 0, n_final = 0, n_slot = 0, ct = 1, data_doubles = 0, meta_ints = 0; bool ok = false;
    };
    std::map<std::tuple<int, int, int, int, int, int>, PipeList> pipes;
    int pipe_cap_doubles = 11000, pipe_meta_ints = 2560, pipe_item_target = 148 * 3;
    // tensor-core kernel: shapes of the fibres, tile programs and work lists
    ShapeTable shapes;
    struct MmaList
    {
        MmaItem * d_items = nullptr; int * d_elem_pool = nullptr; int * d_prog_ints = nullptr;
        int n_item = 0, smem_doubles = 0; bool ok = false;
        std::vector<int> prog_shape;                          // shape id of every program piece of this list
        std::vector<long long> prog_piece;                    // content hash of the piece (entries and pairs): key of its operator fragments
        std::vector<ShapeProg> progs;                         // host copies of the pieces (to build operator values)
        std::map<int, const double **> a_tab;                 // per operator: device table of A pointers
    };
    std::map<std::tuple<int, int, int, int, int, int>, MmaList> mmas;       // key: (dim t * 16 + kf, outer, inner, kt, rel*4+lu, parallel class)
    std::map<std::tuple<int, int, int, long long>, double *> mma_A;         // (op, shape, rel*4+lu, piece hash) -> device operator values
    int mma_cap_doubles = 9 * 1024, mma_item_target = 148 * 8, mma_ent_target = 448, mma_stage_a_max = 64;
    bool tc_force_stage = true; int tc_coarse_ent = 256; int64_t tc_min_doubles = 131072;
    int tc_cap_doubles = 4608, tc_item_target = 148 * 8, tc_ent_target = 48, tc_stage_a_max = 48;      // lean form (kernels_tc.cu)
    bool lean() const { return kernel_variant == 0 || kernel_variant == 5; }
    // column kernel (kernels_col.cu): resolved entry table per (dim t, relation); unit list per (dim t, rel*4+lu, column groups, columns per lane)
    std::map<std::pair<int, int>, int2 *> col_ents;
    struct ColList { ColUnit * d_units = nullptr; int n_unit = 0, n_heavy = 0; };
    std::map<std::tuple<int, int, int, int>, ColList> cols;                    // (dim t, rel*4+lu, column groups, heavy threshold)
    int col_heavy_ent = 24, col_upc = 8, col_force_nc = 0;
    int64_t col_min_block = 64; int col_max_kk = 9;                            // auto mode: column kernel for blocks >= col_min_block doubles with KF*KT <= col_max_kk (measured: profiles/r02_sweep_kernels.md)
    double * d_pts1d = nullptr;                                   // LagrBasis::intep_pt table [T*edge_intp] (amdg_points_set)
    std::map<std::tuple<int, int, int, int, int, int>, std::vector<LeanPiece>> lean_plans;   // (shape, kf, kt, rel*4+lu, outer, inner)
    int n_sm = 148;
    long long * dbg = nullptr;
    // metadata arena: index tables, work lists and operator blocks live in one allocation that is given an L2
    // persisting access-policy window (they are re-read by every sweep while coefficient data streams through L2)
    char * arena = nullptr; size_t arena_cap = 0, arena_lo = 0, arena_hi = 0;   // grid tables grow up from 0, operators down from cap
    bool l2_window = false;
};

static int need_device(amdg_ctx * c)
{
    if (!c) return fail(AMDG_EINVAL, "null context");
    if (c->device < 0) return fail(AMDG_ENODEVICE, "context was created without a device: no CPU compute path exists");
    return AMDG_OK;
}

static void meta_free(amdg_ctx * c, void * p);
static void free_coarse_views(amdg_ctx * c)
{
    for (auto & kv : c->coarse)
    {
        if (kv.second.sub) amdg_ctx_destroy(kv.second.sub);
        cudaFree(kv.second.d_rows); cudaFree(kv.second.buf[0]); cudaFree(kv.second.buf[1]);
    }
    c->coarse.clear();
}
static void free_dev_grid(amdg_ctx * c)
{
    free_coarse_views(c);
    for (auto & D : c->ddims)
    {
        meta_free(c, D.slot_elem); meta_free(c, D.slot_fbase); meta_free(c, D.fibre_ptr);
        for (int k = 0; k < 2; ++k) { meta_free(c, D.nbr_ptr[k]); meta_free(c, D.nbr_split[k]); meta_free(c, D.nbr[k]); }
    }
    c->ddims.clear();
    meta_free(c, c->d_ord1d); c->d_ord1d = nullptr;
    for (auto & kv : c->mmas)
    {
        amdg_ctx::MmaList & L = kv.second;
        meta_free(c, L.d_items); meta_free(c, L.d_elem_pool); meta_free(c, L.d_prog_ints);
        for (auto & at : L.a_tab) meta_free(c, (void *)at.second);
    }
    c->mmas.clear();                     // (operator fragments, mma_A, are per shape and survive a grid change)
    for (auto & kv : c->col_ents) meta_free(c, kv.second);
    c->col_ents.clear();
    for (auto & kv : c->cols) meta_free(c, kv.second.d_units);
    c->cols.clear();
    for (auto & kv : c->pipes)
    {
        amdg_ctx::PipeList & L = kv.second;
        meta_free(c, L.d_rec); meta_free(c, L.d_tab); meta_free(c, L.d_fin); meta_free(c, L.d_fin_ofs); meta_free(c, L.d_counters);
    }
    c->pipes.clear();
    for (auto & kv : c->items)
    {
        amdg_ctx::ItemList & L = kv.second;
        meta_free(c, L.d_items); meta_free(c, L.d_slots); meta_free(c, L.d_pairs); meta_free(c, L.d_rowptr); meta_free(c, L.d_rsplit); meta_free(c, L.d_ent);
    }
    c->items.clear();
    c->arena_lo = 0;
}

template <class T>
static cudaError_t upload(T ** dptr, const T * h, size_t n, cudaStream_t st)
{
    cudaError_t e = cudaMalloc((void **)dptr, std::max<size_t>(n, 1) * sizeof(T));
    if (e != cudaSuccess) return e;
    if (n) e = cudaMemcpyAsync(*dptr, h, n * sizeof(T), cudaMemcpyHostToDevice, st);
    return e;
}


// ---- metadata arena ------------------------------------------------------------------------------------------
static bool in_arena(amdg_ctx * c, const void * p) { return c->arena && (const char *)p >= c->arena && (const char *)p < c->arena + c->arena_cap; }
static void meta_free(amdg_ctx * c, void * p) { if (p && !in_arena(c, p)) cudaFree(p); }
template <class T>
static cudaError_t meta_upload(amdg_ctx * c, T ** dptr, const T * h, size_t n, bool is_op)
{
    const size_t bytes = ((std::max<size_t>(n, 1) * sizeof(T)) + 255) & ~(size_t)255;
    if (c->arena && c->arena_lo + bytes + (c->arena_cap - c->arena_hi) <= c->arena_cap)
    {
        if (is_op) { c->arena_hi -= bytes; *dptr = (T *)(c->arena + c->arena_hi); }
        else { *dptr = (T *)(c->arena + c->arena_lo); c->arena_lo += bytes; }
        return n ? cudaMemcpyAsync(*dptr, h, n * sizeof(T), cudaMemcpyHostToDevice, c->stream) : cudaSuccess;
    }
    return upload(dptr, h, n, c->stream);
}
static void apply_l2_window(amdg_ctx * c)
{
    if (!c->arena || !c->l2_window) return;
    cudaStreamAttrValue attr; std::memset(&attr, 0, sizeof(attr));
    attr.accessPolicyWindow.base_ptr = c->arena;
    attr.accessPolicyWindow.num_bytes = c->arena_cap;
    attr.accessPolicyWindow.hitRatio = 1.0f;
    attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    if (cudaStreamSetAttribute(c->stream, cudaStreamAttributeAccessPolicyWindow, &attr) != cudaSuccess) cudaGetLastError();
}

static int ensure_scratch(amdg_ctx * c, size_t idx, int64_t n)
{
    if (c->scratch.size() <= idx) { c->scratch.resize(idx + 1, nullptr); c->scratch_cap.resize(idx + 1, 0); }
    if (c->scratch_cap[idx] >= n) return AMDG_OK;
    if (c->scratch[idx]) { CU(cudaStreamSynchronize(c->stream)); CU(cudaFree(c->scratch[idx])); c->scratch[idx] = nullptr; c->scratch_cap[idx] = 0; }
    // small buffers grow geometrically: an adaptive run whose grid grows a little every step does not reallocate (and synchronise) every step
    const int64_t want = n < ((int64_t)1 << 24) ? std::max<int64_t>(n + n / 2, 4096) : n;
    cudaError_t e = cudaMalloc((void **)&c->scratch[idx], (size_t)want * sizeof(double));
    if (e != cudaSuccess) return fail(AMDG_ENOMEM, std::string("scratch cudaMalloc: ") + cudaGetErrorString(e));
    c->scratch_cap[idx] = want;
    return AMDG_OK;
}

// Per-shape caches (plans on the host, operator fragments on the device) survive grid changes on purpose; over a long adaptive run the
// set of shapes ever seen keeps growing, so once the known shapes exceed twice the live ones (and a floor, AMDG_CACHE_SHAPES, default 4096)
// everything belonging to shapes the new grid does not have is dropped.  Called from amdg_grid_set after the shape table was rebuilt.
static void evict_shape_caches(amdg_ctx * c)
{
    const size_t floor_shapes = std::getenv("AMDG_CACHE_SHAPES") ? (size_t)std::max(0, atoi(std::getenv("AMDG_CACHE_SHAPES"))) : 4096;
    const std::vector<char> live = c->shapes.live_flags();
    size_t n_live = 0; for (char f : live) n_live += f;
    if (c->shapes.n_known() <= std::max(floor_shapes, 2 * n_live)) return;
    const size_t before = c->shapes.n_known(), frags_before = c->mma_A.size();
    if (c->device >= 0 && !c->mma_A.empty()) { cudaSetDevice(c->device); cudaStreamSynchronize(c->stream); }
    for (auto it = c->mma_A.begin(); it != c->mma_A.end();)
    {
        if (!live[std::get<1>(it->first)]) { cudaFree(it->second); it = c->mma_A.erase(it); } else ++it;
    }
    auto prune = [&](auto & m) { for (auto it = m.begin(); it != m.end();) { if (!live[std::get<0>(it->first)]) it = m.erase(it); else ++it; } };
    prune(c->lean_plans);
    c->shapes.retire(live);
    if (std::getenv("AMDG_VERBOSE"))
        fprintf(stderr, "[amdg] cache eviction: shapes %zu -> %zu, operator fragments %zu -> %zu\n", before, c->shapes.n_known(), frags_before, c->mma_A.size());
}

extern "C" {

const char * amdg_version(void) { return "amdg-b200 0.1 (sm_100a, fp64)"; }
const char * amdg_last_error(void) { return g_err.c_str(); }

int amdg_ctx_create(int dim, int nmax, int pmax_alpt, int pmax_intp, int device, amdg_ctx ** out)
{
    if (!out) return fail(AMDG_EINVAL, "out is null");
    if (dim < 1 || dim > 8) return fail(AMDG_EINVAL, "dim must be in 1..8");
    if (nmax < 0 || nmax > 12) return fail(AMDG_EINVAL, "nmax must be in 0..12");
    if (pmax_alpt < 0 || pmax_alpt > 5 || pmax_intp < 0 || pmax_intp > 5) return fail(AMDG_EINVAL, "pmax must be in 0..5");
    std::unique_ptr<amdg_ctx> c(new amdg_ctx());
    c->dim = dim; c->nmax = nmax; c->edge_alpt = pmax_alpt + 1; c->edge_intp = pmax_intp + 1; c->device = device;
    c->pairs.build(nmax);
    if (const char * e = std::getenv("AMDG_SMEM_DOUBLES")) c->smem_doubles = std::max(256, std::min(atoi(e), fibre_smem_capacity_doubles()));
    if (const char * e = std::getenv("AMDG_ITEM_TARGET")) c->item_target = std::max(1, atoi(e));
    if (const char * e = std::getenv("AMDG_KERNEL")) c->kernel_variant = atoi(e);
    if (const char * e = std::getenv("AMDG_COL_HEAVY")) c->col_heavy_ent = std::max(1, std::min(32, atoi(e)));
    if (const char * e = std::getenv("AMDG_COL_UPC")) c->col_upc = std::max(1, atoi(e));
    if (const char * e = std::getenv("AMDG_COL_MIN_BLOCK")) c->col_min_block = std::max(1, atoi(e));
    if (const char * e = std::getenv("AMDG_COL_MAX_KK")) c->col_max_kk = std::max(1, atoi(e));
    if (const char * e = std::getenv("AMDG_COL_NC")) c->col_force_nc = std::max(0, std::min(4, atoi(e)));
    if (const char * e = std::getenv("AMDG_MMA_CAP")) c->mma_cap_doubles = std::max(512, std::min(atoi(e), mma_smem_capacity_doubles())) & ~1;
    if (const char * e = std::getenv("AMDG_MMA_ITEMS")) c->mma_item_target = std::max(1, atoi(e));
    if (const char * e = std::getenv("AMDG_MMA_ENT")) c->mma_ent_target = std::max(16, atoi(e));
    if (const char * e = std::getenv("AMDG_MMA_STAGE_A")) c->mma_stage_a_max = std::max(0, atoi(e));
    if (const char * e = std::getenv("AMDG_TC_CAP")) c->tc_cap_doubles = std::max(512, std::min(atoi(e), tc_smem_capacity_doubles())) & ~1;
    if (const char * e = std::getenv("AMDG_TC_ITEMS")) c->tc_item_target = std::max(1, atoi(e));
    if (const char * e = std::getenv("AMDG_TC_ENT")) c->tc_ent_target = std::max(16, atoi(e));
    if (const char * e = std::getenv("AMDG_TC_STAGE_A")) c->tc_stage_a_max = std::max(0, atoi(e));
    if (const char * e = std::getenv("AMDG_TC_MIN_DOUBLES")) c->tc_min_doubles = atoll(e);
    if (const char * e = std::getenv("AMDG_ADAPTIVE_LIFE")) c->adaptive_life = atoll(e);
    if (const char * e = std::getenv("AMDG_ADAPTIVE_FIBRE")) c->adaptive_max_fibre = std::max(1, atoi(e));
    if (const char * e = std::getenv("AMDG_TC_FORCE_STAGE")) c->tc_force_stage = atoi(e) != 0;
    if (const char * e = std::getenv("AMDG_TC_COARSE_ENT")) c->tc_coarse_ent = std::max(16, atoi(e));
    if (const char * e = std::getenv("AMDG_PIPE_CAP")) c->pipe_cap_doubles = std::max(256, atoi(e)) & ~1;
    if (const char * e = std::getenv("AMDG_PIPE_META")) c->pipe_meta_ints = (std::max(256, atoi(e)) + 3) & ~3;
    if (const char * e = std::getenv("AMDG_PIPE_ITEMS")) c->pipe_item_target = std::max(1, atoi(e));
    if (device >= 0)
    {
        int count = 0;
        cudaError_t e = cudaGetDeviceCount(&count);
        if (e != cudaSuccess || count == 0) return fail(AMDG_ENODEVICE, std::string("no CUDA device: ") + cudaGetErrorString(e));
        if (device >= count) return fail(AMDG_ENODEVICE, "device ordinal out of range");
        CU(cudaSetDevice(device));
        CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        c->own_stream = true;
        size_t arena_mb = 96; if (const char * e2 = std::getenv("AMDG_ARENA_MB")) arena_mb = (size_t)std::max(0, atoi(e2));
        if (g_arena_mb_override >= 0) arena_mb = (size_t)g_arena_mb_override;
        cudaDeviceProp prop;
        { int v = 0; if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device) == cudaSuccess && v > 0) c->n_sm = v; }
        if (arena_mb > 0 && cudaGetDeviceProperties(&prop, device) == cudaSuccess)
        {
            size_t cap = arena_mb << 20;
            if (prop.accessPolicyMaxWindowSize > 0) cap = std::min(cap, (size_t)prop.accessPolicyMaxWindowSize);
            if (cudaMalloc((void **)&c->arena, cap) == cudaSuccess)
            {
                c->arena_cap = cap; c->arena_lo = 0; c->arena_hi = cap;
                const char * w = std::getenv("AMDG_L2_PERSIST");
                if (!(w && w[0] == '0') && prop.persistingL2CacheMaxSize > 0 && g_arena_mb_override < 0)
                {
                    cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, std::min((size_t)prop.persistingL2CacheMaxSize, (size_t)(32u << 20)));
                    c->l2_window = true; apply_l2_window(c.get());
                }
            }
            else { cudaGetLastError(); c->arena = nullptr; }
        }
    }
    *out = c.release();
    return AMDG_OK;
}

int amdg_ctx_destroy(amdg_ctx * c)
{
    if (!c) return AMDG_OK;
    if (c->device >= 0)
    {
        cudaSetDevice(c->device);
        cudaDeviceSynchronize();
        free_dev_grid(c);
        for (auto & kv : c->mma_A) cudaFree(kv.second);
        c->mma_A.clear();
        for (auto & op : c->ops) meta_free(c, op->d_blocks);
        cudaFree(c->arena);
        for (double * p : c->scratch) cudaFree(p);
        cudaFree(c->h2d); cudaFree(c->d2h); cudaFree(c->d_pts1d);
        if (c->meta_stage) cudaFreeHost(c->meta_stage);
        if (c->own_stream) cudaStreamDestroy(c->stream);
    }
    delete c;
    return AMDG_OK;
}

int amdg_ctx_set_stream(amdg_ctx * c, void * s)
{
    int r = need_device(c); if (r) return r;
    if (c->own_stream) { cudaStreamSynchronize(c->stream); cudaStreamDestroy(c->stream); c->own_stream = false; }
    c->stream = (cudaStream_t)s;
    apply_l2_window(c);
    return AMDG_OK;
}

int amdg_ctx_sync(amdg_ctx * c) { int r = need_device(c); if (r) return r; CU(cudaStreamSynchronize(c->stream)); return AMDG_OK; }
int amdg_ctx_set_schedule(amdg_ctx * c, int s) { if (!c || (s != AMDG_SCHED_LITERAL && s != AMDG_SCHED_SHARED)) return fail(AMDG_EINVAL, "bad schedule"); c->sched = s; return AMDG_OK; }
int amdg_ctx_set_kernel(amdg_ctx * c, int v) { if (!c || v < 0 || v > 8 || v == 6 || v == 7) return fail(AMDG_EINVAL, "bad kernel variant"); c->kernel_variant = v; return AMDG_OK; }
int64_t amdg_ctx_launch_count(amdg_ctx * c) { return c ? c->launches : -1; }
int amdg_ctx_set_debug_buffer(amdg_ctx * c, void * dev_buf) { if (!c) return fail(AMDG_EINVAL, "null context"); c->dbg = (long long *)dev_buf; return AMDG_OK; }

int amdg_hash_key(int dim, const int * level, const int * suppt) { return hash_key(dim, level, suppt); }
int amdg_order_elem(int level, int suppt) { return order_elem(level, suppt); }

int64_t amdg_sparse_grid(int dim, int level_init, int sparse, int * level, int * suppt)
{
    if (dim < 1 || dim > 8 || level_init < 0) return fail(AMDG_EINVAL, "bad grid parameters");
    if (!level) return sparse_grid(dim, level_init, sparse != 0, nullptr, nullptr);
    std::vector<int> l, j;
    const int64_t n = sparse_grid(dim, level_init, sparse != 0, &l, &j);
    std::memcpy(level, l.data(), l.size() * sizeof(int)); std::memcpy(suppt, j.data(), j.size() * sizeof(int));
    return n;
}

// ---- grid ------------------------------------------------------------------------------------------------------
int amdg_grid_set(amdg_ctx * c, int64_t n, const int * level, const int * suppt)
{
    if (!c || n < 1 || !level || !suppt) return fail(AMDG_EINVAL, "bad arguments to amdg_grid_set");
    if (n > 0x7fffffff) return fail(AMDG_EINVAL, "too many elements");
    const auto t0 = std::chrono::steady_clock::now();
    // level / suppt may alias the current grid's own arrays: the spare is built first, then swapped in
    if (c->grid_spare.build(c->dim, c->nmax, n, level, suppt, c->pairs, &c->nbr_caches) != 0) return fail(AMDG_EINVAL, "invalid or duplicate element index");
    const auto t1 = std::chrono::steady_clock::now();
    if (c->have_grid) c->adaptive_mode = c->sweeps_on_grid < c->adaptive_life;
    c->sweeps_on_grid = 0;
    std::swap(c->grid, c->grid_spare); c->have_grid = true;
    c->shapes.build(c->grid);
    evict_shape_caches(c);
    if (std::getenv("AMDG_VERBOSE"))
        fprintf(stderr, "[amdg] grid_set: %lld elements, tables %.2f ms, shapes %.2f ms (%d shapes known)\n", (long long)n,
                std::chrono::duration<double, std::milli>(t1 - t0).count(),
                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t1).count(), (int)c->shapes.ords.size());
    if (c->device < 0) return AMDG_OK;
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    free_dev_grid(c);
    // the device tables are committed as a whole: a failed upload leaves the context without a grid (host tables included), never with partial tables
    // fast path: every table into one pinned staging buffer at the offsets it will have in the arena, one copy.  The staging buffer is
    // reused by the next amdg_grid_set only, which synchronises the stream first.
    auto up_packed = [&]() -> bool
    {
        if (!c->arena || std::getenv("AMDG_GRID_UPLOAD_SPLIT")) return false;
        struct Part { const void * src; size_t bytes; void ** dst; };
        std::vector<Part> parts;
        std::vector<std::vector<int>> fbase(c->dim);
        c->ddims.assign(c->dim, DevDim());
        parts.push_back({ c->grid.ord1d.data(), c->grid.ord1d.size() * sizeof(int), (void **)&c->d_ord1d });
        for (int t = 0; t < c->dim; ++t)
        {
            const DimTables & H = c->grid.dims[t]; DevDim & D = c->ddims[t];
            fbase[t].resize(n);
            for (int64_t s = 0; s < n; ++s) fbase[t][s] = (int)H.fibre_ptr[H.slot_fibre[s]];
            parts.push_back({ H.slot_elem.data(), (size_t)n * sizeof(int), (void **)&D.slot_elem });
            parts.push_back({ fbase[t].data(), (size_t)n * sizeof(int), (void **)&D.slot_fbase });
            parts.push_back({ H.fibre_ptr.data(), H.fibre_ptr.size() * sizeof(int64_t), (void **)&D.fibre_ptr });
            for (int k = 0; k < 2; ++k)
            {
                parts.push_back({ H.nbr_ptr[k].data(), H.nbr_ptr[k].size() * sizeof(int64_t), (void **)&D.nbr_ptr[k] });
                parts.push_back({ H.nbr_split[k].data(), H.nbr_split[k].size() * sizeof(int), (void **)&D.nbr_split[k] });
                parts.push_back({ H.nbr[k].data(), H.nbr[k].size() * sizeof(Nbr), (void **)&D.nbr[k] });
            }
        }
        size_t total = 0;
        for (const Part & p : parts) total += (std::max<size_t>(p.bytes, 1) + 255) & ~(size_t)255;
        if (c->arena_lo + total + (c->arena_cap - c->arena_hi) > c->arena_cap) { c->ddims.clear(); c->d_ord1d = nullptr; return false; }
        if (c->meta_stage_cap < total)
        {
            if (c->meta_stage) cudaFreeHost(c->meta_stage);
            c->meta_stage = nullptr; c->meta_stage_cap = 0;
            // pinned allocations cost milliseconds: grow geometrically from 1 MiB so that an adaptive run reallocates a handful of times at most
            size_t cap = (size_t)1 << 20; while (cap < total) cap <<= 1;
            if (cudaMallocHost((void **)&c->meta_stage, cap) != cudaSuccess) { cudaGetLastError(); c->ddims.clear(); c->d_ord1d = nullptr; return false; }
            c->meta_stage_cap = cap;
        }
        size_t ofs = 0;
        char * base = c->arena + c->arena_lo;
        for (const Part & p : parts)
        {
            if (p.bytes) std::memcpy(c->meta_stage + ofs, p.src, p.bytes);
            *p.dst = base + ofs;
            ofs += (std::max<size_t>(p.bytes, 1) + 255) & ~(size_t)255;
        }
        if (cudaMemcpyAsync(base, c->meta_stage, total, cudaMemcpyHostToDevice, c->stream) != cudaSuccess) { cudaGetLastError(); c->ddims.clear(); c->d_ord1d = nullptr; return false; }
        c->arena_lo += total;
        return true;
    };
    auto up_all = [&]() -> cudaError_t
    {
        if (up_packed()) return cudaSuccess;
        cudaError_t e;
        c->ddims.resize(c->dim);
        if ((e = meta_upload(c, &c->d_ord1d, c->grid.ord1d.data(), c->grid.ord1d.size(), false)) != cudaSuccess) return e;
        for (int t = 0; t < c->dim; ++t)
        {
            const DimTables & H = c->grid.dims[t]; DevDim & D = c->ddims[t];
            std::vector<int> fbase(n);
            for (int64_t s = 0; s < n; ++s) fbase[s] = (int)H.fibre_ptr[H.slot_fibre[s]];
            if ((e = meta_upload(c, &D.slot_elem, H.slot_elem.data(), (size_t)n, false)) != cudaSuccess) return e;
            if ((e = meta_upload(c, &D.slot_fbase, fbase.data(), (size_t)n, false)) != cudaSuccess) return e;
            if ((e = meta_upload(c, &D.fibre_ptr, H.fibre_ptr.data(), H.fibre_ptr.size(), false)) != cudaSuccess) return e;
            for (int k = 0; k < 2; ++k)
            {
                if ((e = meta_upload(c, &D.nbr_ptr[k], H.nbr_ptr[k].data(), H.nbr_ptr[k].size(), false)) != cudaSuccess) return e;
                if ((e = meta_upload(c, &D.nbr_split[k], H.nbr_split[k].data(), H.nbr_split[k].size(), false)) != cudaSuccess) return e;
                if ((e = meta_upload(c, (Nbr **)&D.nbr[k], H.nbr[k].data(), H.nbr[k].size(), false)) != cudaSuccess) return e;
            }
            if ((e = cudaStreamSynchronize(c->stream)) != cudaSuccess) return e;   // fbase is a local
        }
        return cudaSuccess;
    };
    const cudaError_t ue = up_all();
    if (ue != cudaSuccess)
    {
        free_dev_grid(c); c->have_grid = false; cudaGetLastError();
        return fail(AMDG_ECUDA, std::string("amdg_grid_set: table upload failed: ") + cudaGetErrorString(ue));
    }
    return AMDG_OK;
}

int64_t amdg_grid_size(amdg_ctx * c) { return (c && c->have_grid) ? c->grid.n : -1; }

int amdg_grid_keys(amdg_ctx * c, int * hash, int * ord1d)
{
    if (!c || !c->have_grid) return fail(AMDG_ESTATE, "no grid");
    if (hash) { const std::vector<int> & k = c->grid.keys(); std::memcpy(hash, k.data(), k.size() * sizeof(int)); }
    if (ord1d) std::memcpy(ord1d, c->grid.ord1d.data(), c->grid.ord1d.size() * sizeof(int));
    return AMDG_OK;
}

int64_t amdg_grid_relation(amdg_ctx * c, int t, int rel, int64_t * ptr, int * idx)
{
    if (!c || !c->have_grid) return fail(AMDG_ESTATE, "no grid");
    if (t < 0 || t >= c->dim || rel < 0 || rel > 1) return fail(AMDG_EINVAL, "bad dim/relation");
    const DimTables & H = c->grid.dims[t];
    const int64_t nnz = (int64_t)H.nbr[rel].size();
    if (!idx) return nnz;
    const int64_t n = c->grid.n;
    ptr[0] = 0;
    for (int64_t e = 0; e < n; ++e) { const int64_t s = H.elem_slot[e]; ptr[e + 1] = ptr[e] + (H.nbr_ptr[rel][s + 1] - H.nbr_ptr[rel][s]); }
    for (int64_t e = 0; e < n; ++e)
    {
        const int64_t s = H.elem_slot[e]; const int64_t fb = H.fibre_ptr[H.slot_fibre[s]];
        int * out = idx + ptr[e]; int64_t k = 0;
        for (int64_t p = H.nbr_ptr[rel][s]; p < H.nbr_ptr[rel][s + 1]; ++p) out[k++] = H.slot_elem[fb + H.nbr[rel][p].local];
        std::sort(out, out + k);
    }
    return nnz;
}

int64_t amdg_grid_fibres(amdg_ctx * c, int t, int64_t * ptr, int * elems)
{
    if (!c || !c->have_grid) return fail(AMDG_ESTATE, "no grid");
    if (t < 0 || t >= c->dim) return fail(AMDG_EINVAL, "bad dim");
    const DimTables & H = c->grid.dims[t];
    if (!elems) return H.n_fibre;
    std::memcpy(ptr, H.fibre_ptr.data(), H.fibre_ptr.size() * sizeof(int64_t));
    std::memcpy(elems, H.slot_elem.data(), H.slot_elem.size() * sizeof(int));
    return H.n_fibre;
}

// ---- operators ---------------------------------------------------------------------------------------------------
static int push_op(amdg_ctx * c, std::unique_ptr<Op> op, int * out)
{
    if (c->device >= 0)
    {
        CU(cudaSetDevice(c->device));
        CU(meta_upload(c, &op->d_blocks, op->blocks.data(), op->blocks.size(), true));
        CU(cudaStreamSynchronize(c->stream));
    }
    c->ops.push_back(std::move(op));
    *out = (int)c->ops.size() - 1;
    return AMDG_OK;
}

int amdg_op_register(amdg_ctx * c, const double * dense, int rows, int cols, int kf, int kt, int * out)
{
    if (!c || !dense || !out) return fail(AMDG_EINVAL, "null argument");
    const int T = c->pairs.T;
    if (rows != T * kf || cols != T * kt) return fail(AMDG_EINVAL, "operator shape must be (2^nmax*edge_from) x (2^nmax*edge_to)");
    if (!sweep_shape_supported(kf, kt)) return fail(AMDG_EINVAL, "unsupported block edge");
    std::unique_ptr<Op> op(new Op()); op->kf = kf; op->kt = kt;
    op->blocks.resize((size_t)c->pairs.n_pairs * kf * kt);
    for (int p = 0; p < c->pairs.n_pairs; ++p)
    {
        const int f = c->pairs.src[p], e = c->pairs.tgt[p];
        for (int k = 0; k < kf; ++k) for (int q = 0; q < kt; ++q)
            op->blocks[((size_t)p * kf + k) * kt + q] = dense[(size_t)(f * kf + k) * cols + (e * kt + q)];
    }
    return push_op(c, std::move(op), out);
}

int64_t amdg_pairs(amdg_ctx * c, int * src, int * tgt, int * is_vol)
{
    if (!c) return fail(AMDG_EINVAL, "null context");
    if (!src) return c->pairs.n_pairs;
    for (int p = 0; p < c->pairs.n_pairs; ++p) { src[p] = c->pairs.src[p]; tgt[p] = c->pairs.tgt[p]; if (is_vol) is_vol[p] = c->pairs.vol[p]; }
    return c->pairs.n_pairs;
}

int amdg_op_register_compact(amdg_ctx * c, const double * blocks, int64_t n_pairs, int kf, int kt, int hier, int * out)
{
    if (!c || !blocks || !out) return fail(AMDG_EINVAL, "null argument");
    if (n_pairs != c->pairs.n_pairs) return fail(AMDG_EINVAL, "n_pairs does not match the canonical pair enumeration of this nmax");
    if (!sweep_shape_supported(kf, kt)) return fail(AMDG_EINVAL, "unsupported block edge");
    std::unique_ptr<Op> op(new Op()); op->kf = kf; op->kt = kt; op->hier = hier != 0;
    op->blocks.assign(blocks, blocks + (size_t)n_pairs * kf * kt);
    return push_op(c, std::move(op), out);
}

int amdg_op_blocks(amdg_ctx * c, int op, double * blocks)
{
    if (!c || !blocks || op < 0 || op >= (int)c->ops.size()) return fail(AMDG_EINVAL, "bad operator handle");
    std::memcpy(blocks, c->ops[op]->blocks.data(), c->ops[op]->blocks.size() * sizeof(double));
    return AMDG_OK;
}

int amdg_op_register_hier(amdg_ctx * c, const int * anc, const double * wt, int p1, int * out)
{
    if (!c || !anc || !wt || !out) return fail(AMDG_EINVAL, "null argument");
    if (!sweep_shape_supported(p1, p1)) return fail(AMDG_EINVAL, "unsupported block edge");
    const int T = c->pairs.T;
    std::unique_ptr<Op> op(new Op()); op->kf = p1; op->kt = p1; op->hier = true;
    op->blocks.assign((size_t)c->pairs.n_pairs * p1 * p1, 0.0);
    for (int e = 0; e < T; ++e)
    {
        const int self = c->pairs.id[(size_t)e * T + e];
        for (int p = 0; p < p1; ++p) op->blocks[((size_t)self * p1 + p) * p1 + p] = 1.0;   // c_{t+1} = c_t + ...
        if (e == 0) continue;
        for (int ic = 0; ic < p1; ++ic)
        {
            const int f = anc[((size_t)(e - 1) * p1 + ic) * 2], qf = anc[((size_t)(e - 1) * p1 + ic) * 2 + 1];
            if (f < 0 || f >= T || qf < 0 || qf >= p1) return fail(AMDG_EINVAL, "hierarchisation stencil out of range");
            const int pair = c->pairs.id[(size_t)f * T + e];
            if (pair < 0 || !c->pairs.vol[pair] || level_of_order(f) >= level_of_order(e)) return fail(AMDG_EINVAL, "hierarchisation ancestor is not a coarser overlapping element");
            for (int p0 = 0; p0 < p1; ++p0) op->blocks[((size_t)pair * p1 + qf) * p1 + p0] += wt[((size_t)(e - 1) * p1 + p0) * p1 + ic];
        }
    }
    return push_op(c, std::move(op), out);
}

int amdg_op_combine(amdg_ctx * c, int a, double alpha, int b, double beta, int * out)
{
    if (!c || !out || a < 0 || b < 0 || a >= (int)c->ops.size() || b >= (int)c->ops.size()) return fail(AMDG_EINVAL, "bad operator handle");
    const Op & A = *c->ops[a]; const Op & B = *c->ops[b];
    if (A.kf != B.kf || A.kt != B.kt) return fail(AMDG_EINVAL, "operator shapes differ");
    std::unique_ptr<Op> op(new Op()); op->kf = A.kf; op->kt = A.kt;
    op->blocks.resize(A.blocks.size());
    for (size_t i = 0; i < A.blocks.size(); ++i) op->blocks[i] = alpha * A.blocks[i] + beta * B.blocks[i];
    return push_op(c, std::move(op), out);
}

int amdg_internal_fail(int code, const char * msg) { return fail(code, msg ? msg : ""); }     // for the library's other translation units (tables.cu)

int amdg_ctx_info(amdg_ctx * c, int * out)
{
    if (!c || !out) return fail(AMDG_EINVAL, "null argument");
    out[0] = c->dim; out[1] = c->nmax; out[2] = c->edge_alpt - 1; out[3] = c->edge_intp - 1; out[4] = c->device;
    return AMDG_OK;
}

// ---- sweeps ------------------------------------------------------------------------------------------------------
static int check_op(amdg_ctx * c, int op) { return (op >= 0 && op < (int)c->ops.size()) ? AMDG_OK : fail(AMDG_EINVAL, "bad operator handle"); }

static int next_pow2(int x) { int p = 1; while (p < x) p <<= 1; return p; }

// shared-memory pitch: rows handled by one half-warp (16 lanes = (16/cx) rows x cx columns) should fall in distinct banks
static int choose_pitch(int ncol, int kf, int cx)
{
    if (cx >= 16) return ncol;
    int best = ncol, best_conf = 1 << 30;
    for (int P = ncol; P < ncol + 16; ++P)
    {
        int cnt[16] = { 0 }; int conf = 0;
        for (int lane = 0; lane < 16; ++lane) { const int tx = lane % cx, ty = lane / cx; cnt[(ty * kf * P + tx) % 16]++; }
        for (int b = 0; b < 16; ++b) conf = std::max(conf, cnt[b]);
        if (conf < best_conf) { best_conf = conf; best = P; }
    }
    return best;
}

// Work list of the fibre-staged kernel for sweeps along t with W columns, block edges kf -> kt and relation rel.
//
// PACKED item: a set of target rows plus the source rows they need, small enough that the staged rows, the
//   operator blocks of the distinct 1D pairs and the item-local neighbour lists fit in shared memory together.
//   Short fibres are packed several to an item (walked in slot order).  A long fibre is cut at a 1D level k_c:
//   the rows of level >= k_c are grouped by their level-k_c ancestor (a subtree); a subtree's sources are the
//   subtree itself, its ancestor chain and (flx relation) a few adjacent cells -- the union of its rows'
//   neighbour lists, so nothing about the tree shape is assumed.
// STREAMED item: the rows of level < k_c of a long fibre (a prefix in 1D order) have sources all over the fibre:
//   the whole fibre is staged for a narrow column range, one warp per target row with the lanes splitting the
//   neighbour list; operator blocks come from L2.  A fibre that cannot be cut falls back to this form entirely.
static const amdg_ctx::ItemList & get_items(amdg_ctx * c, int t, int W, int kf, int kt, int rel, int par, int lu)
{
    int pcls = 0; while ((1 << (pcls + 1)) <= par && pcls < 5) ++pcls;
    auto key = std::make_tuple(t, W, kf, kt, rel * 4 + lu, pcls);      // the lists of a packed item hold only the entries the L/U/full part uses
    auto it = c->items.find(key);
    if (it != c->items.end()) return it->second;
    amdg_ctx::ItemList L;
    const DimTables & H = c->grid.dims[t];
    const int d = c->dim;
    const int cap = c->smem_doubles;
    int ct = W > 128 ? 4 : (W > 16 ? 2 : 1);
    if (const char * e = std::getenv("AMDG_CT")) { const int v = atoi(e); if (v == 1 || v == 2 || v == 4) ct = v; }
    int lcx_w = 0; while ((1 << lcx_w) < std::min(fibre_threads(), next_pow2((W + ct - 1) / ct))) lcx_w++;
    const int pitch_w = choose_pitch(W, kf, 1 << lcx_w);
    const bool wide = (1 << lcx_w) * ct < W;          // more columns than one pass of the block covers: packed path not possible
    const int64_t total = c->grid.n * (int64_t)kf * pitch_w;
    const int64_t pack_cap = std::max<int64_t>(std::min<int64_t>(cap, total / std::max(1, c->item_target >> pcls)), (int64_t)kf * pitch_w);
    const std::vector<int64_t> & nptr = H.nbr_ptr[rel];
    const std::vector<Nbr> & nbr = H.nbr[rel];
    const std::vector<int> & nsplit = H.nbr_split[rel];
    std::vector<FibreItem> items; std::vector<double> cost;
    std::vector<int> pool_slots, pool_pairs, pool_rowptr, pool_rsplit; std::vector<Nbr> pool_ent;
    std::vector<int> pstamp(c->pairs.n_pairs, -1), plocal(c->pairs.n_pairs, 0);
    int64_t need = 0; bool ok = true; int stamp_id = 0;
    int n_packed = 0, n_streamed = 0;

    auto ord_of = [&](int64_t s) { return c->grid.ord1d[(int64_t)H.slot_elem[s] * d + t]; };
    auto need_p = [&](int64_t nsrc, int64_t npair, int64_t nnz, int64_t ntgt, int pitch) { return nsrc * kf * pitch + npair * kf * kt + (2 * nnz + 3 * ntgt + 2 + 1) / 2; };
    auto need_of = [&](int64_t nsrc, int64_t npair, int64_t nnz, int64_t ntgt) { return need_p(nsrc, npair, nnz, ntgt, pitch_w); };
    // column chunking of a packed item: chunk j covers W / 2^j columns
    auto chunk_cols = [&](int j) { return std::max(1, (W + (1 << j) - 1) >> j); };
    auto chunk_lcx = [&](int ncol) { int l = 0; while ((1 << l) < std::min(fibre_threads(), next_pow2((ncol + ct - 1) / ct))) l++; return l; };

    // emit one packed item: targets (slots) and, per target, its neighbour list; sources = union (kept in first-seen order)
    // fibre_s0: first slot of the fibre the targets belong to (all targets of one call are of one fibre) -- for packing
    // several fibres call begin/add/end.
    struct Build { std::vector<int> src, tgt, rowptr, rsplit; std::vector<Nbr> ent; std::vector<int> pairs; double cost = 0; } B;
    std::vector<int> sstamp(c->grid.n, -1), slocal(c->grid.n, 0);
    auto begin_item = [&]() { B = Build(); B.rowptr.push_back(0); ++stamp_id; };
    auto add_row = [&](int64_t s, int64_t fibre_s0)
    {
        B.tgt.push_back((int)s);
        const int64_t p_lo = (lu == AMDG_LU_L) ? nptr[s] + nsplit[s] : nptr[s];
        const int64_t p_hi = (lu == AMDG_LU_U) ? nptr[s] + nsplit[s] : nptr[s + 1];
        B.rsplit.push_back(lu == AMDG_LU_L ? 0 : nsplit[s]);
        for (int64_t p = p_lo; p < p_hi; ++p)
        {
            const int ss = (int)(fibre_s0 + nbr[p].local), pr = nbr[p].pair;
            if (sstamp[ss] != stamp_id) { sstamp[ss] = stamp_id; slocal[ss] = (int)B.src.size(); B.src.push_back(ss); }
            if (pstamp[pr] != stamp_id) { pstamp[pr] = stamp_id; plocal[pr] = (int)B.pairs.size(); B.pairs.push_back(pr); }
            B.ent.push_back({ slocal[ss], plocal[pr] });
        }
        B.rowptr.push_back((int)B.ent.size());
        B.cost += (double)(p_hi - p_lo) + 1.0;
    };
    auto item_need = [&]() { return need_of((int64_t)B.src.size(), (int64_t)B.pairs.size(), (int64_t)B.ent.size(), (int64_t)B.tgt.size()); };
    auto item_need_chunk = [&](int j) { const int nc = chunk_cols(j); return need_p((int64_t)B.src.size(), (int64_t)B.pairs.size(), (int64_t)B.ent.size(), (int64_t)B.tgt.size(), choose_pitch(nc, kf, 1 << chunk_lcx(nc))); };
    int end_chunks = 0;      // column chunking level used by end_item (0 = all columns)
    auto end_item = [&]()
    {
        if (B.tgt.empty()) return;
        FibreItem x; std::memset(&x, 0, sizeof(x));
        x.col0 = 0; x.ncol = W; x.lcx = lcx_w; x.pitch = pitch_w; x.packed = 1;
        if (end_chunks > 0) { x.ncol = chunk_cols(end_chunks); x.lcx = chunk_lcx(x.ncol); x.pitch = choose_pitch(x.ncol, kf, 1 << x.lcx); }
        x.npair = (int)B.pairs.size(); x.pair_ofs = (int)pool_pairs.size();
        x.nsrc = (int)B.src.size(); x.src_ofs = (int)pool_slots.size();
        for (int ss : B.src) pool_slots.push_back(H.slot_elem[ss]);          // element rows, not slots: one indirection less on the device
        x.ntgt = (int)B.tgt.size(); x.tgt_ofs = (int)pool_slots.size();
        for (int ss : B.tgt) pool_slots.push_back(H.slot_elem[ss]);
        x.ent_ofs = (int)pool_ent.size(); x.row_ofs = (int)pool_rowptr.size();
        pool_pairs.insert(pool_pairs.end(), B.pairs.begin(), B.pairs.end());
        pool_rowptr.insert(pool_rowptr.end(), B.rowptr.begin(), B.rowptr.end());
        pool_rsplit.insert(pool_rsplit.end(), B.rsplit.begin(), B.rsplit.end()); pool_rsplit.push_back(0);   // keep the pools aligned
        pool_ent.insert(pool_ent.end(), B.ent.begin(), B.ent.end());
        need = std::max(need, end_chunks > 0 ? item_need_chunk(end_chunks) : item_need());
        for (int c0 = 0; c0 < W; c0 += x.ncol) { x.col0 = c0; items.push_back(x); cost.push_back(B.cost * std::min(x.ncol, W - c0)); ++n_packed; }
        B = Build();
    };
    auto add_streamed = [&](int64_t s0, int m, int ntgt, double fc, int max_cols)
    {
        const int64_t slab = (int64_t)(fibre_threads() / 32) * 32 * (kf * kt + 2);      // per-warp operator-block slabs
        int ncol = (int)std::min<int64_t>(std::min(std::min(W, 32 * ct), max_cols), (cap - slab) / ((int64_t)m * kf));
        if (ncol < 1) { ok = false; return; }
        const int nchunk = (W + ncol - 1) / ncol;
        ncol = (W + nchunk - 1) / nchunk;
        FibreItem sp; std::memset(&sp, 0, sizeof(sp));
        sp.slot0 = (int)s0; sp.nslot = m; sp.ncol = ncol; sp.ntgt = ntgt;
        while ((1 << sp.lcx) < next_pow2((ncol + ct - 1) / ct)) sp.lcx++;
        sp.pitch = choose_pitch(ncol, kf, 1 << sp.lcx);
        if ((int64_t)m * kf * sp.pitch + slab > cap) sp.pitch = ncol;
        need = std::max(need, (int64_t)m * kf * sp.pitch + slab);
        for (int c0 = 0; c0 < W; c0 += ncol) { sp.col0 = c0; items.push_back(sp); cost.push_back(8.0 * fc * std::min(ncol, W - c0)); ++n_streamed; }
    };

    begin_item();
    for (int64_t f = 0; f < H.n_fibre && ok; ++f)
    {
        const int64_t s0 = H.fibre_ptr[f]; const int m = (int)(H.fibre_ptr[f + 1] - s0);
        const int64_t fnnz = nptr[s0 + m] - nptr[s0];
        if (wide) { add_streamed(s0, m, m, (double)fnnz, W); continue; }
        // (1) whole fibre into the running packed item?
        if (need_of(m, 1, fnnz, m) <= cap)
        {
            Build saved = B; const int saved_stamp = stamp_id;
            for (int64_t s = s0; s < s0 + m; ++s) add_row(s, s0);
            if (item_need() <= pack_cap || saved.tgt.empty()) { if (item_need() <= cap) continue; }
            // does not fit together with what is already there: close the old item and retry alone
            B = saved; (void)saved_stamp;
            end_item(); begin_item();
            for (int64_t s = s0; s < s0 + m; ++s) add_row(s, s0);
            if (item_need() <= cap) continue;
            begin_item();     // discard: falls through to the long-fibre handling
        }
        else { end_item(); begin_item(); }
        // (2) long fibre: cut at level k_c
        int lmax = 0; for (int64_t s = s0; s < s0 + m; ++s) lmax = std::max(lmax, level_of_order(ord_of(s)));
        bool cut_ok = false;
        for (int kc = 1; kc <= lmax && !cut_ok; ++kc)
        {
            // rows of level >= kc grouped by their level-kc ancestor: index (ord - 2^(n-1)) >> (n - kc)
            std::map<int, std::vector<int64_t>> groups; int ntop = 0;
            for (int64_t s = s0; s < s0 + m; ++s)
            {
                const int o = ord_of(s), n = level_of_order(o);
                if (n < kc) { ++ntop; continue; }
                groups[(o - (1 << (n - 1))) >> (n - kc)].push_back(s);
            }
            // feasibility: every group must fit
            bool fits = true;
            const size_t mark_items = items.size(), mark_slots = pool_slots.size(), mark_pairs = pool_pairs.size(), mark_rp = pool_rowptr.size(), mark_rs = pool_rsplit.size(), mark_ent = pool_ent.size();
            const int64_t mark_need = need; const int mark_np = n_packed;
            int jmax = 0;
            for (auto & g : groups)       // pass 1: the column chunking level every group of this cut can live with (at most 4 chunks)
            {
                begin_item();
                for (int64_t s : g.second) add_row(s, s0);
                int j = 0; while (j <= 2 && item_need_chunk(j) > cap) ++j;
                if (j > 2) { fits = false; break; }
                jmax = std::max(jmax, j);
            }
            begin_item();
            if (fits)
                for (auto & g : groups)
                {
                    begin_item();
                    for (int64_t s : g.second) add_row(s, s0);
                    end_chunks = jmax; end_item(); end_chunks = 0;
                }
            if (!fits)
            {
                items.resize(mark_items); cost.resize(mark_items); pool_slots.resize(mark_slots); pool_pairs.resize(mark_pairs);
                pool_rowptr.resize(mark_rp); pool_rsplit.resize(mark_rs); pool_ent.resize(mark_ent); need = mark_need; n_packed = mark_np;
                begin_item();
                continue;
            }
            // the top rows (a prefix of the fibre in 1D order) as streamed items over narrow column ranges
            double tc = 0; for (int64_t s = s0; s < s0 + ntop; ++s) tc += (double)(nptr[s + 1] - nptr[s]);
            if (ntop > 0 && lu == AMDG_LU_U)
            {
                // the sources of a top row under "U" are its ancestors: top rows only -> a packed item
                begin_item();
                for (int64_t s = s0; s < s0 + ntop; ++s) add_row(s, s0);
                int j = 0; while (j <= 2 && item_need_chunk(j) > cap) ++j;
                if (j <= 2) { end_chunks = j; end_item(); end_chunks = 0; }
                else { begin_item(); add_streamed(s0, m, ntop, tc, ct); }
            }
            else if (ntop > 0) add_streamed(s0, m, ntop, tc, ct);
            cut_ok = true;
            begin_item();
        }
        if (!cut_ok) { begin_item(); add_streamed(s0, m, m, (double)fnnz, W); }
    }
    if (ok)
    {
        end_item();
        std::vector<int> order(items.size());
        for (size_t i = 0; i < order.size(); ++i) order[i] = (int)i;
        std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return cost[x] > cost[y]; });
        std::vector<FibreItem> sorted;
        const char * only = std::getenv("AMDG_ONLY");     // experiment switch: "packed" / "streamed" (results are then incomplete)
        for (size_t i = 0; i < order.size(); ++i)
        {
            const FibreItem & x = items[order[i]];
            if (only && only[0] == 'p' && !x.packed) continue;
            if (only && only[0] == 's' && x.packed) continue;
            sorted.push_back(x);
        }
        if (need <= fibre_smem_capacity_doubles() &&
            meta_upload(c, &L.d_items, sorted.data(), sorted.size(), false) == cudaSuccess &&
            meta_upload(c, &L.d_slots, pool_slots.data(), pool_slots.size(), false) == cudaSuccess &&
            meta_upload(c, &L.d_pairs, pool_pairs.data(), pool_pairs.size(), false) == cudaSuccess &&
            meta_upload(c, &L.d_rowptr, pool_rowptr.data(), pool_rowptr.size(), false) == cudaSuccess &&
            meta_upload(c, &L.d_rsplit, pool_rsplit.data(), pool_rsplit.size(), false) == cudaSuccess &&
            meta_upload(c, (Nbr **)&L.d_ent, pool_ent.data(), pool_ent.size(), false) == cudaSuccess &&
            cudaStreamSynchronize(c->stream) == cudaSuccess)
        { L.n = (int)sorted.size(); L.ct = ct; L.smem = (int)need; L.ok = true; }
        if (std::getenv("AMDG_VERBOSE"))
            fprintf(stderr, "[amdg] items t=%d W=%d kf=%d kt=%d rel=%d par=%d lu=%d: %d items (%d packed, %d streamed), smem %lld doubles, ct %d\n",
                    t, W, kf, kt, rel, par, lu, (int)sorted.size(), n_packed, n_streamed, (long long)need, ct);
    }
    return c->items.emplace(key, L).first->second;
}

// ---- tensor-core kernel: work list + tile programs ---------------------------------------------------------------
// shared-memory row of one staged element (kernels.cu, sweep_mma_kernel): X[k][col] with a column pitch = 4 (mod 8), or, for a
// sweep along the last dimension (inner == 1), the element's own [col][k] order
// operator fragments are cached per (operator, shape, relation/part, piece): the piece is identified by its content, because the
// same shape is cut differently for different block shapes (the rows that fit in shared memory depend on outer/inner)
static long long piece_hash(const ShapeProg & P)
{
    unsigned long long h = 1469598103934665603ull;
    auto mix = [&](long long v) { h ^= (unsigned long long)v; h *= 1099511628211ull; };
    mix(P.m); mix(P.n_rt); mix(P.tg); mix(P.nkp);
    for (int v : P.rt_ptr) mix(v);
    for (int v : P.ent_src) mix(v % P.nkp);
    for (int v : P.ent_pair) mix(v);
    return (long long)(h >> 1);
}

static int mma_colpitch(int ncols) { int pk = ncols; while ((pk & 7) != 4) ++pk; return pk; }
static int64_t mma_rowsize(int kf, int no, int ni, int inner) { return inner == 1 ? (int64_t)no * ni * kf : (int64_t)kf * mma_colpitch(no * ni); }

static amdg_ctx::MmaList & get_mma(amdg_ctx * c, int t, int outer, int inner, int kf, int kt, int rel, int par, int lu)
{
    int pcls = 0; while ((1 << (pcls + 1)) <= par && pcls < 5) ++pcls;
    const int cap_doubles = c->mma_cap_doubles, item_target = c->mma_item_target;
    const int ent_target0 = c->mma_ent_target, stage_a_max = c->mma_stage_a_max;
    auto key = std::make_tuple(t * 16 + kf, outer, inner, kt, rel * 4 + lu, pcls);          // kf <= 6: (t, kf) share a field, outer and inner keep their own
    auto it = c->mmas.find(key);
    if (it != c->mmas.end()) return it->second;
    amdg_ctx::MmaList L;
    const std::map<int, std::vector<int>> & sf = c->shapes.shape_fibres[t];
    std::vector<MmaItem> items; std::vector<double> cost; std::vector<int> elem_pool;
    std::vector<int> prog_ints;
    const DimTables & H = c->grid.dims[t];
    bool ok = true; int smem_need = 0;
    const int64_t total = c->grid.n * mma_rowsize(kf, outer, inner, inner);
    const int64_t target = std::max<int64_t>(1, total / std::max(1, item_target >> pcls));
    // tile programs of every shape
    std::map<int, ShapeProg> shape_progs;
    for (auto & kv : sf) build_shape_prog(c->pairs, c->shapes.ords[kv.first], rel, lu, kf, kt, shape_progs[kv.first]);
    // A launch with fewer CTAs than the machine holds is bounded by its longest CTA: cut the programs finer (halve the
    // entry target) until the grid fills the SMs or the pieces reach 32 entries.
    int ent_target = ent_target0;
  retry_finer:
    const int64_t split_above = ent_target == ent_target0 ? (int64_t)4 * ent_target : (int64_t)ent_target * 3 / 2;
    items.clear(); cost.clear(); elem_pool.clear(); prog_ints.clear(); L = amdg_ctx::MmaList(); ok = true; smem_need = 0;
    for (auto & kv : sf)
    {
        const int shape = kv.first; const std::vector<int> & fibres = kv.second;
        if (const char * e = std::getenv("AMDG_MMA_MAXM")) { if ((int)c->shapes.ords[shape].size() > atoi(e)) continue; }     // experiment switch (results incomplete)
        if (const char * e = std::getenv("AMDG_MMA_MINM")) { if ((int)c->shapes.ords[shape].size() < atoi(e)) continue; }
        const ShapeProg & SP = shape_progs[shape];
        const int m = SP.m;
        // pieces: a long program is split so that one CTA walks about ent_target entries
        int np = 1;
        const int64_t row_full = mma_rowsize(kf, outer, inner, inner);
        const int slack = 32 * kf;                                   // tiles past the rectangle are read (never stored): keep them inside the allocation
        const bool whole_fits = (int64_t)m * row_full + (2 * SP.n_rt + 1 + SP.n_ent() + m) / 2 + 4 + slack <= cap_doubles;
        if (!whole_fits || SP.n_ent() > split_above) np = (int)std::min<int64_t>(std::max<int64_t>(1, (SP.n_ent() + ent_target - 1) / ent_target), std::max(1, SP.n_rt));
        std::vector<ShapeProg> pieces; split_shape_prog(SP, np, pieces);
        int max_piece_ints = 0; for (auto & pc : pieces) max_piece_ints = std::max(max_piece_ints, 2 * pc.n_rt + 1 + (int)pc.n_ent());
        const bool stage_a = np == 1 && SP.n_ent() <= stage_a_max;
        const int a_doubles = stage_a ? (int)SP.n_ent() * 32 : 0;
        // column rectangles
        struct Rect { int o0, no, i0, ni, pk; };
        std::vector<Rect> rects;
        int nfib_max = 1;
        const int cap = cap_doubles - ((max_piece_ints + m + 1) / 2 + 4) - a_doubles - slack;      // room for the piece, the element rows, staged A
        if (cap <= 0) { ok = false; break; }
        if ((int64_t)m * row_full <= cap)
        {
            rects.push_back({ 0, outer, 0, inner, mma_colpitch(outer * inner) });
            if (np == 1) nfib_max = (int)std::max<int64_t>(1, std::min<int64_t>(cap, target) / ((int64_t)m * row_full + (m + 1) / 2));
        }
        else if ((int64_t)m * mma_rowsize(kf, 1, inner, inner) <= cap)
        {
            int no = 1; while (no < outer && (int64_t)m * mma_rowsize(kf, no + 1, inner, inner) <= cap) ++no;
            for (int o0 = 0; o0 < outer; o0 += no) { const int n = std::min(no, outer - o0); rects.push_back({ o0, n, 0, inner, mma_colpitch(n * inner) }); }
        }
        else
        {
            int ni = 0;
            for (int cand = (inner / 8) * 8; cand >= 8; cand -= 8) if ((int64_t)m * mma_rowsize(kf, 1, cand, inner) <= cap) { ni = cand; break; }
            if (ni == 0) for (int cand : { 4, 2, 1 }) if (cand <= inner && (int64_t)m * mma_rowsize(kf, 1, cand, inner) <= cap) { ni = cand; break; }
            if (ni == 0) { ok = false; break; }
            for (int o0 = 0; o0 < outer; ++o0) for (int i0 = 0; i0 < inner; i0 += ni) { const int n = std::min(ni, inner - i0); rects.push_back({ o0, 1, i0, n, mma_colpitch(n) }); }
        }
        const int prog0 = (int)L.progs.size();
        std::vector<int> piece_ofs(np);
        for (int q = 0; q < np; ++q)
        {
            piece_ofs[q] = (int)prog_ints.size();
            prog_ints.insert(prog_ints.end(), pieces[q].rt_ptr.begin(), pieces[q].rt_ptr.end());
            prog_ints.insert(prog_ints.end(), pieces[q].rt_order.begin(), pieces[q].rt_order.end());
            prog_ints.insert(prog_ints.end(), pieces[q].ent_src.begin(), pieces[q].ent_src.end());
        }
        for (size_t f0 = 0; f0 < fibres.size(); f0 += nfib_max)
        {
            const int nf = (int)std::min<size_t>(nfib_max, fibres.size() - f0);
            const int eofs = (int)elem_pool.size();
            for (int b = 0; b < nf; ++b) for (int f = 0; f < m; ++f) elem_pool.push_back(H.slot_elem[fibres[f0 + b] + f]);
            for (const Rect & r : rects)
                for (int q = 0; q < np; ++q)
                {
                    MmaItem x; std::memset(&x, 0, sizeof(x));
                    x.prog = prog0 + q; x.elem_ofs = eofs; x.nfib = nf; x.o0 = r.o0; x.no = r.no; x.i0 = r.i0; x.ni = r.ni; x.pk = r.pk;
                    x.m = m; x.n_rt = pieces[q].n_rt; x.prog_ofs = piece_ofs[q]; x.n_ent = (int)pieces[q].n_ent();
                    x.ni_magic = r.ni <= 1 ? 0xffffffffu : (unsigned)((0x100000000ull + (unsigned)r.ni - 1) / (unsigned)r.ni);
                    x.stage_a = stage_a ? 1 : 0; x.nsrc = m; x.src_ofs = eofs;
                    x.nfib_magic = nf <= 1 ? 0xffffffffu : (unsigned)((0x100000000ull + (unsigned)nf - 1) / (unsigned)nf);
                    { const unsigned nrf = (unsigned)(pieces[q].n_rt * nf); x.unit_magic = nrf <= 1 ? 0xffffffffu : (unsigned)((0x100000000ull + nrf - 1) / nrf); }
                    items.push_back(x);
                    cost.push_back(((double)pieces[q].n_ent() + 2.0 * pieces[q].n_rt) * nf * ((r.no * r.ni + 7) / 8) + 0.01 * nf * m * r.no * r.ni);
                    const int n_ints = 2 * pieces[q].n_rt + 1 + (int)pieces[q].n_ent();
                    smem_need = std::max(smem_need, (int)((nf * m * mma_rowsize(kf, r.no, r.ni, inner) + 1) & ~(int64_t)1) + (n_ints + ((nf * m + 1) & ~1) + 1) / 2 + 2 + a_doubles + slack);
                }
        }
        for (int q = 0; q < np; ++q) { L.prog_shape.push_back(shape); L.prog_piece.push_back(piece_hash(pieces[q])); L.progs.push_back(std::move(pieces[q])); }
    }
    if (ok && ent_target > 32 && (int64_t)items.size() * (1 << pcls) < 3 * (int64_t)c->n_sm) { ent_target = std::max(32, ent_target / 2); goto retry_finer; }
    if (ok && !items.empty())
    {
        std::vector<int> order(items.size()); for (size_t i = 0; i < order.size(); ++i) order[i] = (int)i;
        std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return cost[x] > cost[y]; });
        std::vector<MmaItem> sorted(items.size()); for (size_t i = 0; i < order.size(); ++i) sorted[i] = items[order[i]];
        if (prog_ints.empty()) prog_ints.push_back(0);
        bool up = meta_upload(c, &L.d_items, sorted.data(), sorted.size(), false) == cudaSuccess &&
                  meta_upload(c, &L.d_elem_pool, elem_pool.data(), elem_pool.size(), false) == cudaSuccess &&
                  meta_upload(c, &L.d_prog_ints, prog_ints.data(), prog_ints.size(), false) == cudaSuccess &&
                  cudaStreamSynchronize(c->stream) == cudaSuccess;
        if (up) { L.n_item = (int)sorted.size(); L.smem_doubles = (smem_need + 1) & ~1; L.ok = true; }
        if (std::getenv("AMDG_VERBOSE"))
            fprintf(stderr, "[amdg] mma list t=%d outer=%d inner=%d kf=%d kt=%d rel=%d lu=%d par=%d: %d items, %d programs, smem %d doubles\n",
                    t, outer, inner, kf, kt, rel, lu, par, (int)sorted.size(), (int)L.progs.size(), L.smem_doubles);
    }
    return c->mmas.emplace(key, std::move(L)).first->second;
}

// Work list of the lean tensor-core kernel (kernels_tc.cu).  Short fibres: whole fibres x the whole column plane, several
// fibres per item (as get_mma).  A fibre too long for that is cut by TARGETS: row tiles are walked in depth-first order of
// the 1D tree (left end of the support, then level), and consecutive row tiles whose union of source rows still fits in
// shared memory with a 32-column rectangle form a piece -- a subtree plus its chain of ancestors.  Only the few row tiles of
// coarse targets, which read most of the fibre, fall back to narrow rectangles over the union of their sources.
// Plan of one fibre shape for the lean tensor-core kernel: host only (no device needed), cached per (shape, block shape).
static const std::vector<LeanPiece> & lean_plan(amdg_ctx * c, int shape, int kf, int kt, int rel, int lu, int outer, int inner)
{
    auto plan_key = std::make_tuple(shape, kf, kt, rel * 4 + lu, outer, inner);
    auto plan_it = c->lean_plans.find(plan_key);
    if (plan_it != c->lean_plans.end()) return plan_it->second;
    typedef LeanRect Rect;
    typedef LeanPiece Piece;
    const std::vector<int> & ords = c->shapes.ords[shape];
    const int m = (int)ords.size();
    const int cap_doubles = c->tc_cap_doubles, ent_target = c->tc_ent_target, stage_a_max = c->tc_stage_a_max;
    const int W = outer * inner;
    const int64_t row_full = mma_rowsize(kf, outer, inner, inner);
    const int slack = 32 * kf;
    auto ints_doubles = [&](const ShapeProg & P, int nf, int m) { return (2 * P.n_rt + 1 + (int)P.n_ent() + ((nf * m + 1) & ~1) + 1) / 2 + 2; };
    // a piece over the row tiles `rts` of S: piece-local source rows, remapped entries
    auto make_piece = [&](const ShapeProg & S, const std::vector<int> & rts, Piece & P)
    {
        std::vector<int> loc(S.m, -1);
        P.prog = ShapeProg(); P.src.clear();
        ShapeProg & Q = P.prog;
        Q.m = S.m; Q.tg = S.tg; Q.nkp = S.nkp; Q.ktp = S.ktp; Q.n_rt = (int)rts.size(); Q.rt_ptr.assign(1, 0); Q.rt_order = rts;
        for (int rt : rts) for (int p = S.rt_ptr[rt]; p < S.rt_ptr[rt + 1]; ++p) { const int f = S.ent_src[p] / S.nkp; if (loc[f] < 0) loc[f] = 0; }
        for (int f = 0; f < S.m; ++f) if (loc[f] == 0) { loc[f] = (int)P.src.size(); P.src.push_back(f); }
        // longest row tiles first (the warps take them round-robin)
        std::vector<int> order(rts.size()); for (size_t i = 0; i < order.size(); ++i) order[i] = (int)i;
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return S.rt_ptr[rts[a] + 1] - S.rt_ptr[rts[a]] > S.rt_ptr[rts[b] + 1] - S.rt_ptr[rts[b]]; });
        Q.rt_order.clear();
        for (int oi : order)
        {
            const int rt = rts[oi];
            Q.rt_order.push_back(rt);
            for (int p = S.rt_ptr[rt]; p < S.rt_ptr[rt + 1]; ++p)
            {
                const int f = S.ent_src[p] / S.nkp, kp = S.ent_src[p] % S.nkp;
                Q.ent_src.push_back(loc[f] * S.nkp + kp);
                for (int g = 0; g < S.tg; ++g) Q.ent_pair.push_back(S.ent_pair[(size_t)p * S.tg + g]);
            }
            Q.rt_ptr.push_back((int)Q.ent_src.size());
        }
    };
        ShapeProg SP; build_shape_prog(c->pairs, ords, rel, lu, kf, kt, SP);
        std::vector<Piece> pieces;
        const bool whole_fits = (int64_t)m * row_full + ints_doubles(SP, 1, m) + slack + (SP.n_ent() <= stage_a_max ? SP.n_ent() * 32 : 0) <= cap_doubles;
        if (whole_fits && (SP.n_ent() <= stage_a_max || (SP.n_ent() <= 2 * (int64_t)ent_target && !c->tc_force_stage)))
        {
            pieces.emplace_back();
            Piece & P = pieces.back();
            std::vector<int> all(SP.n_rt); for (int i = 0; i < SP.n_rt; ++i) all[i] = i;
            make_piece(SP, all, P);
            P.stage_a = SP.n_ent() <= stage_a_max; P.whole = true;
            P.rects.push_back({ 0, outer, 0, inner, mma_colpitch(W) });
        }
        else
        {
            // 32-column rectangles for the pieces cut by targets
            std::vector<Rect> wide;
            if (inner >= 32) { for (int o0 = 0; o0 < outer; ++o0) for (int i0 = 0; i0 < inner; i0 += 32) { const int n = std::min(32, inner - i0); wide.push_back({ o0, 1, i0, n, mma_colpitch(n) }); } }
            else { const int no = std::max(1, std::min(outer, 32 / inner)); for (int o0 = 0; o0 < outer; o0 += no) { const int n = std::min(no, outer - o0); wide.push_back({ o0, n, 0, inner, mma_colpitch(n * inner) }); } }
            int64_t row_wide = 0; for (const Rect & r : wide) row_wide = std::max(row_wide, mma_rowsize(kf, r.no, r.ni, inner));
            // depth-first key of every fibre-local element and of every row tile (its first target)
            std::vector<std::pair<int64_t, int>> rt_key(SP.n_rt);
            for (int rt = 0; rt < SP.n_rt; ++rt)
            {
                const int o = ords[rt * SP.tg], n = level_of_order(o);
                const int64_t left = n <= 1 ? 0 : (int64_t)(o - (1 << (n - 1))) << (c->nmax - (n - 1));
                rt_key[rt] = { left * 64 + n, rt };
            }
            std::sort(rt_key.begin(), rt_key.end());
            const int fixed = slack + (m + 1) / 2 + 8;                       // slack, element rows, alignment
            auto rows_cap = [&](int64_t rowsize, int n_rt_, int n_ent_) { return (int)((cap_doubles - fixed - (2 * n_rt_ + 1 + n_ent_ + 1) / 2 - (n_ent_ <= stage_a_max ? n_ent_ * 32 : 0)) / rowsize); };
            std::vector<int> coarse;
            std::vector<int> cur; std::vector<char> mark(m, 0); int cur_src = 0, cur_ent = 0;
            auto flush = [&]()
            {
                if (cur.empty()) return;
                pieces.emplace_back(); make_piece(SP, cur, pieces.back()); pieces.back().rects = wide;
                pieces.back().stage_a = pieces.back().prog.n_ent() <= stage_a_max;
                cur.clear(); std::fill(mark.begin(), mark.end(), 0); cur_src = 0; cur_ent = 0;
            };
            for (auto & kr : rt_key)
            {
                const int rt = kr.second;
                const int n_e = SP.rt_ptr[rt + 1] - SP.rt_ptr[rt];
                if (n_e == 0) { cur.push_back(rt); continue; }               // a row tile without sources still writes zeros
                std::vector<int> add;
                for (int p = SP.rt_ptr[rt]; p < SP.rt_ptr[rt + 1]; ++p) { const int f = SP.ent_src[p] / SP.nkp; if (!mark[f] && (add.empty() || add.back() != f)) add.push_back(f); }
                std::sort(add.begin(), add.end()); add.erase(std::unique(add.begin(), add.end()), add.end());
                int own = 0; { std::vector<char> seen(m, 0); for (int p = SP.rt_ptr[rt]; p < SP.rt_ptr[rt + 1]; ++p) { const int f = SP.ent_src[p] / SP.nkp; if (!seen[f]) { seen[f] = 1; ++own; } } }
                if (own > rows_cap(row_wide, 1, n_e)) { coarse.push_back(rt); continue; }
                const bool fits = cur_src + (int)add.size() <= rows_cap(row_wide, (int)cur.size() + 1, cur_ent + n_e) && cur_ent + n_e <= ent_target;
                if (!fits && cur_ent > 0)
                {
                    flush();
                    add.clear();
                    for (int p = SP.rt_ptr[rt]; p < SP.rt_ptr[rt + 1]; ++p) add.push_back(SP.ent_src[p] / SP.nkp);
                    std::sort(add.begin(), add.end()); add.erase(std::unique(add.begin(), add.end()), add.end());
                }
                for (int f : add) mark[f] = 1;
                cur_src += (int)add.size(); cur_ent += n_e; cur.push_back(rt);
            }
            flush();
            // coarse row tiles: balanced groups, narrow rectangles over the union of their sources
            if (!coarse.empty())
            {
                int64_t ce = 0; for (int rt : coarse) ce += SP.rt_ptr[rt + 1] - SP.rt_ptr[rt];
                const int ng = (int)std::min<int64_t>(std::max<int64_t>(1, (ce + c->tc_coarse_ent - 1) / c->tc_coarse_ent), (int64_t)coarse.size());
                std::stable_sort(coarse.begin(), coarse.end(), [&](int a, int b) { return SP.rt_ptr[a + 1] - SP.rt_ptr[a] > SP.rt_ptr[b + 1] - SP.rt_ptr[b]; });
                std::vector<std::vector<int>> groups(ng); std::vector<int64_t> load(ng, 0);
                for (int rt : coarse) { int best = 0; for (int q = 1; q < ng; ++q) if (load[q] < load[best]) best = q; groups[best].push_back(rt); load[best] += SP.rt_ptr[rt + 1] - SP.rt_ptr[rt] + 2; }
                // coarse pieces stage nothing: their sources are streamed from L2 (entries keep fibre-local source indices), 8 columns per item
                std::vector<Rect> narrow;
                if (inner >= 8) { for (int o0 = 0; o0 < outer; ++o0) for (int i0 = 0; i0 < inner; i0 += 8) { const int n = std::min(8, inner - i0); narrow.push_back({ o0, 1, i0, n, mma_colpitch(n) }); } }
                else { const int no = std::max(1, std::min(outer, 8 / inner)); for (int o0 = 0; o0 < outer; o0 += no) { const int n = std::min(no, outer - o0); narrow.push_back({ o0, n, 0, inner, mma_colpitch(n * inner) }); } }
                for (auto & gr : groups)
                {
                    if (gr.empty()) continue;
                    pieces.emplace_back(); Piece & P = pieces.back(); make_piece(SP, gr, P);
                    for (size_t p = 0; p < P.prog.ent_src.size(); ++p) { const int e = P.prog.ent_src[p]; P.prog.ent_src[p] = P.src[e / SP.nkp] * SP.nkp + e % SP.nkp; }
                    P.src.clear(); P.ksplit = true; P.stage_a = false; P.rects = narrow;
                }
            }
        }
        for (Piece & P : pieces) P.hash = piece_hash(P.prog);
    return c->lean_plans.emplace(plan_key, std::move(pieces)).first->second;
}

static amdg_ctx::MmaList & get_mma_lean(amdg_ctx * c, int t, int outer, int inner, int kf, int kt, int rel, int par, int lu)
{
    int pcls = 0; while ((1 << (pcls + 1)) <= par && pcls < 5) ++pcls;
    auto key = std::make_tuple(t * 16 + kf, outer, inner, kt, rel * 4 + lu, pcls + 16);
    auto it = c->mmas.find(key);
    if (it != c->mmas.end()) return it->second;
    amdg_ctx::MmaList L;
    const std::map<int, std::vector<int>> & sf = c->shapes.shape_fibres[t];
    std::vector<MmaItem> items; std::vector<double> cost; std::vector<int> elem_pool, prog_ints;
    const DimTables & H = c->grid.dims[t];
    bool ok = true; int smem_need = 0;
    const int cap_doubles = c->tc_cap_doubles;
    const int64_t row_full = mma_rowsize(kf, outer, inner, inner);
    const int64_t total = c->grid.n * row_full;
    const int64_t target = std::max<int64_t>(1, total / std::max(1, c->tc_item_target >> pcls));
    const int slack = 32 * kf;
    typedef LeanRect Rect;
    typedef LeanPiece Piece;
    for (auto & kv : sf)
    {
        const int shape = kv.first; const std::vector<int> & fibres = kv.second;
        const int m = (int)c->shapes.ords[shape].size();
        const std::vector<Piece> & pieces = lean_plan(c, shape, kf, kt, rel, lu, outer, inner);
        // emit: programs, element rows (whole fibres for the targets, staged rows per piece), items
        const int np = (int)pieces.size();
        const int prog0 = (int)L.progs.size();
        std::vector<int> piece_ofs(np);
        for (int q = 0; q < np; ++q)
        {
            const ShapeProg & Q = pieces[q].prog;
            piece_ofs[q] = (int)prog_ints.size();
            prog_ints.insert(prog_ints.end(), Q.rt_ptr.begin(), Q.rt_ptr.end());
            prog_ints.insert(prog_ints.end(), Q.rt_order.begin(), Q.rt_order.end());
            prog_ints.insert(prog_ints.end(), Q.ent_src.begin(), Q.ent_src.end());
        }
        for (int q = 0; q < np; ++q)
        {
            const Piece & P = pieces[q]; const ShapeProg & Q = P.prog;
            const int nsrc = (int)P.src.size();
            const int a_doubles = P.stage_a ? (int)Q.n_ent() * 32 : 0;
            int nfib_max = 1;                                            // whole-fibre pieces take several fibres per item
            if (P.whole)
            {
                const int64_t room = std::min<int64_t>(cap_doubles - slack - a_doubles - (2 * Q.n_rt + 1 + (int)Q.n_ent()) / 2 - 4, target);
                nfib_max = (int)std::max<int64_t>(1, room / ((int64_t)m * row_full + (m + 1) / 2 + 1));
            }
            for (size_t f0 = 0; f0 < fibres.size(); f0 += nfib_max)
            {
                const int nf = (int)std::min<size_t>(nfib_max, fibres.size() - f0);
                const int eofs = (int)elem_pool.size();
                for (int b = 0; b < nf; ++b) for (int f = 0; f < m; ++f) elem_pool.push_back(H.slot_elem[fibres[f0 + b] + f]);
                int sofs = eofs;
                if (nsrc != m && !P.ksplit)
                {
                    sofs = (int)elem_pool.size();
                    for (int b = 0; b < nf; ++b) for (int s2 = 0; s2 < nsrc; ++s2) elem_pool.push_back(H.slot_elem[fibres[f0 + b] + P.src[s2]]);
                }
                for (const Rect & r : P.rects)
                {
                    MmaItem x; std::memset(&x, 0, sizeof(x));
                    x.prog = prog0 + q; x.elem_ofs = eofs; x.nfib = nf; x.o0 = r.o0; x.no = r.no; x.i0 = r.i0; x.ni = r.ni; x.pk = r.pk;
                    x.m = m; x.n_rt = Q.n_rt; x.prog_ofs = piece_ofs[q]; x.n_ent = (int)Q.n_ent();
                    x.ni_magic = r.ni <= 1 ? 0xffffffffu : (unsigned)((0x100000000ull + (unsigned)r.ni - 1) / (unsigned)r.ni);
                    x.stage_a = P.stage_a ? 1 : 0; x.nsrc = nsrc; x.src_ofs = sofs; x.ksplit = P.ksplit ? 1 : 0;
                    { const unsigned nrun = (unsigned)(r.no * kf); x.nrun_magic = nrun <= 1 ? 0xffffffffu : (unsigned)((0x100000000ull + nrun - 1) / nrun); }
                    x.nfib_magic = nf <= 1 ? 0xffffffffu : (unsigned)((0x100000000ull + (unsigned)nf - 1) / (unsigned)nf);
                    { const unsigned nrf = (unsigned)(Q.n_rt * nf); x.unit_magic = nrf <= 1 ? 0xffffffffu : (unsigned)((0x100000000ull + nrf - 1) / nrf); }
                    items.push_back(x);
                    cost.push_back(((double)Q.n_ent() + 2.0 * Q.n_rt) * nf * ((r.no * r.ni + 7) / 8) + 0.02 * nf * nsrc * r.no * r.ni);
                    const int n_ints = 2 * Q.n_rt + 1 + (int)Q.n_ent();
                    smem_need = std::max(smem_need, (int)((nf * nsrc * mma_rowsize(kf, r.no, r.ni, inner) + 1) & ~(int64_t)1) + (n_ints + ((nf * m + 1) & ~1) + 1) / 2 + 2 + a_doubles + slack + (P.ksplit ? 260 : 0));
                }
            }
        }
        for (int q = 0; q < np; ++q) { L.prog_shape.push_back(shape); L.prog_piece.push_back(pieces[q].hash); L.progs.push_back(pieces[q].prog); }
    }
    if (ok && smem_need > tc_smem_capacity_doubles()) ok = false;
    if (ok && !items.empty())
    {
        std::vector<int> order(items.size()); for (size_t i = 0; i < order.size(); ++i) order[i] = (int)i;
        std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return cost[x] > cost[y]; });
        std::vector<MmaItem> sorted(items.size()); for (size_t i = 0; i < order.size(); ++i) sorted[i] = items[order[i]];
        if (prog_ints.empty()) prog_ints.push_back(0);
        bool up = meta_upload(c, &L.d_items, sorted.data(), sorted.size(), false) == cudaSuccess &&
                  meta_upload(c, &L.d_elem_pool, elem_pool.data(), elem_pool.size(), false) == cudaSuccess &&
                  meta_upload(c, &L.d_prog_ints, prog_ints.data(), prog_ints.size(), false) == cudaSuccess &&
                  cudaStreamSynchronize(c->stream) == cudaSuccess;
        if (up) { L.n_item = (int)sorted.size(); L.smem_doubles = (smem_need + 1) & ~1; L.ok = true; }
        if (std::getenv("AMDG_VERBOSE"))
            fprintf(stderr, "[amdg] lean list t=%d outer=%d inner=%d kf=%d kt=%d rel=%d lu=%d par=%d: %d items, %d programs, smem %d doubles\n",
                    t, outer, inner, kf, kt, rel, lu, par, (int)sorted.size(), (int)L.progs.size(), L.smem_doubles);
    }
    return c->mmas.emplace(key, std::move(L)).first->second;
}

// Diagnostic (host only, works without a device): builds the lean plans of every fibre shape of dimension t for a sweep with the
// given block edges and checks their invariants -- every row tile of a shape's tile program lies in exactly one piece with all its
// entries, staged pieces index their own source rows and fit the shared-memory capacity with at least one rectangle, coarse pieces
// keep fibre-local sources and at most 8 columns, the rectangles of a piece tile the column plane exactly once.
// out[0..5] = shapes, pieces, coarse (streamed) pieces, entries, largest staged row count, largest shared-memory need (doubles).
int amdg_lean_plan_check(amdg_ctx * c, int t, const int * sizes_from, int kf, int kt, int rel, int lu, int64_t * out)
{
    if (!c || !c->have_grid) return fail(AMDG_ESTATE, "no grid");
    if (t < 0 || t >= c->dim || !sizes_from || !out || rel < 0 || rel > 1 || lu < 0 || lu > 2 || !sweep_shape_supported(kf, kt)) return fail(AMDG_EINVAL, "bad arguments");
    int outer = 1, inner = 1;
    for (int k = 0; k < t; ++k) outer *= sizes_from[k];
    for (int k = t + 1; k < c->dim; ++k) inner *= sizes_from[k];
    const int W = outer * inner, slack = 32 * kf;
    for (int i = 0; i < 6; ++i) out[i] = 0;
    for (auto & kv : c->shapes.shape_fibres[t])
    {
        const int shape = kv.first;
        const std::vector<int> & ords = c->shapes.ords[shape];
        const int m = (int)ords.size();
        ShapeProg SP; build_shape_prog(c->pairs, ords, rel, lu, kf, kt, SP);
        const std::vector<LeanPiece> & pieces = lean_plan(c, shape, kf, kt, rel, lu, outer, inner);
        std::vector<int> seen(SP.n_rt, 0);
        out[0]++;
        for (const LeanPiece & P : pieces)
        {
            const ShapeProg & Q = P.prog;
            out[1]++; out[2] += P.ksplit ? 1 : 0; out[3] += Q.n_ent();
            if ((int)Q.rt_order.size() != Q.n_rt || (int)Q.rt_ptr.size() != Q.n_rt + 1) return fail(AMDG_EINVAL, "plan: inconsistent piece");
            const int nsrc = P.ksplit ? m : (int)P.src.size();
            for (int i = 0; i < Q.n_rt; ++i)
            {
                const int rt = Q.rt_order[i];
                if (rt < 0 || rt >= SP.n_rt) return fail(AMDG_EINVAL, "plan: row tile out of range");
                seen[rt]++;
                const int n_e = Q.rt_ptr[i + 1] - Q.rt_ptr[i];
                if (n_e != SP.rt_ptr[rt + 1] - SP.rt_ptr[rt]) return fail(AMDG_EINVAL, "plan: a row tile lost entries");
                for (int p = 0; p < n_e; ++p)
                {
                    const int es = Q.ent_src[Q.rt_ptr[i] + p], eo = SP.ent_src[SP.rt_ptr[rt] + p];
                    const int f = es / SP.nkp;
                    if (f < 0 || f >= nsrc || es % SP.nkp != eo % SP.nkp) return fail(AMDG_EINVAL, "plan: bad source index");
                    const int f_fibre = P.ksplit ? f : P.src[f];
                    if (f_fibre != eo / SP.nkp) return fail(AMDG_EINVAL, "plan: source row mismatch");
                    for (int g = 0; g < SP.tg; ++g)
                        if (Q.ent_pair[(size_t)(Q.rt_ptr[i] + p) * SP.tg + g] != SP.ent_pair[(size_t)(SP.rt_ptr[rt] + p) * SP.tg + g]) return fail(AMDG_EINVAL, "plan: pair mismatch");
                }
            }
            // rectangles tile the column plane
            std::vector<char> col(W, 0);
            for (const LeanRect & r : P.rects)
            {
                if (P.ksplit && r.no * r.ni > 8) return fail(AMDG_EINVAL, "plan: coarse piece wider than 8 columns");
                for (int o = r.o0; o < r.o0 + r.no; ++o) for (int i = r.i0; i < r.i0 + r.ni; ++i)
                {
                    if (o < 0 || o >= outer || i < 0 || i >= inner || col[o * inner + i]) return fail(AMDG_EINVAL, "plan: rectangles overlap or leave the plane");
                    col[o * inner + i] = 1;
                }
                const int64_t need = (P.ksplit ? 0 : (((int64_t)P.src.size() * mma_rowsize(kf, r.no, r.ni, inner) + 1) & ~(int64_t)1)) +
                                     (2 * Q.n_rt + 1 + Q.n_ent() + ((m + 1) & ~1) + 1) / 2 + 2 + (P.stage_a ? Q.n_ent() * 32 : 0) + slack + (P.ksplit ? 260 : 0);
                if (need > tc_smem_capacity_doubles()) return fail(AMDG_EINVAL, "plan: a piece does not fit in shared memory");
                out[5] = std::max<int64_t>(out[5], need);
            }
            for (int i = 0; i < W; ++i) if (!col[i]) return fail(AMDG_EINVAL, "plan: rectangles do not cover the column plane");
            if (!P.ksplit) out[4] = std::max<int64_t>(out[4], (int64_t)P.src.size());
        }
        for (int rt = 0; rt < SP.n_rt; ++rt) if (seen[rt] != 1) return fail(AMDG_EINVAL, "plan: a row tile is not covered exactly once");
    }
    return AMDG_OK;
}

// device table of operator values (fragment order) for every program of a list, for operator `op`
static const double * const * get_mma_a_tab(amdg_ctx * c, amdg_ctx::MmaList & L, int op, int rel, int lu)
{
    auto it = L.a_tab.find(op);
    if (it != L.a_tab.end()) return it->second;
    const Op & O = *c->ops[op];
    std::vector<const double *> tab(L.progs.size());
    for (size_t i = 0; i < L.progs.size(); ++i)
    {
        auto key = std::make_tuple(op, L.prog_shape[i], rel * 4 + lu, L.prog_piece[i]);
        auto f = c->mma_A.find(key);
        if (f == c->mma_A.end())
        {
            std::vector<double> A; build_shape_A(L.progs[i], O.blocks.data(), O.kf, O.kt, A);
            A.resize(A.size() + 32, 0.0);                          // one entry of padding: the streaming kernel fetches one entry ahead
            double * d = nullptr;
            if (upload(&d, A.data(), A.size(), c->stream) != cudaSuccess || cudaStreamSynchronize(c->stream) != cudaSuccess) return nullptr;
            f = c->mma_A.emplace(key, d).first;
        }
        tab[i] = f->second;
    }
    const double ** dtab = nullptr;
    if (meta_upload(c, &dtab, tab.data(), tab.size(), false) != cudaSuccess || cudaStreamSynchronize(c->stream) != cudaSuccess) return nullptr;
    L.a_tab[op] = dtab;
    return dtab;
}

// work list of the pipelined kernel
static const amdg_ctx::PipeList & get_pipe(amdg_ctx * c, int t, int W, int kf, int kt, int rel, int par, int lu)
{
    int pcls = 0; while ((1 << (pcls + 1)) <= par && pcls < 5) ++pcls;
    auto key = std::make_tuple(t, W, kf, kt, rel * 4 + lu, pcls);
    auto it = c->pipes.find(key);
    if (it != c->pipes.end()) return it->second;
    amdg_ctx::PipeList L;
    PipeParams pp; pp.t = t; pp.W = W; pp.kf = kf; pp.kt = kt; pp.rel = rel; pp.lu = lu;
    pp.cap_doubles = c->pipe_cap_doubles; pp.meta_cap_ints = c->pipe_meta_ints - 4; pp.item_target = std::max(1, c->pipe_item_target >> pcls); pp.threads = pipe_threads();
    PipeBuild B;
    build_pipe_list(c->grid, c->pairs, pp, B);
    if (B.ok)
    {
        const int n = (int)B.tab.size() / 2;
        std::vector<int> order(n); for (int i = 0; i < n; ++i) order[i] = i;
        std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return B.cost[x] > B.cost[y]; });
        std::vector<int2> tab(n); for (int i = 0; i < n; ++i) { tab[i].x = B.tab[2 * order[i]]; tab[i].y = B.tab[2 * order[i] + 1]; }
        std::vector<int> zeros((size_t)std::max(1, B.n_final) * 64, 0);
        if (B.fin.empty()) { B.fin.push_back(0); }
        if (B.fin_ofs.empty()) { B.fin_ofs.push_back(0); }
        if (meta_upload(c, &L.d_rec, B.rec.data(), B.rec.size(), false) == cudaSuccess &&
            meta_upload(c, &L.d_tab, tab.data(), tab.size(), false) == cudaSuccess &&
            meta_upload(c, &L.d_fin, B.fin.data(), B.fin.size(), false) == cudaSuccess &&
            meta_upload(c, &L.d_fin_ofs, B.fin_ofs.data(), B.fin_ofs.size(), false) == cudaSuccess &&
            upload(&L.d_counters, zeros.data(), zeros.size(), c->stream) == cudaSuccess &&
            cudaStreamSynchronize(c->stream) == cudaSuccess)
        {
            L.n_item = n; L.n_final = B.n_final; L.n_slot = B.n_slot; L.ct = B.ct;
            L.data_doubles = (B.max_data + 1) & ~1; L.meta_ints = (B.max_meta + 3) & ~3; L.ok = true;
        }
        if (std::getenv("AMDG_VERBOSE"))
            fprintf(stderr, "[amdg] pipe list t=%d W=%d kf=%d kt=%d rel=%d lu=%d par=%d: %d items, %d finals, %d slots, data %d doubles, meta %d ints\n",
                    t, W, kf, kt, rel, lu, par, n, B.n_final, B.n_slot, L.data_doubles, L.meta_ints);
    }
    return c->pipes.emplace(key, L).first->second;
}

// ---- column kernel (kernels_col.cu) ---------------------------------------------------------------------------------
// columns per lane: the candidate with the fewest instructions per entry over all column groups of a block of W columns
static int col_pick_nc(const amdg_ctx * c, int W, int kf, int kt)
{
    const int mx = col_max_nc(kf, kt);
    if (c->col_force_nc >= 1) return std::min(c->col_force_nc, mx);
    int best = 1; double best_cost = 1e300;
    for (int nc = 1; nc <= mx; ++nc)
    {
        const int ng = (W + 32 * nc - 1) / (32 * nc);
        const double cost = ng * (double)(kf * nc + (kf * kt + 1) / 2 + kf * kt * nc + 10);
        if (cost < best_cost - 1e-9) { best_cost = cost; best = nc; }
    }
    return best;
}

// (source element row, pair id) of every neighbour entry of dimension t under relation rel, in slot order ("U" sources first)
static const int2 * get_col_ent(amdg_ctx * c, int t, int rel)
{
    auto it = c->col_ents.find(std::make_pair(t, rel));
    if (it != c->col_ents.end()) return it->second;
    const DimTables & H = c->grid.dims[t];
    std::vector<int2> ent(std::max<size_t>(H.nbr[rel].size(), 1));
    for (int64_t s = 0; s < c->grid.n; ++s)
    {
        const int64_t fb = H.fibre_ptr[H.slot_fibre[s]];
        for (int64_t p = H.nbr_ptr[rel][s]; p < H.nbr_ptr[rel][s + 1]; ++p) { ent[p].x = H.slot_elem[fb + H.nbr[rel][p].local]; ent[p].y = H.nbr[rel][p].pair; }
    }
    int2 * d = nullptr;
    if (meta_upload(c, &d, ent.data(), ent.size(), false) != cudaSuccess || cudaStreamSynchronize(c->stream) != cudaSuccess) return nullptr;
    c->col_ents[std::make_pair(t, rel)] = d;
    return d;
}

// units = (target, column group): heavy ones (long entry lists, longest first) then the rest fibre by fibre, column group by column group
static amdg_ctx::ColList & get_col(amdg_ctx * c, int t, int rel, int lu, int ng)
{
    auto key = std::make_tuple(t, rel * 4 + lu, ng, c->col_heavy_ent);
    auto it = c->cols.find(key);
    if (it != c->cols.end()) return it->second;
    const DimTables & H = c->grid.dims[t];
    std::vector<ColUnit> heavy, normal;
    normal.reserve((size_t)c->grid.n * ng);
    for (int64_t f = 0; f < H.n_fibre; ++f)
        for (int g = 0; g < ng; ++g)
            for (int64_t s = H.fibre_ptr[f]; s < H.fibre_ptr[f + 1]; ++s)
            {
                int64_t p0 = H.nbr_ptr[rel][s], p1 = H.nbr_ptr[rel][s + 1];
                if (lu == AMDG_LU_U) p1 = p0 + H.nbr_split[rel][s]; else if (lu == AMDG_LU_L) p0 += H.nbr_split[rel][s];
                ColUnit u; u.tgt = H.slot_elem[s]; u.ent0 = (int)p0; u.n_ent = (int)(p1 - p0); u.g = g;
                (u.n_ent > c->col_heavy_ent ? heavy : normal).push_back(u);
            }
    std::stable_sort(heavy.begin(), heavy.end(), [](const ColUnit & x, const ColUnit & y) { return x.n_ent > y.n_ent; });
    amdg_ctx::ColList L;
    L.n_heavy = (int)heavy.size(); L.n_unit = (int)(heavy.size() + normal.size());
    heavy.insert(heavy.end(), normal.begin(), normal.end());
    if (meta_upload(c, &L.d_units, heavy.data(), heavy.size(), false) != cudaSuccess || cudaStreamSynchronize(c->stream) != cudaSuccess) { L.d_units = nullptr; L.n_unit = 0; }
    if (std::getenv("AMDG_VERBOSE")) fprintf(stderr, "[amdg] column list t=%d rel=%d lu=%d groups=%d: %d units, %d heavy\n", t, rel, lu, ng, L.n_unit, L.n_heavy);
    return c->cols.emplace(key, L).first->second;
}

// launch one sweep for a batch of jobs sharing (op, rel, lu, t, inner); jobs are grouped by equal `outer`
static int launch_sweep(amdg_ctx * c, int op, int rel, int lu, int t, int inner, const SweepJob * jobs, int n_job, int n_comp)
{
    const Op & O = *c->ops[op];
    const DevDim & D = c->ddims[t];
    // destination maps / accumulate-from exist in the lean, register-direct and streaming kernels only (even map offsets are the caller's contract
    // whenever the block size is even: the kernels keep their 16-byte stores)
    bool mapped = false; for (int i = 0; i < n_job; ++i) mapped = mapped || jobs[i].dst_map || jobs[i].acc_from || jobs[i].dst2;
    // second destinations are written by the column kernel's epilogue; after any other kernel the blocks are copied by a row scatter (same result)
    auto scatter_second = [&](int first, int cnt) -> int
    {
        for (int i = first; i < first + cnt; ++i)
        {
            if (!jobs[i].dst2) continue;
            if (jobs[i].dst_map) return fail(AMDG_EINVAL, "a second destination needs a plain first destination outside the column kernel");
            const int64_t s_to = (int64_t)jobs[i].outer * inner * O.kt;
            cudaError_t e = launch_scatter_rows(jobs[i].dst, c->grid.n, (int)s_to, jobs[i].dst2, jobs[i].dst2_map, c->stream);
            if (e != cudaSuccess) return fail(AMDG_ECUDA, std::string("second-destination scatter: ") + cudaGetErrorString(e));
            c->launches++;
        }
        return AMDG_OK;
    };
    const int variant0 = (mapped && c->kernel_variant < 8) ? 5 : c->kernel_variant;
    c->sweeps_on_grid++;
    const bool young = c->kernel_variant == 0 && !mapped && c->adaptive_mode && c->sweeps_on_grid <= c->adaptive_life && c->grid.dims[t].max_fibre_len <= c->adaptive_max_fibre;
    int done = 0;
    while (done < n_job)
    {
        int cnt = 1;
        while (done + cnt < n_job && cnt < MAX_JOBS && jobs[done + cnt].outer == jobs[done].outer) ++cnt;
        const int W = jobs[done].outer * inner;
        // auto mode picks per block shape (measured, profiles/r02_sweep_kernels.md): the column kernel for large blocks with small operator blocks
        // (the 6-D shapes), the lean tensor-core kernel otherwise
        int variant = variant0;
        if (c->kernel_variant == 0 && W >= 32 && (int64_t)W * O.kf >= c->col_min_block && O.kf * O.kt <= c->col_max_kk) variant = 8;
        if (young && (int64_t)c->grid.n * W * O.kf < c->tc_min_doubles) variant = 1;        // adaptive mode: no work list is built for a grid that will be replaced soon
        const bool lean = variant == 0 || variant == 5;
        if (variant == 8)
        {
            const int nc = col_pick_nc(c, W, O.kf, O.kt);
            const int ng = (W + 32 * nc - 1) / (32 * nc);
            const int2 * ent = get_col_ent(c, t, rel);
            amdg_ctx::ColList & CL = get_col(c, t, rel, lu, ng);
            if (!ent || !CL.d_units) return fail(AMDG_ENOMEM, "column kernel: the work list could not be uploaded");
            ColArgs a;
            a.units = CL.d_units; a.n_unit = CL.n_unit; a.n_heavy = CL.n_heavy; a.upc = c->col_upc; a.ent = ent; a.blocks = O.d_blocks;
            a.n_elem = c->grid.n; a.inner = inner; a.n_comp = n_comp; a.n_job = cnt;
            const bool dual_ok = O.kf <= 4 && O.kt <= 4;          // the column kernel's two-store epilogue is instantiated for these edges
            for (int i = 0; i < cnt; ++i) { a.job[i] = jobs[done + i]; if (!dual_ok) { a.job[i].dst2 = nullptr; a.job[i].dst2_map = nullptr; } }
            cudaError_t e = launch_sweep_col(a, O.kf, O.kt, nc, c->stream);
            if (e != cudaSuccess) return fail(AMDG_ECUDA, std::string("column sweep launch: ") + cudaGetErrorString(e));
            c->launches++;
            if (!dual_ok) { int r2 = scatter_second(done, cnt); if (r2) return r2; }
            done += cnt; continue;
        }
        if (variant == 0 || variant == 4 || variant == 5)
        {
            // lean tensor-core kernel first (variants 0 and 5); the whole-fibre tensor-core kernel when its list cannot be built (0) or on request (4)
            // (auto mode keeps the whole-fibre form for sweeps of a few KB, which are bounded by the latency of one CTA, not by throughput)
            bool launched = false;
            const bool tiny = variant == 0 && (int64_t)c->grid.n * W * O.kf < c->tc_min_doubles;
            for (int form = (lean && !tiny) ? 1 : 0; form >= 0 && !launched; --form)
            {
                amdg_ctx::MmaList & ML = form ? get_mma_lean(c, t, jobs[done].outer, inner, O.kf, O.kt, rel, cnt * n_comp, lu) : get_mma(c, t, jobs[done].outer, inner, O.kf, O.kt, rel, cnt * n_comp, lu);
                const double * const * atab = ML.ok ? get_mma_a_tab(c, ML, op, rel, lu) : nullptr;
                if (!(ML.ok && atab)) { if (form && variant == 5) break; continue; }
                MmaArgs a;
                a.items = ML.d_items; a.n_item = ML.n_item; a.prog_pool = ML.d_prog_ints; a.a_tab = atab; a.elem_pool = ML.d_elem_pool; a.dbg = c->dbg;
                a.n_elem = c->grid.n; a.inner = inner; a.n_comp = n_comp; a.n_job = cnt;
                for (int i = 0; i < cnt; ++i) a.job[i] = jobs[done + i];
                cudaError_t e = form ? launch_sweep_tc(a, O.kf, O.kt, ML.smem_doubles, c->stream) : launch_sweep_mma(a, O.kf, O.kt, ML.smem_doubles, c->stream);
                if (e != cudaSuccess) return fail(AMDG_ECUDA, std::string("tensor-core sweep launch: ") + cudaGetErrorString(e));
                launched = true;
            }
            if (launched) { c->launches++; int r2 = scatter_second(done, cnt); if (r2) return r2; done += cnt; continue; }
            if (variant == 4 || variant == 5) return fail(AMDG_EINVAL, "tensor-core kernel requested but the work list could not be built");
        }
        if (variant == 3)
        {
            while (cnt * n_comp > 64 && cnt > 1) --cnt;
            const amdg_ctx::PipeList & PL = get_pipe(c, t, W, O.kf, O.kt, rel, cnt * n_comp, lu);
            if (PL.ok && cnt * n_comp <= 64)
            {
                PipeArgs a;
                a.rec = PL.d_rec; a.tab = PL.d_tab; a.n_item = PL.n_item; a.fin = PL.d_fin; a.fin_ofs = PL.d_fin_ofs; a.counters = PL.d_counters;
                a.n_slot = PL.n_slot; a.partial = nullptr;
                if (PL.n_slot > 0)
                {
                    int64_t s_to = (int64_t)W * O.kt;
                    int r2 = ensure_scratch(c, 200, (int64_t)cnt * n_comp * PL.n_slot * s_to); if (r2) return r2;
                    a.partial = c->scratch[200];
                }
                a.blocks = O.d_blocks; a.n_elem = c->grid.n; a.inner = inner; a.n_comp = n_comp; a.n_job = cnt;
                a.inner_magic = inner <= 1 ? 0xffffffffu : (unsigned)((0x100000000ull + (unsigned)inner - 1) / (unsigned)inner);
                a.data_doubles = PL.data_doubles; a.meta_ints = PL.meta_ints; a.dbg = c->dbg;
                for (int i = 0; i < cnt; ++i) a.job[i] = jobs[done + i];
                cudaError_t e = launch_sweep_pipe(a, O.kf, O.kt, PL.ct, c->n_sm, c->stream);
                if (e != cudaSuccess) return fail(AMDG_ECUDA, std::string("pipelined sweep launch: ") + cudaGetErrorString(e));
                c->launches++;
                { int r2 = scatter_second(done, cnt); if (r2) return r2; }
                done += cnt;
                continue;
            }
            if (variant == 3) return fail(AMDG_EINVAL, "pipelined kernel requested but the work list could not be built");
        }
        const amdg_ctx::ItemList * L = nullptr;
        if (variant != 1) { L = &get_items(c, t, W, O.kf, O.kt, rel, cnt * n_comp, lu); if (!L->ok) L = nullptr; }
        if (variant == 2 && !L) return fail(AMDG_EINVAL, "fibre-staged kernel requested but a fibre does not fit in shared memory");
        cudaError_t e;
        if (L)
        {
            FibreSweepArgs a;
            a.slot_elem = D.slot_elem; a.slot_fbase = D.slot_fbase; a.nbr_ptr = D.nbr_ptr[rel]; a.nbr_split = D.nbr_split[rel]; a.nbr = D.nbr[rel];
            a.pool_slots = L->d_slots; a.pool_pairs = L->d_pairs; a.pool_rowptr = L->d_rowptr; a.pool_rsplit = L->d_rsplit; a.pool_ent = L->d_ent;
            a.blocks = O.d_blocks; a.items = L->d_items; a.n_item = L->n; a.n_elem = c->grid.n; a.inner = inner; a.lu = lu; a.n_comp = n_comp;
            a.n_job = cnt; a.smem_doubles = L->smem; a.dbg = c->dbg;
            for (int i = 0; i < cnt; ++i) a.job[i] = jobs[done + i];
            e = launch_sweep_fibre(a, O.kf, O.kt, L->ct, c->stream);
        }
        else
        {
            SweepArgs a;
            a.slot_elem = D.slot_elem; a.slot_fbase = D.slot_fbase; a.nbr_ptr = D.nbr_ptr[rel]; a.nbr_split = D.nbr_split[rel]; a.nbr = D.nbr[rel];
            a.blocks = O.d_blocks; a.n_elem = c->grid.n; a.inner = inner; a.lu = lu; a.n_comp = n_comp; a.n_job = cnt;
            for (int i = 0; i < cnt; ++i) a.job[i] = jobs[done + i];
            e = launch_sweep_gather(a, O.kf, O.kt, c->stream);
        }
        if (e != cudaSuccess) return fail(AMDG_ECUDA, std::string("sweep launch: ") + cudaGetErrorString(e));
        c->launches++;
        { int r2 = scatter_second(done, cnt); if (r2) return r2; }
        done += cnt;
    }
    return AMDG_OK;
}

int amdg_sweep1d(amdg_ctx * c, int op, int rel, int lu, int t, const int * sizes_from, const double * src, double * dst,
                 int n_comp, double coef, int accumulate)
{
    int r = need_device(c); if (r) return r;
    if (!c->have_grid) return fail(AMDG_ESTATE, "no grid");
    if ((r = check_op(c, op))) return r;
    if (t < 0 || t >= c->dim || rel < 0 || rel > 1 || lu < 0 || lu > 2 || !sizes_from || !src || !dst || n_comp < 1) return fail(AMDG_EINVAL, "bad sweep arguments");
    const Op & O = *c->ops[op];
    if (sizes_from[t] != O.kf) return fail(AMDG_EINVAL, "sizes_from[t] does not match the operator's source edge");
    if (src == dst) return fail(AMDG_EINVAL, "a sweep cannot run in place");
    int outer = 1, inner = 1;
    for (int k = 0; k < t; ++k) outer *= sizes_from[k];
    for (int k = t + 1; k < c->dim; ++k) inner *= sizes_from[k];
    SweepJob j; j.src = src; j.dst = dst; j.outer = outer; j.accumulate = accumulate; j.coef = coef;
    CU(cudaSetDevice(c->device));
    return launch_sweep(c, op, rel, lu, t, inner, &j, 1, n_comp);
}

static int sweep1d_batch_impl(amdg_ctx * c, int op, int rel, int lu, int t, const int * sizes_from, const double * const * src, double * const * dst,
                             const double * coef, const int * accumulate, const long long * const * dst_map, const double * const * acc_from, int n_job, int n_comp,
                             double * const * dst2 = nullptr, const long long * const * dst2_map = nullptr)
{
    int r = need_device(c); if (r) return r;
    if (!c->have_grid) return fail(AMDG_ESTATE, "no grid");
    if ((r = check_op(c, op))) return r;
    if (t < 0 || t >= c->dim || rel < 0 || rel > 1 || lu < 0 || lu > 2 || !sizes_from || !src || !dst || n_comp < 1 || n_job < 1) return fail(AMDG_EINVAL, "bad sweep arguments");
    const Op & O = *c->ops[op];
    std::vector<SweepJob> jobs(n_job);
    int inner0 = -1;
    for (int i = 0; i < n_job; ++i)
    {
        const int * sz = sizes_from + (size_t)i * c->dim;
        if (sz[t] != O.kf) return fail(AMDG_EINVAL, "sizes_from[t] does not match the operator's source edge");
        if (!src[i] || !dst[i] || src[i] == dst[i]) return fail(AMDG_EINVAL, "a sweep cannot run in place");
        int outer = 1, inner = 1;
        for (int k = 0; k < t; ++k) outer *= sz[k];
        for (int k = t + 1; k < c->dim; ++k) inner *= sz[k];
        if (inner0 < 0) inner0 = inner; else if (inner != inner0) return fail(AMDG_EINVAL, "the jobs of a batch must agree in the edges of the dims after t");
        jobs[i].src = src[i]; jobs[i].dst = dst[i]; jobs[i].outer = outer; jobs[i].accumulate = accumulate ? accumulate[i] : 0; jobs[i].coef = coef ? coef[i] : 1.0;
        jobs[i].dst_map = dst_map ? dst_map[i] : nullptr; jobs[i].acc_from = acc_from ? acc_from[i] : nullptr;
        if (dst2 && dst2[i]) { if (!dst2_map || !dst2_map[i]) return fail(AMDG_EINVAL, "a second destination needs its map"); jobs[i].dst2 = dst2[i]; jobs[i].dst2_map = dst2_map[i]; }
        if ((jobs[i].dst_map || jobs[i].acc_from || jobs[i].dst2) && n_comp != 1) return fail(AMDG_EINVAL, "mapped destinations take one component per job");
        if (jobs[i].acc_from && !jobs[i].accumulate) return fail(AMDG_EINVAL, "acc_from needs accumulate");
    }
    std::stable_sort(jobs.begin(), jobs.end(), [](const SweepJob & a, const SweepJob & b) { return a.outer < b.outer; });
    CU(cudaSetDevice(c->device));
    return launch_sweep(c, op, rel, lu, t, inner0, jobs.data(), n_job, n_comp);
}

int amdg_sweep1d_batch(amdg_ctx * c, int op, int rel, int lu, int t, const int * sizes_from, const double * const * src, double * const * dst,
                       const double * coef, const int * accumulate, int n_job, int n_comp)
{
    return sweep1d_batch_impl(c, op, rel, lu, t, sizes_from, src, dst, coef, accumulate, nullptr, nullptr, n_job, n_comp);
}

int amdg_sweep1d_batch_mapped(amdg_ctx * c, int op, int rel, int lu, int t, const int * sizes_from, const double * const * src, double * const * dst,
                              const double * coef, const int * accumulate, const int64_t * const * dst_map, const double * const * acc_from, int n_job)
{
    return sweep1d_batch_impl(c, op, rel, lu, t, sizes_from, src, dst, coef, accumulate, reinterpret_cast<const long long * const *>(dst_map), acc_from, n_job, 1);
}

int amdg_sweep1d_batch_dual(amdg_ctx * c, int op, int rel, int lu, int t, const int * sizes_from, const double * const * src, double * const * dst,
                            const double * coef, const int * accumulate, const int64_t * const * dst_map, const double * const * acc_from,
                            double * const * dst2, const int64_t * const * dst2_map, int n_job)
{
    return sweep1d_batch_impl(c, op, rel, lu, t, sizes_from, src, dst, coef, accumulate, reinterpret_cast<const long long * const *>(dst_map), acc_from, n_job, 1,
                              dst2, reinterpret_cast<const long long * const *>(dst2_map));
}

static int64_t ipow(int b, int e) { int64_t r = 1; while (e-- > 0) r *= b; return r; }

// the reference's schedule: source/FastMultiplyLU.cpp:614-664 (orderings), :121-142 (chain), :596-612 (sizes)
static int apply_tensor_literal(amdg_ctx * c, const int * ops, const int * rels, const double * src, double * dst, int n_comp, double coef, int accumulate)
{
    const int d = c->dim; const int64_t n = c->grid.n;
    const int kf = c->ops[ops[0]]->kf, kt = c->ops[ops[0]]->kt;
    const int64_t cap = n * n_comp * ipow(std::max(kf, kt), d);
    int r;
    if ((r = ensure_scratch(c, 0, cap)) || (r = ensure_scratch(c, 1, cap))) return r;
    const int n_chain = d == 1 ? 1 : (1 << (d - 1));
    for (int ch = 0; ch < n_chain; ++ch)
    {
        // choice of dim k < d-1: bit (d-2-k) of ch (row-major IterativeNestedLoop), 0 = L, 1 = U
        std::vector<int> order, lus;
        for (int k = 0; k < d - 1; ++k) if (((ch >> (d - 2 - k)) & 1) == 0) { order.push_back(k); lus.push_back(AMDG_LU_L); }
        order.push_back(d - 1); lus.push_back(AMDG_LU_FULL);
        for (int k = 0; k < d - 1; ++k) if (((ch >> (d - 2 - k)) & 1) == 1) { order.push_back(k); lus.push_back(AMDG_LU_U); }
        std::vector<int> sizes(d, kf);
        const double * cur = src;
        for (int step = 0; step < d; ++step)
        {
            const int t = order[step];
            int outer = 1, inner = 1;
            for (int k = 0; k < t; ++k) outer *= sizes[k];
            for (int k = t + 1; k < d; ++k) inner *= sizes[k];
            const bool last = (step == d - 1);
            SweepJob j;
            j.src = cur; j.dst = last ? dst : c->scratch[step & 1]; j.outer = outer;
            j.accumulate = last ? ((ch > 0) || accumulate) : 0;
            j.coef = (step == 0) ? coef : 1.0;
            if ((r = launch_sweep(c, ops[t], rels[t], lus[step], t, inner, &j, 1, n_comp))) return r;
            cur = j.dst; sizes[t] = kt;
        }
    }
    return AMDG_OK;
}

// Same sum with shared prefixes/suffixes.  With X_S = (prod_{k in S} L_k) x for S subset of {0..d-2},
//   Y_S = F_{d-1} X_S,   R_{d-1}(S) = Y_S,   R_k(S) = U_k R_{k+1}(S) + R_{k+1}(S + {k})  (S subset of {0..k-1}),
// the result is R_0({}).  Sweeps: (2^(d-1)-1) L + 2^(d-1) full + (2^(d-1)-1) U instead of d*2^(d-1); all sweeps of
// one level go out as one batched launch.  Buffers: one per subset S (index = bitmask of S).
static int apply_tensor_shared(amdg_ctx * c, const int * ops, const int * rels, const double * src, double * dst, int n_comp, double coef, int accumulate)
{
    const int d = c->dim; const int64_t n = c->grid.n;
    if (d == 1) return apply_tensor_literal(c, ops, rels, src, dst, n_comp, coef, accumulate);
    const int kf = c->ops[ops[0]]->kf, kt = c->ops[ops[0]]->kt;
    const int nsub = 1 << (d - 1);
    if (nsub > MAX_JOBS * 4) return fail(AMDG_EINVAL, "dimension too large for the shared schedule");
    const int64_t cap = n * n_comp * ipow(std::max(kf, kt), d);
    int r;
    // X buffers 0..nsub-1 (X_0 = src itself), Y buffers nsub..2*nsub-1
    for (int s = 1; s < 2 * nsub; ++s) if ((r = ensure_scratch(c, (size_t)s, cap))) return r;
    auto xbuf = [&](int S) -> const double * { return S == 0 ? src : c->scratch[S]; };
    // R_1({0}) ends up in the Y buffer of the full set S = {0..d-2} (every level accumulates into the buffer of S + {k}); its blocks
    // have the destination's shape, so when the caller's dst is overwritten anyway that buffer IS dst and the closing
    // "dst += R_1({0})" pass disappears
    const bool y_in_dst = !accumulate && src != dst;
    auto ybuf = [&](int S) -> double * { return (y_in_dst && S == nsub - 1) ? dst : c->scratch[nsub + S]; };
    auto edge = [&](int S, int k) { return ((S >> k) & 1) ? kt : kf; };   // dims in S already have the target edge
    std::vector<SweepJob> jobs;
    // down: L_k applied to every X_S with S subset of {0..k-1}
    for (int k = 0; k < d - 1; ++k)
    {
        jobs.clear();
        int inner = 1; for (int q = k + 1; q < d; ++q) inner *= kf;
        for (int S = 0; S < (1 << k); ++S)
        {
            int outer = 1; for (int q = 0; q < k; ++q) outer *= edge(S, q);
            SweepJob j; j.src = xbuf(S); j.dst = c->scratch[S | (1 << k)]; j.outer = outer; j.accumulate = 0; j.coef = 1.0;
            jobs.push_back(j);
        }
        std::stable_sort(jobs.begin(), jobs.end(), [](const SweepJob & a, const SweepJob & b) { return a.outer < b.outer; });
        if ((r = launch_sweep(c, ops[k], rels[k], AMDG_LU_L, k, inner, jobs.data(), (int)jobs.size(), n_comp))) return r;
    }
    // full sweep along d-1 for every S; coef enters here (each chain passes through exactly one full sweep)
    {
        jobs.clear();
        for (int S = 0; S < nsub; ++S)
        {
            int outer = 1; for (int q = 0; q < d - 1; ++q) outer *= edge(S, q);
            SweepJob j; j.src = xbuf(S); j.dst = ybuf(S); j.outer = outer; j.accumulate = 0; j.coef = coef;
            jobs.push_back(j);
        }
        std::stable_sort(jobs.begin(), jobs.end(), [](const SweepJob & a, const SweepJob & b) { return a.outer < b.outer; });
        if ((r = launch_sweep(c, ops[d - 1], rels[d - 1], AMDG_LU_FULL, d - 1, 1, jobs.data(), (int)jobs.size(), n_comp))) return r;
    }
    // up: R_k(S) = U_k R_{k+1}(S) + R_{k+1}(S+{k}); the sum is accumulated into the buffer of S+{k}
    // (for k = 0 into the caller's dst).  After level k the live buffers are those of S subset of {0..k-1},
    // stored at index S + {k}... we keep R_{k}(S) in ybuf(S | bit k) to avoid a copy, tracked by `where`.
    std::vector<int> where(nsub);
    for (int S = 0; S < nsub; ++S) where[S] = S;      // R_{d-1}(S) lives in ybuf(S)
    for (int k = d - 2; k >= 0; --k)
    {
        jobs.clear();
        int inner = 1; for (int q = k + 1; q < d; ++q) inner *= kt;
        std::vector<int> nw(1 << k);
        for (int S = 0; S < (1 << k); ++S)
        {
            int outer = 1; for (int q = 0; q < k; ++q) outer *= edge(S, q);
            const int lo = where[S], hi = where[S | (1 << k)];
            SweepJob j; j.src = ybuf(lo); j.outer = outer; j.coef = 1.0;
            if (k == 0)
            {
                // final: dst (+)= U_0 R_1({}) ; then dst += R_1({0}) (already there when y_in_dst)
                j.dst = dst; j.accumulate = y_in_dst ? 1 : accumulate;
            }
            else { j.dst = ybuf(hi); j.accumulate = 1; }
            nw[S] = hi;
            jobs.push_back(j);
        }
        std::stable_sort(jobs.begin(), jobs.end(), [](const SweepJob & a, const SweepJob & b) { return a.outer < b.outer; });
        if ((r = launch_sweep(c, ops[k], rels[k], AMDG_LU_U, k, inner, jobs.data(), (int)jobs.size(), n_comp))) return r;
        for (int S = 0; S < (1 << k); ++S) where[S] = nw[S];
    }
    // dst += R_1({0})
    if (!y_in_dst)
    {
        const int64_t total = n * n_comp * ipow(kt, d);
        cudaError_t e = launch_axpby(total, 1.0, ybuf(where[0]), 1.0, dst, c->stream);
        if (e != cudaSuccess) return fail(AMDG_ECUDA, std::string("axpby launch: ") + cudaGetErrorString(e));
        c->launches++;
    }
    return AMDG_OK;
}

int amdg_apply_tensor(amdg_ctx * c, const int * ops, const int * rels, const double * src, double * dst, int n_comp, double coef, int accumulate)
{
    int r = need_device(c); if (r) return r;
    if (!c->have_grid) return fail(AMDG_ESTATE, "no grid");
    if (!ops || !rels || !src || !dst || n_comp < 1) return fail(AMDG_EINVAL, "bad arguments");
    for (int t = 0; t < c->dim; ++t)
    {
        if ((r = check_op(c, ops[t]))) return r;
        if (rels[t] < 0 || rels[t] > 1) return fail(AMDG_EINVAL, "bad relation");
        if (c->ops[ops[t]]->kf != c->ops[ops[0]]->kf || c->ops[ops[t]]->kt != c->ops[ops[0]]->kt) return fail(AMDG_EINVAL, "operators of one tensor product must share block edges");
    }
    CU(cudaSetDevice(c->device));
    if (c->sched == AMDG_SCHED_LITERAL) return apply_tensor_literal(c, ops, rels, src, dst, n_comp, coef, accumulate);
    return apply_tensor_shared(c, ops, rels, src, dst, n_comp, coef, accumulate);
}

// The *_coarse_grid forms (FastMultiplyLU::transform_1D_coarse_grid, reference source/FastMultiplyLU.cpp:514-594, and the drivers
// :18-30, 144-165, 760-779, 926-946): every sweep skips targets and sources whose levels sum to more than mesh_nmax and leaves the skipped
// targets at zero.  The kept elements are a downward-closed grid and relations are pairwise, so this is the plain transform on that
// sub-grid: rows are gathered, transformed by a sub-context that owns the sub-grid's tables, and added back into their rows.
int amdg_apply_tensor_coarse(amdg_ctx * c, const int * ops, const int * rels, const double * src, double * dst, int n_comp, double coef, int accumulate, int mesh_nmax)
{
    int r = need_device(c); if (r) return r;
    if (!c->have_grid) return fail(AMDG_ESTATE, "no grid");
    if (!ops || !rels || !src || !dst || n_comp < 1) return fail(AMDG_EINVAL, "bad arguments");
    for (int t = 0; t < c->dim; ++t)
    {
        if ((r = check_op(c, ops[t]))) return r;
        if (c->ops[ops[t]]->kf != c->ops[ops[0]]->kf || c->ops[ops[t]]->kt != c->ops[ops[0]]->kt) return fail(AMDG_EINVAL, "operators of one tensor product must share block edges");
    }
    CU(cudaSetDevice(c->device));
    const int64_t n = c->grid.n;
    auto it = c->coarse.find(mesh_nmax);
    if (it == c->coarse.end())
    {
        amdg_ctx::CoarseView V;
        std::vector<int> rows, lev, sup;
        for (int64_t e = 0; e < n; ++e)
        {
            int sum = 0; for (int t = 0; t < c->dim; ++t) sum += c->grid.level[e * c->dim + t];
            if (sum > mesh_nmax) continue;
            rows.push_back((int)e);
            for (int t = 0; t < c->dim; ++t) { lev.push_back(c->grid.level[e * c->dim + t]); sup.push_back(c->grid.suppt[e * c->dim + t]); }
        }
        V.n = (int64_t)rows.size();
        if (V.n > 0 && V.n < n)
        {
            g_arena_mb_override = 8;
            r = amdg_ctx_create(c->dim, c->nmax, c->edge_alpt - 1, c->edge_intp - 1, c->device, &V.sub);
            g_arena_mb_override = -1;
            if (r) return r;
            if ((r = amdg_grid_set(V.sub, V.n, lev.data(), sup.data()))) { amdg_ctx_destroy(V.sub); return r; }
            cudaError_t e = upload(&V.d_rows, rows.data(), rows.size(), c->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);     // rows is a local
            if (e != cudaSuccess) { amdg_ctx_destroy(V.sub); cudaFree(V.d_rows); return fail(AMDG_ECUDA, std::string("coarse view: ") + cudaGetErrorString(e)); }
        }
        it = c->coarse.emplace(mesh_nmax, std::move(V)).first;
    }
    amdg_ctx::CoarseView & V = it->second;
    const int kf = c->ops[ops[0]]->kf, kt = c->ops[ops[0]]->kt;
    const int64_t s_from = ipow(kf, c->dim), s_to = ipow(kt, c->dim);
    if (V.n == n) return amdg_apply_tensor(c, ops, rels, src, dst, n_comp, coef, accumulate);
    if (!accumulate) CU(cudaMemsetAsync(dst, 0, (size_t)n_comp * n * s_to * sizeof(double), c->stream));
    if (V.n == 0) return AMDG_OK;
    amdg_ctx * sub = V.sub;
    if (sub->stream != c->stream) { if ((r = amdg_ctx_set_stream(sub, (void *)c->stream))) return r; }
    sub->sched = c->sched; sub->kernel_variant = c->kernel_variant;
    V.op_map.resize(c->ops.size(), -1);
    std::vector<int> sub_ops(c->dim);
    for (int t = 0; t < c->dim; ++t)
    {
        int & m = V.op_map[ops[t]];
        if (m < 0)
        {
            const Op & O = *c->ops[ops[t]];
            if ((r = amdg_op_register_compact(sub, O.blocks.data(), c->pairs.n_pairs, O.kf, O.kt, O.hier ? 1 : 0, &m))) return r;
        }
        sub_ops[t] = m;
    }
    const int64_t need[2] = { (int64_t)n_comp * V.n * s_from, (int64_t)n_comp * V.n * s_to };
    for (int k = 0; k < 2; ++k)
    {
        if (V.cap[k] >= need[k]) continue;
        CU(cudaStreamSynchronize(c->stream));
        cudaFree(V.buf[k]); V.buf[k] = nullptr; V.cap[k] = 0;
        if (cudaMalloc((void **)&V.buf[k], (size_t)need[k] * sizeof(double)) != cudaSuccess) { cudaGetLastError(); return fail(AMDG_ENOMEM, "coarse view buffers"); }
        V.cap[k] = need[k];
    }
    for (int v = 0; v < n_comp; ++v)
        CU(launch_rows_gather(src + (int64_t)v * n * s_from, V.d_rows, V.n, (int)s_from, V.buf[0] + (int64_t)v * V.n * s_from, c->stream));
    const int64_t l0 = sub->launches;
    if ((r = amdg_apply_tensor(sub, sub_ops.data(), rels, V.buf[0], V.buf[1], n_comp, coef, 0))) return r;
    for (int v = 0; v < n_comp; ++v)
        CU(launch_rows_scatter_add(V.buf[1] + (int64_t)v * V.n * s_to, V.d_rows, V.n, (int)s_to, dst + (int64_t)v * n * s_to, c->stream));
    c->launches += (sub->launches - l0) + 2 * n_comp;
    return AMDG_OK;
}

int amdg_hierarchize(amdg_ctx * c, int hop, const double * src, double * dst, int n_comp)
{
    int r = need_device(c); if (r) return r;
    if (!c->have_grid) return fail(AMDG_ESTATE, "no grid");
    if ((r = check_op(c, hop))) return r;
    const Op & O = *c->ops[hop];
    if (!O.hier) return fail(AMDG_EINVAL, "operator is not a hierarchisation stencil");
    if (!src || !dst || n_comp < 1) return fail(AMDG_EINVAL, "bad arguments");
    const int d = c->dim; const int b = O.kf; const int64_t n = c->grid.n;
    const int64_t cap = n * n_comp * ipow(b, d);
    if ((r = ensure_scratch(c, 0, cap)) || (r = ensure_scratch(c, 1, cap))) return r;
    CU(cudaSetDevice(c->device));
    // pass t reads stage t and writes stage t+1 (source/Interplation.cpp:1260-1400); ancestors are "U" sources
    const double * cur = src;
    for (int t = 0; t < d; ++t)
    {
        int outer = 1, inner = 1;
        for (int k = 0; k < t; ++k) outer *= b;
        for (int k = t + 1; k < d; ++k) inner *= b;
        SweepJob j; j.src = cur; j.outer = outer; j.accumulate = 0; j.coef = 1.0;
        const bool last = (t == d - 1);
        // ping-pong so that the last pass lands in dst even when dst == src
        j.dst = last ? dst : c->scratch[t & 1];
        if (last && cur == dst)
        {
            j.dst = c->scratch[t & 1];
            if ((r = launch_sweep(c, hop, AMDG_REL_VOL, AMDG_LU_U, t, inner, &j, 1, n_comp))) return r;
            CU(cudaMemcpyAsync(dst, j.dst, (size_t)cap * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
            return AMDG_OK;
        }
        if ((r = launch_sweep(c, hop, AMDG_REL_VOL, AMDG_LU_U, t, inner, &j, 1, n_comp))) return r;
        cur = j.dst;
    }
    return AMDG_OK;
}

// ---- point-wise, RK ------------------------------------------------------------------------------------------------
int amdg_pointwise(amdg_ctx * c, int n_flux, const int * flux_id, const double * params, const double * up, double * fp, const double * pts)
{
    int r = need_device(c); if (r) return r;
    if (!c->have_grid) return fail(AMDG_ESTATE, "no grid");
    if (n_flux < 1 || n_flux > 8 || !flux_id || !up || !fp) return fail(AMDG_EINVAL, "bad arguments");
    PointwiseArgs a; a.up = up; a.fp = fp; a.pts = pts; a.n_points = c->grid.n * ipow(c->edge_intp, c->dim); a.n_flux = n_flux; a.dim = c->dim;
    for (int i = 0; i < n_flux; ++i)
    {
        a.flux_id[i] = flux_id[i];
        if (flux_id[i] == AMDG_FLUX_VLASOV_SMOOTH_E && !pts) return fail(AMDG_EINVAL, "the Vlasov product needs point coordinates");
        for (int k = 0; k < 4; ++k) a.params[i][k] = params ? params[i * 4 + k] : 0.0;
    }
    CU(cudaSetDevice(c->device));
    cudaError_t e = launch_pointwise(a, c->stream);
    if (e != cudaSuccess) return fail(AMDG_ECUDA, std::string("pointwise launch: ") + cudaGetErrorString(e));
    c->launches++;
    return AMDG_OK;
}

int amdg_pointwise_hermite2d(amdg_ctx * c, int n_flux, const int * flux_id, const double * params, const double * up, double * fp)
{
    int r = need_device(c); if (r) return r;
    if (!c->have_grid) return fail(AMDG_ESTATE, "no grid");
    if (c->dim != 2 || c->edge_intp != 4) return fail(AMDG_EINVAL, "the Hermite point-wise flux exists for DIM == 2 and HermBasis::PMAX == 3 only (as in the reference)");
    if (n_flux < 1 || n_flux > 8 || !flux_id || !up || !fp) return fail(AMDG_EINVAL, "bad arguments");
    PointwiseArgs a; a.up = up; a.fp = fp; a.pts = nullptr; a.n_points = c->grid.n * 16; a.n_flux = n_flux; a.dim = 2;
    for (int i = 0; i < n_flux; ++i)
    {
        if (flux_id[i] < AMDG_FLUX_LINEAR || flux_id[i] > AMDG_FLUX_COS) return fail(AMDG_EINVAL, "flux kind has no Hermite (derivative) form");
        a.flux_id[i] = flux_id[i];
        for (int k = 0; k < 4; ++k) a.params[i][k] = params ? params[i * 4 + k] : 0.0;
    }
    CU(cudaSetDevice(c->device));
    cudaError_t e = launch_pointwise_herm2d(a, c->stream);
    if (e != cudaSuccess) return fail(AMDG_ECUDA, std::string("pointwise launch: ") + cudaGetErrorString(e));
    c->launches++;
    return AMDG_OK;
}

int amdg_point_coords(amdg_ctx * c, const double * host_pts1d, double * dev_pts)
{
    int r = need_device(c); if (r) return r;
    if (!c->have_grid) return fail(AMDG_ESTATE, "no grid");
    if (!host_pts1d || !dev_pts) return fail(AMDG_EINVAL, "null argument");
    CU(cudaSetDevice(c->device));
    double * d1 = nullptr; const size_t n1 = (size_t)c->pairs.T * c->edge_intp;
    CU(upload(&d1, host_pts1d, n1, c->stream));
    cudaError_t e = launch_point_coords(d1, c->d_ord1d, c->grid.n, c->dim, c->edge_intp, dev_pts, c->stream);
    c->launches++;
    cudaStreamSynchronize(c->stream); cudaFree(d1);
    if (e != cudaSuccess) return fail(AMDG_ECUDA, std::string("point_coords launch: ") + cudaGetErrorString(e));
    return AMDG_OK;
}

int amdg_rk_stage(amdg_ctx * c, int scheme, int stage, double dt, const double * u_tn, double * u, const double * rhs, int64_t n)
{
    int r = need_device(c); if (r) return r;
    if (!u_tn || !u || !rhs || n < 0) return fail(AMDG_EINVAL, "bad arguments");
    CU(cudaSetDevice(c->device));
    cudaError_t e = launch_rk_stage(scheme, stage, dt, u_tn, u, rhs, n, c->stream);
    if (e != cudaSuccess) return fail(e == cudaErrorInvalidValue ? AMDG_EINVAL : AMDG_ECUDA, std::string("rk_stage: ") + cudaGetErrorString(e));
    c->launches++;
    return AMDG_OK;
}

int amdg_rk4_ode2nd_stage(amdg_ctx * c, int stage, double dt, const double * u_tn, const double * v_tn, double * u, double * v,
                          const double * rhs, double * ku, double * kv, int64_t n)
{
    int r = need_device(c); if (r) return r;
    if (!u_tn || !v_tn || !u || !v || !rhs || !ku || !kv || n < 0) return fail(AMDG_EINVAL, "bad arguments");
    CU(cudaSetDevice(c->device));
    cudaError_t e = launch_rk4_ode2nd_stage(stage, dt, u_tn, v_tn, u, v, rhs, ku, kv, n, c->stream);
    if (e != cudaSuccess) return fail(e == cudaErrorInvalidValue ? AMDG_EINVAL : AMDG_ECUDA, std::string("rk4_ode2nd_stage: ") + cudaGetErrorString(e));
    c->launches++;
    return AMDG_OK;
}

int amdg_axpby(amdg_ctx * c, int64_t n, double alpha, const double * x, double beta, double * y)
{
    int r = need_device(c); if (r) return r;
    if (!x || !y || n < 0) return fail(AMDG_EINVAL, "bad arguments");
    CU(cudaSetDevice(c->device));
    cudaError_t e = launch_axpby(n, alpha, x, beta, y, c->stream);
    if (e != cudaSuccess) return fail(AMDG_ECUDA, std::string("axpby: ") + cudaGetErrorString(e));
    c->launches++;
    return AMDG_OK;
}

// ---- point-wise expressions ---------------------------------------------------------------------------------------------
int amdg_points_set(amdg_ctx * c, const double * host_pts1d)
{
    int r = need_device(c); if (r) return r;
    if (!host_pts1d) return fail(AMDG_EINVAL, "null argument");
    CU(cudaSetDevice(c->device));
    const size_t n1 = (size_t)c->pairs.T * c->edge_intp;
    if (!c->d_pts1d) CU(cudaMalloc((void **)&c->d_pts1d, n1 * sizeof(double)));
    CU(cudaMemcpyAsync(c->d_pts1d, host_pts1d, n1 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return AMDG_OK;
}

int amdg_pointwise_expr(amdg_ctx * c, int n_var, const double * const * up, int n_other, const double * const * other, const int * other_map,
                        int n_out, double * const * out, const int * prog, int n_prog, const int * out_ptr, const double * consts, int n_const)
{
    int r = need_device(c); if (r) return r;
    if (!c->have_grid) return fail(AMDG_ESTATE, "no grid");
    if (n_var < 0 || n_var > PW_MAX_IO || n_other < 0 || n_other > PW_MAX_IO || n_out < 1 || n_out > PW_MAX_IO || !out || !prog || !out_ptr ||
        n_prog < 1 || n_prog > PW_MAX_OPS || n_const < 0 || n_const > PW_MAX_CONST || (n_var && !up) || (n_other && !other) || (n_const && !consts))
        return fail(AMDG_EINVAL, "bad point-wise expression arguments");
    PwExprArgs a; std::memset(&a, 0, sizeof(a));
    for (int i = 0; i < n_var; ++i) a.up[i] = up[i];
    for (int i = 0; i < n_other; ++i) a.other[i] = other[i];
    for (int i = 0; i < n_out; ++i) { if (!out[i]) return fail(AMDG_EINVAL, "null output"); a.out[i] = out[i]; }
    a.other_map = other_map; a.pts1d = c->d_pts1d; a.ord1d = c->d_ord1d;
    a.edge = c->edge_intp; a.dim = c->dim; a.n_out = n_out;
    a.block = (int)ipow(c->edge_intp, c->dim); a.n_points = c->grid.n * a.block;
    for (int t = 0; t < c->dim; ++t) a.stride[t] = (int)ipow(c->edge_intp, c->dim - 1 - t);
    if (out_ptr[0] != 0 || out_ptr[n_out] != n_prog) return fail(AMDG_EINVAL, "out_ptr does not cover the program");
    for (int i = 0; i < n_const; ++i) a.consts[i] = consts[i];
    // validate: operand ranges and stack discipline (every output leaves exactly one value)
    for (int cidx = 0; cidx < n_out; ++cidx)
    {
        a.out_ptr[cidx] = out_ptr[cidx]; a.out_ptr[cidx + 1] = out_ptr[cidx + 1];
        int sp = 0;
        for (int i = out_ptr[cidx]; i < out_ptr[cidx + 1]; ++i)
        {
            const int op = prog[2 * i], arg = prog[2 * i + 1];
            a.op[i] = (short)op; a.arg[i] = (short)arg;
            switch (op)
            {
                case PW_VAR: if (arg < 0 || arg >= n_var) return fail(AMDG_EINVAL, "expression: variable index out of range"); ++sp; break;
                case PW_X: if (arg < 0 || arg >= c->dim) return fail(AMDG_EINVAL, "expression: coordinate index out of range");
                           if (!c->d_pts1d) return fail(AMDG_ESTATE, "expression uses point coordinates: call amdg_points_set first"); ++sp; break;
                case PW_OTHER: if (arg < 0 || arg >= n_other) return fail(AMDG_EINVAL, "expression: field index out of range"); ++sp; break;
                case PW_CONST: if (arg < 0 || arg >= n_const) return fail(AMDG_EINVAL, "expression: constant index out of range"); ++sp; break;
                case PW_ADD: case PW_SUB: case PW_MUL: case PW_DIV: case PW_POW: case PW_MIN: case PW_MAX:
                    if (sp < 2) return fail(AMDG_EINVAL, "expression: stack underflow"); --sp; break;
                case PW_NEG: case PW_SIN: case PW_COS: case PW_SQR: case PW_EXP: case PW_SQRT: case PW_ABS: case PW_TANH:
                    if (sp < 1) return fail(AMDG_EINVAL, "expression: stack underflow"); break;
                default: return fail(AMDG_EINVAL, "expression: unknown operation");
            }
            if (sp > PW_STACK) return fail(AMDG_EINVAL, "expression: stack too deep");
        }
        if (sp != 1) return fail(AMDG_EINVAL, "expression: an output must leave exactly one value");
    }
    CU(cudaSetDevice(c->device));
    cudaError_t e = launch_pointwise_expr(a, c->stream);
    if (e != cudaSuccess) return fail(AMDG_ECUDA, std::string("pointwise_expr launch: ") + cudaGetErrorString(e));
    c->launches++;
    return AMDG_OK;
}

// ---- multi-GPU plumbing (one process per GPU): peer-mapped device memory, rows to mapped destinations, device-side barrier ---------
int amdg_peer_export(amdg_ctx * c, const void * dev_ptr, void * handle64)
{
    int r = need_device(c); if (r) return r;
    if (!dev_ptr || !handle64) return fail(AMDG_EINVAL, "null argument");
    CU(cudaSetDevice(c->device));
    cudaIpcMemHandle_t h;
    static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
    CU(cudaIpcGetMemHandle(&h, const_cast<void *>(dev_ptr)));
    std::memcpy(handle64, &h, 64);
    return AMDG_OK;
}

int amdg_peer_open(amdg_ctx * c, const void * handle64, void ** dev_ptr_out)
{
    int r = need_device(c); if (r) return r;
    if (!handle64 || !dev_ptr_out) return fail(AMDG_EINVAL, "null argument");
    CU(cudaSetDevice(c->device));
    cudaIpcMemHandle_t h; std::memcpy(&h, handle64, 64);
    CU(cudaIpcOpenMemHandle(dev_ptr_out, h, cudaIpcMemLazyEnablePeerAccess));
    return AMDG_OK;
}

int amdg_peer_close(amdg_ctx * c, void * dev_ptr)
{
    int r = need_device(c); if (r) return r;
    CU(cudaSetDevice(c->device));
    CU(cudaIpcCloseMemHandle(dev_ptr));
    return AMDG_OK;
}

int amdg_peer_barrier(amdg_ctx * c, void * const * flag_ptrs, int world, int rank, void * dev_epoch, void * dev_error)
{
    int r = need_device(c); if (r) return r;
    if (!flag_ptrs || world < 1 || world > 16 || rank < 0 || rank >= world || !dev_epoch || !dev_error) return fail(AMDG_EINVAL, "bad barrier arguments");
    PeerBarrierArgs a; std::memset(&a, 0, sizeof(a));
    for (int i = 0; i < world; ++i) { if (!flag_ptrs[i]) return fail(AMDG_EINVAL, "null flag pointer"); a.flags_of[i] = (unsigned *)flag_ptrs[i]; }
    a.epoch = (unsigned *)dev_epoch; a.error = (unsigned *)dev_error; a.world = world; a.rank = rank; a.timeout_cycles = 6000000000ll;
    CU(cudaSetDevice(c->device));
    cudaError_t e = launch_peer_barrier(a, c->stream);
    if (e != cudaSuccess) return fail(AMDG_ECUDA, std::string("peer barrier launch: ") + cudaGetErrorString(e));
    c->launches++;
    return AMDG_OK;
}

int amdg_scatter_rows(amdg_ctx * c, const double * dev_src, int64_t n_rows, int width, double * dev_dst_base, const int64_t * dev_map)
{
    int r = need_device(c); if (r) return r;
    if (!dev_src || !dev_dst_base || !dev_map || n_rows < 0 || width < 1) return fail(AMDG_EINVAL, "bad arguments");
    if (n_rows == 0) return AMDG_OK;
    CU(cudaSetDevice(c->device));
    cudaError_t e = launch_scatter_rows(dev_src, n_rows, width, dev_dst_base, reinterpret_cast<const long long *>(dev_map), c->stream);
    if (e != cudaSuccess) return fail(AMDG_ECUDA, std::string("scatter_rows launch: ") + cudaGetErrorString(e));
    c->launches++;
    return AMDG_OK;
}

int amdg_lincomb(amdg_ctx * c, int64_t n, int k, const double * coefs, const double * const * dev_x, double beta, double * dev_y)
{
    int r = need_device(c); if (r) return r;
    if (n < 0 || k < 1 || k > 16 || !coefs || !dev_x || !dev_y) return fail(AMDG_EINVAL, "bad arguments");
    LincombArgs a; std::memset(&a, 0, sizeof(a));
    for (int i = 0; i < k; ++i) { if (!dev_x[i]) return fail(AMDG_EINVAL, "null input"); a.x[i] = dev_x[i]; a.c[i] = coefs[i]; }
    a.y = dev_y; a.beta = beta; a.n = n; a.k = k;
    CU(cudaSetDevice(c->device));
    cudaError_t e = launch_lincomb(a, c->stream);
    if (e != cudaSuccess) return fail(AMDG_ECUDA, std::string("lincomb: ") + cudaGetErrorString(e));
    c->launches++;
    return AMDG_OK;
}

int amdg_moment(amdg_ctx * c, int64_t n_field, const int * dev_map, int n_vdim, const int * order, double weight, const double * dev_f, double * dev_rhs_field)
{
    int r = need_device(c); if (r) return r;
    if (n_field < 0 || !dev_map || n_vdim < 1 || n_vdim >= c->dim || n_vdim > 4 || !order || !dev_f || !dev_rhs_field) return fail(AMDG_EINVAL, "bad moment arguments");
    const int a = c->edge_alpt;
    MomentArgs m; std::memset(&m, 0, sizeof(m));
    m.f = dev_f; m.rhs = dev_rhs_field; m.map = dev_map; m.n_field = n_field; m.weight = weight;
    m.x_block = (int)ipow(a, c->dim - n_vdim); m.v_block = (int)ipow(a, n_vdim);
    // combinations of velocity degrees: dv_i in 0..order_i; block index of dv = sum dv_i a^(n_vdim-1-i) (last dimension fastest)
    m.n_combo = 1;
    for (int i = 0; i < n_vdim; ++i)
    {
        if (order[i] < 0 || order[i] > 1 || (order[i] == 1 && a < 2)) return fail(AMDG_EINVAL, "moment order must be 0 or 1 (1 needs polynomial degree >= 1)");
        m.n_combo *= order[i] + 1;
    }
    for (int combo = 0; combo < m.n_combo; ++combo)
    {
        int rest = combo, off = 0; double cf = 1.0;
        for (int i = n_vdim - 1; i >= 0; --i)
        {
            const int dv = rest % (order[i] + 1); rest /= order[i] + 1;
            off += dv * (int)ipow(a, n_vdim - 1 - i);
            cf *= order[i] == 0 ? 1.0 : (dv == 0 ? 0.5 : 1.0 / (2.0 * std::sqrt(3.0)));
        }
        m.offset[combo] = off; m.coef[combo] = cf;
    }
    CU(cudaSetDevice(c->device));
    cudaError_t e = launch_moment(m, c->stream);
    if (e != cudaSuccess) return fail(AMDG_ECUDA, std::string("moment: ") + cudaGetErrorString(e));
    c->launches++;
    return AMDG_OK;
}

int amdg_indicator_norm(amdg_ctx * c, int n_var, const double * const * dev_u, double * dev_norm)
{
    int r = need_device(c); if (r) return r;
    if (!c->have_grid) return fail(AMDG_ESTATE, "no grid");
    if (n_var < 1 || n_var > 16 || !dev_u || !dev_norm) return fail(AMDG_EINVAL, "bad indicator arguments");
    IndicatorArgs a; std::memset(&a, 0, sizeof(a));
    for (int v = 0; v < n_var; ++v) { if (!dev_u[v]) return fail(AMDG_EINVAL, "null coefficient array"); a.u[v] = dev_u[v]; }
    a.norm = dev_norm; a.n_elem = c->grid.n; a.block = (int)ipow(c->edge_alpt, c->dim); a.n_var = n_var;
    CU(cudaSetDevice(c->device));
    cudaError_t e = launch_indicator_norm(a, c->stream);
    if (e != cudaSuccess) return fail(AMDG_ECUDA, std::string("indicator norm: ") + cudaGetErrorString(e));
    c->launches++;
    return AMDG_OK;
}

// ---- device memory helpers -------------------------------------------------------------------------------------------
int amdg_dev_alloc(amdg_ctx * c, int64_t n, double ** out)
{
    int r = need_device(c); if (r) return r;
    if (!out || n < 0) return fail(AMDG_EINVAL, "bad arguments");
    CU(cudaSetDevice(c->device));
    cudaError_t e = cudaMalloc((void **)out, std::max<int64_t>(n, 1) * sizeof(double));
    if (e != cudaSuccess) return fail(AMDG_ENOMEM, cudaGetErrorString(e));
    return AMDG_OK;
}
int amdg_dev_free(amdg_ctx * c, double * p) { int r = need_device(c); if (r) return r; CU(cudaSetDevice(c->device)); CU(cudaStreamSynchronize(c->stream)); CU(cudaFree(p)); return AMDG_OK; }
int amdg_dev_upload(amdg_ctx * c, double * dst, const double * src, int64_t n)
{
    int r = need_device(c); if (r) return r;
    CU(cudaSetDevice(c->device));
    CU(cudaMemcpyAsync(dst, src, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return AMDG_OK;
}
int amdg_dev_download(amdg_ctx * c, double * dst, const double * src, int64_t n)
{
    int r = need_device(c); if (r) return r;
    CU(cudaSetDevice(c->device));
    CU(cudaMemcpyAsync(dst, src, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return AMDG_OK;
}
int amdg_dev_zero(amdg_ctx * c, double * p, int64_t n)
{
    int r = need_device(c); if (r) return r;
    CU(cudaSetDevice(c->device));
    CU(cudaMemsetAsync(p, 0, (size_t)n * sizeof(double), c->stream));
    return AMDG_OK;
}

// ---- host-buffer entry points ------------------------------------------------------------------------------------------
static int ensure_stage(amdg_ctx * c, double ** buf, int64_t * cap, int64_t n)
{
    if (*cap >= n) return AMDG_OK;
    if (*buf) { CU(cudaStreamSynchronize(c->stream)); CU(cudaFree(*buf)); *buf = nullptr; *cap = 0; }
    cudaError_t e = cudaMalloc((void **)buf, (size_t)n * sizeof(double));
    if (e != cudaSuccess) return fail(AMDG_ENOMEM, cudaGetErrorString(e));
    *cap = n;
    return AMDG_OK;
}

int amdg_host_apply_tensor(amdg_ctx * c, const int * ops, const int * rels, const double * hsrc, double * hdst, int n_comp, double coef, int accumulate)
{
    int r = need_device(c); if (r) return r;
    if (!c->have_grid || !ops || !hsrc || !hdst) return fail(AMDG_EINVAL, "bad arguments");
    if ((r = check_op(c, ops[0]))) return r;
    const int64_t nf = c->grid.n * n_comp * ipow(c->ops[ops[0]]->kf, c->dim), nt = c->grid.n * n_comp * ipow(c->ops[ops[0]]->kt, c->dim);
    CU(cudaSetDevice(c->device));
    if ((r = ensure_stage(c, &c->h2d, &c->h2d_cap, nf)) || (r = ensure_stage(c, &c->d2h, &c->d2h_cap, nt))) return r;
    CU(cudaMemcpyAsync(c->h2d, hsrc, (size_t)nf * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    if (accumulate) CU(cudaMemcpyAsync(c->d2h, hdst, (size_t)nt * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    if ((r = amdg_apply_tensor(c, ops, rels, c->h2d, c->d2h, n_comp, coef, accumulate))) return r;
    CU(cudaMemcpyAsync(hdst, c->d2h, (size_t)nt * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return AMDG_OK;
}

int amdg_host_sweep1d(amdg_ctx * c, int op, int rel, int lu, int t, const int * sizes_from, const double * hsrc, double * hdst, int n_comp, double coef, int accumulate)
{
    int r = need_device(c); if (r) return r;
    if (!c->have_grid || !sizes_from || !hsrc || !hdst) return fail(AMDG_EINVAL, "bad arguments");
    if ((r = check_op(c, op))) return r;
    if (t < 0 || t >= c->dim) return fail(AMDG_EINVAL, "bad dim");
    int64_t bf = 1; for (int k = 0; k < c->dim; ++k) bf *= sizes_from[k];
    const int64_t bt = bf / sizes_from[t] * c->ops[op]->kt;
    const int64_t nf = c->grid.n * n_comp * bf, nt = c->grid.n * n_comp * bt;
    CU(cudaSetDevice(c->device));
    if ((r = ensure_stage(c, &c->h2d, &c->h2d_cap, nf)) || (r = ensure_stage(c, &c->d2h, &c->d2h_cap, nt))) return r;
    CU(cudaMemcpyAsync(c->h2d, hsrc, (size_t)nf * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    if (accumulate) CU(cudaMemcpyAsync(c->d2h, hdst, (size_t)nt * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    if ((r = amdg_sweep1d(c, op, rel, lu, t, sizes_from, c->h2d, c->d2h, n_comp, coef, accumulate))) return r;
    CU(cudaMemcpyAsync(hdst, c->d2h, (size_t)nt * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return AMDG_OK;
}

int amdg_host_hierarchize(amdg_ctx * c, int hop, const double * hsrc, double * hdst, int n_comp)
{
    int r = need_device(c); if (r) return r;
    if (!c->have_grid || !hsrc || !hdst) return fail(AMDG_EINVAL, "bad arguments");
    if ((r = check_op(c, hop))) return r;
    const int64_t nn = c->grid.n * n_comp * ipow(c->ops[hop]->kf, c->dim);
    CU(cudaSetDevice(c->device));
    if ((r = ensure_stage(c, &c->h2d, &c->h2d_cap, nn)) || (r = ensure_stage(c, &c->d2h, &c->d2h_cap, nn))) return r;
    CU(cudaMemcpyAsync(c->h2d, hsrc, (size_t)nn * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    if ((r = amdg_hierarchize(c, hop, c->h2d, c->d2h, n_comp))) return r;
    CU(cudaMemcpyAsync(hdst, c->d2h, (size_t)nn * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return AMDG_OK;
}

int amdg_host_roundtrip(amdg_ctx * c, int op_fwd, int hop, int op_inv, const double * hin, double * hout, int n_comp)
{
    int r = need_device(c); if (r) return r;
    if (!c->have_grid || !hin || !hout) return fail(AMDG_EINVAL, "bad arguments");
    if ((r = check_op(c, op_fwd)) || (r = check_op(c, hop)) || (r = check_op(c, op_inv))) return r;
    const int d = c->dim;
    const int64_t na = c->grid.n * n_comp * ipow(c->ops[op_fwd]->kf, d), nb = c->grid.n * n_comp * ipow(c->ops[op_fwd]->kt, d);
    CU(cudaSetDevice(c->device));
    if ((r = ensure_stage(c, &c->h2d, &c->h2d_cap, std::max(na, nb))) || (r = ensure_stage(c, &c->d2h, &c->d2h_cap, std::max(na, nb)))) return r;
    std::vector<int> ops(d, op_fwd), rels(d, AMDG_REL_VOL), ops2(d, op_inv);
    CU(cudaMemcpyAsync(c->h2d, hin, (size_t)na * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    if ((r = amdg_apply_tensor(c, ops.data(), rels.data(), c->h2d, c->d2h, n_comp, 1.0, 0))) return r;       // ucoe_alpt -> up_intp
    if ((r = amdg_hierarchize(c, hop, c->d2h, c->d2h, n_comp))) return r;                                       // up_intp -> ucoe_intp
    if ((r = amdg_apply_tensor(c, ops2.data(), rels.data(), c->d2h, c->h2d, n_comp, 1.0, 0))) return r;         // ucoe_intp -> ucoe_alpt
    CU(cudaMemcpyAsync(hout, c->h2d, (size_t)na * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return AMDG_OK;
}

}  // extern "C"
