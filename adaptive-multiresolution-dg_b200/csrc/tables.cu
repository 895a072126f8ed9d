// C ABI of the table generator (tables.hpp): the library's own OperatorMatrix1D / point tables / hierarchisation stencils, registered in compact
// form through the public entry points of capi.cu.  Host code only; kept in its own translation unit so that it builds in seconds.
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>

#include "../../include/amdg.h"
#include "tables.hpp"

using namespace amdg;

extern "C" int amdg_internal_fail(int code, const char * msg);     // capi.cu: records the message amdg_last_error() returns
static int fail(int code, const std::string & msg) { return amdg_internal_fail(code, msg.c_str()); }

// the canonical pair enumeration depends on nmax only: one copy per nmax for all contexts
static const Pairs1D & pairs_of(int nmax)
{
    static std::mutex mu; static std::map<int, std::unique_ptr<Pairs1D>> cache;
    std::lock_guard<std::mutex> lk(mu);
    auto & p = cache[nmax];
    if (!p) { p.reset(new Pairs1D()); p->build(nmax); }
    return *p;
}
#define CTX_INFO() int info_[5]; { int r_ = amdg_ctx_info(c, info_); if (r_) return r_; } const int nmax = info_[1], pmax_alpt = info_[2], pmax_intp = info_[3], device = info_[4]; \
    (void)pmax_intp; (void)device; const Pairs1D & pairs = pairs_of(nmax)

// ---- tables generated on our side, straight into the compact form (tables.hpp) ---------------------------------------------------
static int make_intp_basis(int basis, int pmax, int msh_case, std::unique_ptr<tab::IntpBasis> & out)
{
    if (basis == AMDG_BASIS_LAGRANGE)
    {
        std::unique_ptr<tab::LagrangeBasis> b(new tab::LagrangeBasis(pmax, msh_case));
        if (!b->ok) return fail(AMDG_EINVAL, "no Lagrange point set for this (pmax, mesh case) (source/LagrBasis.cpp:33-154)");
        out = std::move(b); return AMDG_OK;
    }
    if (basis == AMDG_BASIS_HERMITE)
    {
        std::unique_ptr<tab::HermiteBasis> b(new tab::HermiteBasis(pmax));
        if (!b->ok) return fail(AMDG_EINVAL, "Hermite interpolation needs pmax 3 or 5 (source/HermBasis.cpp:41-69)");
        out = std::move(b); return AMDG_OK;
    }
    return fail(AMDG_EINVAL, "basis must be AMDG_BASIS_LAGRANGE or AMDG_BASIS_HERMITE");
}

extern "C" {

static int op_generate_impl(amdg_ctx * c, int basis_u, int pmax_u, int msh_case_u, int table, int boundary, int * out);
int amdg_op_generate(amdg_ctx * c, int basis_u, int pmax_u, int msh_case_u, int table, int * out) { return op_generate_impl(c, basis_u, pmax_u, msh_case_u, table, AMDG_BC_PERIOD, out); }
int amdg_op_generate_bc(amdg_ctx * c, int basis_u, int pmax_u, int msh_case_u, int table, int boundary, int * out) { return op_generate_impl(c, basis_u, pmax_u, msh_case_u, table, boundary, out); }

}  // extern "C"

static int op_generate_impl(amdg_ctx * c, int basis_u, int pmax_u, int msh_case_u, int table, int boundary, int * out)
{
    if (!c || !out) return fail(AMDG_EINVAL, "null argument");
    CTX_INFO();
    if (table < 0 || table >= tab::N_TABLE) return fail(AMDG_EINVAL, "unknown table");
    if (boundary < 0 || boundary > 2) return fail(AMDG_EINVAL, "boundary must be AMDG_BC_PERIOD, _ZERO or _INSIDE");
    if (pmax_u < 0 || pmax_u > 5) return fail(AMDG_EINVAL, "pmax must be in 0..5");
    const tab::AlpertBasis V(pmax_alpt);
    std::unique_ptr<tab::Basis1D> U;
    if (basis_u == AMDG_BASIS_ALPERT) U.reset(new tab::AlpertBasis(pmax_u));
    else
    {
        if (table >= tab::UX_VX) return fail(AMDG_EINVAL, "tables with derivatives of u exist for Alpert x Alpert only (include/OperatorMatrix1D.h:216)");
        std::unique_ptr<tab::IntpBasis> I; int r = make_intp_basis(basis_u, pmax_u, msh_case_u, I); if (r) return r;
        U = std::move(I);
    }
    if ((table == tab::UX_V || table == tab::UJP_VXAVE) && (basis_u != AMDG_BASIS_ALPERT || pmax_u != pmax_alpt)) return fail(AMDG_EINVAL, "transposed tables need U == V");
    std::vector<double> blocks;
    tab::operator_blocks(pairs, *U, V, table, blocks, boundary);
    return amdg_op_register_compact(c, blocks.data(), pairs.n_pairs, pmax_u + 1, (pmax_alpt + 1), 0, out);
}

extern "C" {

int amdg_op_generate_points(amdg_ctx * c, int basis, int pmax, int msh_case, int derivative, int * out)
{
    if (!c || !out) return fail(AMDG_EINVAL, "null argument");
    CTX_INFO();
    if (derivative < 0 || derivative > 1 || (basis == AMDG_BASIS_HERMITE && derivative != 0)) return fail(AMDG_EINVAL, "bad derivative order");
    std::unique_ptr<tab::IntpBasis> I; int r = make_intp_basis(basis, pmax, msh_case, I); if (r) return r;
    const tab::AlpertBasis A(pmax_alpt);
    std::vector<double> blocks;
    tab::point_blocks(pairs, A, *I, derivative, blocks);
    return amdg_op_register_compact(c, blocks.data(), pairs.n_pairs, (pmax_alpt + 1), pmax + 1, 0, out);
}

int amdg_op_generate_hier(amdg_ctx * c, int basis, int pmax, int msh_case, int * out)
{
    if (!c || !out) return fail(AMDG_EINVAL, "null argument");
    CTX_INFO(); (void)pairs; (void)pmax_alpt;
    std::unique_ptr<tab::IntpBasis> I; int r = make_intp_basis(basis, pmax, msh_case, I); if (r) return r;
    std::vector<int> anc; std::vector<double> wt;
    if (!tab::hier_stencils(nmax, *I, basis == AMDG_BASIS_HERMITE, anc, wt)) return fail(AMDG_EINVAL, "the point set is not hierarchical");
    return amdg_op_register_hier(c, anc.data(), wt.data(), pmax + 1, out);
}

int amdg_points_generate(amdg_ctx * c, int basis, int pmax, int msh_case, double * host_pts1d)
{
    if (!c) return fail(AMDG_EINVAL, "null context");
    CTX_INFO(); (void)pairs; (void)pmax_alpt;
    std::unique_ptr<tab::IntpBasis> I; int r = make_intp_basis(basis, pmax, msh_case, I); if (r) return r;
    std::vector<double> pts;
    tab::point_table(nmax, *I, pts);
    if (host_pts1d) std::memcpy(host_pts1d, pts.data(), pts.size() * sizeof(double));
    if (device >= 0 && pmax == pmax_intp) return amdg_points_set(c, pts.data());
    return AMDG_OK;
}

// the grid of a field solution with auxiliary dimensions (grid.hpp: aux_grid); level == NULL returns the count
int64_t amdg_aux_grid(int dim, int level_init, int aux_dim, int * level, int * suppt)
{
    if (dim < 1 || dim > 8 || level_init < 0 || level_init > 12 || aux_dim < 0 || aux_dim >= dim) return fail(AMDG_EINVAL, "bad grid parameters");
    if (!level) return aux_grid(dim, level_init, aux_dim, nullptr, nullptr);
    std::vector<int> l, j;
    const int64_t n = aux_grid(dim, level_init, aux_dim, &l, &j);
    std::memcpy(level, l.data(), l.size() * sizeof(int)); std::memcpy(suppt, j.data(), j.size() * sizeof(int));
    return n;
}

}  // extern "C"
