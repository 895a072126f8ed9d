// C++ host mirror of the reference's class surface for the fast-transform path, on top of the C ABI
// (include/amdg.h).  Names, call order and argument meaning follow the reference so that an example's time loop
// reads the same; storage moves from the hash-keyed Element map to device-resident flat arrays.
//
//   reference class (file:line)                                    mirror here
//   DGSolution / DGAdapt          include/DGSolution.h:11, DGAdapt.h:6     amdg::DGSolution  (device arrays + grid tables)
//   OperatorMatrix1D<U,V>         include/OperatorMatrix1D.h:10            amdg::OperatorMatrix1D (handles of registered tables)
//   FastLagrIntp / FastHermIntp   include/FastMultiplyLU.h:190,224         amdg::FastLagrIntp, amdg::FastHermIntp
//   FastLagrInit / FastHermInit   include/FastMultiplyLU.h:344,320         amdg::FastLagrInit, amdg::FastHermInit
//   LagrInterpolation (fast part) include/Interpolation.h:36               amdg::LagrInterpolation
//   HyperbolicLagrRHS/HermRHS     include/FastMultiplyLU.h:571,597         amdg::HyperbolicLagrRHS (Hermite: same class, Hermite tables)
//   HyperbolicAlptRHS             include/FastMultiplyLU.h:619             amdg::HyperbolicAlptRHS
//   HyperbolicSameFlux{Lagr,Herm}RHS, HyperbolicDiffFlux{Lagr,Herm}RHS   include/FastMultiplyLU.h:477-570   same names
//   SourceFastLagr                include/FastMultiplyLU.h:632             amdg::SourceFastLagr
//   ForwardEuler/RK2SSP/RK2Midpoint/RK3SSP/RK3HeunLinear  include/ODESolver.h:115-230   amdg::ExplicitRK + the scheme classes
//   RK4ODE2nd                     include/ODESolver.h (source/ODESolver.cpp:543-615)   amdg::RK4ODE2nd
//
// Error convention: the reference prints and exit(1)s; the mirror throws amdg::Error carrying amdg_last_error().
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <cmath>
#include <vector>

#include "../../include/amdg.h"

namespace amdg {

struct Error : std::runtime_error { explicit Error(const std::string & m) : std::runtime_error(m) {} };
inline void check(int rc) { if (rc < 0) throw Error(std::string("amdg: ") + amdg_last_error()); }

// device array of doubles owned by a context
class DeviceArray
{
public:
    DeviceArray() {}
    DeviceArray(amdg_ctx * c, int64_t n) : ctx_(c), n_(n), cap_(n) { check(amdg_dev_alloc(c, n, &p_)); check(amdg_dev_zero(c, p_, n)); }
    DeviceArray(const DeviceArray &) = delete;
    DeviceArray & operator=(const DeviceArray &) = delete;
    DeviceArray(DeviceArray && o) noexcept { *this = std::move(o); }
    DeviceArray & operator=(DeviceArray && o) noexcept { release(); ctx_ = o.ctx_; p_ = o.p_; n_ = o.n_; cap_ = o.cap_; o.p_ = nullptr; o.n_ = 0; o.cap_ = 0; return *this; }
    ~DeviceArray() { release(); }
    double * data() const { return p_; }
    int64_t size() const { return n_; }
    void upload(const double * h) { check(amdg_dev_upload(ctx_, p_, h, n_)); }
    void download(double * h) const { check(amdg_dev_download(ctx_, h, p_, n_)); }
    void set_zero() { check(amdg_dev_zero(ctx_, p_, n_)); }
    // n zeroed doubles in place: the allocation is kept while it is large enough and grows geometrically, so an adaptive run whose grid changes
    // every step (DGAdapt::refine / coarsen) does not pay a cudaFree (which synchronises) and a cudaMalloc per array and step
    void resize(amdg_ctx * c, int64_t n)
    {
        if (c != ctx_ || n > cap_)
        {
            release();
            ctx_ = c; cap_ = n + n / 2 + 1024;
            check(amdg_dev_alloc(c, cap_, &p_));
        }
        n_ = n;
        check(amdg_dev_zero(ctx_, p_, n_));
    }
private:
    void release() { if (p_) amdg_dev_free(ctx_, p_); p_ = nullptr; cap_ = 0; }
    amdg_ctx * ctx_ = nullptr; double * p_ = nullptr; int64_t n_ = 0, cap_ = 0;
};

// The device-side stand-in of DGSolution: the statics Element::DIM / PMAX_alpt / PMAX_intp / VEC_NUM
// (include/Element.h:21-24) become constructor arguments; the element arrays of include/Element.h:76-124 become
// flat device arrays [vec][elem][block] (fp_intp / fucoe_intp: [vec][dim][elem][block]).
class DGSolution
{
public:
    DGSolution(int dim, int nmax, int pmax_alpt, int pmax_intp, int vec_num, int device = 0)
        : DIM(dim), NMAX(nmax), PMAX_alpt(pmax_alpt), PMAX_intp(pmax_intp), VEC_NUM(vec_num)
    { check(amdg_ctx_create(dim, nmax, pmax_alpt, pmax_intp, device, &ctx)); }
    ~DGSolution() { ucoe_alpt = DeviceArray(); up_intp = DeviceArray(); ucoe_intp = DeviceArray(); fp_intp = DeviceArray(); fucoe_intp = DeviceArray(); rhs = DeviceArray(); rk_u_tn = DeviceArray(); amdg_ctx_destroy(ctx); }
    DGSolution(const DGSolution &) = delete;

    // (re)build the index tables from the element list (call after construction / DGAdapt::refine / coarsen);
    // rows in the caller's order, e.g. the iteration order of the reference's DGSolution::dg
    void set_elements(int64_t n, const int * level, const int * suppt)
    {
        check(amdg_grid_set(ctx, n, level, suppt));
        n_elem = n;
        max_mesh = 0;                                                    // DGSolution::max_mesh_level (source/DGSolution.cpp): largest 1D level of any element
        for (int64_t i = 0; i < n * DIM; ++i) max_mesh = std::max(max_mesh, level[i]);
        ucoe_alpt.resize(ctx, VEC_NUM * n * size_alpt());
        rhs.resize(ctx, VEC_NUM * n * size_alpt());
        up_intp.resize(ctx, VEC_NUM * n * size_intp());
        ucoe_intp.resize(ctx, VEC_NUM * n * size_intp());
        fp_intp.resize(ctx, (int64_t)VEC_NUM * DIM * n * size_intp());
        fucoe_intp.resize(ctx, (int64_t)VEC_NUM * DIM * n * size_intp());
    }
    // DGSolution(sparse, level_init, ...) initial grid (source/DGSolution.cpp:10-57)
    void init_sparse_grid(int level_init, bool sparse = true)
    {
        const int64_t n = amdg_sparse_grid(DIM, level_init, sparse, nullptr, nullptr);
        std::vector<int> l(n * DIM), j(n * DIM);
        amdg_sparse_grid(DIM, level_init, sparse, l.data(), j.data());
        set_elements(n, l.data(), j.data());
    }
    int64_t size_alpt() const { int64_t s = 1; for (int d = 0; d < DIM; ++d) s *= PMAX_alpt + 1; return s; }
    int64_t size_intp() const { int64_t s = 1; for (int d = 0; d < DIM; ++d) s *= PMAX_intp + 1; return s; }
    int64_t size_basis_alpt() const { return n_elem * size_alpt(); }                       // source/DGSolution.cpp:841
    int64_t get_dof() const { return size_basis_alpt() * VEC_NUM; }                        // source/DGSolution.cpp:846 (all variables evolve)
    void set_rhs_zero() { rhs.set_zero(); }                                                  // source/DGSolution.cpp:1067
    double * ucoe(int v) const { return ucoe_alpt.data() + (int64_t)v * n_elem * size_alpt(); }
    double * rhs_v(int v) const { return rhs.data() + (int64_t)v * n_elem * size_alpt(); }
    double * up(int v) const { return up_intp.data() + (int64_t)v * n_elem * size_intp(); }
    double * ucoe_i(int v) const { return ucoe_intp.data() + (int64_t)v * n_elem * size_intp(); }
    double * fp(int v, int d) const { return fp_intp.data() + ((int64_t)v * DIM + d) * n_elem * size_intp(); }
    double * fucoe(int v, int d) const { return fucoe_intp.data() + ((int64_t)v * DIM + d) * n_elem * size_intp(); }

    const int DIM, NMAX, PMAX_alpt, PMAX_intp, VEC_NUM;
    amdg_ctx * ctx = nullptr;
    int64_t n_elem = 0;
    int max_mesh = 0;
    int max_mesh_level() const { return max_mesh; }
    DeviceArray ucoe_alpt, up_intp, ucoe_intp, fp_intp, fucoe_intp, rhs;
    DeviceArray rk_u_tn;                 // the u^n snapshot of ExplicitRK (ODESolver::ucoe_tn): lives with the solution, not with the short-lived solver objects
};

// Handles of the registered 1D tables of one (U,V) basis pair: the members the path uses
// (include/OperatorMatrix1D.h:24-71).  Built from the reference's dense tables (row-major [from][to]).
struct OperatorMatrix1D
{
    int u_v = -1, u_vx = -1, ulft_vjp = -1, urgt_vjp = -1, ujp_vjp = -1, uave2_vjp = -1;   // uave2 = ulft_vjp + urgt_vjp (source/FastMultiplyLU.cpp:1165)
    int edge_from = 0, edge_to = 0;
    OperatorMatrix1D() {}
    OperatorMatrix1D(DGSolution & dg, int edge_from_, int edge_to_, const double * t_u_v, const double * t_u_vx, const double * t_ulft_vjp,
                     const double * t_urgt_vjp, const double * t_ujp_vjp = nullptr) : edge_from(edge_from_), edge_to(edge_to_)
    {
        const int T = 1 << dg.NMAX, rows = T * edge_from, cols = T * edge_to;
        check(amdg_op_register(dg.ctx, t_u_v, rows, cols, edge_from, edge_to, &u_v));
        check(amdg_op_register(dg.ctx, t_u_vx, rows, cols, edge_from, edge_to, &u_vx));
        check(amdg_op_register(dg.ctx, t_ulft_vjp, rows, cols, edge_from, edge_to, &ulft_vjp));
        check(amdg_op_register(dg.ctx, t_urgt_vjp, rows, cols, edge_from, edge_to, &urgt_vjp));
        check(amdg_op_combine(dg.ctx, ulft_vjp, 1.0, urgt_vjp, 1.0, &uave2_vjp));
        if (t_ujp_vjp) check(amdg_op_register(dg.ctx, t_ujp_vjp, rows, cols, edge_from, edge_to, &ujp_vjp));
        // uave_vjp = (urgt_vjp + ulft_vjp) / 2 (include/OperatorMatrix1D.h:199)
        check(amdg_op_combine(dg.ctx, ulft_vjp, 0.5, urgt_vjp, 0.5, &uave_vjp));
    }
    // the same tables without any table of the reference: the library generates them in compact form (amdg_op_generate, csrc/tables.hpp).
    // basis = AMDG_BASIS_ALPERT / _LAGRANGE / _HERMITE of degree pmax (U, the row basis); V is the Alpert basis of the context.
    OperatorMatrix1D(DGSolution & dg, int basis, int pmax, int msh_case = 1) : edge_from(pmax + 1), edge_to(dg.PMAX_alpt + 1)
    {
        check(amdg_op_generate(dg.ctx, basis, pmax, msh_case, AMDG_TAB_U_V, &u_v));
        check(amdg_op_generate(dg.ctx, basis, pmax, msh_case, AMDG_TAB_U_VX, &u_vx));
        check(amdg_op_generate(dg.ctx, basis, pmax, msh_case, AMDG_TAB_ULFT_VJP, &ulft_vjp));
        check(amdg_op_generate(dg.ctx, basis, pmax, msh_case, AMDG_TAB_URGT_VJP, &urgt_vjp));
        check(amdg_op_generate(dg.ctx, basis, pmax, msh_case, AMDG_TAB_UJP_VJP, &ujp_vjp));
        check(amdg_op_generate(dg.ctx, basis, pmax, msh_case, AMDG_TAB_UJP_VXLFT, &ujp_vxlft));
        check(amdg_op_generate(dg.ctx, basis, pmax, msh_case, AMDG_TAB_UJP_VXRGT, &ujp_vxrgt));
        check(amdg_op_combine(dg.ctx, ulft_vjp, 1.0, urgt_vjp, 1.0, &uave2_vjp));
        check(amdg_op_combine(dg.ctx, ulft_vjp, 0.5, urgt_vjp, 0.5, &uave_vjp));
        check(amdg_op_combine(dg.ctx, ujp_vxlft, 1.0, ujp_vxrgt, 1.0, &ujp_vxave2));
        if (basis == AMDG_BASIS_ALPERT && pmax == dg.PMAX_alpt)
        {
            // the tables of the interior-penalty diffusion operator (include/OperatorMatrix1D.h:216-226), Alpert x Alpert only
            check(amdg_op_generate(dg.ctx, basis, pmax, msh_case, AMDG_TAB_UX_VX, &ux_vx));
            check(amdg_op_generate(dg.ctx, basis, pmax, msh_case, AMDG_TAB_UXAVE_VJP, &uxave_vjp));
            check(amdg_op_generate(dg.ctx, basis, pmax, msh_case, AMDG_TAB_UJP_VXAVE, &ujp_vxave));
        }
    }
    // the same three tables from the reference's dense ones
    void set_diffusion_tables(DGSolution & dg, const double * t_ux_vx, const double * t_uxave_vjp, const double * t_ujp_vxave)
    {
        const int T = 1 << dg.NMAX, rows = T * edge_from, cols = T * edge_to;
        check(amdg_op_register(dg.ctx, t_ux_vx, rows, cols, edge_from, edge_to, &ux_vx));
        check(amdg_op_register(dg.ctx, t_uxave_vjp, rows, cols, edge_from, edge_to, &uxave_vjp));
        check(amdg_op_register(dg.ctx, t_ujp_vxave, rows, cols, edge_from, edge_to, &ujp_vxave));
    }
    int ux_vx = -1, uxave_vjp = -1, ujp_vxave = -1;
    // the [u] * v_x^-, [u] * v_x^+ tables of DiffusionRHS (include/OperatorMatrix1D.h:209-214)
    void set_ujp_vx(DGSolution & dg, const double * t_ujp_vxlft, const double * t_ujp_vxrgt)
    {
        const int T = 1 << dg.NMAX, rows = T * edge_from, cols = T * edge_to;
        check(amdg_op_register(dg.ctx, t_ujp_vxlft, rows, cols, edge_from, edge_to, &ujp_vxlft));
        check(amdg_op_register(dg.ctx, t_ujp_vxrgt, rows, cols, edge_from, edge_to, &ujp_vxrgt));
        check(amdg_op_combine(dg.ctx, ujp_vxlft, 1.0, ujp_vxrgt, 1.0, &ujp_vxave2));
    }
    int uave_vjp = -1, ujp_vxlft = -1, ujp_vxrgt = -1, ujp_vxave2 = -1;   // ujp_vxave2 = ujp_vxlft + ujp_vxrgt (source/FastMultiplyLU.cpp:1742)
};

// FastLagrIntp / FastHermIntp (source/FastMultiplyLU.cpp:1316-1390): the constructor takes the point table in the
// reference's orientation Lag_pt_Alpt_1D[point][alpert] and transposes it like the reference does (:1350-1358).
class FastLagrIntp
{
public:
    FastLagrIntp(DGSolution & dg, const std::vector<std::vector<double>> & Lag_pt_Alpt_1D, const std::vector<std::vector<double>> & Lag_pt_Alpt_1D_d1)
        : dg_(&dg) { op_pt_ = reg(Lag_pt_Alpt_1D); if (!Lag_pt_Alpt_1D_d1.empty()) op_d1_ = reg(Lag_pt_Alpt_1D_d1); }
    // point tables generated by the library (amdg_op_generate_points); also installs the point coordinates (amdg_points_generate)
    FastLagrIntp(DGSolution & dg, int basis, int msh_case = 1) : dg_(&dg)
    {
        check(amdg_op_generate_points(dg.ctx, basis, dg.PMAX_intp, msh_case, 0, &op_pt_));
        if (basis == AMDG_BASIS_LAGRANGE) check(amdg_op_generate_points(dg.ctx, basis, dg.PMAX_intp, msh_case, 1, &op_d1_));
        check(amdg_points_generate(dg.ctx, basis, dg.PMAX_intp, msh_case, nullptr));
    }
    void eval_up_Lagr() { for (int v = 0; v < dg_->VEC_NUM; ++v) eval_up_Lagr(v); }
    // FastLagrIntp::eval_up_Lagr_coarse_grid (source/FastMultiplyLU.cpp:1367-1370): elements above the cut are skipped and left at zero
    void eval_up_Lagr_coarse_grid(int mesh_nmax)
    {
        const int d = dg_->DIM; std::vector<int> ops(d, op_pt_), rels(d, AMDG_REL_VOL);
        for (int v = 0; v < dg_->VEC_NUM; ++v)
            check(amdg_apply_tensor_coarse(dg_->ctx, ops.data(), rels.data(), dg_->ucoe(v), dg_->up(v), 1, 1.0, 0, mesh_nmax));
    }
    void eval_up_Lagr(int vec_index)
    {
        std::vector<int> ops(dg_->DIM, op_pt_), rels(dg_->DIM, AMDG_REL_VOL);
        check(amdg_apply_tensor(dg_->ctx, ops.data(), rels.data(), dg_->ucoe(vec_index), dg_->up(vec_index), 1, 1.0, 0));
    }
    void eval_der_up_Lagr(int d0)
    {
        std::vector<int> ops(dg_->DIM, op_pt_), rels(dg_->DIM, AMDG_REL_VOL); ops[d0] = op_d1_;
        for (int v = 0; v < dg_->VEC_NUM; ++v) check(amdg_apply_tensor(dg_->ctx, ops.data(), rels.data(), dg_->ucoe(v), dg_->up(v), 1, 1.0, 0));
    }
protected:
    int reg(const std::vector<std::vector<double>> & pt_alpt)
    {
        const size_t np = pt_alpt.size(), na = pt_alpt[0].size();
        std::vector<double> t(na * np);
        for (size_t r = 0; r < na; ++r) for (size_t c = 0; c < np; ++c) t[r * np + c] = pt_alpt[c][r];
        int op; check(amdg_op_register(dg_->ctx, t.data(), (int)na, (int)np, dg_->PMAX_alpt + 1, dg_->PMAX_intp + 1, &op)); return op;
    }
    DGSolution * dg_; int op_pt_ = -1, op_d1_ = -1;
};
class FastHermIntp : public FastLagrIntp
{
public:
    FastHermIntp(DGSolution & dg, const std::vector<std::vector<double>> & Her_pt_Alpt_1D) : FastLagrIntp(dg, Her_pt_Alpt_1D, {}) {}
    explicit FastHermIntp(DGSolution & dg) : FastLagrIntp(dg, AMDG_BASIS_HERMITE) {}     // Her_pt_Alpt_1D generated by the library
    void eval_up_Herm() { eval_up_Lagr(); }
};

// FastLagrInit / FastHermInit (source/FastMultiplyLU.cpp:1572-1625): ucoe_intp -> ucoe_alpt with the u_v table
class FastLagrInit
{
public:
    FastLagrInit(DGSolution & dg, const OperatorMatrix1D & matrix) : dg_(&dg), op_(matrix.u_v) {}
    void eval_ucoe_Alpt_Lagr()
    {
        std::vector<int> ops(dg_->DIM, op_), rels(dg_->DIM, AMDG_REL_VOL);
        check(amdg_apply_tensor(dg_->ctx, ops.data(), rels.data(), dg_->ucoe_intp.data(), dg_->ucoe_alpt.data(), dg_->VEC_NUM, 1.0, 0));
    }
    void eval_ucoe_Alpt_Herm() { eval_ucoe_Alpt_Lagr(); }
private:
    DGSolution * dg_; int op_;
};
typedef FastLagrInit FastHermInit;

// LagrInterpolation / HermInterpolation, fast wrappers only (source/Interplation.cpp:4033-4046, 4609-4621, 891-1048,
// 1222-1430).  The std::function flux of the reference becomes an enumerated flux id per dimension.
class LagrInterpolation
{
public:
    // pw_anc / pw_wt: the pwts stencils of every 1D element (include/Interpolation.h:5-11) in amdg_op_register_hier layout
    LagrInterpolation(DGSolution & dg, const int * pw_anc, const double * pw_wt) : dg_(&dg)
    { check(amdg_op_register_hier(dg.ctx, pw_anc, pw_wt, dg.PMAX_intp + 1, &op_hier_)); }
    // stencils generated by the library (amdg_op_generate_hier)
    explicit LagrInterpolation(DGSolution & dg, int msh_case = 1) : dg_(&dg)
    { check(amdg_op_generate_hier(dg.ctx, AMDG_BASIS_LAGRANGE, dg.PMAX_intp, msh_case, &op_hier_)); }
    // nonlinear_Lagr_fast: fastLagr.eval_up_Lagr(); eval_fp_Lag(func, is_intp); eval_fp_to_coe_D_Lag(is_intp)
    void nonlinear_Lagr_fast(const std::vector<int> & flux_id, const std::vector<double> & params, const std::vector<std::vector<bool>> & is_intp, FastLagrIntp & fastLagr)
    {
        fastLagr.eval_up_Lagr();
        for (int v = 0; v < dg_->VEC_NUM; ++v)
            for (int d = 0; d < dg_->DIM; ++d)
            {
                if (!is_intp[v][d]) continue;
                const double * prm = params.empty() ? nullptr : &params[(size_t)d * 4];
                check(amdg_pointwise(dg_->ctx, 1, &flux_id[d], prm, dg_->up(v), dg_->fp(v, d), nullptr));
                check(amdg_hierarchize(dg_->ctx, op_hier_, dg_->fp(v, d), dg_->fucoe(v, d), 1));
            }
    }
    // eval_up_to_coe_D_Lag: up_intp -> ucoe_intp
    void eval_up_to_coe_D_Lag() { check(amdg_hierarchize(dg_->ctx, op_hier_, dg_->up_intp.data(), dg_->ucoe_intp.data(), dg_->VEC_NUM)); }
    int hier_op() const { return op_hier_; }
private:
    DGSolution * dg_; int op_hier_ = -1;
};

// HermInterpolation, fast wrapper (source/Interplation.cpp:4609-4621): fastHerm.eval_up_Herm(); eval_fp_Her_2D(func, func_d1, func_d2, is_intp);
// eval_fp_to_coe_D_Her(is_intp).  DIM == 2, HermBasis::PMAX == 3, scalar, as in the reference; the flux and its derivatives are an enumerated id.
class HermInterpolation
{
public:
    HermInterpolation(DGSolution & dg, const int * pw_anc, const double * pw_wt) : dg_(&dg)
    { check(amdg_op_register_hier(dg.ctx, pw_anc, pw_wt, dg.PMAX_intp + 1, &op_hier_)); }
    explicit HermInterpolation(DGSolution & dg) : dg_(&dg)
    { check(amdg_op_generate_hier(dg.ctx, AMDG_BASIS_HERMITE, dg.PMAX_intp, 1, &op_hier_)); }
    void nonlinear_Herm_2D_fast(const std::vector<int> & flux_id, const std::vector<std::vector<bool>> & is_intp, FastLagrIntp & fastHerm, const std::vector<double> & params = {})
    {
        fastHerm.eval_up_Lagr();
        for (int d = 0; d < dg_->DIM; ++d)
        {
            if (!is_intp[0][d]) continue;
            const double * prm = params.empty() ? nullptr : &params[(size_t)d * 4];
            check(amdg_pointwise_hermite2d(dg_->ctx, 1, &flux_id[d], prm, dg_->up(0), dg_->fp(0, d)));
            check(amdg_hierarchize(dg_->ctx, op_hier_, dg_->fp(0, d), dg_->fucoe(0, d), 1));
        }
    }
    void eval_up_to_coe_D_Her() { check(amdg_hierarchize(dg_->ctx, op_hier_, dg_->up_intp.data(), dg_->ucoe_intp.data(), dg_->VEC_NUM)); }
private:
    DGSolution * dg_; int op_hier_ = -1;
};

// HyperbolicLagrRHS / HyperbolicHermRHS (source/FastMultiplyLU.cpp:1125-1267)
class HyperbolicLagrRHS
{
public:
    HyperbolicLagrRHS(DGSolution & dg, OperatorMatrix1D & oper) : dg_(&dg), m_(&oper) {}
    void rhs_vol_scalar()
    {
        const int d = dg_->DIM; std::vector<int> rels(d, AMDG_REL_VOL);
        for (int t = 0; t < d; ++t)
        {
            std::vector<int> ops(d, m_->u_v); ops[t] = m_->u_vx;
            check(amdg_apply_tensor(dg_->ctx, ops.data(), rels.data(), dg_->fucoe(0, t), dg_->rhs_v(0), 1, 1.0, 1));
        }
    }
    void rhs_flx_intp_scalar()
    {
        const int d = dg_->DIM;
        for (int t = 0; t < d; ++t)
        {
            std::vector<int> ops(d, m_->u_v), rels(d, AMDG_REL_VOL); ops[t] = m_->uave2_vjp; rels[t] = AMDG_REL_FLX;
            check(amdg_apply_tensor(dg_->ctx, ops.data(), rels.data(), dg_->fucoe(0, t), dg_->rhs_v(0), 1, 0.5, 1));
        }
    }
    // rhs_vol_scalar_coarse_grid / rhs_flx_intp_scalar_coarse_grid (source/FastMultiplyLU.cpp:1143-1158, 1191-1219)
    void rhs_vol_scalar_coarse_grid(int mesh_nmax)
    {
        const int d = dg_->DIM; std::vector<int> rels(d, AMDG_REL_VOL);
        for (int t = 0; t < d; ++t)
        {
            std::vector<int> ops(d, m_->u_v); ops[t] = m_->u_vx;
            check(amdg_apply_tensor_coarse(dg_->ctx, ops.data(), rels.data(), dg_->fucoe(0, t), dg_->rhs_v(0), 1, 1.0, 1, mesh_nmax));
        }
    }
    void rhs_flx_intp_scalar_coarse_grid(int mesh_nmax)
    {
        const int d = dg_->DIM;
        for (int t = 0; t < d; ++t)
        {
            std::vector<int> ops(d, m_->u_v), rels(d, AMDG_REL_VOL); ops[t] = m_->uave2_vjp; rels[t] = AMDG_REL_FLX;
            check(amdg_apply_tensor_coarse(dg_->ctx, ops.data(), rels.data(), dg_->fucoe(0, t), dg_->rhs_v(0), 1, 0.5, 1, mesh_nmax));
        }
    }
private:
    DGSolution * dg_; OperatorMatrix1D * m_;
};
typedef HyperbolicLagrRHS HyperbolicHermRHS;
// the DIM == 2 forms with one flux component per direction are the same compositions (rhs_2D(.., .., 0) and (.., .., 1),
// source/FastMultiplyLU.cpp:1053-1058, 1089-1123)
typedef HyperbolicLagrRHS HyperbolicDiffFluxLagrRHS;
typedef HyperbolicLagrRHS HyperbolicDiffFluxHermRHS;

// HyperbolicSameFluxHermRHS / HyperbolicSameFluxLagrRHS (source/FastMultiplyLU.cpp:970-1087): every direction uses the
// flux component fucoe_intp[0][0]; the per-direction overloads rhs_vol_scalar(dim) / rhs_flx_intp_scalar(dim) as well
class HyperbolicSameFluxLagrRHS
{
public:
    HyperbolicSameFluxLagrRHS(DGSolution & dg, OperatorMatrix1D & oper) : dg_(&dg), m_(&oper) {}
    void rhs_vol_scalar() { for (int t = 0; t < dg_->DIM; ++t) rhs_vol_scalar(t); }
    void rhs_vol_scalar(int t)
    {
        const int d = dg_->DIM; std::vector<int> rels(d, AMDG_REL_VOL), ops(d, m_->u_v); ops[t] = m_->u_vx;
        check(amdg_apply_tensor(dg_->ctx, ops.data(), rels.data(), dg_->fucoe(0, 0), dg_->rhs_v(0), 1, 1.0, 1));
    }
    void rhs_flx_intp_scalar() { for (int t = 0; t < dg_->DIM; ++t) rhs_flx_intp_scalar(t); }
    void rhs_flx_intp_scalar(int t)
    {
        const int d = dg_->DIM; std::vector<int> ops(d, m_->u_v), rels(d, AMDG_REL_VOL); ops[t] = m_->uave2_vjp; rels[t] = AMDG_REL_FLX;
        check(amdg_apply_tensor(dg_->ctx, ops.data(), rels.data(), dg_->fucoe(0, 0), dg_->rhs_v(0), 1, 0.5, 1));
    }
private:
    DGSolution * dg_; OperatorMatrix1D * m_;
};
typedef HyperbolicSameFluxLagrRHS HyperbolicSameFluxHermRHS;

// SourceFastLagr::rhs_source (source/FastMultiplyLU.cpp:1304-1314): rhs[v] += (u_v x ... x u_v) fucoe_intp[v][0]
class SourceFastLagr
{
public:
    SourceFastLagr(DGSolution & dg, OperatorMatrix1D & oper) : dg_(&dg), m_(&oper) {}
    void rhs_source()
    {
        const int d = dg_->DIM; std::vector<int> rels(d, AMDG_REL_VOL), ops(d, m_->u_v);
        for (int v = 0; v < dg_->VEC_NUM; ++v)
            check(amdg_apply_tensor(dg_->ctx, ops.data(), rels.data(), dg_->fucoe(v, 0), dg_->rhs_v(v), 1, 1.0, 1));
    }
private:
    DGSolution * dg_; OperatorMatrix1D * m_;
};

// DiffusionRHS (source/FastMultiplyLU.cpp:1691-1821): the interior-penalty terms of a nonlinear diffusion flux interpolated in the Lagrange basis.
// rhs_vol and rhs_flx_gradu read one flux component per direction (fucoe_intp[0][t]); the three [u]-terms read fucoe_intp[0][0].
class DiffusionRHS
{
public:
    DiffusionRHS(DGSolution & dg, OperatorMatrix1D & oper) : dg_(&dg), m_(&oper) {}
    void rhs_vol() { for (int t = 0; t < dg_->DIM; ++t) term(m_->u_vx, AMDG_REL_VOL, t, t, -1.0); }
    void rhs_flx_gradu() { for (int t = 0; t < dg_->DIM; ++t) term(m_->uave_vjp, AMDG_REL_FLX, t, t, -1.0); }
    void rhs_flx_u() { for (int t = 0; t < dg_->DIM; ++t) term(m_->ujp_vxave2, AMDG_REL_FLX, t, 0, -0.5); }
    void rhs_flx_k_minus_u() { for (int t = 0; t < dg_->DIM; ++t) term(m_->ujp_vxlft, AMDG_REL_FLX, t, 0, -0.5); }
    void rhs_flx_k_plus_u() { for (int t = 0; t < dg_->DIM; ++t) term(m_->ujp_vxrgt, AMDG_REL_FLX, t, 0, -0.5); }
private:
    void term(int op_t, int rel_t, int t, int dim_interp, double coef)
    {
        if (op_t < 0) throw Error("DiffusionRHS: table not registered (OperatorMatrix1D::set_ujp_vx)");
        const int d = dg_->DIM; std::vector<int> ops(d, m_->u_v), rels(d, AMDG_REL_VOL); ops[t] = op_t; rels[t] = rel_t;
        check(amdg_apply_tensor(dg_->ctx, ops.data(), rels.data(), dg_->fucoe(0, dim_interp), dg_->rhs_v(0), 1, coef, 1));
    }
    DGSolution * dg_; OperatorMatrix1D * m_;
};

// FastRHSHamiltonJacobi::rhs_nonlinear (source/FastMultiplyLU.cpp:426-434): rhs += (u_v x ... x u_v) fucoe_intp[0][0]
class FastRHSHamiltonJacobi
{
public:
    FastRHSHamiltonJacobi(DGSolution & dg, OperatorMatrix1D & oper) : dg_(&dg), m_(&oper) {}
    void rhs_nonlinear()
    {
        const int d = dg_->DIM; std::vector<int> rels(d, AMDG_REL_VOL), ops(d, m_->u_v);
        check(amdg_apply_tensor(dg_->ctx, ops.data(), rels.data(), dg_->fucoe(0, 0), dg_->rhs_v(0), 1, 1.0, 1));
    }
private:
    DGSolution * dg_; OperatorMatrix1D * m_;
};

// HyperbolicAlptRHS::rhs_flx_penalty_scalar (source/FastMultiplyLU.cpp:1269-1276)
class HyperbolicAlptRHS
{
public:
    HyperbolicAlptRHS(DGSolution & dg, OperatorMatrix1D & oper_alpt) : dg_(&dg), m_(&oper_alpt) {}
    void rhs_flx_penalty_scalar(const std::vector<double> & lax_alpha)
    {
        const int d = dg_->DIM;
        if (d == 1) return;     // the reference's single-matrix form never sweeps for DIM == 1 (source/FastMultiplyLU.cpp:206-212, 249)
        std::vector<int> sizes(d, dg_->PMAX_alpt + 1);
        for (int t = 0; t < d; ++t)
            check(amdg_sweep1d(dg_->ctx, m_->ujp_vjp, AMDG_REL_FLX, AMDG_LU_FULL, t, sizes.data(), dg_->ucoe(0), dg_->rhs_v(0), 1, -lax_alpha[t] / 2., 1));
    }
private:
    DGSolution * dg_; OperatorMatrix1D * m_;
};

// BilinearFormAlpt / HyperbolicAlpt / DiffusionAlpt (include/BilinearForm.h; source/BilinearForm.cpp:25-87, 691-779, 877-929).  The reference assembles an
// Eigen sparse matrix from the 1D tables (assemble_matrix_alpt: the 1D operator along `dim`, the mass matrix elsewhere, "vol" or "flx" relation) and
// ExplicitRK::add_rhs_matrix / step_rk multiply it.  Here every assemble call adds its scaled 1D table to ONE pre-merged operator per dimension
// (amdg_op_combine; a volume table has zero blocks on the pairs that are only flux-related, so the merged operator runs under the flx relation, and
// the Alpert mass matrix is the identity) and the product is one full sweep per dimension -- no matrix is ever assembled.
class BilinearFormAlpt
{
public:
    BilinearFormAlpt(DGSolution & dg, OperatorMatrix1D & oper_alpt) : dg_(&dg), m_(&oper_alpt), merged_(dg.DIM, -1) {}
    // mat += coef * (table along dim (x) identity elsewhere)
    void assemble_matrix_alpt(double coef, int dim, int table)
    {
        if (table < 0) throw Error("BilinearFormAlpt: table not registered");
        int out = -1;
        if (merged_[dim] < 0) check(amdg_op_combine(dg_->ctx, table, coef, table, 0.0, &out));
        else check(amdg_op_combine(dg_->ctx, merged_[dim], 1.0, table, coef, &out));
        merged_[dim] = out;
    }
    // rhs += mat * ucoe_alpt (ODESolver::add_rhs_matrix, source/ODESolver.cpp:129-137)
    void add_to_rhs(int vec_index = 0) const
    {
        std::vector<int> sizes(dg_->DIM, dg_->PMAX_alpt + 1);
        for (int t = 0; t < dg_->DIM; ++t)
            if (merged_[t] >= 0)
                check(amdg_sweep1d(dg_->ctx, merged_[t], AMDG_REL_FLX, AMDG_LU_FULL, t, sizes.data(), dg_->ucoe(vec_index), dg_->rhs_v(vec_index), 1, 1.0, 1));
    }
    DGSolution & solution() const { return *dg_; }
protected:
    DGSolution * dg_; OperatorMatrix1D * m_; std::vector<int> merged_;
};

class HyperbolicAlpt : public BilinearFormAlpt
{
public:
    using BilinearFormAlpt::BilinearFormAlpt;
    // u_t + sum_d c_d u_{x_d} = 0 with upwind fluxes (source/BilinearForm.cpp:691-710)
    void assemble_matrix_scalar(const std::vector<double> & eqnCoefficient)
    {
        for (int dim = 0; dim < dg_->DIM; ++dim)
        {
            assemble_matrix_alpt(eqnCoefficient[dim], dim, m_->u_vx);
            assemble_matrix_alpt(eqnCoefficient[dim], dim, eqnCoefficient[dim] >= 0 ? m_->ulft_vjp : m_->urgt_vjp);
        }
    }
    // one-sided flux term (source/BilinearForm.cpp:760-769): sign -1 takes the left limit, +1 the right limit
    void assemble_matrix_flx_scalar(int dim, int sign, double coefficient = 1.) { assemble_matrix_alpt(coefficient, dim, sign == -1 ? m_->ulft_vjp : m_->urgt_vjp); }
};

class DiffusionAlpt : public BilinearFormAlpt
{
public:
    DiffusionAlpt(DGSolution & dg, OperatorMatrix1D & oper_alpt, double sigma_ipdg_) : BilinearFormAlpt(dg, oper_alpt), sigma_ipdg(sigma_ipdg_) {}
    // interior-penalty Laplacian (source/BilinearForm.cpp:877-929): -(u_x, v_x) - {u_x}[v] - [u]{v_x} - sigma / dx [u][v], dx = 2^-max_mesh_level
    void assemble_matrix_scalar(const std::vector<double> & eqnCoefficient)
    {
        const double dx = 1. / std::pow(2., dg_->max_mesh_level());
        for (int dim = 0; dim < dg_->DIM; ++dim)
        {
            assemble_matrix_alpt(-eqnCoefficient[dim], dim, m_->ux_vx);
            assemble_matrix_alpt(-eqnCoefficient[dim], dim, m_->uxave_vjp);
            assemble_matrix_alpt(-eqnCoefficient[dim], dim, m_->ujp_vxave);
            assemble_matrix_alpt(-sigma_ipdg / dx, dim, m_->ujp_vjp);
        }
    }
    const double sigma_ipdg;
};

// ExplicitRK (include/ODESolver.h:76-112).  The reference packs Element arrays into Eigen vectors
// (source/ODESolver.cpp:25-127); here ucoe_alpt / rhs already are the flat vectors, so init() only snapshots u_tn
// and add_rhs_to_eigenvec()/final() are no-ops kept for call-order compatibility.
class ExplicitRK
{
public:
    ExplicitRK(DGSolution & dg, double dt_, int scheme, int stages) : num_stage(stages), dt(dt_), dg_(&dg), scheme_(scheme) { dg.rk_u_tn.resize(dg.ctx, dg.get_dof()); }
    // the reference's ODESolver(BilinearForm &) constructor (include/ODESolver.h:14): the solver keeps the linear operator for step_rk
    ExplicitRK(BilinearFormAlpt & linear, double dt_, int scheme, int stages) : ExplicitRK(linear.solution(), dt_, scheme, stages) { linear_ = &linear; }
    // rhs += mat * ucoe (ODESolver::add_rhs_matrix): here d sweeps with the operator's pre-merged 1D tables
    void add_rhs_matrix(const BilinearFormAlpt & linear) { linear.add_to_rhs(); }
    // one whole step of a linear problem (ExplicitRK::step_rk, e.g. RK3SSP::step_rk source/ODESolver.cpp:261-271): every stage evaluates rhs = mat * u
    void step_rk()
    {
        if (!linear_) throw Error("step_rk needs the solver to be constructed from a linear operator");
        check(amdg_axpby(dg_->ctx, dg_->get_dof(), 1.0, dg_->ucoe_alpt.data(), 0.0, dg_->rk_u_tn.data()));
        for (int stage = 0; stage < num_stage; ++stage) { dg_->set_rhs_zero(); linear_->add_to_rhs(); step_stage(stage); }
    }
    virtual ~ExplicitRK() {}
    virtual void init() { check(amdg_axpby(dg_->ctx, dg_->get_dof(), 1.0, dg_->ucoe_alpt.data(), 0.0, dg_->rk_u_tn.data())); }
    void set_rhs_zero() {}
    void add_rhs_to_eigenvec() {}
    virtual void step_stage(int stage) { check(amdg_rk_stage(dg_->ctx, scheme_, stage, dt, dg_->rk_u_tn.data(), dg_->ucoe_alpt.data(), dg_->rhs.data(), dg_->get_dof())); }
    virtual void final() {}
    const int num_stage;
    const double dt;
protected:
    DGSolution * dg_; int scheme_; BilinearFormAlpt * linear_ = nullptr;
};
#define AMDG_RK_SCHEME(NAME, ID, STAGES) struct NAME : ExplicitRK { NAME(DGSolution & dg, double dt) : ExplicitRK(dg, dt, ID, STAGES) {} \
                                                                    NAME(BilinearFormAlpt & linear, double dt) : ExplicitRK(linear, dt, ID, STAGES) {} };
AMDG_RK_SCHEME(ForwardEuler, AMDG_RK_EULER, 1)
AMDG_RK_SCHEME(RK2SSP, AMDG_RK_RK2SSP, 2)
AMDG_RK_SCHEME(RK2Midpoint, AMDG_RK_RK2MID, 2)
AMDG_RK_SCHEME(RK3SSP, AMDG_RK_RK3SSP, 3)
AMDG_RK_SCHEME(RK3HeunLinear, AMDG_RK_RK3HEUN, 3)
#undef AMDG_RK_SCHEME

// RK4ODE2nd (include/ODESolver.h, source/ODESolver.cpp:543-615): u_tt = L u as the pair (ucoe_alpt, ucoe_ut); the caller
// evaluates rhs = L u before every step_stage, exactly as with the reference's stage interface.
class RK4ODE2nd
{
public:
    RK4ODE2nd(DGSolution & dg, DeviceArray & ucoe_ut, double dt_) : num_stage(4), dt(dt_), dg_(&dg), v_(&ucoe_ut), u_tn_(dg.ctx, dg.get_dof()),
        v_tn_(dg.ctx, dg.get_dof()), ku_(dg.ctx, 4 * dg.get_dof()), kv_(dg.ctx, 4 * dg.get_dof()) {}
    void init()
    {
        check(amdg_axpby(dg_->ctx, dg_->get_dof(), 1.0, dg_->ucoe_alpt.data(), 0.0, u_tn_.data()));
        check(amdg_axpby(dg_->ctx, dg_->get_dof(), 1.0, v_->data(), 0.0, v_tn_.data()));
    }
    void step_stage(int stage)
    {
        check(amdg_rk4_ode2nd_stage(dg_->ctx, stage, dt, u_tn_.data(), v_tn_.data(), dg_->ucoe_alpt.data(), v_->data(), dg_->rhs.data(),
                                    ku_.data(), kv_.data(), dg_->get_dof()));
    }
    void final() {}
    const int num_stage;
    const double dt;
private:
    DGSolution * dg_; DeviceArray * v_; DeviceArray u_tn_, v_tn_, ku_, kv_;
};

}  // namespace amdg
