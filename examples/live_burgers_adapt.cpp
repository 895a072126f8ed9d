// Live drop-in test: the reference's own adaptive Burgers example (example/02_hyperbolic_05_burgers_adapt.cpp:223-470, run as
// `-imex 0 -v 0`: Hermite flux interpolation, Lax-Friedrichs penalty, Euler predictor -> DGAdapt::refine -> RK3SSP -> DGAdapt::coarsen)
// executed twice in lockstep:
//   reference arm: the UNMODIFIED reference classes from oracle/_ref/libsgdg_ref.a (DGAdapt, HermInterpolation, FastHermIntp,
//                  HyperbolicSameFluxHermRHS, HyperbolicAlpt assembled penalty matrix, ForwardEuler / RK3SSP), stock code path;
//   device arm:    the same DGAdapt class keeps the hash-keyed element map and does refine()/coarsen() on the host, every fast class is
//                  replaced by the mirror of adaptive-multiresolution-dg_b200/host/amdg_host.hpp over the C ABI (libamdg_b200.so).  The glue
//                  (a)-(d) of INTEGRATION.md is compiled here for real: element list in DGSolution::dg order -> amdg_grid_set after every
//                  grid change, the live OperatorMatrix1D / Her_pt_Alpt_1D tables and the pwts stencils -> operator handles, coefficient
//                  upload / download around the host-side adaptivity.
// After every time step the two element sets must be identical and the coefficients agree to 1e-10 (relative L2); at the end the L2 error
// against the exact Burgers solution is evaluated by the reference's own error routine on both coefficient sets.
// Built only where the reference sources exist (examples/Makefile.live, run by __graft_entry__.build()); the binary travels to the GPU box.
//
// Usage: live_burgers_adapt [-NM 6] [-N0 2] [-steps 10] [-r 1e-2] [-timing 1] [-static 1 -tf 0.05]   (static: sparse grid of level N0 = NM, no adaptivity -- convergence study)
#include <iostream>
#include <iomanip>
#include <fstream>
#include <sstream>
#include <algorithm>
#include <map>
#include <vector>
#include <string>
#include <cstring>
#include <cstdint>
#include <cmath>
#include <numeric>
#include <functional>
#include <iterator>
#include <unordered_map>
#include <unordered_set>
#include <random>
#include <array>
#include <cassert>
#include <chrono>
#include <set>
#include <omp.h>
#include <Eigen/Eigen>

// the hierarchisation stencils are computed by a protected member (HermInterpolation::set_pts_wts_1d_ada_Her, source/Interplation.cpp:3166-3315);
// the glue reaches it without editing a reference header
#define private public
#define protected public
#include "DGAdaptIntp.h"
#include "Interpolation.h"
#include "FastMultiplyLU.h"
#include "ODESolver.h"
#include "OperatorMatrix1D.h"
#include "BilinearForm.h"
#include "ExactSolution.h"
#undef private
#undef protected

#include "../adaptive-multiresolution-dg_b200/host/amdg_host.hpp"

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// ---- glue (a): element list in DGSolution::dg iteration order
static void export_elements(DGSolution & dg, amdg::DGSolution & dev)
{
    std::vector<int> level, suppt;
    for (auto & it : dg.dg) for (int d = 0; d < DGSolution::DIM; ++d) { level.push_back(it.second.level[d]); suppt.push_back(it.second.suppt[d]); }
    dev.set_elements((int64_t)dg.dg.size(), level.data(), suppt.data());
}
// ---- glue (b): dense tables, VecMultiD<double> is row-major mat.at(from, to)
static std::vector<double> dense(const VecMultiD<double> & m) { std::vector<double> v(m.size()); for (int i = 0; i < m.size(); ++i) v[i] = m.at(i); return v; }
// ---- glue (d): coefficients in the same order
static void upload_ucoe(DGSolution & dg, amdg::DGSolution & dev)
{
    std::vector<double> h; h.reserve(dev.get_dof());
    for (auto & it : dg.dg) for (int i = 0; i < it.second.ucoe_alpt[0].size(); ++i) h.push_back(it.second.ucoe_alpt[0].at(i));
    dev.ucoe_alpt.upload(h.data());
}
static void download_ucoe(DGSolution & dg, amdg::DGSolution & dev)
{
    std::vector<double> h(dev.get_dof());
    dev.ucoe_alpt.download(h.data());
    size_t p = 0;
    for (auto & it : dg.dg) for (int i = 0; i < it.second.ucoe_alpt[0].size(); ++i) it.second.ucoe_alpt[0].at(i) = h[p++];
}

int main(int argc, char ** argv)
{
    int NMAX = 6, N_init = 2, n_steps = 10, timing = 0;
    bool is_static = false;
    bool trace = false;              // -trace 1: host issue time and device drain time of the three RK3 stages, separately
    bool gen_tables = false;         // -gen 1: the device arm takes no table from the reference objects, the library generates them (amdg_op_generate*)
    double refine_eps = 1e-2, final_time = -1.;
    for (int i = 1; i + 1 < argc; i += 2)
    {
        std::string k = argv[i];
        if (k == "-NM") NMAX = std::atoi(argv[i + 1]); else if (k == "-N0") N_init = std::atoi(argv[i + 1]);
        else if (k == "-steps") n_steps = std::atoi(argv[i + 1]); else if (k == "-r") refine_eps = std::atof(argv[i + 1]);
        else if (k == "-timing") timing = std::atoi(argv[i + 1]);
        else if (k == "-trace") trace = std::atoi(argv[i + 1]) != 0;
        else if (k == "-gen") gen_tables = std::atoi(argv[i + 1]) != 0;
        else if (k == "-static") is_static = std::atoi(argv[i + 1]) != 0;      // no predictor / refine / coarsen: the sparse grid of level N0 for the whole run (convergence study)
        else if (k == "-tf") final_time = std::atof(argv[i + 1]);              // run to this time (the last step is shortened), overrides -steps
    }
    const double coarsen_eta = refine_eps / 10.;
    // ---- statics exactly as example/02_hyperbolic_05_burgers_adapt.cpp:34-63
    const int DIM = 2;
    AlptBasis::PMAX = 2;
    LagrBasis::PMAX = 3; LagrBasis::msh_case = 1;
    HermBasis::PMAX = 3; HermBasis::msh_case = 1;
    Element::PMAX_alpt = AlptBasis::PMAX; Element::PMAX_intp = HermBasis::PMAX;
    Element::DIM = DIM; Element::VEC_NUM = 1;
    DGSolution::DIM = DIM; DGSolution::VEC_NUM = 1;
    Interpolation::DIM = DIM; Interpolation::VEC_NUM = 1;
    DGSolution::ind_var_vec = { 0 };
    DGAdapt::indicator_var_adapt = { 0 };
    Element::is_intp.resize(1); Element::is_intp[0] = std::vector<bool>(DIM, true);
    const std::string boundary_type = "period";
    const double cfl_hyper = 0.2;

    Hash hash;
    LagrBasis::set_interp_msh01();
    HermBasis::set_interp_msh01();
    AllBasis<LagrBasis> all_bas_lagr(NMAX);
    AllBasis<HermBasis> all_bas_herm(NMAX);
    AllBasis<AlptBasis> all_bas_alpt(NMAX);
    OperatorMatrix1D<AlptBasis, AlptBasis> oper_matx_alpt(all_bas_alpt, all_bas_alpt, boundary_type);
    OperatorMatrix1D<HermBasis, AlptBasis> oper_matx_herm(all_bas_herm, all_bas_alpt, boundary_type);

    auto init_func_1 = [](double x, int d) { return (d == 0) ? (sin(2. * Const::PI * x)) : (cos(2. * Const::PI * x)); };
    auto init_func_2 = [](double x, int d) { return (d == 0) ? (cos(2. * Const::PI * x)) : (sin(2. * Const::PI * x)); };
    std::vector<std::function<double(double, int)>> init_func{ init_func_1, init_func_2 };
    auto func_flux = [&](std::vector<double> u, int i, int d) -> double { return FluxFunction::burgers_flux_scalar(u[0]); };
    auto func_flux_d1 = [&](std::vector<double> u, int i, int d, int i1) -> double { return FluxFunction::burgers_flux_1st_derivative_scalar(u[0]); };
    auto func_flux_d2 = [&](std::vector<double> u, int i, int d, int i1, int i2) -> double { return FluxFunction::burgers_flux_2nd_derivative_scalar(u[0]); };
    const std::vector<double> lxf_alpha{ 1.2, 1.2 };
    const std::vector<double> wave_speed{ 1., 1. };
    std::vector<std::vector<bool>> is_intp_herm; is_intp_herm.push_back(std::vector<bool>{ true, false });

    // ---- reference arm
    DGAdapt dg_ref(true, N_init, NMAX, all_bas_alpt, all_bas_lagr, all_bas_herm, hash, refine_eps, coarsen_eta, true, true);
    dg_ref.init_separable_scalar_sum(init_func);
    HyperbolicSameFluxHermRHS fast_rhs_herm(dg_ref, oper_matx_herm);
    HermInterpolation interp_herm(dg_ref);
    FastHermIntp fast_herm_intp(dg_ref, interp_herm.Her_pt_Alpt_1D);

    // ---- device arm: the same host-side DGAdapt, fast classes from the mirror
    DGAdapt dg_dev(true, N_init, NMAX, all_bas_alpt, all_bas_lagr, all_bas_herm, hash, refine_eps, coarsen_eta, true, true);
    dg_dev.init_separable_scalar_sum(init_func);
    try
    {
        amdg::DGSolution dev(DIM, NMAX, AlptBasis::PMAX, HermBasis::PMAX, 1, 0);
        export_elements(dg_dev, dev);
        amdg::OperatorMatrix1D d_oper_herm = gen_tables ? amdg::OperatorMatrix1D(dev, AMDG_BASIS_HERMITE, HermBasis::PMAX)
            : amdg::OperatorMatrix1D(dev, HermBasis::PMAX + 1, AlptBasis::PMAX + 1, dense(oper_matx_herm.u_v).data(), dense(oper_matx_herm.u_vx).data(),
                                     dense(oper_matx_herm.ulft_vjp).data(), dense(oper_matx_herm.urgt_vjp).data());
        amdg::OperatorMatrix1D d_oper_alpt = gen_tables ? amdg::OperatorMatrix1D(dev, AMDG_BASIS_ALPERT, AlptBasis::PMAX)
            : amdg::OperatorMatrix1D(dev, AlptBasis::PMAX + 1, AlptBasis::PMAX + 1, dense(oper_matx_alpt.u_v).data(), dense(oper_matx_alpt.u_vx).data(),
                                     dense(oper_matx_alpt.ulft_vjp).data(), dense(oper_matx_alpt.urgt_vjp).data(), dense(oper_matx_alpt.ujp_vjp).data());
        // glue (c): pwts stencils of every 1D element with level > 0 (include/Interpolation.h:5-11); not needed with -gen 1
        std::vector<int> anc; std::vector<double> wt;
        if (!gen_tables)
        {
            HermInterpolation tmp(dg_dev);
            for (int n = 1; n <= NMAX; ++n)
                for (int j = 1; j < std::max(2, pow_int(2, n)); j += 2)
                {
                    tmp.pw1d.clear(); tmp.set_pts_wts_1d_ada_Her(n, j);
                    const pwts & p = tmp.pw1d.begin()->second;
                    for (size_t ic = 0; ic < p.p_k.size(); ++ic) { anc.push_back(tmp.hash_key1d(p.p_k[ic], p.p_i[ic])); anc.push_back(p.p_num[ic]); }
                    for (size_t p0 = 0; p0 < p.wt.size(); ++p0) for (size_t ic = 0; ic < p.wt[p0].size(); ++ic) wt.push_back(p.wt[p0][ic]);
                }
        }
        amdg::HermInterpolation d_interp_herm = gen_tables ? amdg::HermInterpolation(dev) : amdg::HermInterpolation(dev, anc.data(), wt.data());
        std::vector<std::vector<double>> her_pt(interp_herm.Her_pt_Alpt_1D);
        amdg::FastHermIntp d_fast_herm_intp = gen_tables ? amdg::FastHermIntp(dev) : amdg::FastHermIntp(dev, her_pt);
        amdg::HyperbolicSameFluxHermRHS d_fast_rhs_herm(dev, d_oper_herm);
        amdg::HyperbolicAlptRHS d_fast_rhs_alpt(dev, d_oper_alpt);
        upload_ucoe(dg_dev, dev);
        const std::vector<int> flux_id(DIM, AMDG_FLUX_BURGERS);

        double curr_time = 0., worst = 0., t_ref = 0., t_dev = 0., t_grid = 0., t_copy = 0., t_adapt = 0., first_ref = 0., first_dev = 0.;
        int n_skipped = 0;
        int n_done = 0;
        for (int step = 0; final_time > 0. ? curr_time < final_time * (1. - 1e-14) : step < n_steps; ++step)
        {
            // ---- part 1: dt (both arms must agree)
            auto dt_of = [&](DGAdapt & dg) { const std::vector<int> & mm = dg.max_mesh_level_vec(); double s = 0.; for (int d = 0; d < DIM; ++d) s += std::abs(wave_speed[d]) * std::pow(2., mm[d]); return cfl_hyper / s; };
            double dt = dt_of(dg_ref);
            if (dt != dt_of(dg_dev)) { std::printf("FAIL: the arms disagree on dt at step %d\n", step); return 1; }

            if (final_time > 0. && curr_time + dt > final_time) dt = final_time - curr_time;
            n_done = step + 1;

            // =============================== reference arm (stock code of the example)
            double t0 = now();
            if (!is_static)
            {
                dg_ref.copy_ucoe_to_predict();
                HyperbolicAlpt linear(dg_ref, oper_matx_alpt);
                linear.assemble_matrix_flx_scalar(0, -1, lxf_alpha[0] / 2); linear.assemble_matrix_flx_scalar(0, 1, -lxf_alpha[0] / 2);
                linear.assemble_matrix_flx_scalar(1, -1, lxf_alpha[1] / 2); linear.assemble_matrix_flx_scalar(1, 1, -lxf_alpha[1] / 2);
                ForwardEuler odeSolver(linear, dt);
                odeSolver.init();
                interp_herm.nonlinear_Herm_2D_fast(func_flux, func_flux_d1, func_flux_d2, is_intp_herm, fast_herm_intp);
                dg_ref.set_rhs_zero();
                fast_rhs_herm.rhs_vol_scalar(); fast_rhs_herm.rhs_flx_intp_scalar();
                odeSolver.set_rhs_zero(); odeSolver.add_rhs_to_eigenvec(); odeSolver.add_rhs_matrix(linear);
                odeSolver.step_stage(0); odeSolver.final();
            }
            if (!is_static) { dg_ref.refine(); dg_ref.copy_predict_to_ucoe(); }
            {
                HyperbolicAlpt linear(dg_ref, oper_matx_alpt);
                linear.assemble_matrix_flx_scalar(0, -1, lxf_alpha[0] / 2); linear.assemble_matrix_flx_scalar(0, 1, -lxf_alpha[0] / 2);
                linear.assemble_matrix_flx_scalar(1, -1, lxf_alpha[1] / 2); linear.assemble_matrix_flx_scalar(1, 1, -lxf_alpha[1] / 2);
                RK3SSP odeSolver(linear, dt);
                odeSolver.init();
                for (int stage = 0; stage < odeSolver.num_stage; ++stage)
                {
                    interp_herm.nonlinear_Herm_2D_fast(func_flux, func_flux_d1, func_flux_d2, is_intp_herm, fast_herm_intp);
                    dg_ref.set_rhs_zero();
                    fast_rhs_herm.rhs_vol_scalar(); fast_rhs_herm.rhs_flx_intp_scalar();
                    odeSolver.set_rhs_zero(); odeSolver.add_rhs_to_eigenvec(); odeSolver.add_rhs_matrix(linear);
                    odeSolver.step_stage(stage); odeSolver.final();
                }
            }
            if (!is_static) dg_ref.coarsen();
            t_ref += now() - t0;

            // =============================== device arm (mirror classes; DGAdapt on the host)
            t0 = now();
            double t1 = now();
            if (!is_static)
            {
            const double tp0 = now();
            dg_dev.copy_ucoe_to_predict();
            const double tp1 = now();
            {
                amdg::ForwardEuler odeSolver(dev, dt);
                odeSolver.init();
                d_interp_herm.nonlinear_Herm_2D_fast(flux_id, is_intp_herm, d_fast_herm_intp);
                dev.set_rhs_zero();
                d_fast_rhs_herm.rhs_vol_scalar(); d_fast_rhs_herm.rhs_flx_intp_scalar();
                d_fast_rhs_alpt.rhs_flx_penalty_scalar(lxf_alpha);             // = add_rhs_matrix(linear): the assembled Lax-Friedrichs jump terms as two sweeps
                odeSolver.set_rhs_zero(); odeSolver.add_rhs_to_eigenvec();
                odeSolver.step_stage(0); odeSolver.final();
            }
            if (trace)
            {
                const double tp2 = now();
                amdg_ctx_sync(dev.ctx);
                std::printf("trace step %d: predictor: copy_ucoe_to_predict %.3f ms, issue %.3f ms, drain %.3f ms more\n", step, 1e3 * (tp1 - tp0), 1e3 * (tp2 - tp1), 1e3 * (now() - tp2));
            }
            t1 = now();
            download_ucoe(dg_dev, dev);                                         // DGAdapt::refine reads the predicted coefficients on the host
            t_copy += now() - t1; t1 = now();
            dg_dev.refine();
            dg_dev.copy_predict_to_ucoe();
            t_adapt += now() - t1; t1 = now();
            export_elements(dg_dev, dev);                                       // amdg_grid_set + device arrays for the new element list
            t_grid += now() - t1; t1 = now();
            upload_ucoe(dg_dev, dev);
            t_copy += now() - t1;
            }
            {
                if (trace) amdg_ctx_sync(dev.ctx);
                const double ti0 = now();
                const int64_t l0 = amdg_ctx_launch_count(dev.ctx);
                amdg::RK3SSP odeSolver(dev, dt);
                odeSolver.init();
                for (int stage = 0; stage < odeSolver.num_stage; ++stage)
                {
                    d_interp_herm.nonlinear_Herm_2D_fast(flux_id, is_intp_herm, d_fast_herm_intp);
                    dev.set_rhs_zero();
                    d_fast_rhs_herm.rhs_vol_scalar(); d_fast_rhs_herm.rhs_flx_intp_scalar();
                    d_fast_rhs_alpt.rhs_flx_penalty_scalar(lxf_alpha);
                    odeSolver.set_rhs_zero(); odeSolver.add_rhs_to_eigenvec();
                    odeSolver.step_stage(stage); odeSolver.final();
                }
                if (trace)
                {
                    const double ti1 = now();
                    amdg_ctx_sync(dev.ctx);
                    std::printf("trace step %d: three RK3 stages, %lld launches: issue %.3f ms, drain %.3f ms more\n", step,
                                (long long)(amdg_ctx_launch_count(dev.ctx) - l0), 1e3 * (ti1 - ti0), 1e3 * (now() - ti1));
                }
            }
            t1 = now();
            download_ucoe(dg_dev, dev);
            t_copy += now() - t1; t1 = now();
            if (!is_static)
            {
                dg_dev.coarsen();
                t_adapt += now() - t1; t1 = now();
                export_elements(dg_dev, dev);
                t_grid += now() - t1; t1 = now();
                upload_ucoe(dg_dev, dev);
                t_copy += now() - t1;
            }
            t_dev += now() - t0;
            if (step == 0 && (final_time > 0. || n_steps > 1))
            {
                // the first step carries one-time costs of the process (CUDA module load of the library, work lists of the first, not yet adaptive,
                // grid): reported on its own, the per-step averages start at the second step
                first_ref = t_ref; first_dev = t_dev;
                t_ref = t_dev = t_grid = t_copy = t_adapt = 0.;
                n_skipped = 1;
            }

            // =============================== compare
            if (dg_ref.dg.size() != dg_dev.dg.size()) { std::printf("FAIL: element counts differ at step %d: %zu vs %zu\n", step, dg_ref.dg.size(), dg_dev.dg.size()); return 1; }
            double num = 0., den = 0.;
            for (auto & it : dg_ref.dg)
            {
                auto jt = dg_dev.dg.find(it.first);
                if (jt == dg_dev.dg.end()) { std::printf("FAIL: element %d of the reference grid is missing in the device arm at step %d\n", it.first, step); return 1; }
                for (int i = 0; i < it.second.ucoe_alpt[0].size(); ++i)
                {
                    const double a = it.second.ucoe_alpt[0].at(i), b = jt->second.ucoe_alpt[0].at(i);
                    num += (a - b) * (a - b); den += a * a;
                }
            }
            const double e = std::sqrt(num / den);
            worst = std::max(worst, e);
            curr_time += dt;
            std::printf("step %2d  dt %.3e  elements %5zu  DoF %6d  rel-L2(device, reference) %.3e\n", step, dt, dg_ref.dg.size(), dg_ref.size_basis_alpt(), e);
            if (trace)
            {
                // longest fibre of the current grid per dimension (what decides the list-free kernel in the context's adaptive mode)
                for (int t = 0; t < DIM; ++t)
                {
                    const int64_t nf = amdg_grid_fibres(dev.ctx, t, nullptr, nullptr);
                    std::vector<int64_t> ptr(nf + 1); std::vector<int> el(dev.n_elem);
                    amdg_grid_fibres(dev.ctx, t, ptr.data(), el.data());
                    int64_t longest = 0; for (int64_t f = 0; f < nf; ++f) longest = std::max(longest, ptr[f + 1] - ptr[f]);
                    std::printf("trace step %d: dimension %d: %lld fibres, longest %lld elements\n", step, t, (long long)nf, (long long)longest);
                }
            }
        }
        // ---- L2 error against the exact solution (the reference's own routine on both coefficient sets)
        BurgersExact burgers(0., 1., 0.);
        auto final_func = [&](std::vector<double> x) -> double { return burgers.exact_2d(x[0], x[1], curr_time); };
        const int num_gauss_pt = 3;
        std::vector<double> err_ref = dg_ref.get_error_no_separable_scalar(final_func, num_gauss_pt);
        std::vector<double> err_dev = dg_dev.get_error_no_separable_scalar(final_func, num_gauss_pt);
        std::printf("L1 / L2 / Linf error vs exact Burgers at t = %.4f: reference %.6e %.6e %.6e | device %.6e %.6e %.6e\n", curr_time,
                    err_ref[0], err_ref[1], err_ref[2], err_dev[0], err_dev[1], err_dev[2]);
        const int n_avg = std::max(1, n_done - n_skipped);
        std::printf("wall per step: reference arm %.2f ms (host cores: %d) | device arm %.2f ms = kernels+launch %.2f + amdg_grid_set/realloc %.2f + coefficient copies %.2f + DGAdapt refine/coarsen %.2f\n",
                    1e3 * t_ref / n_avg, omp_get_max_threads(), 1e3 * t_dev / n_avg, 1e3 * (t_dev - t_grid - t_copy - t_adapt) / n_avg, 1e3 * t_grid / n_avg,
                    1e3 * t_copy / n_avg, 1e3 * t_adapt / n_avg);
        if (n_skipped) std::printf("(averages over steps 1..%d; the first step, with the one-time costs of the process: reference arm %.2f ms, device arm %.2f ms)\n", n_done - 1, 1e3 * first_ref, 1e3 * first_dev);
        std::printf("LIVE OK worst rel-L2 %.3e over %d %s steps (NMAX %d)\n", worst, n_done, is_static ? "static-grid" : "adaptive", NMAX);
    }
    catch (const std::exception & e) { std::cerr << e.what() << std::endl; return 1; }
    return 0;
}
