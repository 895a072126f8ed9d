// Full-run validation and convergence table: the reference's linear advection example (example/02_hyperbolic_01_scalar_const_coefficient.cpp:
// u_t + u_x + u_y = 0, u(x,0) = cos 2 pi (x + y), periodic, Alpert k = 2, full sparse grid of level N, RK3SSP, dt = cfl dx / DIM with cfl = 0.1,
// run to t = 0.1) executed twice for every N:
//   reference arm: the UNMODIFIED reference classes (oracle/_ref/libsgdg_ref.a): HyperbolicAlpt::assemble_matrix_scalar + RK3SSP::step_rk on the
//                  assembled sparse matrix (:154-166), the shipped path;
//   device arm:    amdg::HyperbolicAlpt + amdg::RK3SSP::step_rk of the mirror (library-generated tables): per stage d sweeps with u_vx + ulft_vjp (upwind flux, c >= 0:
//                  source/BilinearForm.cpp:700-703) merged into one operator, then ExplicitRK::step_stage on the device.
// After the last step the coefficients must agree to 1e-10 (relative L2; north_star's full-run bound) and the L2 errors against the exact solution
// (the reference's own error routine, DGSolution::get_error_no_separable_scalar, on both coefficient sets) must coincide; the table of errors and
// observed orders is printed -- the reference's order for k = 2 on sparse grids is about k + 1/2 .. k + 1 (example/02_hyperbolic_01...cpp:190-214).
// Built only where the reference sources exist (examples/Makefile.live); the binary travels to the GPU box.
//
// Usage: live_advection_convergence [-Nmin 3] [-Nmax 7] [-tf 0.1]
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

#include "DGAdaptIntp.h"
#include "ODESolver.h"
#include "OperatorMatrix1D.h"
#include "BilinearForm.h"

#include "../adaptive-multiresolution-dg_b200/host/amdg_host.hpp"

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
static std::vector<double> dense(const VecMultiD<double> & m) { std::vector<double> v(m.size()); for (int i = 0; i < m.size(); ++i) v[i] = m.at(i); return v; }

int main(int argc, char ** argv)
{
    int Nmin = 3, Nmax = 7;
    double final_time = 0.1;
    for (int i = 1; i + 1 < argc; i += 2)
    {
        std::string k = argv[i];
        if (k == "-Nmin") Nmin = std::atoi(argv[i + 1]); else if (k == "-Nmax") Nmax = std::atoi(argv[i + 1]); else if (k == "-tf") final_time = std::atof(argv[i + 1]);
    }
    const int DIM = 2;
    AlptBasis::PMAX = 2;
    LagrBasis::PMAX = 3; LagrBasis::msh_case = 1;
    HermBasis::PMAX = 3; HermBasis::msh_case = 1;
    Element::PMAX_alpt = AlptBasis::PMAX; Element::PMAX_intp = LagrBasis::PMAX;
    Element::DIM = DIM; Element::VEC_NUM = 1;
    DGSolution::DIM = DIM; DGSolution::VEC_NUM = 1;
    Interpolation::DIM = DIM; Interpolation::VEC_NUM = 1;
    DGSolution::ind_var_vec = { 0 };
    DGAdapt::indicator_var_adapt = { 0 };
    Element::is_intp.resize(1); Element::is_intp[0] = std::vector<bool>(DIM, true);
    const double cfl = 0.1;
    LagrBasis::set_interp_msh01();
    HermBasis::set_interp_msh01();
    // cos 2 pi (x + y) = cos cos - sin sin
    auto f1 = [](double x, int d) { return cos(2. * Const::PI * x); };
    auto f2 = [](double x, int d) { return (d == 0) ? (-sin(2. * Const::PI * x)) : (sin(2. * Const::PI * x)); };
    std::vector<std::function<double(double, int)>> init_func{ f1, f2 };
    const std::vector<double> c(DIM, 1.);
    std::vector<double> e_ref_all, e_dev_all; std::vector<int> Ns;
    bool ok = true;
    try
    {
        for (int N = Nmin; N <= Nmax; ++N)
        {
            Hash hash;
            AllBasis<AlptBasis> all_bas_alpt(N);
            AllBasis<LagrBasis> all_bas_lagr(N);
            AllBasis<HermBasis> all_bas_herm(N);
            OperatorMatrix1D<AlptBasis, AlptBasis> oper(all_bas_alpt, all_bas_alpt, "period");
            DGAdapt dg_ref(true, N, N, all_bas_alpt, all_bas_lagr, all_bas_herm, hash, 1e10, -1., true, false);
            dg_ref.init_separable_scalar_sum(init_func);
            DGAdapt dg_dev(true, N, N, all_bas_alpt, all_bas_lagr, all_bas_herm, hash, 1e10, -1., true, false);
            dg_dev.init_separable_scalar_sum(init_func);
            // time step of the example (:147-151)
            const double dx = 1. / std::pow(2., dg_ref.max_mesh_level());
            double dt = dx * cfl / DIM;
            const int n_steps = (int)std::ceil(final_time / dt) + 1;
            dt = final_time / n_steps;

            // ---- reference arm
            double t0 = now();
            HyperbolicAlpt op(dg_ref, oper);
            op.assemble_matrix_scalar(c);
            {
                RK3SSP ode(op, dt);
                ode.init();
                for (int s = 0; s < n_steps; ++s) ode.step_rk();
                ode.final();
            }
            const double t_ref = now() - t0;

            // ---- device arm
            t0 = now();
            amdg::DGSolution dev(DIM, N, AlptBasis::PMAX, LagrBasis::PMAX, 1, 0);
            {
                std::vector<int> level, suppt;
                for (auto & it : dg_dev.dg) for (int d = 0; d < DIM; ++d) { level.push_back(it.second.level[d]); suppt.push_back(it.second.suppt[d]); }
                dev.set_elements((int64_t)dg_dev.dg.size(), level.data(), suppt.data());
            }
            // the device arm reads like the stock example: tables, the linear operator, RK3SSP(operator, dt), step_rk -- only the namespace differs.
            // The tables are generated by the library (no table of the reference is read); the operator is one pre-merged 1D table per dimension
            amdg::OperatorMatrix1D d_oper(dev, AMDG_BASIS_ALPERT, AlptBasis::PMAX);
            amdg::HyperbolicAlpt d_op(dev, d_oper);
            d_op.assemble_matrix_scalar(c);
            {
                std::vector<double> h; h.reserve(dev.get_dof());
                for (auto & it : dg_dev.dg) for (int i = 0; i < it.second.ucoe_alpt[0].size(); ++i) h.push_back(it.second.ucoe_alpt[0].at(i));
                dev.ucoe_alpt.upload(h.data());
            }
            {
                amdg::RK3SSP ode(d_op, dt);
                ode.init();
                for (int s = 0; s < n_steps; ++s) ode.step_rk();
                ode.final();
            }
            {
                std::vector<double> h(dev.get_dof());
                dev.ucoe_alpt.download(h.data());
                size_t p = 0;
                for (auto & it : dg_dev.dg) for (int i = 0; i < it.second.ucoe_alpt[0].size(); ++i) it.second.ucoe_alpt[0].at(i) = h[p++];
            }
            const double t_dev = now() - t0;

            // ---- compare coefficients (full-run bound 1e-10) and errors against the exact solution
            double num = 0., den = 0.;
            for (auto & it : dg_ref.dg)
            {
                auto jt = dg_dev.dg.find(it.first);
                for (int i = 0; i < it.second.ucoe_alpt[0].size(); ++i)
                {
                    const double x = it.second.ucoe_alpt[0].at(i), y = jt->second.ucoe_alpt[0].at(i);
                    num += (x - y) * (x - y); den += x * x;
                }
            }
            const double rel = std::sqrt(num / std::max(den, 1e-300));
            auto exact = [&](std::vector<double> x) -> double { return cos(2. * Const::PI * (x[0] + x[1] - DIM * final_time)); };
            const std::vector<double> er = dg_ref.get_error_no_separable_scalar(exact, 4), ed = dg_dev.get_error_no_separable_scalar(exact, 4);
            std::printf("N %d  elements %5zu  DoF %6d  steps %4d  rel-L2(device, reference) %.3e  L2 error: reference %.6e device %.6e  wall: reference %.2f s, device %.2f s\n",
                        N, dg_ref.dg.size(), dg_ref.size_basis_alpt(), n_steps, rel, er[1], ed[1], t_ref, t_dev);
            Ns.push_back(N); e_ref_all.push_back(er[1]); e_dev_all.push_back(ed[1]);
            if (!(rel < 1e-10) || std::abs(er[1] - ed[1]) > 1e-10 * std::max(1., er[1])) ok = false;
        }
        std::printf("\n| N | L2 error (reference) | order | L2 error (device) | order |\n|---|---|---|---|---|\n");
        for (size_t i = 0; i < Ns.size(); ++i)
        {
            if (i == 0) std::printf("| %d | %.4e | - | %.4e | - |\n", Ns[i], e_ref_all[i], e_dev_all[i]);
            else std::printf("| %d | %.4e | %.2f | %.4e | %.2f |\n", Ns[i], e_ref_all[i], std::log2(e_ref_all[i - 1] / e_ref_all[i]), e_dev_all[i], std::log2(e_dev_all[i - 1] / e_dev_all[i]));
        }
        std::printf(ok ? "CONVERGENCE OK\n" : "CONVERGENCE FAIL\n");
    }
    catch (const std::exception & e) { std::cerr << e.what() << std::endl; return 1; }
    return ok ? 0 : 1;
}
