// One RK3SSP step of 2D Burgers through the C++ host mirror (amdg_host.hpp), written like the reference's time
// loop (example/02_hyperbolic_05_burgers_adapt.cpp:223-470, explicit branch) and checked against a dump of the
// compiled reference (tests/golden/*.dump, format of oracle/ref_harness.cpp).  Usage: burgers_stage <dump file>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include "../adaptive-multiresolution-dg_b200/host/amdg_host.hpp"

struct Arr { char dtype; std::vector<int64_t> dims; std::vector<char> raw;
    const double * d() const { return reinterpret_cast<const double *>(raw.data()); }
    const int * i() const { return reinterpret_cast<const int *>(raw.data()); } };

static std::map<std::string, Arr> load_dump(const char * path)
{
    std::ifstream f(path, std::ios::binary); std::map<std::string, Arr> out;
    char magic[8]; f.read(magic, 8);
    while (f.peek() != EOF)
    {
        uint32_t nl; f.read((char *)&nl, 4); std::string name(nl, ' '); f.read(&name[0], nl);
        Arr a; f.read(&a.dtype, 1); uint32_t nd; f.read((char *)&nd, 4); a.dims.resize(nd); f.read((char *)a.dims.data(), 8 * nd);
        int64_t n = 1; for (auto x : a.dims) n *= x;
        a.raw.resize(n * (a.dtype == 'i' ? 4 : 8)); f.read(a.raw.data(), a.raw.size());
        out[name] = std::move(a);
    }
    return out;
}
static std::vector<std::vector<double>> rows_of(const Arr & a)
{
    std::vector<std::vector<double>> m(a.dims[0], std::vector<double>(a.dims[1]));
    for (int64_t r = 0; r < a.dims[0]; ++r) for (int64_t c = 0; c < a.dims[1]; ++c) m[r][c] = a.d()[r * a.dims[1] + c];
    return m;
}
static double rel_l2(const std::vector<double> & x, const double * y)
{
    double a = 0, b = 0; for (size_t i = 0; i < x.size(); ++i) { a += (x[i] - y[i]) * (x[i] - y[i]); b += y[i] * y[i]; }
    return std::sqrt(a / b);
}

int main(int argc, char ** argv)
{
    if (argc < 2) { std::cerr << "usage: burgers_stage <dump> [--generated]" << std::endl; return 2; }
    // --generated: every 1D table comes from the library's own generator (amdg_op_generate*); the dump then only supplies the grid, the
    // initial coefficients and the reference's stage results
    const bool generated = argc > 2 && std::string(argv[2]) == "--generated";
    // --wave / --wave-generated: the dump is a wave fixture (cfg3): the interior-penalty Laplacian of amdg::DiffusionAlpt (one pre-merged 1D operator per
    // dimension, no assembled matrix) applied to the dumped coefficients against the reference's assembled SpMV (wave.rhs_spmv)
    const bool wave = argc > 2 && std::string(argv[2]).rfind("--wave", 0) == 0;
    if (wave)
    {
        auto D = load_dump(argv[1]);
        const int * cfg = D["config"].i();
        const int DIM = cfg[0], NMAX = cfg[1], PA = cfg[4], PL = cfg[5];
        try
        {
            amdg::DGSolution dg_solu(DIM, NMAX, PA, PL, 1, 0);
            dg_solu.set_elements(cfg[9], D["level"].i(), D["suppt"].i());
            amdg::OperatorMatrix1D oper_matx_alpt;
            if (std::string(argv[2]) == "--wave-generated") oper_matx_alpt = amdg::OperatorMatrix1D(dg_solu, AMDG_BASIS_ALPERT, PA);
            else
            {
                oper_matx_alpt = amdg::OperatorMatrix1D(dg_solu, PA + 1, PA + 1, D["alpt.u_v"].d(), D["alpt.u_vx"].d(), D["alpt.ulft_vjp"].d(), D["alpt.urgt_vjp"].d(), D["alpt.ujp_vjp"].d());
                oper_matx_alpt.set_diffusion_tables(dg_solu, D["alpt.ux_vx"].d(), D["alpt.uxave_vjp"].d(), D["alpt.ujp_vxave"].d());
            }
            const double sigma = DIM == 2 ? 10. : 20.;                           // example/03_wave_01_const_coeff_periodic.cpp:64
            amdg::DiffusionAlpt linear(dg_solu, oper_matx_alpt, sigma);
            linear.assemble_matrix_scalar(std::vector<double>(DIM, 1.));
            dg_solu.ucoe_alpt.upload(D["ucoe_alpt.in"].d());
            dg_solu.set_rhs_zero();
            amdg::RK3SSP ode(linear, 1e-3);
            ode.add_rhs_matrix(linear);
            std::vector<double> host(dg_solu.get_dof());
            dg_solu.rhs.download(host.data());
            const double e = rel_l2(host, D["wave.rhs_spmv"].d());
            std::printf("DiffusionAlpt (max mesh level %d): rel-L2 vs the reference's assembled operator %.3e\n", dg_solu.max_mesh_level(), e);
            if (!(e < 1e-12)) { std::printf("FAIL\n"); return 1; }
            std::printf("OK\n");
        }
        catch (const std::exception & e) { std::cerr << e.what() << std::endl; return 1; }
        return 0;
    }
    auto D = load_dump(argv[1]);
    const int * cfg = D["config"].i();
    const int DIM = cfg[0], NMAX = cfg[1], PA = cfg[4], PL = cfg[5];
    const int64_t ne = cfg[9];
    try
    {
        amdg::DGSolution dg_solu(DIM, NMAX, PA, PL, 1, 0);
        dg_solu.set_elements(ne, D["level"].i(), D["suppt"].i());
        amdg::OperatorMatrix1D oper_matx_lagr = generated ? amdg::OperatorMatrix1D(dg_solu, AMDG_BASIS_LAGRANGE, PL)
            : amdg::OperatorMatrix1D(dg_solu, PL + 1, PA + 1, D["lagr.u_v"].d(), D["lagr.u_vx"].d(), D["lagr.ulft_vjp"].d(), D["lagr.urgt_vjp"].d());
        amdg::OperatorMatrix1D oper_matx_alpt = generated ? amdg::OperatorMatrix1D(dg_solu, AMDG_BASIS_ALPERT, PA)
            : amdg::OperatorMatrix1D(dg_solu, PA + 1, PA + 1, D["alpt.u_v"].d(), D["alpt.u_vx"].d(), D["alpt.ulft_vjp"].d(), D["alpt.urgt_vjp"].d(), D["alpt.ujp_vjp"].d());
        amdg::LagrInterpolation interp_lagr = generated ? amdg::LagrInterpolation(dg_solu) : amdg::LagrInterpolation(dg_solu, D["lagr.pw_anc"].i(), D["lagr.pw_wt"].d());
        amdg::FastLagrIntp fast_lagr_intp = generated ? amdg::FastLagrIntp(dg_solu, AMDG_BASIS_LAGRANGE)
            : amdg::FastLagrIntp(dg_solu, rows_of(D["Lag_pt_Alpt_1D"]), rows_of(D["Lag_pt_Alpt_1D_d1"]));
        amdg::HyperbolicLagrRHS fast_rhs_lagr(dg_solu, oper_matx_lagr);
        amdg::HyperbolicAlptRHS fast_rhs_alpt(dg_solu, oper_matx_alpt);
        dg_solu.ucoe_alpt.upload(D["ucoe_alpt.in"].d());

        const double dt = 0.002;
        const std::vector<double> lxf_alpha(DIM, 1.2);
        std::vector<std::vector<bool>> is_intp(1, std::vector<bool>(DIM, false)); is_intp[0][0] = true;     // Burgers: one flux component
        const std::vector<int> flux(DIM, AMDG_FLUX_BURGERS);
        amdg::RK3SSP odeSolver(dg_solu, dt);
        odeSolver.init();
        std::vector<double> host(dg_solu.get_dof());
        double worst = 0;
        for (int stage = 0; stage < odeSolver.num_stage; ++stage)
        {
            interp_lagr.nonlinear_Lagr_fast(flux, {}, is_intp, fast_lagr_intp);
            dg_solu.set_rhs_zero();
            fast_rhs_lagr.rhs_vol_scalar();
            fast_rhs_lagr.rhs_flx_intp_scalar();
            fast_rhs_alpt.rhs_flx_penalty_scalar(lxf_alpha);
            odeSolver.set_rhs_zero();
            odeSolver.add_rhs_to_eigenvec();
            odeSolver.step_stage(stage);
            odeSolver.final();
            dg_solu.ucoe_alpt.download(host.data());
            const double e = rel_l2(host, D["stage" + std::to_string(stage) + ".ucoe_alpt"].d());
            std::printf("stage %d rel-L2 vs reference %.3e\n", stage, e);
            worst = std::max(worst, e);
        }
        if (!(worst < 1e-12)) { std::printf("FAIL\n"); return 1; }
        std::printf("OK\n");
    }
    catch (const std::exception & e) { std::cerr << e.what() << std::endl; return 1; }
    return 0;
}
