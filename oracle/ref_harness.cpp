// TEST INFRASTRUCTURE ONLY (oracle).  Never linked into, imported by or executed from the product path.
//
// Driver around the UNMODIFIED reference sources (compiled from /root/reference/source/*.cpp by
// oracle/Makefile into oracle/_ref/).  It sets the reference's process-wide statics the way its examples do
// (example/02_hyperbolic_01_scalar_const_coefficient.cpp:40-65), builds a DGAdapt on the requested grid,
// fills Element::ucoe_alpt with a stateless pseudo-random field, calls the reference's own hot-path functions
// (FastLagrIntp::eval_up_Lagr, LagrInterpolation::eval_fp_Lag / eval_fp_to_coe_D_Lag / eval_up_to_coe_D_Lag,
// FastLagrInit::eval_ucoe_Alpt_Lagr, HyperbolicLagrRHS / HyperbolicHermRHS / HyperbolicAlptRHS,
// FastRHS::transform_ucoe_alpt_to_rhs, RK3SSP / ForwardEuler / RK4ODE2nd, HyperbolicAlpt / DiffusionAlpt SpMV)
// and writes every phase's per-element arrays, sorted by ascending Hash::hash_key, into one binary dump that
// tests/ read (tests/refdump.py).  With --time it prints per-phase wall times as one JSON line (the
// cpu_baseline of bench.py).
//
// Usage: ref_harness --dim D --nmax N [--n0 N0] [--sparse 1] [--pa K] [--pl M] [--ph M] [--intp lagr|herm]
//                    [--vecnum V] [--seed S] [--flux burgers|linear|kpp|vlasov] [--run a,b,c] [--out FILE]
//                    [--time REPS] [--threads T] [--dump-tables 1]
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

// the oracle needs to reach protected/private phases (eval_fp_Lag, eval_fp_to_coe_D_Lag, pw1d, eval_fp_Her_2D ...)
#define private public
#define protected public
#include "DGAdaptIntp.h"
#include "Interpolation.h"
#include "FastMultiplyLU.h"
#include "ODESolver.h"
#include "OperatorMatrix1D.h"
#include "BilinearForm.h"
#undef private
#undef protected

// ---------------------------------------------------------------------------------------------------------
// dump container: sequence of records  [u32 name_len][name][u8 dtype: 'd' f64 | 'i' i32 | 'q' i64][u32 ndim][i64 dims...][raw]
// ---------------------------------------------------------------------------------------------------------
struct Dump
{
    FILE * f = nullptr;
    bool open(const std::string & path) { f = fopen(path.c_str(), "wb"); if (f) fwrite("AMDGDUMP", 1, 8, f); return f != nullptr; }
    void close() { if (f) fclose(f); f = nullptr; }
    void header(const std::string & name, char dtype, const std::vector<int64_t> & dims)
    {
        uint32_t nl = name.size(); fwrite(&nl, 4, 1, f); fwrite(name.data(), 1, nl, f);
        fwrite(&dtype, 1, 1, f);
        uint32_t nd = dims.size(); fwrite(&nd, 4, 1, f); fwrite(dims.data(), 8, nd, f);
    }
    void put(const std::string & name, const std::vector<double> & v, std::vector<int64_t> dims = {})
    {
        if (!f) return;
        if (dims.empty()) dims = { (int64_t)v.size() };
        header(name, 'd', dims); fwrite(v.data(), 8, v.size(), f);
    }
    void put(const std::string & name, const std::vector<int> & v, std::vector<int64_t> dims = {})
    {
        if (!f) return;
        if (dims.empty()) dims = { (int64_t)v.size() };
        header(name, 'i', dims); fwrite(v.data(), 4, v.size(), f);
    }
};

// stateless pseudo random field (restated identically in tests/refdump.py)
static inline uint64_t splitmix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}
static inline double field_value(uint64_t seed, int hash_key, int vec, int idx, int sum_level)
{
    uint64_t h = splitmix64(seed ^ splitmix64(((uint64_t)(uint32_t)hash_key << 20) + ((uint64_t)vec << 16) + (uint64_t)idx));
    double u = (double)(h >> 11) * (1.0 / 9007199254740992.0);   // [0,1)
    return (2.0 * u - 1.0) * std::ldexp(1.0, -sum_level);
}

struct Args
{
    int dim = 2, nmax = 4, n0 = -1, sparse = 1, pa = 2, pl = 3, ph = 3, vecnum = 1, time_reps = 0, threads = 0, dump_tables = 0;
    int msh_lagr = 1, msh_herm = 1, steps = 1, rounds = 0;
    double eps = 1e10, eta = -1.0;
    uint64_t seed = 20240901ULL;
    double dt = 1e-3;
    std::string intp = "lagr", flux = "burgers", run = "grid", out = "", boundary = "period";
};

static Args parse(int argc, char ** argv)
{
    Args a;
    for (int i = 1; i + 1 < argc; i += 2)
    {
        std::string k = argv[i], v = argv[i + 1];
        if (k == "--dim") a.dim = std::stoi(v); else if (k == "--nmax") a.nmax = std::stoi(v);
        else if (k == "--n0") a.n0 = std::stoi(v); else if (k == "--sparse") a.sparse = std::stoi(v);
        else if (k == "--pa") a.pa = std::stoi(v); else if (k == "--pl") a.pl = std::stoi(v);
        else if (k == "--ph") a.ph = std::stoi(v); else if (k == "--vecnum") a.vecnum = std::stoi(v);
        else if (k == "--seed") a.seed = std::stoull(v); else if (k == "--intp") a.intp = v;
        else if (k == "--flux") a.flux = v; else if (k == "--run") a.run = v; else if (k == "--out") a.out = v;
        else if (k == "--time") a.time_reps = std::stoi(v); else if (k == "--threads") a.threads = std::stoi(v);
        else if (k == "--dump-tables") a.dump_tables = std::stoi(v); else if (k == "--dt") a.dt = std::stod(v);
        else if (k == "--msh-lagr") a.msh_lagr = std::stoi(v); else if (k == "--msh-herm") a.msh_herm = std::stoi(v);
        else if (k == "--steps") a.steps = std::stoi(v);
        else if (k == "--boundary") a.boundary = v;
        else if (k == "--adapt-eps") a.eps = std::stod(v); else if (k == "--adapt-eta") a.eta = std::stod(v);
        else if (k == "--adapt-rounds") a.rounds = std::stoi(v);
        else { std::cerr << "unknown option " << k << std::endl; exit(2); }
    }
    if (a.n0 < 0) a.n0 = a.nmax;
    return a;
}

static std::vector<double> flat(const VecMultiD<double> & m)
{
    std::vector<double> v(m.size());
    for (int i = 0; i < m.size(); ++i) v[i] = m.at(i);
    return v;
}

struct Harness
{
    Args a;
    Dump dump;
    DGAdapt * dg = nullptr;
    std::vector<Element *> sorted;              // elements in ascending hash key
    std::map<std::string, double> timing;

    double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

    void sort_elements()
    {
        sorted.clear();
        for (auto & it : dg->dg) sorted.push_back(&it.second);
        std::sort(sorted.begin(), sorted.end(), [](Element * x, Element * y) { return x->hash_key < y->hash_key; });
    }

    // ---- dumps -------------------------------------------------------------------------------------------
    void dump_grid()
    {
        const int d = a.dim; const int64_t ne = sorted.size();
        std::vector<int> key(ne), lev(ne * d), sup(ne * d), ord(ne * d), iter_order;
        std::unordered_map<Element *, int> row;
        for (int64_t e = 0; e < ne; ++e)
        {
            key[e] = sorted[e]->hash_key; row[sorted[e]] = e;
            for (int t = 0; t < d; ++t) { lev[e * d + t] = sorted[e]->level[t]; sup[e * d + t] = sorted[e]->suppt[t]; ord[e * d + t] = sorted[e]->order_elem[t]; }
        }
        for (auto & it : dg->dg) iter_order.push_back(it.first);
        dump.put("hash_key", key); dump.put("level", lev, { ne, d }); dump.put("suppt", sup, { ne, d });
        dump.put("order_elem", ord, { ne, d }); dump.put("iter_order", iter_order);
        for (int t = 0; t < d; ++t)
        {
            for (int kind = 0; kind < 2; ++kind)
            {
                std::vector<int> ptr(1, 0), idx;
                for (int64_t e = 0; e < ne; ++e)
                {
                    const auto & s = kind == 0 ? sorted[e]->ptr_vol_alpt[t] : sorted[e]->ptr_flx_alpt[t];
                    std::vector<int> r; for (Element * p : s) r.push_back(row[p]);
                    std::sort(r.begin(), r.end()); idx.insert(idx.end(), r.begin(), r.end()); ptr.push_back(idx.size());
                }
                std::string nm = std::string(kind == 0 ? "vol" : "flx") + "_d" + std::to_string(t);
                dump.put(nm + "_ptr", ptr); dump.put(nm + "_idx", idx);
            }
        }
    }

    enum Field { UCOE_ALPT, UP_INTP, UCOE_INTP, RHS };
    void dump_field(const std::string & name, Field fld)
    {
        if (!dump.f) return;
        std::vector<double> all; int64_t blk = 0;
        for (Element * e : sorted)
            for (int v = 0; v < a.vecnum; ++v)
            {
                const VecMultiD<double> & m = fld == UCOE_ALPT ? e->ucoe_alpt[v] : fld == UP_INTP ? e->up_intp[v] : fld == UCOE_INTP ? e->ucoe_intp[v] : e->rhs[v];
                blk = m.size();
                for (int i = 0; i < m.size(); ++i) all.push_back(m.at(i));
            }
        dump.put(name, all, { (int64_t)sorted.size(), a.vecnum, blk });
    }
    // fp_intp / fucoe_intp : [elem][vec][dim][block]
    void dump_flux_field(const std::string & name, bool coe)
    {
        if (!dump.f) return;
        std::vector<double> all; int64_t blk = 0;
        for (Element * e : sorted)
            for (int v = 0; v < a.vecnum; ++v)
                for (int t = 0; t < a.dim; ++t)
                {
                    const VecMultiD<double> & m = coe ? e->fucoe_intp[v][t] : e->fp_intp[v][t];
                    blk = m.size();
                    for (int i = 0; i < m.size(); ++i) all.push_back(m.at(i));
                }
        dump.put(name, all, { (int64_t)sorted.size(), a.vecnum, a.dim, blk });
    }

    void fill_ucoe(uint64_t seed)
    {
        for (Element * e : sorted)
        {
            int sl = 0; for (int t = 0; t < a.dim; ++t) sl += e->level[t];
            for (int v = 0; v < a.vecnum; ++v)
                for (int i = 0; i < e->ucoe_alpt[v].size(); ++i)
                    e->ucoe_alpt[v].at(i) = field_value(seed, e->hash_key, v, i, sl);
        }
    }
};

static void dump_matrix(Dump & dump, const std::string & name, const VecMultiD<double> & m)
{
    dump.put(name, flat(m), { m.vec_size()[0], m.vec_size()[1] });
}
static void dump_matrix(Dump & dump, const std::string & name, const std::vector<std::vector<double>> & m)
{
    std::vector<double> v; for (auto & r : m) v.insert(v.end(), r.begin(), r.end());
    dump.put(name, v, { (int64_t)m.size(), (int64_t)m[0].size() });
}

int main(int argc, char ** argv)
{
    Harness H; H.a = parse(argc, argv); Args & a = H.a;
    const int DIM = a.dim;
    const bool herm = (a.intp == "herm");

    // ---- statics, as the examples set them ---------------------------------------------------------------
    AlptBasis::PMAX = a.pa;
    LagrBasis::PMAX = a.pl; LagrBasis::msh_case = a.msh_lagr;
    HermBasis::PMAX = a.ph; HermBasis::msh_case = a.msh_herm;
    Element::PMAX_alpt = AlptBasis::PMAX;
    Element::PMAX_intp = herm ? HermBasis::PMAX : LagrBasis::PMAX;
    Element::DIM = DIM; Element::VEC_NUM = a.vecnum;
    DGSolution::DIM = DIM; DGSolution::VEC_NUM = a.vecnum;
    Interpolation::DIM = DIM; Interpolation::VEC_NUM = a.vecnum;
    DGSolution::ind_var_vec.clear(); for (int v = 0; v < a.vecnum; ++v) DGSolution::ind_var_vec.push_back(v);
    DGAdapt::indicator_var_adapt = { 0 };
    Element::is_intp.resize(a.vecnum);
    for (int v = 0; v < a.vecnum; ++v) Element::is_intp[v] = std::vector<bool>(DIM, true);
    if (a.threads > 0) omp_set_num_threads(a.threads);
    const std::string boundary_type = a.boundary;

    std::set<std::string> run; { std::stringstream ss(a.run); std::string tok; while (std::getline(ss, tok, ',')) run.insert(tok); }
    auto has = [&](const char * s) { return run.count(s) > 0; };

    if (!a.out.empty() && !H.dump.open(a.out)) { std::cerr << "cannot open " << a.out << std::endl; return 2; }

    double t0 = H.now();
    Hash hash;
    LagrBasis::set_interp_msh01();
    HermBasis::set_interp_msh01();
    AllBasis<AlptBasis> all_bas_alpt(a.nmax);
    AllBasis<LagrBasis> all_bas_lagr(a.nmax);
    AllBasis<HermBasis> all_bas_herm(a.nmax);

    OperatorMatrix1D<AlptBasis, AlptBasis> oper_alpt(all_bas_alpt, all_bas_alpt, boundary_type);
    OperatorMatrix1D<LagrBasis, AlptBasis> oper_lagr(all_bas_lagr, all_bas_alpt, boundary_type);
    OperatorMatrix1D<HermBasis, AlptBasis> oper_herm(all_bas_herm, all_bas_alpt, boundary_type);
    H.timing["setup_tables"] = H.now() - t0;

    t0 = H.now();
    DGAdapt dg(a.sparse == 1, a.n0, a.nmax, all_bas_alpt, all_bas_lagr, all_bas_herm, hash, a.eps, a.eta, true, false);
    H.dg = &dg; H.sort_elements();
    // adaptive (irregular) grids: DGAdapt::refine / coarsen driven by the norms of the pseudo-random field
    // (source/DGAdapt.cpp:371-404, 675-688); the neighbour sets are then maintained by add_elem / del_elem (:1073-1248)
    for (int r = 0; r < a.rounds; ++r) { H.fill_ucoe(a.seed); dg.refine(); H.sort_elements(); }
    if (a.rounds > 0 && a.eta > 0) { H.fill_ucoe(a.seed); dg.coarsen(); H.sort_elements(); }
    H.timing["setup_grid"] = H.now() - t0;

    {
        std::vector<int> cfg = { DIM, a.nmax, a.n0, a.sparse, a.pa, a.pl, a.ph, a.vecnum, herm ? 1 : 0, (int)H.sorted.size() };
        H.dump.put("config", cfg);
    }
    if (has("grid")) H.dump_grid();

    LagrInterpolation interp_lagr(dg);
    HermInterpolation interp_herm(dg);

    if (a.dump_tables)
    {
        dump_matrix(H.dump, "alpt.u_v", oper_alpt.u_v); dump_matrix(H.dump, "alpt.u_vx", oper_alpt.u_vx);
        dump_matrix(H.dump, "alpt.ulft_vjp", oper_alpt.ulft_vjp); dump_matrix(H.dump, "alpt.urgt_vjp", oper_alpt.urgt_vjp);
        dump_matrix(H.dump, "alpt.ujp_vjp", oper_alpt.ujp_vjp); dump_matrix(H.dump, "alpt.ux_vx", oper_alpt.ux_vx);
        dump_matrix(H.dump, "alpt.uxave_vjp", oper_alpt.uxave_vjp); dump_matrix(H.dump, "alpt.ujp_vxave", oper_alpt.ujp_vxave);
        dump_matrix(H.dump, "lagr.u_v", oper_lagr.u_v); dump_matrix(H.dump, "lagr.u_vx", oper_lagr.u_vx);
        dump_matrix(H.dump, "lagr.ulft_vjp", oper_lagr.ulft_vjp); dump_matrix(H.dump, "lagr.urgt_vjp", oper_lagr.urgt_vjp);
        if (a.dump_tables > 1)   // the tables DiffusionRHS uses, and the remaining Alpert x Alpert tables (table-generation parity)
        {
            dump_matrix(H.dump, "lagr.uave_vjp", oper_lagr.uave_vjp); dump_matrix(H.dump, "lagr.ujp_vxlft", oper_lagr.ujp_vxlft);
            dump_matrix(H.dump, "lagr.ujp_vxrgt", oper_lagr.ujp_vxrgt); dump_matrix(H.dump, "lagr.ujp_vjp", oper_lagr.ujp_vjp);
            dump_matrix(H.dump, "alpt.uave_vjp", oper_alpt.uave_vjp); dump_matrix(H.dump, "alpt.ujp_vxlft", oper_alpt.ujp_vxlft);
            dump_matrix(H.dump, "alpt.ujp_vxrgt", oper_alpt.ujp_vxrgt); dump_matrix(H.dump, "alpt.ux_v", oper_alpt.ux_v);
            dump_matrix(H.dump, "alpt.uxrgt_vjp", oper_alpt.uxrgt_vjp); dump_matrix(H.dump, "alpt.uxlft_vjp", oper_alpt.uxlft_vjp);
            dump_matrix(H.dump, "herm.uave_vjp", oper_herm.uave_vjp); dump_matrix(H.dump, "herm.ujp_vjp", oper_herm.ujp_vjp);
        }
        dump_matrix(H.dump, "herm.u_v", oper_herm.u_v); dump_matrix(H.dump, "herm.u_vx", oper_herm.u_vx);
        dump_matrix(H.dump, "herm.ulft_vjp", oper_herm.ulft_vjp); dump_matrix(H.dump, "herm.urgt_vjp", oper_herm.urgt_vjp);
        dump_matrix(H.dump, "Lag_pt_Alpt_1D", interp_lagr.Lag_pt_Alpt_1D);
        dump_matrix(H.dump, "Lag_pt_Alpt_1D_d1", interp_lagr.Lag_pt_Alpt_1D_d1);
        dump_matrix(H.dump, "Her_pt_Alpt_1D", interp_herm.Her_pt_Alpt_1D);
        H.dump.put("LagrBasis.intp_msh0", LagrBasis::intp_msh0); H.dump.put("LagrBasis.intp_msh1", LagrBasis::intp_msh1);
        H.dump.put("HermBasis.intp_msh0", HermBasis::intp_msh0); H.dump.put("HermBasis.intp_msh1", HermBasis::intp_msh1);
        // interpolation point coordinate of every 1D Lagrange basis function, in AllBasis order
        std::vector<double> pts; for (int i = 0; i < all_bas_lagr.size(); ++i) pts.push_back(all_bas_lagr.at(i).intep_pt);
        H.dump.put("lagr.intep_pt", pts);
        std::vector<double> hpts; for (int i = 0; i < all_bas_herm.size(); ++i) hpts.push_back(all_bas_herm.at(i).intep_pt);
        H.dump.put("herm.intep_pt", hpts);
        // hierarchisation stencils (pwts) of every 1D element with level > 0: Lagrange and Hermite
        {
            const int P1 = LagrBasis::PMAX + 1;
            std::vector<int> anc; std::vector<double> wt;
            for (int n = 1; n <= a.nmax; ++n)
                for (int j = 1; j < std::max(2, pow_int(2, n)); j += 2)
                {
                    interp_lagr.pw1d.clear(); interp_lagr.set_pts_wts_1d_ada_Lag(n, j);
                    const pwts & p = interp_lagr.pw1d.begin()->second;
                    for (int ic = 0; ic < P1; ++ic) { anc.push_back(interp_lagr.hash_key1d(p.p_k[ic], p.p_i[ic])); anc.push_back(p.p_num[ic]); }
                    for (int p0 = 0; p0 < P1; ++p0) for (int ic = 0; ic < P1; ++ic) wt.push_back(p.wt[p0][ic]);
                }
            interp_lagr.pw1d.clear();
            const int64_t T1 = pow_int(2, a.nmax) - 1;
            H.dump.put("lagr.pw_anc", anc, { T1, P1, 2 }); H.dump.put("lagr.pw_wt", wt, { T1, P1, P1 });
        }
        {
            const int P1 = HermBasis::PMAX + 1;
            std::vector<int> anc; std::vector<double> wt;
            for (int n = 1; n <= a.nmax; ++n)
                for (int j = 1; j < std::max(2, pow_int(2, n)); j += 2)
                {
                    interp_herm.pw1d.clear(); interp_herm.set_pts_wts_1d_ada_Her(n, j);
                    const pwts & p = interp_herm.pw1d.begin()->second;
                    for (size_t ic = 0; ic < p.p_k.size(); ++ic) { anc.push_back(interp_herm.hash_key1d(p.p_k[ic], p.p_i[ic])); anc.push_back(p.p_num[ic]); }
                    for (size_t p0 = 0; p0 < p.wt.size(); ++p0) for (size_t ic = 0; ic < p.wt[p0].size(); ++ic) wt.push_back(p.wt[p0][ic]);
                }
            interp_herm.pw1d.clear();
            const int64_t T1 = pow_int(2, a.nmax) - 1;
            H.dump.put("herm.pw_anc", anc, { T1, (int64_t)anc.size() / (2 * T1), 2 });
            H.dump.put("herm.pw_wt", wt, { T1, P1, (int64_t)wt.size() / (T1 * P1) });
        }
    }

    H.fill_ucoe(a.seed);
    H.dump_field("ucoe_alpt.in", Harness::UCOE_ALPT);

    // ---- flux functions ------------------------------------------------------------------------------------
    const std::vector<double> lin_coef = { 1.0, 0.7, -0.5, 0.3, 1.3, -0.9 };
    const std::string fname = (a.flux == "burgers1") ? std::string("burgers") : a.flux;
    auto flux = [&](std::vector<double> u, int i, int d) -> double
    {
        if (fname == "burgers") return u[i] * u[i] / 2.;
        if (fname == "linear") return lin_coef[d] * u[i];
        if (fname == "kpp") return d == 0 ? std::sin(u[i]) : std::cos(u[i]);
        return u[i];
    };
    auto flux_d1 = [&](std::vector<double> u, int i, int d, int i1) -> double
    {
        if (fname == "burgers") return u[i];
        if (fname == "linear") return lin_coef[d];
        if (fname == "kpp") return d == 0 ? std::cos(u[i]) : -std::sin(u[i]);
        return 1.;
    };
    auto flux_d2 = [&](std::vector<double> u, int i, int d, int i1, int i2) -> double
    {
        if (fname == "burgers") return 1.;
        if (fname == "linear") return 0.;
        if (fname == "kpp") return d == 0 ? -std::sin(u[i]) : -std::cos(u[i]);
        return 0.;
    };
    std::vector<std::vector<bool>> is_intp(a.vecnum, std::vector<bool>(DIM, true));
    // Burgers in the shipped example interpolates one flux component only (example/02_hyperbolic_05_burgers_adapt.cpp:205)
    if (a.flux == "burgers1") { for (int v = 0; v < a.vecnum; ++v) for (int t = 1; t < DIM; ++t) is_intp[v][t] = false; }

    FastLagrIntp fast_lagr_intp(dg, interp_lagr.Lag_pt_Alpt_1D, interp_lagr.Lag_pt_Alpt_1D_d1);
    FastHermIntp fast_herm_intp(dg, interp_herm.Her_pt_Alpt_1D);
    FastLagrInit fast_lagr_init(dg, oper_lagr);
    FastHermInit fast_herm_init(dg, oper_herm);
    HyperbolicLagrRHS rhs_lagr(dg, oper_lagr);
    HyperbolicHermRHS rhs_herm(dg, oper_herm);
    HyperbolicAlptRHS rhs_alpt(dg, oper_alpt);
    const std::vector<double> lax_alpha(DIM, 1.2);

    const int reps = std::max(1, a.time_reps);
    auto timed = [&](const std::string & name, std::function<void()> fn)
    {
        std::vector<double> ts;
        for (int r = 0; r < reps; ++r) { double s = H.now(); fn(); ts.push_back(H.now() - s); }
        std::sort(ts.begin(), ts.end());
        H.timing[name] = ts[ts.size() / 2]; H.timing[name + ".min"] = ts[0];
    };

    // the generalised Vlasov point-wise product (SURVEY cfg5): fp[0][t] = v_t f for t < DIM/2, E_t(x) f for t >= DIM/2
    auto vlasov_pointwise = [&]()
    {
        const int hd = DIM / 2;
        for (auto & it : dg.dg)
        {
            Element & e = it.second;
            for (auto const & p : e.order_local_intp)
            {
                std::vector<double> pos(DIM);
                for (int t = 0; t < DIM; ++t) pos[t] = dg.all_bas_Lag.at(e.level[t], e.suppt[t], p[t]).intep_pt;
                const double f = e.up_intp[0].at(p);
                for (int t = 0; t < DIM; ++t)
                {
                    double c;
                    if (t < hd) c = pos[hd + t];                                    // v_t
                    else { c = 0.; for (int s = 0; s < hd; ++s) c += std::sin(2. * Const::PI * (pos[s] + 0.125 * (t - hd + 1))); }  // prescribed smooth E_t(x)
                    e.fp_intp[0][t].at(p) = c * f;
                }
            }
        }
    };

    // ---- one nonlinear right-hand side ------------------------------------------------------------
    auto nonlinear_rhs = [&](bool dumpit, const std::string & tag)
    {
        if (!herm)
        {
            if (a.flux == "vlasov")
            {
                timed("intp", [&]() { fast_lagr_intp.eval_up_Lagr(0); });
                if (dumpit) H.dump_field("up_intp" + tag, Harness::UP_INTP);
                timed("pointwise", [&]() { vlasov_pointwise(); });
                std::vector<std::vector<bool>> isv; isv.push_back(std::vector<bool>(DIM, true));
                for (int v = 1; v < a.vecnum; ++v) isv.push_back(std::vector<bool>(DIM, false));
                if (dumpit) H.dump_flux_field("fp_intp" + tag, false);
                timed("hier", [&]() { interp_lagr.pw1d.clear(); interp_lagr.eval_fp_to_coe_D_Lag(isv); });
            }
            else
            {
                timed("intp", [&]() { fast_lagr_intp.eval_up_Lagr(); });
                if (dumpit) H.dump_field("up_intp" + tag, Harness::UP_INTP);
                timed("pointwise", [&]() { interp_lagr.eval_fp_Lag(flux, is_intp); });
                if (dumpit) H.dump_flux_field("fp_intp" + tag, false);
                timed("hier", [&]() { interp_lagr.pw1d.clear(); interp_lagr.eval_fp_to_coe_D_Lag(is_intp); });
            }
        }
        else
        {
            timed("intp", [&]() { fast_herm_intp.eval_up_Herm(); });
            if (dumpit) H.dump_field("up_intp" + tag, Harness::UP_INTP);
            timed("pointwise", [&]() { interp_herm.eval_fp_Her_2D(flux, flux_d1, flux_d2, is_intp); });
            if (dumpit) H.dump_flux_field("fp_intp" + tag, false);
            timed("hier", [&]() { interp_herm.pw1d.clear(); interp_herm.eval_fp_to_coe_D_Her(is_intp); });
        }
        if (dumpit) H.dump_flux_field("fucoe_intp" + tag, true);
        dg.set_rhs_zero();
        timed("rhs_vol", [&]() { if (herm) rhs_herm.rhs_vol_scalar(); else rhs_lagr.rhs_vol_scalar(); });
        if (dumpit) H.dump_field("rhs_vol" + tag, Harness::RHS);     // with --time REPS>1 the rhs accumulates REPS times: dump only with reps==1
        timed("rhs_flx", [&]() { if (herm) rhs_herm.rhs_flx_intp_scalar(); else rhs_lagr.rhs_flx_intp_scalar(); });
        if (dumpit) H.dump_field("rhs_vol_flx" + tag, Harness::RHS);
        timed("rhs_penalty", [&]() { rhs_alpt.rhs_flx_penalty_scalar(lax_alpha); });
        if (dumpit) H.dump_field("rhs_all" + tag, Harness::RHS);
    };

    // ---- scenarios ---------------------------------------------------------------------------------------------
    if (has("roundtrip"))   // cfg2: Alpert -> point values -> hierarchical interpolation coefficients -> Alpert
    {
        if (!herm)
        {
            timed("intp", [&]() { fast_lagr_intp.eval_up_Lagr(); });
            H.dump_field("rt.up_intp", Harness::UP_INTP);
            timed("hier", [&]() { interp_lagr.pw1d.clear(); interp_lagr.eval_up_to_coe_D_Lag(); });
            H.dump_field("rt.ucoe_intp", Harness::UCOE_INTP);
            timed("init", [&]() { fast_lagr_init.eval_ucoe_Alpt_Lagr(); });
            H.dump_field("rt.ucoe_alpt", Harness::UCOE_ALPT);
        }
        else
        {
            timed("intp", [&]() { fast_herm_intp.eval_up_Herm(); });
            H.dump_field("rt.up_intp", Harness::UP_INTP);
            timed("hier", [&]() { interp_herm.pw1d.clear(); interp_herm.eval_up_to_coe_D_Her(); });
            H.dump_field("rt.ucoe_intp", Harness::UCOE_INTP);
            timed("init", [&]() { fast_herm_init.eval_ucoe_Alpt_Herm(); });
            H.dump_field("rt.ucoe_alpt", Harness::UCOE_ALPT);
        }
        H.fill_ucoe(a.seed);
    }
    if (has("der"))          // FastLagrIntp::eval_der_up_Lagr(d0) for every d0
    {
        for (int d0 = 0; d0 < DIM; ++d0)
        {
            fast_lagr_intp.eval_der_up_Lagr(d0);
            H.dump_field("der" + std::to_string(d0) + ".up_intp", Harness::UP_INTP);
        }
    }
    if (has("rhs")) { nonlinear_rhs(true, ""); }
    if (has("variants") && !herm)   // the other FastRHS compositions over the same interpolated flux: same flux in every
    {                                // direction, one flux per direction (DIM == 2 forms), source term
        nonlinear_rhs(false, "");
        H.dump_flux_field("var.fucoe_intp", true);
        if (DIM == 2)
        {
            dg.set_rhs_zero();
            HyperbolicSameFluxLagrRHS same(dg, oper_lagr); same.rhs_vol_scalar(); same.rhs_flx_intp_scalar();
            H.dump_field("var.rhs_sameflux", Harness::RHS);
            dg.set_rhs_zero();
            HyperbolicDiffFluxLagrRHS diff(dg, oper_lagr); diff.rhs_vol_scalar(); diff.rhs_flx_intp_scalar();
            H.dump_field("var.rhs_diffflux", Harness::RHS);
        }
        dg.set_rhs_zero();
        SourceFastLagr srcf(dg, oper_lagr); srcf.rhs_source();
        H.dump_field("var.rhs_source", Harness::RHS);
    }
    if (has("variants") && herm && DIM == 2)
    {
        nonlinear_rhs(false, "");
        H.dump_flux_field("var.fucoe_intp", true);
        dg.set_rhs_zero();
        HyperbolicSameFluxHermRHS same(dg, oper_herm); same.rhs_vol_scalar(); same.rhs_flx_intp_scalar();
        H.dump_field("var.rhs_sameflux", Harness::RHS);
        dg.set_rhs_zero();
        HyperbolicDiffFluxHermRHS diff(dg, oper_herm); diff.rhs_vol_scalar(); diff.rhs_flx_intp_scalar();
        H.dump_field("var.rhs_diffflux", Harness::RHS);
    }
    if (has("f4") && !herm)   // DiffusionRHS, FastRHSHamiltonJacobi, the *_coarse_grid transforms, the adapt indicator
    {
        nonlinear_rhs(false, "");
        H.dump_flux_field("f4.fucoe_intp", true);
        DiffusionRHS diff(dg, oper_lagr);
        dg.set_rhs_zero(); diff.rhs_vol(); H.dump_field("f4.diff_vol", Harness::RHS);
        dg.set_rhs_zero(); diff.rhs_flx_gradu(); H.dump_field("f4.diff_flx_gradu", Harness::RHS);
        dg.set_rhs_zero(); diff.rhs_flx_u(); H.dump_field("f4.diff_flx_u", Harness::RHS);
        dg.set_rhs_zero(); diff.rhs_flx_k_minus_u(); H.dump_field("f4.diff_flx_k_minus_u", Harness::RHS);
        dg.set_rhs_zero(); diff.rhs_flx_k_plus_u(); H.dump_field("f4.diff_flx_k_plus_u", Harness::RHS);
        FastRHSHamiltonJacobi hj(dg, oper_lagr);
        dg.set_rhs_zero(); hj.rhs_nonlinear(); H.dump_field("f4.hj", Harness::RHS);
        for (int cut = 1; cut <= 2 && a.nmax - cut >= 0; ++cut)
        {
            const int M = a.nmax - cut;
            const std::string tag = "f4.cg" + std::to_string(cut);
            H.dump.put(tag + ".mesh_nmax", std::vector<int>{ M });
            fast_lagr_intp.eval_up_Lagr_coarse_grid(M);
            H.dump_field(tag + ".up_intp", Harness::UP_INTP);
            dg.set_rhs_zero(); rhs_lagr.rhs_vol_scalar_coarse_grid(M); H.dump_field(tag + ".rhs_vol", Harness::RHS);
            rhs_lagr.rhs_flx_intp_scalar_coarse_grid(M); H.dump_field(tag + ".rhs_vol_flx", Harness::RHS);
        }
        {
            std::vector<double> norms;
            for (Element * e : H.sorted) norms.push_back(dg.indicator_norm(*e));
            H.dump.put("f4.indicator_norm", norms);
        }
        dg.set_rhs_zero();
    }
    if (has("pw") && !herm)   // the point-wise kernels of the reference beyond a scalar flux
    {
        H.fill_ucoe(a.seed);
        // (A) system flux: LagrInterpolation::eval_fp_Lag passes all VEC_NUM unknowns to the flux (source/Interplation.cpp:256-295)
        if (a.vecnum >= 2)
        {
            auto sys_flux = [&](std::vector<double> u, int i, int d) -> double
            {
                return i == 0 ? (d + 1.) * u[0] * u[1] : 0.5 * u[1] * u[1] - (d + 1.) * u[0];
            };
            std::vector<std::vector<bool>> all_intp(a.vecnum, std::vector<bool>(DIM, true));
            interp_lagr.pw1d.clear();
            interp_lagr.nonlinear_Lagr_fast(sys_flux, all_intp, fast_lagr_intp);
            H.dump_field("pw.sys.up_intp", Harness::UP_INTP);
            H.dump_flux_field("pw.sys.fp_intp", false);
            H.dump_flux_field("pw.sys.fucoe_intp", true);
        }
        // (B) coefficient of position: LagrInterpolation::var_coeff_u_Lagr_fast -> eval_coe_u_Lag (:648-698, 4199-4213)
        {
            auto coe = [&](std::vector<double> x, int d) -> double
            {
                double s = 0.3 * (d + 1.);
                for (int t = 0; t < DIM; ++t) s += std::sin(2. * Const::PI * (x[t] + 0.1 * t)) * (t == d ? 1. : 0.25);
                return s;
            };
            std::vector<std::vector<bool>> all_intp(a.vecnum, std::vector<bool>(DIM, true));
            interp_lagr.pw1d.clear();
            interp_lagr.var_coeff_u_Lagr_fast(coe, all_intp, fast_lagr_intp);
            H.dump_flux_field("pw.coe.fp_intp", false);
            H.dump_flux_field("pw.coe.fucoe_intp", true);
        }
        // (C) the 2D2V Vlasov body with the field values of a second solution broadcast by DGSolution::copy_up_intp_to_f
        //     (LagrInterpolation::interp_Vlasov_2D2V, :4508-4580; source/DGSolution.cpp:1024-1065)
        if (DIM == 4 && a.vecnum == 2)
        {
            DGAdapt E(false, a.nmax, a.nmax, 2, all_bas_alpt, all_bas_lagr, all_bas_herm, hash, a.eps, a.eta, true, false);   // swept: needs its Alpert neighbour sets
            std::vector<Element *> es;
            for (auto & it : E.dg) es.push_back(&it.second);
            std::sort(es.begin(), es.end(), [](Element * x, Element * y) { return x->hash_key < y->hash_key; });
            std::vector<int> key, lev, sup; std::vector<double> ue;
            for (Element * e : es)
            {
                int sl = 0; for (int t = 0; t < DIM; ++t) sl += e->level[t];
                key.push_back(e->hash_key);
                for (int t = 0; t < DIM; ++t) { lev.push_back(e->level[t]); sup.push_back(e->suppt[t]); }
                for (int v = 0; v < a.vecnum; ++v)
                    for (int i = 0; i < e->ucoe_alpt[v].size(); ++i) { e->ucoe_alpt[v].at(i) = field_value(a.seed + 7, e->hash_key, v, i, sl); ue.push_back(e->ucoe_alpt[v].at(i)); }
            }
            H.dump.put("pw.vl.E.hash_key", key);
            H.dump.put("pw.vl.E.level", lev, { (int64_t)es.size(), DIM });
            H.dump.put("pw.vl.E.suppt", sup, { (int64_t)es.size(), DIM });
            H.dump.put("pw.vl.E.ucoe_alpt", ue, { (int64_t)es.size(), a.vecnum, (int64_t)es[0]->ucoe_alpt[0].size() });
            FastLagrIntp fast_lagr_E(E, interp_lagr.Lag_pt_Alpt_1D, interp_lagr.Lag_pt_Alpt_1D_d1);
            interp_lagr.pw1d.clear();
            interp_lagr.interp_Vlasov_2D2V(E, fast_lagr_intp, fast_lagr_E);
            H.dump_field("pw.vl.up_intp", Harness::UP_INTP);
            H.dump_flux_field("pw.vl.fp_intp", false);
            H.dump_flux_field("pw.vl.fucoe_intp", true);
        }
        H.fill_ucoe(a.seed);
    }
    if (has("pw2") && !herm)   // more of the reference's point-wise wrappers: coefficient times gradient, source terms, the shipped 1D2V Vlasov-Maxwell body
    {
        H.fill_ucoe(a.seed);
        std::vector<std::vector<bool>> all_intp(a.vecnum, std::vector<bool>(DIM, true));
        auto coe = [&](std::vector<double> x, int d) -> double
        {
            double s = 0.3 * (d + 1.);
            for (int t = 0; t < DIM; ++t) s += std::sin(2. * Const::PI * (x[t] + 0.1 * t)) * (t == d ? 1. : 0.25);
            return s;
        };
        // (D) LagrInterpolation::var_coeff_gradu_Lagr_fast (source/Interplation.cpp:4159-4175): fp[v][d] = coe(x, d) * d/dx_d u_v, then hierarchisation
        interp_lagr.pw1d.clear();
        interp_lagr.var_coeff_gradu_Lagr_fast(coe, all_intp, fast_lagr_intp);
        H.dump_flux_field("pw2.gradu.fp_intp", false);
        H.dump_flux_field("pw2.gradu.fucoe_intp", true);
        // (E) LagrInterpolation::source_from_lagr_to_rhs (:4102-4123): a source function sampled at the points, hierarchised, projected, added to rhs
        {
            auto src = [&](std::vector<double> x, int i) -> double
            {
                double s = 1. + 0.5 * i;
                for (int t = 0; t < DIM; ++t) s *= std::cos(2. * Const::PI * (x[t] - 0.05 * (t + 1)));
                return s;
            };
            dg.set_rhs_zero();
            interp_lagr.pw1d.clear();
            interp_lagr.source_from_lagr_to_rhs(src, fast_lagr_init);
            H.dump_field("pw2.source.rhs", Harness::RHS);
            H.dump_field("pw2.source.ucoe_after", Harness::UCOE_ALPT);      // must be the coefficients from before the call
        }
        // (F) the body of the shipped example/07_vlasov_maxwell_sparse.cpp: LagrInterpolation::interp_Vlasov_1D2V with the Maxwell coefficient functions
        //     (:4435-4505; coe_x2 = v2, coe_v1 = E1 + v2 B3, coe_v2 = E2 - v1 B3, example lines 236-239), fields (B3, E1, E2) broadcast from a second solution
        if (DIM == 3 && a.vecnum == 3)
        {
            DGAdapt BE(false, a.nmax, a.nmax, 2, all_bas_alpt, all_bas_lagr, all_bas_herm, hash, a.eps, a.eta, true, false);
            std::vector<Element *> es;
            for (auto & it : BE.dg) es.push_back(&it.second);
            std::sort(es.begin(), es.end(), [](Element * x, Element * y) { return x->hash_key < y->hash_key; });
            std::vector<int> key, lev, sup; std::vector<double> ue;
            for (Element * e : es)
            {
                int sl = 0; for (int t = 0; t < DIM; ++t) sl += e->level[t];
                key.push_back(e->hash_key);
                for (int t = 0; t < DIM; ++t) { lev.push_back(e->level[t]); sup.push_back(e->suppt[t]); }
                for (int v = 0; v < a.vecnum; ++v)
                    for (int i = 0; i < e->ucoe_alpt[v].size(); ++i) { e->ucoe_alpt[v].at(i) = field_value(a.seed + 7, e->hash_key, v, i, sl); ue.push_back(e->ucoe_alpt[v].at(i)); }
            }
            H.dump.put("pw2.vm.BE.hash_key", key);
            H.dump.put("pw2.vm.BE.level", lev, { (int64_t)es.size(), DIM });
            H.dump.put("pw2.vm.BE.suppt", sup, { (int64_t)es.size(), DIM });
            H.dump.put("pw2.vm.BE.ucoe_alpt", ue, { (int64_t)es.size(), a.vecnum, (int64_t)es[0]->ucoe_alpt[0].size() });
            FastLagrIntp fast_lagr_BE(BE, interp_lagr.Lag_pt_Alpt_1D, interp_lagr.Lag_pt_Alpt_1D_d1);
            auto coe_x2 = [](double v2) -> double { return v2; };
            auto coe_v1 = [](double v2, double E1, double B3) -> double { return E1 + v2 * B3; };
            auto coe_v2 = [](double v1, double E2, double B3) -> double { return E2 - v1 * B3; };
            interp_lagr.pw1d.clear();
            interp_lagr.interp_Vlasov_1D2V(BE, coe_x2, coe_v1, coe_v2, fast_lagr_intp, fast_lagr_BE);
            H.dump_field("pw2.vm.up_intp", Harness::UP_INTP);
            H.dump_flux_field("pw2.vm.fp_intp", false);
            H.dump_flux_field("pw2.vm.fucoe_intp", true);
        }
        H.fill_ucoe(a.seed);
    }
    if (has("vlasov_ampere") && !herm && DIM == 4 && a.vecnum == 2)
    {
        // one RK3SSP step of the coupled 2D2V Vlasov-Ampere system on static grids, stage by stage as example/07_vlasov_ampere_02_2D2V_accuracy.cpp:255-318
        // (without its manufactured source): f through interp_Vlasov_2D2V + HyperbolicLagrRHS + penalty, E_t = -J through compute_moment_2D2V
        H.fill_ucoe(a.seed);
        DGAdapt E(false, a.nmax, a.nmax, 2, all_bas_alpt, all_bas_lagr, all_bas_herm, hash, a.eps, a.eta, true, false);
        std::vector<Element *> es;
        for (auto & it : E.dg) es.push_back(&it.second);
        std::sort(es.begin(), es.end(), [](Element * x, Element * y) { return x->hash_key < y->hash_key; });
        auto dump_E = [&](const std::string & name)
        {
            std::vector<double> ue;
            for (Element * e : es) for (int v = 0; v < a.vecnum; ++v) for (int i = 0; i < e->ucoe_alpt[v].size(); ++i) ue.push_back(e->ucoe_alpt[v].at(i));
            H.dump.put(name, ue, { (int64_t)es.size(), a.vecnum, (int64_t)es[0]->ucoe_alpt[0].size() });
        };
        {
            std::vector<int> key, lev, sup;
            for (Element * e : es)
            {
                int sl = 0; for (int t = 0; t < DIM; ++t) sl += e->level[t];
                key.push_back(e->hash_key);
                for (int t = 0; t < DIM; ++t) { lev.push_back(e->level[t]); sup.push_back(e->suppt[t]); }
                for (int v = 0; v < a.vecnum; ++v)
                    for (int i = 0; i < e->ucoe_alpt[v].size(); ++i) e->ucoe_alpt[v].at(i) = field_value(a.seed + 7, e->hash_key, v, i, sl);
            }
            H.dump.put("va.E.hash_key", key);
            H.dump.put("va.E.level", lev, { (int64_t)es.size(), DIM });
            H.dump.put("va.E.suppt", sup, { (int64_t)es.size(), DIM });
            dump_E("va.E.ucoe_alpt.in");
        }
        FastLagrIntp fast_lagr_E(E, interp_lagr.Lag_pt_Alpt_1D, interp_lagr.Lag_pt_Alpt_1D_d1);
        RK3SSP ode_f(dg, a.dt); ode_f.init();
        RK3SSP ode_E(E, a.dt); ode_E.init();
        for (int stage = 0; stage < ode_f.num_stage; ++stage)
        {
            interp_lagr.pw1d.clear();
            interp_lagr.interp_Vlasov_2D2V(E, fast_lagr_intp, fast_lagr_E);
            dg.set_rhs_zero();
            rhs_lagr.rhs_vol_scalar(); rhs_lagr.rhs_flx_intp_scalar(); rhs_alpt.rhs_flx_penalty_scalar(lax_alpha);
            if (stage == 0) H.dump_field("va.stage0.rhs_f", Harness::RHS);
            ode_f.set_rhs_zero(); ode_f.add_rhs_to_eigenvec(); ode_f.step_stage(stage);
            E.set_rhs_zero();
            E.compute_moment_2D2V(dg, { 1, 0 }, -1.0, 0, 0);      // - int f v1 dv -> rhs of E1
            E.compute_moment_2D2V(dg, { 0, 1 }, -1.0, 1, 0);      // - int f v2 dv -> rhs of E2
            ode_E.set_rhs_zero(); ode_E.add_rhs_to_eigenvec(); ode_E.step_stage(stage);
            ode_f.final(); ode_E.final();
            H.dump_field("va.stage" + std::to_string(stage) + ".f", Harness::UCOE_ALPT);
            dump_E("va.stage" + std::to_string(stage) + ".E");
        }
        H.fill_ucoe(a.seed);
    }
    if (has("stage"))        // full RK3SSP step with the nonlinear right-hand side, a.steps steps
    {
        for (int step = 0; step < a.steps; ++step)
        {
            RK3SSP ode(dg, a.dt);
            ode.init();
            for (int stage = 0; stage < ode.num_stage; ++stage)
            {
                nonlinear_rhs(false, "");
                timed("rk_pack", [&]() { ode.set_rhs_zero(); ode.add_rhs_to_eigenvec(); });
                double s = H.now(); ode.step_stage(stage); ode.final(); H.timing["rk_axpy"] = H.now() - s;
                if (step == 0) H.dump_field("stage" + std::to_string(stage) + ".ucoe_alpt", Harness::UCOE_ALPT);
            }
        }
        H.dump_field("final.ucoe_alpt", Harness::UCOE_ALPT);
        H.fill_ucoe(a.seed);
    }
    if (has("advection"))    // cfg1: linear advection; shipped path = assembled SpMV + RK3SSP::step_rk; sweep path alongside
    {
        const std::vector<double> c(DIM, 1.);
        // (i) single 1D sweeps, the FastRHS form of the same operator
        dg.set_rhs_zero();
        FastRHS fr(dg);
        timed("adv_sweeps", [&]()
        {
            for (int t = 0; t < DIM; ++t)
            {
                fr.transform_ucoe_alpt_to_rhs(&oper_alpt.u_vx, "vol", t, c[t], 0);
                fr.transform_ucoe_alpt_to_rhs(&oper_alpt.ulft_vjp, "flx", t, c[t], 0);    // upwind: c>=0 takes the left trace (source/BilinearForm.cpp:700-703)
            }
        });
        H.dump_field("adv.rhs_sweep", Harness::RHS);
        // (ii) the shipped assembled operator
        HyperbolicAlpt op(dg, oper_alpt);
        double s = H.now(); op.assemble_matrix_scalar(c); H.timing["adv_assemble"] = H.now() - s;
        {
            ODESolver pack(op); pack.ucoe_to_eigenvec();
            Eigen::VectorXd y = op.mat * pack.ucoe;
            pack.rhs = y; pack.eigenvec_to_rhs();
            H.dump_field("adv.rhs_spmv", Harness::RHS);
        }
        RK3SSP ode(op, a.dt);
        ode.init();
        timed("adv_step_rk", [&]() { ode.step_rk(); });
        ode.final();
        H.dump_field("adv.ucoe_alpt", Harness::UCOE_ALPT);   // after `reps` RK3 steps
        H.fill_ucoe(a.seed);
    }
    if (has("wave"))         // cfg3: u_tt = Laplace u, IPDG, RK4ODE2nd::step_rk
    {
        const double sigma = (DIM == 2) ? 10. : 20.;
        const double dx = 1. / std::pow(2., dg.max_mesh_level());
        for (Element * e : H.sorted)
        {
            int sl = 0; for (int t = 0; t < DIM; ++t) sl += e->level[t];
            for (int v = 0; v < a.vecnum; ++v) for (int i = 0; i < e->ucoe_ut[v].size(); ++i) e->ucoe_ut[v].at(i) = field_value(a.seed + 1, e->hash_key, v, i, sl);
        }
        DiffusionAlpt op(dg, oper_alpt, sigma);
        double s = H.now(); op.assemble_matrix_scalar(std::vector<double>(DIM, 1.)); H.timing["wave_assemble"] = H.now() - s;
        {
            ODESolver pack(op); pack.ucoe_to_eigenvec();
            Eigen::VectorXd y = op.mat * pack.ucoe;
            pack.rhs = y; pack.eigenvec_to_rhs();
            H.dump_field("wave.rhs_spmv", Harness::RHS);
        }
        // the same operator as single sweeps (source/BilinearForm.cpp:877-929)
        dg.set_rhs_zero();
        FastRHS fr(dg);
        timed("wave_sweeps", [&]()
        {
            for (int t = 0; t < DIM; ++t)
            {
                fr.transform_ucoe_alpt_to_rhs(&oper_alpt.ux_vx, "vol", t, -1., 0);
                fr.transform_ucoe_alpt_to_rhs(&oper_alpt.uxave_vjp, "flx", t, -1., 0);
                fr.transform_ucoe_alpt_to_rhs(&oper_alpt.ujp_vxave, "flx", t, -1., 0);
                fr.transform_ucoe_alpt_to_rhs(&oper_alpt.ujp_vjp, "flx", t, -sigma / dx, 0);
            }
        });
        H.dump_field("wave.rhs_sweep", Harness::RHS);
        RK4ODE2nd ode(op, a.dt);
        ode.init();
        timed("wave_step_rk", [&]() { ode.step_rk(); });
        ode.final();
        H.dump_field("wave.ucoe_alpt", Harness::UCOE_ALPT);
        {
            std::vector<double> all;
            for (Element * e : H.sorted) for (int v = 0; v < a.vecnum; ++v) for (int i = 0; i < e->ucoe_ut[v].size(); ++i) all.push_back(e->ucoe_ut[v].at(i));
            H.dump.put("wave.ucoe_ut", all, { (int64_t)H.sorted.size(), a.vecnum, (int64_t)H.sorted[0]->ucoe_ut[0].size() });
        }
        H.fill_ucoe(a.seed);
    }

    if (has("moment") && (DIM == 3 || DIM == 4))
    {
        // Vlasov coupling: moments of f in the two velocity dimensions accumulated into the right-hand side of a field solution E that lives on the
        // elements with level 0 in the velocity dimensions (DGAdapt::compute_moment_1D2V / _2D2V, source/DGAdapt.cpp:243-338); three calls:
        // order (0,0) weight 1.25, (1,0) weight -0.5, (0,1) weight 2
        H.fill_ucoe(a.seed);
        DGAdapt E(false, a.nmax, a.nmax, 2, all_bas_alpt, all_bas_lagr, all_bas_herm, hash, a.eps, a.eta, false, false);
        E.set_rhs_zero();
        const std::vector<std::vector<int>> orders = { {0, 0}, {1, 0}, {0, 1} };
        const std::vector<double> weights = { 1.25, -0.5, 2.0 };
        for (size_t q = 0; q < orders.size(); ++q)
        {
            if (DIM == 3) E.compute_moment_1D2V(dg, orders[q], weights[q], 0, 0); else E.compute_moment_2D2V(dg, orders[q], weights[q], 0, 0);
        }
        std::vector<Element *> es;
        for (auto & it : E.dg) es.push_back(&it.second);
        std::sort(es.begin(), es.end(), [](Element * x, Element * y) { return x->hash_key < y->hash_key; });
        std::vector<int> key, lev, sup; std::vector<double> rhs;
        for (Element * e : es)
        {
            key.push_back(e->hash_key);
            for (int t = 0; t < DIM; ++t) { lev.push_back(e->level[t]); sup.push_back(e->suppt[t]); }
            for (int i = 0; i < e->rhs[0].size(); ++i) rhs.push_back(e->rhs[0].at(i));
        }
        H.dump.put("moment.hash_key", key);
        H.dump.put("moment.level", lev, { (int64_t)es.size(), DIM });
        H.dump.put("moment.suppt", sup, { (int64_t)es.size(), DIM });
        H.dump.put("moment.rhs", rhs, { (int64_t)es.size(), (int64_t)(es.empty() ? 0 : es[0]->rhs[0].size()) });
    }

    H.dump.close();

    if (a.time_reps > 0)
    {
        std::cout << std::setprecision(9) << "{\"n_elem\": " << H.sorted.size() << ", \"dof\": " << dg.get_dof()
                  << ", \"threads\": " << omp_get_max_threads() << ", \"reps\": " << reps;
        for (auto & kv : H.timing) std::cout << ", \"" << kv.first << "\": " << kv.second;
        std::cout << "}" << std::endl;
    }
    return 0;
}
