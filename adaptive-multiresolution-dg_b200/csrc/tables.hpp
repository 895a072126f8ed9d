// 1D tables of the path, generated on our side directly in their compact form.
//
// The reference fills dense (T*p_u) x (T*p_v) tables, T = 2^NMAX, by evaluating every basis pair through std::function quadrature
// (OperatorMatrix1D, include/OperatorMatrix1D.h:124-264; LagrInterpolation::eval_Lag_pt_at_Alpt_1D, source/Interplation.cpp:16-99;
// HermInterpolation twin :1886-1965; hierarchisation stencils set_pts_wts_1d_ada_Lag / _Her, :775-887, 3166-3315): O(T^2 p^2) evaluations
// for tables that are > 99 % structural zeros.  Here only the blocks of the related 1D element pairs (Pairs1D, grid.hpp) are computed:
// O(T log T) blocks, milliseconds instead of seconds, and nothing dense is ever stored.
//
// What is restated, and how:
//   * Alpert's multiwavelets are CONSTRUCTED from their defining properties (orthonormal, piecewise degree <= P on (-1,0),(0,1), parity
//     (-1)^(p+P+1), p+P+1 vanishing moments, f(1-) > 0) by a null-space solve in long double; the reference hard-codes the resulting
//     polynomials for P <= 4 (source/AlptBasis.cpp:255-520).  Level 0 is the normalised Legendre basis (:63-69).
//   * the hierarchical Lagrange basis is the Lagrange polynomial of degree P over the P+1 interpolation points of one half interval
//     (level-0 points and level-1 points lying in that half); the reference spells the polynomials out per (P, mesh case)
//     (source/LagrBasis.cpp:190-850).  The point sets themselves (:33-154) are data and are listed below as (value, side shift).
//   * the hierarchical Hermite basis is the two-point Hermite polynomial of degree 3 / 5 on a half interval (source/HermBasis.cpp:347-1150).
//   * inner products use the reference's rule (two 10-point Gauss panels over the finer support, source/Basis.cpp:62-83) and one-sided
//     limits are taken at x -/+ 1e-13 exactly like Basis::val (source/AlptBasis.cpp:16-29), so tables agree to ~1e-15, not just to O(1e-13).
//   * boundary types "period" (the examples of the path; source/Basis.cpp:145-173, 207-235), "zero" and "inside" (:131-147, 193-205).
// Host code only; no device needed.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <string>
#include <thread>
#include <vector>

#include "grid.hpp"

namespace amdg {
namespace tab {

static const double RO = 1e-13;   // Const::ROUND_OFF, include/libs.h:28

enum Family { ALPERT = 0, LAGRANGE = 1, HERMITE = 2 };
enum Table { U_V = 0, U_VX = 1, ULFT_VJP = 2, URGT_VJP = 3, UJP_VJP = 4, UAVE_VJP = 5, UJP_VXLFT = 6, UJP_VXRGT = 7,
             UX_VX = 8, UXAVE_VJP = 9, UJP_VXAVE = 10, UX_V = 11, N_TABLE = 12 };

// ---- dense helpers (long double) ---------------------------------------------------------------------------
// solve A x = b (n x n, partial pivoting); returns false when singular
inline bool solve(std::vector<std::vector<long double>> A, std::vector<long double> b, std::vector<long double> & x)
{
    const int n = (int)b.size();
    for (int c = 0; c < n; ++c)
    {
        int piv = c;
        for (int r = c + 1; r < n; ++r) if (fabsl(A[r][c]) > fabsl(A[piv][c])) piv = r;
        if (fabsl(A[piv][c]) < 1e-30L) return false;
        std::swap(A[piv], A[c]); std::swap(b[piv], b[c]);
        for (int r = c + 1; r < n; ++r)
        {
            const long double f = A[r][c] / A[c][c];
            for (int k = c; k < n; ++k) A[r][k] -= f * A[c][k];
            b[r] -= f * b[c];
        }
    }
    x.assign(n, 0.0L);
    for (int r = n - 1; r >= 0; --r)
    {
        long double s = b[r];
        for (int k = r + 1; k < n; ++k) s -= A[r][k] * x[k];
        x[r] = s / A[r][r];
    }
    return true;
}

inline double horner(const std::vector<double> & c, double x, int der)
{
    // der-th derivative of sum_i c_i x^i
    const int n = (int)c.size();
    double r = 0.0;
    for (int i = n - 1; i >= der; --i)
    {
        double f = 1.0;
        for (int k = 0; k < der; ++k) f *= (double)(i - k);
        r = r * x + f * c[i];
    }
    return r;
}

// ---- Gauss-Legendre rule on [-1, 1] -----------------------------------------------------------------------------
struct Gauss
{
    std::vector<double> x, w;
    explicit Gauss(int n = 10) : x(n), w(n)
    {
        for (int i = 0; i < n; ++i)
        {
            long double z = cosl(3.14159265358979323846264338327950288L * (i + 0.75L) / (n + 0.5L)), pp = 1.0L;
            for (int it = 0; it < 100; ++it)
            {
                long double p1 = 1.0L, p2 = 0.0L;
                for (int j = 1; j <= n; ++j) { const long double p3 = p2; p2 = p1; p1 = ((2.0L * j - 1.0L) * z * p2 - (j - 1.0L) * p3) / j; }
                pp = n * (z * p1 - p2) / (z * z - 1.0L);
                const long double dz = p1 / pp;
                z -= dz;
                if (fabsl(dz) < 1e-19L) break;
            }
            x[n - 1 - i] = (double)z; w[n - 1 - i] = (double)(2.0L / ((1.0L - z * z) * pp * pp));
        }
    }
};

// ---- a family of hierarchical 1D basis functions ------------------------------------------------------------------
struct Basis1D
{
    int P = 0;                                      // polynomial degree: P + 1 functions per 1D element
    virtual ~Basis1D() {}
    // AlptBasis::val0..val4 / LagrBasis::val0 / HermBasis::val0..val3: der-th derivative at x, no limit handling
    virtual double raw(double x, int n, int j, int p, int der) const = 0;
    // Basis::val(x, derivative, sgn): one-sided limits are taken RO away from x (source/AlptBasis.cpp:16-29)
    double val(double x, int n, int j, int p, int der, int sgn) const
    {
        double xlim = x;
        if (sgn == -1) { if (std::abs(x) <= RO) return 0.; xlim -= RO; }
        else if (sgn == 1) { if (std::abs(x - 1) <= RO) return 0.; xlim += RO; }
        return raw(xlim, n, j, p, der);
    }
};

// Alpert's multiwavelets (B. Alpert, SIAM J. Math. Anal. 24 (1993)): f_p, p = 0..P, on [-1, 1]
struct AlpertBasis : Basis1D
{
    std::vector<std::vector<double>> w;     // f_p(xi) = sum_i w[p][i] xi^i for xi in [0, 1]
    std::vector<std::vector<double>> leg;   // Legendre polynomial P_p(t) = sum_i leg[p][i] t^i
    explicit AlpertBasis(int P_)
    {
        P = P_;
        const int k = P + 1;
        leg.assign(k, std::vector<double>(k, 0.0));
        leg[0][0] = 1.0;
        if (k > 1) leg[1][1] = 1.0;
        for (int n = 1; n + 1 < k; ++n)
            for (int i = 0; i < k; ++i)
                leg[n + 1][i] = ((2.0 * n + 1.0) * (i > 0 ? leg[n][i - 1] : 0.0) - n * leg[n - 1][i]) / (n + 1.0);
        // top-down: f_p is the null vector of its vanishing-moment conditions and the orthogonality to the f_q (q > p) of equal parity
        std::vector<std::vector<long double>> W(k, std::vector<long double>(k, 0.0L));
        for (int p = k - 1; p >= 0; --p)
        {
            std::vector<std::vector<long double>> rows;
            for (int i = p + k - 2; i >= 0; i -= 2)
            {
                std::vector<long double> r(k);
                for (int m = 0; m < k; ++m) r[m] = 1.0L / (m + i + 1);
                rows.push_back(r);
            }
            for (int q = p + 2; q < k; q += 2)
            {
                std::vector<long double> r(k, 0.0L);
                for (int m = 0; m < k; ++m) for (int l = 0; l < k; ++l) r[m] += W[q][l] / (m + l + 1);
                rows.push_back(r);
            }
            // k - 1 conditions on k coefficients: fix one coefficient to 1 (try each until the system is regular) and solve for the rest
            std::vector<long double> c(k, 0.0L);
            bool ok = false;
            for (int fix = k - 1; fix >= 0 && !ok; --fix)
            {
                if (k == 1) { c[0] = 1.0L; ok = true; break; }
                std::vector<std::vector<long double>> A(k - 1, std::vector<long double>(k - 1));
                std::vector<long double> b(k - 1), x;
                for (int r = 0; r < k - 1; ++r)
                {
                    int cc = 0;
                    for (int m = 0; m < k; ++m) { if (m == fix) continue; A[r][cc++] = rows[r][m]; }
                    b[r] = -rows[r][fix];
                }
                if (!solve(A, b, x)) continue;
                int cc = 0;
                for (int m = 0; m < k; ++m) c[m] = (m == fix) ? 1.0L : x[cc++];
                ok = true;
            }
            long double nrm = 0.0L, at1 = 0.0L;
            for (int m = 0; m < k; ++m) { at1 += c[m]; for (int l = 0; l < k; ++l) nrm += c[m] * c[l] / (m + l + 1); }
            const long double s = (at1 < 0 ? -1.0L : 1.0L) / sqrtl(2.0L * nrm);     // int_0^1 f^2 = 1/2, f(1-) > 0
            for (int m = 0; m < k; ++m) W[p][m] = c[m] * s;
        }
        w.assign(k, std::vector<double>(k));
        for (int p = 0; p < k; ++p) for (int m = 0; m < k; ++m) w[p][m] = (double)W[p][m];
    }
    // AlptBasis::phi (source/AlptBasis.cpp:230-520): odd / even extension to xi < 0
    double phi(double xi, int p, int der) const
    {
        if (xi < 0)
        {
            const double sgn = ((p + P + 1 + der) % 2 != 0) ? -1. : 1.;
            return sgn * horner(w[p], -xi, der);
        }
        return horner(w[p], xi, der);
    }
    // 2^(m/2): the reference's pow(2, .) of an integer or half-integer exponent (exact powers of two times sqrt(2))
    static double pow2_half(int m) { return std::ldexp((m & 1) ? 1.4142135623730951 : 1.0, (m - (m & 1)) / 2); }
    double raw(double x, int n, int j, int p, int der) const override
    {
        if (n == 0) return horner(leg[p], 2 * x - 1, der) * std::ldexp(1.0, der) * std::sqrt(2 * p + 1.);     // :63-69, 88-92
        if (n == 1) { if (x >= 0 && x <= 1) return pow2_half(1 + 2 * der) * phi(2 * x - 1, p, der); return 0.; }
        const int j_odd = (j - 1) / 2;
        const double shift_x = std::ldexp(1.0, n - 1) * x - j_odd;
        return pow2_half((n - 1) * (1 + 2 * der)) * raw(shift_x, 1, 0, p, der);
    }
};

// interpolation point sets on [-1, 1]: value = num / den, shifted by side * RO (source/LagrBasis.cpp:33-154)
struct Pt { double num, den; int side; };
inline bool lagrange_points(int P, int msh_case, std::vector<Pt> & m0, std::vector<Pt> & m1)
{
    m0.clear(); m1.clear();
    auto S = [](std::vector<Pt> & v, std::initializer_list<Pt> l) { v.assign(l); };
    if (P == 1 && msh_case == 1) { S(m0, { {-1, 3, 0}, {1, 3, 0} }); S(m1, { {-2, 3, 0}, {2, 3, 0} }); }
    else if (P == 1 && msh_case == 2) { S(m0, { {-1, 1, 1}, {1, 1, -1} }); S(m1, { {0, 1, -1}, {0, 1, 1} }); }
    else if (P == 2 && msh_case == 1) { S(m0, { {-2, 3, 0}, {-1, 3, 0}, {1, 3, 0} }); S(m1, { {-5, 6, 0}, {1, 6, 0}, {2, 3, 0} }); }
    else if (P == 2 && msh_case == 2) { S(m0, { {-1, 1, 1}, {0, 1, -1}, {1, 1, -1} }); S(m1, { {-0.5, 1, -1}, {0, 1, 1}, {0.5, 1, -1} }); }
    else if (P == 3 && msh_case == 1) { S(m0, { {-3, 5, 0}, {-1, 5, 0}, {1, 5, 0}, {3, 5, 0} }); S(m1, { {-4, 5, 0}, {-2, 5, 0}, {2, 5, 0}, {4, 5, 0} }); }
    else if (P == 3 && msh_case == 2) { S(m0, { {-1, 1, 1}, {-1, 2, -1}, {0, 1, -1}, {1, 1, -1} }); S(m1, { {-3, 4, -1}, {0, 1, 1}, {1, 4, -1}, {1, 2, -1} }); }
    else if (P == 3 && msh_case == 3) { S(m0, { {-1, 1, 1}, {-1, 3, 0}, {1, 3, 0}, {1, 1, -1} }); S(m1, { {-2, 3, 0}, {0, 1, -1}, {0, 1, 1}, {2, 3, 0} }); }
    else if (P == 4 && msh_case == 1) { S(m0, { {-2, 3, 0}, {-5, 12, 0}, {-1, 3, 0}, {1, 6, 0}, {1, 3, 0} }); S(m1, { {-5, 6, 0}, {-17, 24, 0}, {7, 24, 0}, {7, 12, 0}, {2, 3, 0} }); }
    else if (P == 4 && msh_case == 2) { S(m0, { {-1, 1, 1}, {-0.5, 1, -1}, {0, 1, -1}, {0.5, 1, -1}, {1, 1, -1} }); S(m1, { {-0.75, 1, -1}, {-0.25, 1, -1}, {0, 1, 1}, {0.25, 1, -1}, {0.75, 1, -1} }); }
    else if (P == 5 && msh_case == 1) { S(m0, { {-1, 1, 1}, {-0.6, 1, 0}, {-0.2, 1, 0}, {0.2, 1, 0}, {0.6, 1, 0}, {1, 1, -1} }); S(m1, { {-0.8, 1, 0}, {-0.4, 1, 0}, {0, 1, -1}, {0, 1, 1}, {0.4, 1, 0}, {0.8, 1, 0} }); }
    else if (P == 5 && msh_case == 2) { S(m0, { {-5, 6, 0}, {-2, 3, 0}, {-5, 12, 0}, {-1, 3, 0}, {1, 6, 0}, {1, 3, 0} }); S(m1, { {-17, 24, 0}, {-11, 12, 0}, {7, 24, 0}, {7, 12, 0}, {2, 3, 0}, {1, 12, 0} }); }
    else return false;
    return true;
}
inline double pt_exact(const Pt & p) { return p.num / p.den; }
inline double pt_shifted01(const Pt & p) { const double v = p.side == 0 ? pt_exact(p) : (p.side > 0 ? pt_exact(p) + RO : pt_exact(p) - RO); return (v + 1) / 2.; }   // then mapped to [0, 1] (:149-153)
inline int pt_half(const Pt & p) { const double v = pt_exact(p) + p.side * RO; return v < 0 ? 0 : 1; }

// a family with interpolation points: Lagrange and Hermite
struct IntpBasis : Basis1D
{
    std::vector<double> msh0, msh1;                 // intp_msh0 / intp_msh1 on [0, 1]
    // interpolation point of (n, j, p): LagrBasis / HermBasis constructor (source/LagrBasis.cpp:19-28), same floating-point expressions
    double point(int n, int j, int p) const
    {
        if (n == 0) return msh0[p];
        double xl = 0., xr = 1.;
        if (n > 1) { xl = std::pow(2., -n + 1.) * (j - 1.) / 2.; xr = std::pow(2., -n + 1.) * (j + 1.) / 2.; }
        const double h = xr - xl;
        return xl + h * msh1[p];
    }
    // derivative order of dof p (Hermite: value, value, first, first, second, second; Lagrange: 0)
    virtual int dof_order(int) const { return 0; }
};

struct LagrangeBasis : IntpBasis
{
    // per function: nodes (on [-1, 1], exact), own node index; level 1 also the half it lives on
    struct Fn { std::vector<double> nodes; int own; int half; };
    std::vector<Fn> f0, f1;
    bool ok = false;
    LagrangeBasis(int P_, int msh_case)
    {
        P = P_;
        std::vector<Pt> m0, m1;
        if (!lagrange_points(P, msh_case, m0, m1)) return;
        for (auto & p : m0) msh0.push_back(pt_shifted01(p));
        for (auto & p : m1) msh1.push_back(pt_shifted01(p));
        for (int p = 0; p <= P; ++p)
        {
            Fn f; f.own = p; f.half = -1;
            for (auto & q : m0) f.nodes.push_back(pt_exact(q));
            f0.push_back(f);
        }
        for (int p = 0; p <= P; ++p)
        {
            Fn f; f.half = pt_half(m1[p]); f.own = -1;
            for (auto & q : m0) if (pt_half(q) == f.half) f.nodes.push_back(pt_exact(q));
            for (int r = 0; r <= P; ++r) if (pt_half(m1[r]) == f.half) { if (r == p) f.own = (int)f.nodes.size(); f.nodes.push_back(pt_exact(m1[r])); }
            if ((int)f.nodes.size() != P + 1) return;        // not a hierarchical point set
            f1.push_back(f);
        }
        ok = true;
    }
    static double lagr(const Fn & f, double x)
    {
        double v = 1.0;
        for (int m = 0; m < (int)f.nodes.size(); ++m) if (m != f.own) v *= (x - f.nodes[m]) / (f.nodes[f.own] - f.nodes[m]);
        return v;
    }
    // LagrBasis::phi (source/LagrBasis.cpp:190-850)
    double phi(double xt, int msh, int p) const
    {
        const double x = 2 * xt - 1;
        if (x < -1 || x > 1) return 0.;
        if (msh == 0) return lagr(f0[p], x);
        const Fn & f = f1[p];
        if (f.half == 0 ? (x > 0) : (x < 0)) return 0.;
        return lagr(f, x);
    }
    double raw(double x, int n, int j, int p, int der) const override
    {
        if (der != 0) return 0.;                                  // LagrBasis::val has no derivatives (:172-176)
        if (n <= 1) return phi(x, n, p);
        const int odd_j = (j - 1) / 2;
        return phi(std::ldexp(1.0, n - 1) * x - odd_j, 1, p);
    }
};

struct HermiteBasis : IntpBasis
{
    int L = 1;                                       // highest derivative interpolated: P = 2 L + 1
    std::vector<std::vector<double>> c0, c1;          // monomial coefficients (in x on [0,1]) of the level-0 / level-1 functions
    bool ok = false;
    explicit HermiteBasis(int P_)
    {
        P = P_;
        if (P != 3 && P != 5) return;
        L = (P - 1) / 2;
        // HermBasis::set_interp_msh01 (source/HermBasis.cpp:41-69): both points carry all derivative orders
        for (int p = 0; p <= P; ++p)
        {
            const int q = p % 2;
            msh0.push_back(((q == 0 ? -1.0 + RO : 1.0 - RO) + 1) / 2.);
            msh1.push_back(((q == 0 ? 0.0 - RO : 0.0 + RO) + 1) / 2.);
        }
        for (int p = 0; p <= P; ++p) { c0.push_back(hermite(0.0, 1.0, p % 2, p / 2)); }
        for (int p = 0; p <= P; ++p) { const int q = p % 2; c1.push_back(q == 0 ? hermite(0.0, 0.5, 1, p / 2) : hermite(0.5, 1.0, 0, p / 2)); }
        ok = true;
    }
    int dof_order(int p) const override { return p / 2; }      // HermBasis::deg_pt_deri_1d (:111-155)
    // polynomial of degree 2L+1 on [a, b] whose derivatives of order 0..L vanish at both ends except order l at end `node` (0: a, 1: b), which is 1
    std::vector<double> hermite(double a, double b, int node, int l) const
    {
        const int n = 2 * L + 2;
        std::vector<std::vector<long double>> A(n, std::vector<long double>(n, 0.0L));
        std::vector<long double> rhs(n, 0.0L), x;
        int r = 0;
        for (int e = 0; e < 2; ++e)
            for (int m = 0; m <= L; ++m, ++r)
            {
                const long double xe = e == 0 ? a : b;
                for (int i = m; i < n; ++i)
                {
                    long double f = 1.0L;
                    for (int k = 0; k < m; ++k) f *= (i - k);
                    A[r][i] = f * powl(xe, i - m);
                }
                if (e == node && m == l) rhs[r] = 1.0L;
            }
        solve(A, rhs, x);
        std::vector<double> c(n);
        for (int i = 0; i < n; ++i) c[i] = (double)x[i];
        return c;
    }
    double raw(double x, int n, int j, int p, int der) const override
    {
        if (n <= 1)
        {
            if (x < 0 || x > 1) return 0.;
            if (n == 0) return horner(c0[p], x, der);
            const int q = p % 2;
            if (q == 0 ? (x > 0.5) : (x < 0.5)) return 0.;
            return horner(c1[p], x, der);
        }
        // level >= 2: the level-1 function on the rescaled support; derivative dofs are scaled so that they stay unit derivatives in x
        // (source/HermBasis.cpp:158-345)
        const int odd_j = (j - 1) / 2;
        const double xtrans = std::ldexp(1.0, n - 1) * x - odd_j;
        return raw(xtrans, 1, 0, p, der) * std::ldexp(1.0, -(n - 1) * (dof_order(p) - der));
    }
};

// ---- 1D element geometry --------------------------------------------------------------------------------------------
struct Elem1D { int n, j; double xl, xr, dis[3]; };
inline Elem1D elem_of_order(int o)
{
    Elem1D e;
    e.n = level_of_order(o);
    e.j = e.n == 0 ? 1 : 2 * (o - (1 << (e.n - 1))) + 1;
    support(e.n, e.j, e.xl, e.xr);
    e.dis[0] = e.xl; e.dis[1] = (e.xl + e.xr) / 2; e.dis[2] = e.xr;
    return e;
}

// ---- inner products of Basis (source/Basis.cpp) ------------------------------------------------------------------------
enum Boundary { PERIOD = 0, ZERO = 1, INSIDE = 2 };     // boundary_type "period" / "zero" / "inside" of Basis::product_edge_dis_v / _u
struct Products
{
    const Basis1D & U; const Basis1D & V; Gauss g; int boundary = PERIOD;
    // the reference uses 10 points per panel; any rule exact for degree P_u + P_v gives the same integrals up to rounding
    Products(const Basis1D & u, const Basis1D & v, int boundary_ = PERIOD) : U(u), V(v), g(10), boundary(boundary_) {}
    double gl(const Elem1D & eu, int pu, int du, const Elem1D & ev, int pv, int dv, double tl, double tr) const
    {
        double s = 0;
        for (size_t i = 0; i < g.x.size(); ++i)
        {
            const double x = tl + (g.x[i] + 1.) * (tr - tl) / 2.;
            s += g.w[i] * U.val(x, eu.n, eu.j, pu, du, 0) * V.val(x, ev.n, ev.j, pv, dv, 0);
        }
        return s * (tr - tl) / 2.;
    }
    // Basis::product_volume (:62-83)
    double volume(const Elem1D & eu, int pu, int du, const Elem1D & ev, int pv, int dv) const
    {
        if ((ev.xl >= eu.xr) || (ev.xr <= eu.xl)) return 0.;
        double xl = eu.xl, xr = eu.xr;
        if (eu.n <= ev.n) { xl = ev.xl; xr = ev.xr; }
        const double xm = (xl + xr) / 2.;
        return gl(eu, pu, du, ev, pv, dv, xl, xm) + gl(eu, pu, du, ev, pv, dv, xm, xr);
    }
    // Basis::product_edge_dis_v / product_edge_dis_u, boundary type "period" (:145-173, 207-235): sum over the discontinuity points of v (of u)
    double edge(bool over_u, const Elem1D & eu, int pu, int su, int du, const Elem1D & ev, int pv, int sv, int dv) const
    {
        const Elem1D & e = over_u ? eu : ev;
        double s = 0;
        for (int i = 0; i < 3; ++i)
        {
            const double pt = e.dis[i];
            if (boundary == ZERO) { s += U.val(pt, eu.n, eu.j, pu, du, su) * V.val(pt, ev.n, ev.j, pv, dv, sv); continue; }                     // :131-137
            if (boundary == INSIDE)                                                                                                                     // :139-147
            {
                if (std::abs(pt - 0.) < RO || std::abs(pt - 1.) < RO) continue;
                s += U.val(pt, eu.n, eu.j, pu, du, su) * V.val(pt, ev.n, ev.j, pv, dv, sv);
                continue;
            }
            if (e.n <= 1 && i == 2) continue;
            double vu = U.val(pt, eu.n, eu.j, pu, du, su), vv = V.val(pt, ev.n, ev.j, pv, dv, sv);
            if (std::abs(pt - 0.) < RO)
            {
                if (su == -1) vu = U.val(1., eu.n, eu.j, pu, du, su);
                if (sv == -1) vv = V.val(1., ev.n, ev.j, pv, dv, sv);
            }
            else if (std::abs(pt - 1.) < RO)
            {
                if (su == 1) vu = U.val(0., eu.n, eu.j, pu, du, su);
                if (sv == 1) vv = V.val(0., ev.n, ev.j, pv, dv, sv);
            }
            s += vu * vv;
        }
        return s;
    }
    // one entry table.at(u, v) of OperatorMatrix1D<U, V> (include/OperatorMatrix1D.h:166-262)
    double entry(int table, const Elem1D & eu, int pu, const Elem1D & ev, int pv) const
    {
        auto jv = [&](int su, int du) { return edge(false, eu, pu, su, du, ev, pv, 1, 0) - edge(false, eu, pu, su, du, ev, pv, -1, 0); };   // u^{su} [v]
        switch (table)
        {
        case U_V: return volume(eu, pu, 0, ev, pv, 0);
        case U_VX: return volume(eu, pu, 0, ev, pv, 1);
        case UX_VX: return volume(eu, pu, 1, ev, pv, 1);
        case ULFT_VJP: return jv(-1, 0);
        case URGT_VJP: return jv(1, 0);
        case UJP_VJP: return jv(1, 0) - jv(-1, 0);
        case UAVE_VJP: return (jv(1, 0) + jv(-1, 0)) / 2.;
        case UJP_VXLFT: return edge(true, eu, pu, 1, 0, ev, pv, -1, 1) - edge(true, eu, pu, -1, 0, ev, pv, -1, 1);
        case UJP_VXRGT: return edge(true, eu, pu, 1, 0, ev, pv, 1, 1) - edge(true, eu, pu, -1, 0, ev, pv, 1, 1);
        case UXAVE_VJP: return (jv(-1, 1) + jv(1, 1)) / 2.;
        default: return 0.;
        }
    }
};

// the blocks of different pairs are independent: host threads over ranges of pairs
template <class F>
inline void parallel_pairs(int n, F body)
{
    const int nt = (int)std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    if (n < 256 || nt == 1) { for (int p = 0; p < n; ++p) body(p); return; }
    std::vector<std::thread> th;
    for (int t = 0; t < nt; ++t)
        th.emplace_back([=, &body]() { for (int p = (int)((int64_t)n * t / nt); p < (int)((int64_t)n * (t + 1) / nt); ++p) body(p); });
    for (auto & x : th) x.join();
}

// blocks[n_pairs][ku][kv] of one table of OperatorMatrix1D<U, V>, pairs in the canonical order (source element carries U, target element V)
inline void operator_blocks(const Pairs1D & P1, const Basis1D & U, const Basis1D & V, int table, std::vector<double> & blocks, int boundary = PERIOD)
{
    const int ku = U.P + 1, kv = V.P + 1;
    blocks.assign((size_t)P1.n_pairs * ku * kv, 0.0);
    // transposed tables: ux_v.at(v, u) = u_vx.at(u, v), ujp_vxave.at(v, u) = uxave_vjp.at(u, v) (both square, U == V)
    const bool transposed = (table == UX_V || table == UJP_VXAVE);
    const int base = table == UX_V ? (int)U_VX : (table == UJP_VXAVE ? (int)UXAVE_VJP : table);
    const Products pr(U, V, boundary);
    std::vector<Elem1D> el(P1.T);
    for (int o = 0; o < P1.T; ++o) el[o] = elem_of_order(o);
    parallel_pairs(P1.n_pairs, [&](int p)
    {
        const Elem1D & ef = el[P1.src[p]]; const Elem1D & ee = el[P1.tgt[p]];
        const bool vol_only = (base == U_V || base == U_VX || base == UX_VX);
        if (vol_only && !P1.vol[p]) return;
        for (int k = 0; k < ku; ++k)
            for (int q = 0; q < kv; ++q)
                blocks[((size_t)p * ku + k) * kv + q] = transposed ? pr.entry(base, ee, q, ef, k) : pr.entry(base, ef, k, ee, q);
    });
}

// the transposed point table of FastLagrIntp / FastHermIntp (source/FastMultiplyLU.cpp:1316-1360 over eval_Lag_pt_at_Alpt_1D / eval_Her_pt_at_Alpt_1D):
// blocks[pair (Alpert element f -> point element e)][k][q] = (d/dx)^m alpert_{f,k}(point_{e,q}) if the point lies in the closed support of f,
// m = derivative for Lagrange points (0: Lag_pt_Alpt_1D, 1: Lag_pt_Alpt_1D_d1), m = the dof's own order for Hermite points
inline void point_blocks(const Pairs1D & P1, const AlpertBasis & A, const IntpBasis & I, int derivative, std::vector<double> & blocks)
{
    const int ka = A.P + 1, kb = I.P + 1;
    blocks.assign((size_t)P1.n_pairs * ka * kb, 0.0);
    std::vector<Elem1D> el(P1.T);
    for (int o = 0; o < P1.T; ++o) el[o] = elem_of_order(o);
    parallel_pairs(P1.n_pairs, [&](int p)
    {
        const Elem1D & ef = el[P1.src[p]]; const Elem1D & ee = el[P1.tgt[p]];
        for (int q = 0; q < kb; ++q)
        {
            const double pos = I.point(ee.n, ee.j, q);
            if (pos < ef.xl || pos > ef.xr) continue;
            const int m = derivative + I.dof_order(q);
            for (int k = 0; k < ka; ++k) blocks[((size_t)p * ka + k) * kb + q] = A.val(pos, ef.n, ef.j, k, m, 0);
        }
    });
}

// pts1d[T * (P+1)]: interpolation point of every 1D basis function (LagrBasis::intep_pt, source/LagrBasis.cpp:19-28)
inline void point_table(int nmax, const IntpBasis & I, std::vector<double> & pts)
{
    const int T = 1 << nmax, kb = I.P + 1;
    pts.resize((size_t)T * kb);
    for (int o = 0; o < T; ++o) { const Elem1D e = elem_of_order(o); for (int q = 0; q < kb; ++q) pts[(size_t)o * kb + q] = I.point(e.n, e.j, q); }
}

// hierarchisation stencils pwts of every 1D element of level > 0 (LagrInterpolation::set_pts_wts_1d_ada_Lag, source/Interplation.cpp:775-887;
// HermInterpolation::set_pts_wts_1d_ada_Her, :3166-3315): anc[T-1][P1][2] = (ancestor ord1d, point index), wt[T-1][P1][P1] = wt[p0][ic]
inline bool hier_stencils(int nmax, const IntpBasis & I, bool hermite, std::vector<int> & anc, std::vector<double> & wt)
{
    const int T = 1 << nmax, P1 = I.P + 1;
    anc.assign((size_t)(T - 1) * P1 * 2, 0); wt.assign((size_t)(T - 1) * P1 * P1, 0.0);
    for (int o = 1; o < T; ++o)
    {
        const Elem1D e = elem_of_order(o);
        std::vector<double> p_pos(P1); std::vector<int> p_ord(P1), p_num(P1);
        int ic = 0;
        for (int k = 0; k < e.n && ic < P1; ++k)
            for (int i = 1; i < std::max(2, 1 << k); i += 2)
                for (int q0 = 0; q0 < P1; ++q0)
                {
                    const double pos1 = I.point(k, i, q0);
                    if (pos1 > e.xl && pos1 < e.xr)
                    {
                        if (ic == P1) return false;
                        p_pos[ic] = pos1; p_ord[ic] = order_elem(k, i); p_num[ic] = q0; ++ic;
                    }
                }
        if (ic != P1) return false;
        for (int c = 0; c < P1; ++c) { anc[((size_t)(o - 1) * P1 + c) * 2] = p_ord[c]; anc[((size_t)(o - 1) * P1 + c) * 2 + 1] = p_num[c]; }
        std::vector<double> pos(p_pos); std::sort(pos.begin(), pos.end());
        for (int p0 = 0; p0 < P1; ++p0)
        {
            int i10 = 0, i20 = 1;
            for (int c = 0; c < P1; ++c)
            {
                int ic0 = 0; double scale = 1.0; int l1 = 0;
                if (!hermite) { for (int i2 = 0; i2 < P1; ++i2) if (std::fabs(p_pos[c] - pos[i2]) < 1.0e-15) ic0 = i2; }
                else
                {
                    if (std::fabs(p_pos[c] - pos[0]) < 1.0e-14) { ic0 = i10; i10 += 2; } else { ic0 = i20; i20 += 2; }
                    l1 = I.dof_order(p0);
                    scale = std::pow(2, (e.n - 1) * (l1 - I.dof_order(ic0)));
                }
                wt[((size_t)(o - 1) * P1 + p0) * P1 + c] = -1.0 * scale * I.val(I.msh1[p0], 0, 1, ic0, l1, 0);
            }
        }
    }
    return true;
}

}  // namespace tab
}  // namespace amdg
