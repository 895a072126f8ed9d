"""The library's own 1D tables (csrc/tables.hpp through amdg_op_generate / _points / _hier / amdg_points_generate) against the
reference's: every OperatorMatrix1D table, the point tables Lag_pt_Alpt_1D / _d1 / Her_pt_Alpt_1D, the interpolation point coordinates
(bit exact) and the hierarchisation stencils dumped by the compiled reference (tests/golden/*.dump.xz and the benchmark-size bundles
tests/golden/tables/*.npz).  Host only: no GPU needed."""
import os

import numpy as np
import pytest

from conftest import ROOT, golden_names, load_golden

TOL = 1e-12          # relative to the largest entry of the table; observed <= 2e-13 (quintic Hermite), typically 1e-15


def _rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def _msh_case(name):
    return 2 if name == "vlasov_d4_k3_m4_n2" else 1     # tests/golden/make_golden.py: --msh-lagr 2


def _boundary(name):
    return "zero" if "_bc_zero_" in name else ("inside" if "_bc_inside_" in name else "period")


@pytest.mark.parametrize("name", [n for n in golden_names() if "alpt.u_v" in load_golden(n)])
def test_generated_tables_match_reference_dump(amdg, name):
    d = load_golden(name)
    dim, nmax, n0, sparse, pa, pl, ph, vecnum, herm, ne = [int(x) for x in d["config"]]
    a, msh, bc = pa + 1, _msh_case(name), _boundary(name)
    ctx = amdg.Context(1, nmax, pa, pl, device=-1)
    fam = {"alpt": (amdg.BASIS_ALPERT, pa, 1), "lagr": (amdg.BASIS_LAGRANGE, pl, msh), "herm": (amdg.BASIS_HERMITE, ph, 1)}
    checked = 0
    for key in d:
        pre, _, tab = key.partition(".")
        if pre in fam and tab in amdg.TABLES:
            basis, p, m = fam[pre]
            if pre == "herm" and p not in (3, 5):
                continue
            gen = ctx.op_blocks(ctx.op_generate(basis, p, tab, m, bc), p + 1, a)
            ref = ctx.op_blocks(ctx.op_register(d[key], p + 1, a), p + 1, a)
            assert _rel(gen, ref) < TOL, key
            checked += 1
    assert checked >= 12
    for der, key in ((0, "Lag_pt_Alpt_1D"), (1, "Lag_pt_Alpt_1D_d1")):
        gen = ctx.op_blocks(ctx.op_generate_points(amdg.BASIS_LAGRANGE, pl, msh, der), a, pl + 1)
        ref = ctx.op_blocks(ctx.op_register(d[key].T.copy(), a, pl + 1), a, pl + 1)
        assert _rel(gen, ref) < TOL, key
    assert np.array_equal(ctx.points_generate(amdg.BASIS_LAGRANGE, pl, msh), d["lagr.intep_pt"])            # same floating-point expressions
    gen = ctx.op_blocks(ctx.op_generate_hier(amdg.BASIS_LAGRANGE, pl, msh), pl + 1, pl + 1)
    ref = ctx.op_blocks(ctx.op_register_hier(d["lagr.pw_anc"], d["lagr.pw_wt"]), pl + 1, pl + 1)
    assert _rel(gen, ref) < TOL
    if ph in (3, 5):
        gen = ctx.op_blocks(ctx.op_generate_points(amdg.BASIS_HERMITE, ph), a, ph + 1)
        ref = ctx.op_blocks(ctx.op_register(d["Her_pt_Alpt_1D"].T.copy(), a, ph + 1), a, ph + 1)
        assert _rel(gen, ref) < TOL
        assert np.array_equal(ctx.points_generate(amdg.BASIS_HERMITE, ph), d["herm.intep_pt"])
        gen = ctx.op_blocks(ctx.op_generate_hier(amdg.BASIS_HERMITE, ph), ph + 1, ph + 1)
        ref = ctx.op_blocks(ctx.op_register_hier(d["herm.pw_anc"], d["herm.pw_wt"]), ph + 1, ph + 1)
        assert _rel(gen, ref) < TOL
    ctx.close()


@pytest.mark.parametrize("k,m,nmax", [(1, 2, 7), (2, 3, 7), (3, 3, 8)])
def test_generated_tables_match_reference_bundles(amdg, k, m, nmax):
    """the benchmark sizes: bundles made from the compiled reference by tests/golden/tables/make_tables.py"""
    ref = np.load(os.path.join(ROOT, "tests", "golden", "tables", "tables_k%d_m%d_n%d.npz" % (k, m, nmax)))
    gen = amdg.generate_tables(nmax, k, m)
    keys = [key for key in ref if key in gen]
    assert len(keys) >= 14
    for key in keys:
        assert _rel(gen[key], ref[key]) < TOL, key


@pytest.mark.parametrize("P", [0, 1, 2, 3, 4, 5])
def test_alpert_construction_is_orthonormal(amdg, P):
    """the multiwavelets are constructed from their defining properties (the reference hard-codes P <= 4): the mass matrix of the whole
    hierarchical basis is the identity, for every degree the context accepts"""
    ctx = amdg.Context(1, 5, P, max(P, 1), device=-1)
    src, tgt, vol = ctx.pairs()
    uv = ctx.op_blocks(ctx.op_generate(amdg.BASIS_ALPERT, P, "u_v"), P + 1, P + 1)
    expect = np.zeros_like(uv)
    expect[src == tgt] = np.eye(P + 1)
    assert np.abs(uv - expect).max() < 1e-13
    ctx.close()


@pytest.mark.parametrize("P", [0, 1, 2, 3, 4, 5])
def test_alpert_tables_satisfy_integration_by_parts(amdg, P):
    """a check that does not involve the reference (whose hard-coded multiwavelets stop at P = 4): on the periodic interval
        (u, v_x) + (u_x, v) = - sum over all discontinuity points of [[u v]],   [[u v]] = u+ [[v]] + [[u]] v-,
    i.e. for every related pair (f, e) and its transposed partner (e, f):  u_vx[f,e] + u_vx[e,f]^T + urgt_vjp[f,e] + ulft_vjp[e,f]^T = 0.
    The residual is the reference's own convention of taking one-sided limits 1e-13 away from the point (Basis::val), amplified by the derivative
    scale of the fine levels: <= 2e-8 of the largest entry at NMAX = 5, exact for P = 0"""
    ctx = amdg.Context(1, 5, P, max(P, 1), device=-1)
    src, tgt, vol = ctx.pairs()
    gen = lambda t: ctx.op_blocks(ctx.op_generate(amdg.BASIS_ALPERT, P, t), P + 1, P + 1)
    uvx, ur, ul = gen("u_vx"), gen("urgt_vjp"), gen("ulft_vjp")
    pid = {(int(a), int(b)): i for i, (a, b) in enumerate(zip(src, tgt))}
    worst = 0.0
    for i, (a, b) in enumerate(zip(src, tgt)):
        j = pid[(int(b), int(a))]                                     # the relation is symmetric: the transposed pair exists
        r = uvx[i] + uvx[j].T + ur[i] + ul[j].T
        worst = max(worst, float(np.abs(r).max() / max(np.abs(uvx[i]).max(), np.abs(ur[i]).max(), 1.0)))
    assert worst < 1e-7
    if P == 0:
        assert worst == 0.0
    ctx.close()


@pytest.mark.parametrize("basis,P,msh", [("lagr", 1, 1), ("lagr", 1, 2), ("lagr", 2, 1), ("lagr", 2, 2), ("lagr", 3, 1), ("lagr", 3, 2), ("lagr", 3, 3),
                                         ("lagr", 4, 1), ("lagr", 4, 2), ("lagr", 5, 1), ("lagr", 5, 2), ("herm", 3, 1), ("herm", 5, 1)])
def test_generated_stencils_annihilate_polynomials(amdg, basis, P, msh):
    """independent of the reference, for every point set the generator knows: a polynomial of degree <= P is reproduced by the level-0 interpolant,
    so its hierarchical surplus (I + W applied to its point values; for Hermite dofs the point values of its derivatives) vanishes on every 1D element of
    level >= 1 and equals the point values on level 0"""
    nmax = 5
    b = amdg.BASIS_LAGRANGE if basis == "lagr" else amdg.BASIS_HERMITE
    ctx = amdg.Context(1, nmax, 1, P, device=-1)
    src, tgt, vol = ctx.pairs()
    pts = ctx.points_generate(b, P, msh).reshape(-1, P + 1)
    B = ctx.op_blocks(ctx.op_generate_hier(b, P, msh), P + 1, P + 1)
    poly = np.polynomial.Polynomial(np.random.default_rng(P * 10 + msh).standard_normal(P + 1))
    v = poly(pts)
    if basis == "herm":                                          # dof p: derivative of order p // 2 at point p % 2 (HermBasis::deg_pt_deri_1d)
        for p in range(2, P + 1):
            v[:, p] = poly.deriv(p // 2)(pts[:, p])
    c = np.zeros_like(v)
    for i in range(len(src)):
        c[tgt[i]] += v[src[i]] @ B[i]
    level = np.array([0 if o == 0 else int(np.floor(np.log2(o))) + 1 for o in range(1 << nmax)])
    assert np.abs(c[level >= 1]).max() < 1e-11 * np.abs(v).max()
    assert np.array_equal(c[0], v[0])
    ctx.close()


@pytest.mark.parametrize("k,basis,m,msh", [(1, "lagr", 1, 1), (1, "lagr", 1, 2), (1, "lagr", 2, 1), (2, "lagr", 2, 2), (2, "lagr", 3, 1), (2, "lagr", 3, 2),
                                           (3, "lagr", 3, 3), (2, "lagr", 4, 1), (3, "lagr", 4, 2), (3, "lagr", 5, 1), (4, "lagr", 5, 2), (5, "lagr", 5, 1),
                                           (1, "herm", 3, 1), (3, "herm", 3, 1), (2, "herm", 5, 1), (5, "herm", 5, 1)])
def test_generated_tables_round_trip_in_1d(amdg, k, basis, m, msh):
    """the three generated table families together, independent of the reference and for combinations it never dumped (Alpert degree 5 included):
    Alpert coefficients -> values at the interpolation points (point table) -> hierarchical interpolation coefficients (stencils) -> L2 projection
    back (u_v of the interpolation basis against the Alpert basis) is the identity on the full 1D grid whenever m >= k.  The residual comes from the
    interface points sitting 1e-13 off the cell boundaries (the reference's convention), amplified on the fine levels"""
    nmax = 4
    b_id = amdg.BASIS_LAGRANGE if basis == "lagr" else amdg.BASIS_HERMITE
    ctx = amdg.Context(1, nmax, k, m, device=-1)
    src, tgt, vol = ctx.pairs()
    a, b = k + 1, m + 1
    PT = ctx.op_blocks(ctx.op_generate_points(b_id, m, msh), a, b)
    H = ctx.op_blocks(ctx.op_generate_hier(b_id, m, msh), b, b)
    UV = ctx.op_blocks(ctx.op_generate(b_id, m, "u_v", msh), b, a)
    T = 1 << nmax

    def apply(x, B, kt):
        y = np.zeros((T, kt))
        for i in range(len(src)):
            if vol[i]:
                y[tgt[i]] += x[src[i]] @ B[i]
        return y
    u = np.random.default_rng(3).standard_normal((T, a))
    u2 = apply(apply(apply(u, PT, b), H, b), UV, a)
    assert np.abs(u2 - u).max() < 1e-9 * np.abs(u).max()
    ctx.close()


def test_table_generator_rejects_bad_requests(amdg):
    ctx = amdg.Context(1, 3, 2, 3, device=-1)
    with pytest.raises(amdg.AmdgError, match="Alpert x Alpert only"):
        ctx.op_generate(amdg.BASIS_LAGRANGE, 3, "ux_vx")
    with pytest.raises(amdg.AmdgError, match="no Lagrange point set"):
        ctx.op_generate(amdg.BASIS_LAGRANGE, 3, "u_v", 7)
    with pytest.raises(amdg.AmdgError, match="pmax 3 or 5"):
        ctx.op_generate_points(amdg.BASIS_HERMITE, 4)
    with pytest.raises(amdg.AmdgError):
        ctx.op_generate(amdg.BASIS_ALPERT, 2, 99)
    ctx.close()
