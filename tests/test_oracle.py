"""The numpy restatement (oracle/amdg_oracle.py) against the outputs of the compiled reference (tests/golden/).
CPU only.  Tolerance: the reference's own summation order is address dependent, so floating
point phases are compared at 1e-12 relative L2 (north_star's per-stage bound); index tables are exact."""
import numpy as np
import pytest

import amdg_oracle as O
import refdump
from conftest import golden_names, load_golden

TOL = 1e-12


def rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


class Case:
    def __init__(self, name):
        d = self.d = load_golden(name)
        (self.dim, self.nmax, self.n0, self.sparse, self.pa, self.pl, self.ph, self.vecnum, self.herm, self.ne) = [int(x) for x in d["config"]]
        self.a = self.pa + 1
        self.b = (self.ph if self.herm else self.pl) + 1
        self.lev, self.sup, self.ord1d = d["level"], d["suppt"], d["order_elem"]
        self.rels = None

    def relations(self):
        if self.rels is None:
            self.rels = {k: [O.relations(self.lev, self.sup, t, k) for t in range(self.dim)] for k in ("vol", "flx")}
        return self.rels


@pytest.mark.parametrize("name", golden_names())
def test_grid_tables(name):
    c = Case(name)
    if name.startswith("adapt_"):
        # grid produced by DGAdapt::refine / coarsen: take the element list from the dump, check the keys
        keys = np.array([O.hash_key(l, s) for l, s in zip(c.lev, c.sup)])
        assert (keys == c.d["hash_key"]).all() and (np.diff(keys) > 0).all()
    else:
        lev, sup = O.sparse_grid(c.dim, c.n0, c.sparse == 1)
        keys = np.array([O.hash_key(l, s) for l, s in zip(lev, sup)])
        o = np.argsort(keys, kind="stable")
        assert len(set(keys.tolist())) == len(keys)
        assert (keys[o] == c.d["hash_key"]).all()
        assert (lev[o] == c.lev).all() and (sup[o] == c.sup).all()
    ord1d = np.array([[O.order_elem(int(n), int(j)) for n, j in zip(l, s)] for l, s in zip(c.lev, c.sup)])
    assert (ord1d == c.ord1d).all()
    rels = c.relations()
    for k in ("vol", "flx"):
        for t in range(c.dim):
            ptr, idx = c.d["%s_d%d_ptr" % (k, t)], c.d["%s_d%d_idx" % (k, t)]
            for e in range(c.ne):
                assert rels[k][t][e] == list(idx[ptr[e]:ptr[e + 1]])


@pytest.mark.parametrize("name", golden_names())
def test_input_field(name):
    c = Case(name)
    f = refdump.field(20240901, c.d["hash_key"], c.lev, c.vecnum, c.a ** c.dim)
    assert np.array_equal(f, c.d["ucoe_alpt.in"])


def _tables(c):
    d = c.d
    if c.herm:
        return d["Her_pt_Alpt_1D"].T.copy(), d["herm.u_v"], d["herm.u_vx"], d["herm.ulft_vjp"] + d["herm.urgt_vjp"], d["herm.pw_anc"], d["herm.pw_wt"]
    return d["Lag_pt_Alpt_1D"].T.copy(), d["lagr.u_v"], d["lagr.u_vx"], d["lagr.ulft_vjp"] + d["lagr.urgt_vjp"], d["lagr.pw_anc"], d["lagr.pw_wt"]


@pytest.mark.parametrize("name", [n for n in golden_names() if "rt." + "up_intp" in load_golden(n)])
def test_roundtrip(name):
    c = Case(name)
    if c.ne * c.b ** c.dim > 3e5:
        pytest.skip("numpy restatement too slow for this fixture; covered by the C-ABI parity tests")
    pt, u_v, _, _, anc, wt = _tables(c)
    rels = c.relations()
    u = c.d["ucoe_alpt.in"][:, 0, :]
    up = O.apply_tensor(u, c.a, c.b, [pt] * c.dim, ["vol"] * c.dim, rels, c.lev, c.ord1d)
    assert rel(up, c.d["rt.up_intp"][:, 0, :]) < TOL
    uc = O.hierarchize(c.d["rt.up_intp"][:, 0, :], c.b, c.lev, c.sup, c.ord1d, anc, wt)
    assert rel(uc, c.d["rt.ucoe_intp"][:, 0, :]) < TOL
    ua = O.apply_tensor(c.d["rt.ucoe_intp"][:, 0, :], c.b, c.a, [u_v] * c.dim, ["vol"] * c.dim, rels, c.lev, c.ord1d)
    assert rel(ua, c.d["rt.ucoe_alpt"][:, 0, :]) < TOL


@pytest.mark.parametrize("name", ["cfg4_burgers_lagr_d2_k2_n4", "kpp_lagr_d2_k1_n4", "full_d2_k2_n3", "line_d1_k2_n5", "adapt_d3_k1_n4"])
def test_nonlinear_rhs_lagrange(name):
    c = Case(name)
    d = c.d
    pt, u_v, u_vx, uave, anc, wt = _tables(c)
    rels = c.relations()
    flux = {"cfg4_burgers_lagr_d2_k2_n4": "burgers", "kpp_lagr_d2_k1_n4": "kpp", "full_d2_k2_n3": "linear", "line_d1_k2_n5": "burgers", "adapt_d3_k1_n4": "kpp"}[name]
    n_flux = 1 if name.startswith("cfg4") else c.dim
    u = d["ucoe_alpt.in"][:, 0, :]
    up = O.apply_tensor(u, c.a, c.b, [pt] * c.dim, ["vol"] * c.dim, rels, c.lev, c.ord1d)
    assert rel(up, d["up_intp"][:, 0, :]) < TOL
    rhs = np.zeros_like(u)
    fuc = []
    for t in range(n_flux):
        fp = O.flux_pointwise(d["up_intp"][:, 0, :], flux, t)
        assert rel(fp, d["fp_intp"][:, 0, t, :]) < 1e-15
        fc = O.hierarchize(fp, c.b, c.lev, c.sup, c.ord1d, anc, wt)
        assert rel(fc, d["fucoe_intp"][:, 0, t, :]) < TOL
        fuc.append(fc)
    # HyperbolicLagrRHS::rhs_vol_scalar reads fucoe_intp[0][t] for every t (also the components the example did
    # not interpolate, which stay zero): take them from the dump
    for t in range(c.dim):
        src = d["fucoe_intp"][:, 0, t, :]
        mats = [u_vx if s == t else u_v for s in range(c.dim)]
        rhs += O.apply_tensor(src, c.b, c.a, mats, ["vol"] * c.dim, rels, c.lev, c.ord1d)
    assert rel(rhs, d["rhs_vol"][:, 0, :]) < TOL
    for t in range(c.dim):
        src = d["fucoe_intp"][:, 0, t, :]
        mats = [uave if s == t else u_v for s in range(c.dim)]
        kinds = ["flx" if s == t else "vol" for s in range(c.dim)]
        rhs += O.apply_tensor(src, c.b, c.a, mats, kinds, rels, c.lev, c.ord1d, 0.5)
    assert rel(rhs, d["rhs_vol_flx"][:, 0, :]) < TOL
    if c.dim == 1:
        # reference quirk: for DIM == 1 the single-matrix form copies twice and never sweeps
        # (source/FastMultiplyLU.cpp:206-212 with is_first_step[0] tested first at :249), so the penalty is a no-op
        assert rel(rhs, d["rhs_all"][:, 0, :]) < TOL
        return
    for t in range(c.dim):
        rhs += O.single_sweep(u, c.a, d["alpt.ujp_vjp"], "flx", rels, c.lev, c.ord1d, t, -1.2 / 2)
    assert rel(rhs, d["rhs_all"][:, 0, :]) < TOL


def test_linear_sweeps_equal_assembled_operator():
    """cfg1 / cfg3: the shipped assembled SpMV equals the sum of single 1D sweeps"""
    c = Case("cfg1_adv_d2_k2_n4")
    d = c.d
    rels = c.relations()
    u = d["ucoe_alpt.in"][:, 0, :]
    rhs = np.zeros_like(u)
    for t in range(c.dim):
        rhs += O.single_sweep(u, c.a, d["alpt.u_vx"], "vol", rels, c.lev, c.ord1d, t, 1.0)
        rhs += O.single_sweep(u, c.a, d["alpt.ulft_vjp"], "flx", rels, c.lev, c.ord1d, t, 1.0)
    assert rel(rhs, d["adv.rhs_sweep"][:, 0, :]) < TOL
    assert rel(rhs, d["adv.rhs_spmv"][:, 0, :]) < TOL
    c = Case("cfg3_wave_d3_k2_n3")
    d = c.d
    rels = c.relations()
    u = d["ucoe_alpt.in"][:, 0, :]
    sigma_dx = 20.0 * 2 ** c.nmax
    rhs = np.zeros_like(u)
    for t in range(c.dim):
        for nm, kind, cf in (("alpt.ux_vx", "vol", -1.0), ("alpt.uxave_vjp", "flx", -1.0), ("alpt.ujp_vxave", "flx", -1.0), ("alpt.ujp_vjp", "flx", -sigma_dx)):
            rhs += O.single_sweep(u, c.a, d[nm], kind, rels, c.lev, c.ord1d, t, cf)
    assert rel(rhs, d["wave.rhs_sweep"][:, 0, :]) < TOL
    assert rel(rhs, d["wave.rhs_spmv"][:, 0, :]) < TOL


def test_rk3ssp_and_schedule():
    orders, lus = O.transform_order(3)
    assert orders == [[0, 1, 2], [0, 2, 1], [1, 2, 0], [2, 0, 1]]
    assert lus == [["L", "L", "full"], ["L", "full", "U"], ["L", "full", "U"], ["full", "U", "U"]]
    u_tn, u, r = np.array([1.0, 2.0]), np.array([0.5, -1.0]), np.array([3.0, 4.0])
    assert np.allclose(O.rk3ssp_stage(1, u_tn, u, r, 0.1), 0.75 * u_tn + 0.25 * (u + 0.1 * r))


def test_rk_schemes_and_rk4_stage_form():
    """step_stage of every explicit scheme; RK4ODE2nd stage form == its step_rk form (the golden wave step)"""
    u_tn, u, r = np.array([1.0, 2.0]), np.array([0.5, -1.0]), np.array([3.0, 4.0])
    assert np.allclose(O.rk_stage("rk2ssp", 1, u_tn, u, r, 0.1), 0.5 * u_tn + 0.5 * (u + 0.1 * r))
    assert np.allclose(O.rk_stage("rk2mid", 0, u_tn, u, r, 0.1), u_tn + 0.05 * r)
    assert np.allclose(O.rk_stage("rk3heun", 1, u_tn, u, r, 0.1), u_tn + 0.05 * r)
    c = Case("cfg3_wave_d3_k2_n3")
    d = c.d
    rels = c.relations()
    sigma_dx = 20.0 * 2 ** c.nmax

    def L(x):
        out = np.zeros_like(x)
        for t in range(c.dim):
            for nm, kind, cf in (("ux_vx", "vol", -1.0), ("uxave_vjp", "flx", -1.0), ("ujp_vxave", "flx", -1.0), ("ujp_vjp", "flx", -sigma_dx)):
                out += O.single_sweep(x, c.a, d["alpt." + nm], kind, rels, c.lev, c.ord1d, t, cf)
        return out
    u0 = d["ucoe_alpt.in"][:, 0, :]
    v0 = refdump.field(20240901 + 1, d["hash_key"], c.lev, c.vecnum, c.a ** c.dim)[:, 0, :]
    u1, v1 = O.rk4_ode2nd_stages(u0, v0, L, 0.0005)
    assert rel(u1, d["wave.ucoe_alpt"][:, 0, :]) < 1e-10 and rel(v1, d["wave.ucoe_ut"][:, 0, :]) < 1e-10


@pytest.mark.parametrize("name", ["variants_lagr_d2_k2_n4", "variants_herm_d2_k2_n4", "variants_lagr_d3_k1_n3"])
def test_rhs_variants(name):
    """HyperbolicSameFlux*/DiffFlux* (DIM == 2) and SourceFastLagr::rhs_source as compositions of the tensor application
    (source/FastMultiplyLU.cpp:970-1123, 1304-1314)"""
    c = Case(name)
    d = c.d
    pt, u_v, u_vx, uave, anc, wt = _tables(c)
    rels = c.relations()
    fuc = d["var.fucoe_intp"]

    def hyperbolic(comp_of_dim):
        rhs = np.zeros((c.ne, c.a ** c.dim))
        for t in range(c.dim):
            src = fuc[:, 0, comp_of_dim(t), :]
            rhs += O.apply_tensor(src, c.b, c.a, [u_vx if s == t else u_v for s in range(c.dim)], ["vol"] * c.dim, rels, c.lev, c.ord1d)
            rhs += O.apply_tensor(src, c.b, c.a, [uave if s == t else u_v for s in range(c.dim)], ["flx" if s == t else "vol" for s in range(c.dim)],
                                  rels, c.lev, c.ord1d, 0.5)
        return rhs
    if "var.rhs_sameflux" in d:
        assert rel(hyperbolic(lambda t: 0), d["var.rhs_sameflux"][:, 0, :]) < TOL
        assert rel(hyperbolic(lambda t: t), d["var.rhs_diffflux"][:, 0, :]) < TOL
    if "var.rhs_source" in d:
        for v in range(c.vecnum):
            out = O.apply_tensor(fuc[:, v, 0, :], c.b, c.a, [u_v] * c.dim, ["vol"] * c.dim, rels, c.lev, c.ord1d)
            assert rel(out, d["var.rhs_source"][:, v, :]) < TOL


def f4_terms(d, dim):
    """the compositions of SURVEY 8(f4) as (dump key, [(flux component, table per dim, relation per dim, coef)], mesh_nmax):
    DiffusionRHS (source/FastMultiplyLU.cpp:1691-1821), FastRHSHamiltonJacobi::rhs_nonlinear (:426-434)"""
    uv, uvx = d["lagr.u_v"], d["lagr.u_vx"]
    jx = d["lagr.ujp_vxlft"] + d["lagr.ujp_vxrgt"]
    def per_dim(table, kind, coef, comp):
        return [(comp(t), [table if s == t else uv for s in range(dim)], [kind if s == t else "vol" for s in range(dim)], coef) for t in range(dim)]
    return {
        "f4.diff_vol": per_dim(uvx, "vol", -1.0, lambda t: t),
        "f4.diff_flx_gradu": per_dim(d["lagr.uave_vjp"], "flx", -1.0, lambda t: t),
        "f4.diff_flx_u": per_dim(jx, "flx", -0.5, lambda t: 0),
        "f4.diff_flx_k_minus_u": per_dim(d["lagr.ujp_vxlft"], "flx", -0.5, lambda t: 0),
        "f4.diff_flx_k_plus_u": per_dim(d["lagr.ujp_vxrgt"], "flx", -0.5, lambda t: 0),
        "f4.hj": [(0, [uv] * dim, ["vol"] * dim, 1.0)],
    }


@pytest.mark.parametrize("name", ["f4_lagr_d2_k2_n4", "f4_lagr_d3_k1_n3"])
def test_f4_compositions(name):
    """DiffusionRHS, FastRHSHamiltonJacobi, the *_coarse_grid transforms and DGAdapt::indicator_norm restated over the oracle's tensor application"""
    c = Case(name)
    d = c.d
    pt, u_v, u_vx, uave, anc, wt = _tables(c)
    rels = c.relations()
    fuc = d["f4.fucoe_intp"]
    for key, terms in f4_terms(d, c.dim).items():
        rhs = np.zeros((c.ne, c.a ** c.dim))
        for comp, mats, kinds, coef in terms:
            rhs += O.apply_tensor(fuc[:, 0, comp, :], c.b, c.a, mats, kinds, rels, c.lev, c.ord1d, coef)
        assert rel(rhs, d[key][:, 0, :]) < TOL, key
    u = d["ucoe_alpt.in"][:, 0, :]
    for cut in (1, 2):
        tag = "f4.cg%d" % cut
        M = int(d[tag + ".mesh_nmax"][0])
        up = O.apply_tensor(u, c.a, c.b, [pt] * c.dim, ["vol"] * c.dim, rels, c.lev, c.ord1d, mesh_nmax=M)
        assert rel(up, d[tag + ".up_intp"][:, 0, :]) < TOL
        assert np.all(up[c.lev.sum(axis=1) > M] == 0) and np.any(c.lev.sum(axis=1) > M)
        rhs = np.zeros((c.ne, c.a ** c.dim))
        for t in range(c.dim):
            rhs += O.apply_tensor(fuc[:, 0, t, :], c.b, c.a, [u_vx if s == t else u_v for s in range(c.dim)], ["vol"] * c.dim, rels, c.lev, c.ord1d, mesh_nmax=M)
        assert rel(rhs, d[tag + ".rhs_vol"][:, 0, :]) < TOL
        for t in range(c.dim):
            rhs += O.apply_tensor(fuc[:, 0, t, :], c.b, c.a, [uave if s == t else u_v for s in range(c.dim)], ["flx" if s == t else "vol" for s in range(c.dim)],
                                  rels, c.lev, c.ord1d, 0.5, mesh_nmax=M)
        assert rel(rhs, d[tag + ".rhs_vol_flx"][:, 0, :]) < TOL
    assert rel(np.linalg.norm(u, axis=1), d["f4.indicator_norm"]) < 1e-14


@pytest.mark.parametrize("name", ["f4_lagr_d2_k2_n4", "adapt_d3_k1_n4"])
def test_coarse_grid_transform_is_the_transform_on_the_sub_grid(name):
    """the argument behind amdg_apply_tensor_coarse: skipping, in every sweep, the targets and sources whose levels sum to more than mesh_nmax
    (FastMultiplyLU::transform_1D_coarse_grid, source/FastMultiplyLU.cpp:514-594) is the plain transform on the sub-grid of the kept elements --
    relations are pairwise and the kept set is closed under taking coarser elements -- with the skipped rows left at zero.  Sparse and adaptive grid"""
    c = Case(name)
    d = c.d
    pt, u_v, u_vx, uave, anc, wt = _tables(c)
    rels = c.relations()
    u = d["ucoe_alpt.in"][:, 0, :]
    for M in (c.nmax - 1, c.nmax - 2):
        keep = np.nonzero(c.lev.sum(axis=1) <= M)[0]
        assert 0 < len(keep) < c.ne
        masked = O.apply_tensor(u, c.a, c.b, [pt] * c.dim, ["vol"] * c.dim, rels, c.lev, c.ord1d, mesh_nmax=M)
        sub_rels = {"vol": [O.relations(c.lev[keep], c.sup[keep], t, "vol") for t in range(c.dim)]}
        sub = O.apply_tensor(u[keep], c.a, c.b, [pt] * c.dim, ["vol"] * c.dim, sub_rels, c.lev[keep], c.ord1d[keep])
        full = np.zeros_like(masked)
        full[keep] = sub
        assert rel(full, masked) < 1e-14
        # and it is NOT the plain transform with the result masked afterwards: values of skipped elements feed kept ones in later sweeps
        plain = O.apply_tensor(u, c.a, c.b, [pt] * c.dim, ["vol"] * c.dim, rels, c.lev, c.ord1d)
        plain[c.lev.sum(axis=1) > M] = 0.0
        assert rel(plain, masked) > 1e-6


@pytest.mark.parametrize("name", ["pw_d2_k2_n4_v2", "pw_vlasov_d4_k1_n3_v2"])
def test_pointwise_bodies_beyond_scalar_flux(name):
    """the reference's point-wise kernels with several unknowns, a coefficient of position and a broadcast field (eval_fp_Lag with VEC_NUM = 2,
    var_coeff_u_Lagr_fast / eval_coe_u_Lag, interp_Vlasov_2D2V over copy_up_intp_to_f) restated; also pins what the GPU test feeds amdg_pointwise_expr"""
    c = Case(name)
    d = c.d
    pt, u_v, u_vx, uave, anc, wt = _tables(c)
    rels = c.relations()
    up = np.stack([O.apply_tensor(d["ucoe_alpt.in"][:, v, :], c.a, c.b, [pt] * c.dim, ["vol"] * c.dim, rels, c.lev, c.ord1d) for v in range(2)], axis=1)
    assert rel(up, d["pw.sys.up_intp"]) < TOL
    f = O.system_flux(up)
    for i in range(2):
        for t in range(c.dim):
            assert rel(f(i, t), d["pw.sys.fp_intp"][:, i, t, :]) < 1e-14
    assert rel(O.hierarchize(f(1, 0), c.b, c.lev, c.sup, c.ord1d, anc, wt), d["pw.sys.fucoe_intp"][:, 1, 0, :]) < TOL
    X = O.point_coordinates(c.ord1d, d["lagr.intep_pt"], c.b)
    for v in range(2):
        for t in range(c.dim):
            assert rel(O.position_coefficient(X, t) * up[:, v], d["pw.coe.fp_intp"][:, v, t, :]) < 1e-13
    if "pw.vl.fp_intp" in d:
        le, se = d["pw.vl.E.level"], d["pw.vl.E.suppt"]
        orde = np.array([[O.order_elem(int(n), int(j)) for n, j in zip(l, s)] for l, s in zip(le, se)])
        rels_e = {"vol": [O.relations(le, se, t, "vol") for t in range(c.dim)]}
        Eup = np.stack([O.apply_tensor(d["pw.vl.E.ucoe_alpt"][:, v, :], c.a, c.b, [pt] * c.dim, ["vol"] * c.dim, rels_e, le, orde) for v in range(2)], axis=1)
        rows = O.field_rows_of(c.lev, c.sup, le, se, (2, 3))
        fpt = d["pw.vl.up_intp"][:, 0, :]
        assert rel(fpt, up[:, 0]) < TOL
        want = [X[..., 2] * fpt, X[..., 3] * fpt, Eup[rows, 0] * fpt, Eup[rows, 1] * fpt]
        for t in range(4):
            assert rel(want[t], d["pw.vl.fp_intp"][:, 0, t, :]) < TOL, t


def test_pointwise_wrappers_gradient_source_and_1d2v_body():
    """more of the reference's wrappers over the same primitives, restated (CPU only -- on the device they are compositions of amdg_apply_tensor with the
    derivative point table, amdg_pointwise_expr with X / OTHER operands, amdg_hierarchize): var_coeff_gradu_Lagr_fast (coefficient times the derivative
    transform of FastLagrIntp::eval_der_up_Lagr, source/Interplation.cpp:4159-4175), source_from_lagr_to_rhs (:4102-4123) and the body of the shipped
    example/07_vlasov_maxwell_sparse.cpp, interp_Vlasov_1D2V with the Maxwell coefficient functions (:4435-4505) and (B3, E1, E2) broadcast by
    DGSolution::copy_up_intp_to_f"""
    c = Case("pw2_vm_d3_k2_n3_v3")
    d = c.d
    pt, u_v, u_vx, uave, anc, wt = _tables(c)
    pt_d1 = d["Lag_pt_Alpt_1D_d1"].T.copy()
    rels = c.relations()
    X = O.point_coordinates(c.ord1d, d["lagr.intep_pt"], c.b)
    u = d["ucoe_alpt.in"]
    # (D) fp[v][t] = coe(x, t) * d/dx_t u_v: the derivative table along t, the value table elsewhere
    for v in range(c.vecnum):
        for t in range(c.dim):
            der = O.apply_tensor(u[:, v, :], c.a, c.b, [pt_d1 if s == t else pt for s in range(c.dim)], ["vol"] * c.dim, rels, c.lev, c.ord1d)
            fp = O.position_coefficient(X, t) * der
            assert rel(fp, d["pw2.gradu.fp_intp"][:, v, t, :]) < TOL
            if v == 0:
                assert rel(O.hierarchize(fp, c.b, c.lev, c.sup, c.ord1d, anc, wt), d["pw2.gradu.fucoe_intp"][:, v, t, :]) < TOL
    # (E) rhs[v] += (u_v x ... x u_v) hierarchise(src(x, v)), coefficients untouched
    for v in range(c.vecnum):
        src = (1.0 + 0.5 * v) * np.prod([np.cos(2.0 * np.pi * (X[..., t] - 0.05 * (t + 1))) for t in range(c.dim)], axis=0)
        proj = O.apply_tensor(O.hierarchize(src, c.b, c.lev, c.sup, c.ord1d, anc, wt), c.b, c.a, [u_v] * c.dim, ["vol"] * c.dim, rels, c.lev, c.ord1d)
        assert rel(proj, d["pw2.source.rhs"][:, v, :]) < TOL
    assert np.array_equal(d["pw2.source.ucoe_after"], u)
    # (F) f_t + v2 f_x2 + (E1 + v2 B3) f_v1 + (E2 - v1 B3) f_v2 = 0, pos = (x2, v1, v2), fields = (B3, E1, E2) on the elements with velocity level 0
    le, se = d["pw2.vm.BE.level"], d["pw2.vm.BE.suppt"]
    orde = np.array([[O.order_elem(int(n), int(j)) for n, j in zip(l, s)] for l, s in zip(le, se)])
    rels_e = {"vol": [O.relations(le, se, t, "vol") for t in range(3)]}
    F = [O.apply_tensor(d["pw2.vm.BE.ucoe_alpt"][:, v, :], c.a, c.b, [pt] * 3, ["vol"] * 3, rels_e, le, orde) for v in range(3)]
    rows = O.field_rows_of(c.lev, c.sup, le, se, (1, 2))
    f = O.apply_tensor(u[:, 0, :], c.a, c.b, [pt] * 3, ["vol"] * 3, rels, c.lev, c.ord1d)
    assert rel(f, d["pw2.vm.up_intp"][:, 0, :]) < TOL
    B3, E1, E2 = F[0][rows], F[1][rows], F[2][rows]
    v1, v2 = X[..., 1], X[..., 2]
    want = [v2 * f, (E1 + v2 * B3) * f, (E2 - v1 * B3) * f]
    for t in range(3):
        assert rel(want[t], d["pw2.vm.fp_intp"][:, 0, t, :]) < TOL, t
        assert rel(O.hierarchize(want[t], c.b, c.lev, c.sup, c.ord1d, anc, wt), d["pw2.vm.fucoe_intp"][:, 0, t, :]) < TOL, t


def test_stack_programs_restated():
    """the (operation, argument) programs of amdg_pointwise_expr interpreted by the oracle: the benchmark's own Vlasov program (bench.vlasov_program)
    on the d = 6 fixture against the reference's fp_intp, the operand order of the binary operations, and the header's opcode numbering"""
    import importlib
    import os
    import re
    import sys
    from conftest import ROOT
    sys.path.insert(0, ROOT)
    import bench
    A = importlib.import_module("adaptive-multiresolution-dg_b200")
    assert A.PW == O.PW_OPS
    hdr = open(os.path.join(ROOT, "include", "amdg.h")).read()
    for name, val in O.PW_OPS.items():
        assert re.search(r"AMDG_PW_%s = %d\b" % (name, val), hdr), name
    c = Case("cfg5_vlasov_d6_k1_n2")
    d = c.d
    X = O.point_coordinates(c.ord1d, d["lagr.intep_pt"], c.b)
    up = d["up_intp"][:, 0, :]
    prog, ptr, consts = bench.vlasov_program(A, c.dim)
    outs = O.pointwise_expr(prog, ptr, consts, [up], X)
    for t in range(c.dim):
        assert rel(outs[t], d["fp_intp"][:, 0, t, :]) < 1e-14
    P = O.PW_OPS
    two = O.pointwise_expr([(P["VAR"], 0), (P["CONST"], 0), (P["SUB"], 0), (P["VAR"], 0), (P["CONST"], 1), (P["DIV"], 0)], [0, 3, 6], [2.0, 4.0], [up])
    assert np.array_equal(two[0], up - 2.0) and np.array_equal(two[1], up / 4.0)


def test_vlasov_ampere_2d2v_step_restated():
    """one RK3SSP step of the coupled 2D2V Vlasov-Ampere system, stage by stage (example/07_vlasov_ampere_02_2D2V_accuracy.cpp:255-318 without its
    manufactured source): f through interp_Vlasov_2D2V (field broadcast by copy_up_intp_to_f), HyperbolicLagrRHS vol + flx, penalty; E_t = -J through
    compute_moment_2D2V; both advanced by ExplicitRK::step_stage"""
    c = Case("vlasov_ampere_d4_k1_n3_v2")
    d = c.d
    dt, alpha = 0.002, 1.2
    pt, u_v, u_vx, uave, anc, wt = _tables(c)
    rels = c.relations()
    le, se = d["va.E.level"], d["va.E.suppt"]
    orde = np.array([[O.order_elem(int(n), int(j)) for n, j in zip(l, s)] for l, s in zip(le, se)])
    rels_e = {"vol": [O.relations(le, se, t, "vol") for t in range(4)]}
    rows = O.field_rows_of(c.lev, c.sup, le, se, (2, 3))
    partner = O.field_partner(c.lev, c.sup, le, se)
    X = O.point_coordinates(c.ord1d, d["lagr.intep_pt"], c.b)
    f = d["ucoe_alpt.in"][:, 0, :].copy(); f_tn = f.copy()
    E = d["va.E.ucoe_alpt.in"].copy(); E_tn = E.copy()
    for stage in range(3):
        up = O.apply_tensor(f, c.a, c.b, [pt] * 4, ["vol"] * 4, rels, c.lev, c.ord1d)
        Eup = [O.apply_tensor(E[:, v, :], c.a, c.b, [pt] * 4, ["vol"] * 4, rels_e, le, orde) for v in range(2)]
        fp = [X[..., 2] * up, X[..., 3] * up, Eup[0][rows] * up, Eup[1][rows] * up]
        rhs = np.zeros_like(f)
        for t in range(4):
            fuc = O.hierarchize(fp[t], c.b, c.lev, c.sup, c.ord1d, anc, wt)
            rhs += O.apply_tensor(fuc, c.b, c.a, [u_vx if s == t else u_v for s in range(4)], ["vol"] * 4, rels, c.lev, c.ord1d)
            rhs += O.apply_tensor(fuc, c.b, c.a, [uave if s == t else u_v for s in range(4)], ["flx" if s == t else "vol" for s in range(4)], rels, c.lev, c.ord1d, 0.5)
        for t in range(4):
            rhs += O.single_sweep(f, c.a, d["alpt.ujp_vjp"], "flx", rels, c.lev, c.ord1d, t, -alpha / 2.0)
        if stage == 0:
            assert rel(rhs, d["va.stage0.rhs_f"][:, 0, :]) < TOL
        rhs_e = np.zeros_like(E)
        rhs_e[:, 0, :] = O.moments(f, partner, c.a, 4, 2, (1, 0), -1.0, np.zeros_like(E[:, 0, :]))
        rhs_e[:, 1, :] = O.moments(f, partner, c.a, 4, 2, (0, 1), -1.0, np.zeros_like(E[:, 1, :]))
        f = O.rk3ssp_stage(stage, f_tn, f, rhs, dt)
        E = O.rk3ssp_stage(stage, E_tn, E, rhs_e, dt)
        assert rel(f, d["va.stage%d.f" % stage][:, 0, :]) < TOL and rel(E, d["va.stage%d.E" % stage]) < TOL


def test_hierarchisation_stencil_restated():
    """set_pts_wts_1d_ada_Lag restated from point coordinates and level-0 basis values (Lagrange)"""
    c = Case("cfg4_burgers_lagr_d2_k2_n4")
    d = c.d
    P1 = c.b
    pts = d["lagr.intep_pt"]
    msh0, msh1 = d["LagrBasis.intp_msh0"], d["LagrBasis.intp_msh1"]
    # level-0 Lagrange basis r at x: the interpolating polynomial on the (sorted) level-0 nodes
    nodes = np.sort(msh0)
    def phi0(r, x):
        v = 1.0
        for s in range(P1):
            if s != r:
                v *= (x - nodes[s]) / (nodes[r] - nodes[s])
        return v
    # AllBasis order of level-0 functions follows msh0 as stored; rank -> stored index
    stored = [int(np.argmin(np.abs(msh0 - nodes[r]))) for r in range(P1)]
    lvl0 = [[phi0(stored.index(r) if False else list(nodes).index(msh0[r]), msh1[p0]) for p0 in range(P1)] for r in range(P1)]
    pts_of = lambda k, i, q: pts[O.order_elem(k, i) * P1 + q]
    row = 0
    for n in range(1, c.nmax + 1):
        for j in range(1, max(2, 1 << n), 2):
            anc, wt = O.lagr_pw1d(n, j, pts_of, lvl0, P1)
            assert [list(x) for x in anc] == d["lagr.pw_anc"][row].tolist()
            assert np.allclose(wt, d["lagr.pw_wt"][row], atol=1e-13)
            row += 1


@pytest.mark.parametrize("name", ["moment_d3_k2_n4", "moment_d4_k1_n3"])
def test_velocity_moments(name):
    """DGAdapt::compute_moment_1D2V / _2D2V (reference source/DGAdapt.cpp:243-338): three accumulated calls, as the harness made them"""
    c = Case(name)
    d = c.d
    partner = O.field_partner(c.lev, c.sup, d["moment.level"], d["moment.suppt"])
    assert (partner >= 0).any()
    rhs = np.zeros_like(d["moment.rhs"])
    for order, w in (((0, 0), 1.25), ((1, 0), -0.5), ((0, 1), 2.0)):
        O.moments(d["ucoe_alpt.in"][:, 0, :], partner, c.a, c.dim, 2, order, w, rhs)
    assert rel(rhs, d["moment.rhs"]) < TOL
    assert np.array_equal(rhs != 0, d["moment.rhs"] != 0)
