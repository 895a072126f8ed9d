"""The batched / fibre-partitioned stage program (adaptive-multiresolution-dg_b200/stage.py) executed with the numpy oracle.

Every rank's plan is built in this process; buffers live in per-rank numpy slabs laid out by stage.SlabLayout with fake base addresses, pushes
and scatters go through the very destination maps the device path uses, sweeps are the oracle's transform_1d on the rank's local grid.  The
result (point values, hierarchical flux coefficients, right-hand side, stage update) must equal the single-grid oracle computation with the
reference's literal schedule, for 1, 2, 3 and 8 ranks -- including ranks that own no element."""
import importlib

import numpy as np
import pytest

import amdg_oracle as O
from conftest import load_golden

A = importlib.import_module("adaptive-multiresolution-dg_b200")
D = importlib.import_module("adaptive-multiresolution-dg_b200.dist")
S = importlib.import_module("adaptive-multiresolution-dg_b200.stage")

PEN = -0.6
DT = 1e-3


def dense_from_blocks(ctx, blocks, kf, kt):
    src, tgt, vol = ctx.pairs()
    T = 2 ** ctx.nmax
    m = np.zeros((T * kf, T * kt))
    for p in range(len(src)):
        m[src[p] * kf:(src[p] + 1) * kf, tgt[p] * kt:(tgt[p] + 1) * kt] = blocks[p]
    return m


def setup(name):
    d = load_golden(name)
    dim, nmax, n0, sparse, pa, pl = [int(x) for x in d["config"][:6]]
    a, b = pa + 1, pl + 1
    ctx = A.Context(dim, nmax, pa, pl, device=-1)
    ctx.grid_set(d["level"], d["suppt"])
    hop = ctx.op_register_hier(d["lagr.pw_anc"], d["lagr.pw_wt"])
    mats = {"pt": (d["Lag_pt_Alpt_1D"].T.copy(), a, b), "uv": (d["lagr.u_v"], b, a),
            "volflx": (d["lagr.u_vx"] + 0.5 * (d["lagr.ulft_vjp"] + d["lagr.urgt_vjp"]), b, a),
            "pen": (d["alpt.ujp_vjp"], a, a), "hier": (dense_from_blocks(ctx, ctx.op_blocks(hop, b, b), b, b), b, b)}
    ctx.close()
    return d, dim, nmax, a, b, mats


def flux(c, up):
    return (c + 1.0) * up * up / 2.0


def reference(d, dim, a, b, mats, u):
    lev, sup, ordv = d["level"], d["suppt"], d["order_elem"]
    rels = {k: [O.relations(lev, sup, t, k) for t in range(dim)] for k in ("vol", "flx")}
    up = O.apply_tensor(u, a, b, [mats["pt"][0]] * dim, ["vol"] * dim, rels, lev, ordv)
    fuc = []
    for c in range(dim):
        cur, sizes = flux(c, up), [b] * dim
        for t in range(dim):
            cur, _ = O.transform_1d(cur, sizes, mats["hier"][0], "U", rels["vol"][t], lev, ordv, t, b - 1, b - 1)
        fuc.append(cur)
    rhs = np.zeros_like(u)
    for t in range(dim):
        m = [mats["volflx"][0] if s == t else mats["uv"][0] for s in range(dim)]
        kinds = ["flx" if s == t else "vol" for s in range(dim)]
        rhs += O.apply_tensor(fuc[t], b, a, m, kinds, rels, lev, ordv)
    if dim > 1:
        for t in range(dim):
            rhs += O.single_sweep(u, a, mats["pen"][0], "flx", rels, lev, ordv, t, PEN)
    return up, fuc, rhs, 0.75 * u + 0.25 * (u + DT * rhs)           # RK3SSP stage 1 with u_tn = u


class Emulator:
    """all ranks of a partitioned stage in one process"""

    SYM = {"pen": PEN, "rk_a": 0.75, "rk_b": 0.25, "rk_c": 0.25 * DT}        # RK3SSP stage 1

    @classmethod
    def coef(cls, c):
        if isinstance(c, str):
            v = 1.0
            for f in c.split("*"):
                v *= cls.SYM[f]
            return v
        return c

    def __init__(self, d, dim, a, b, mats, world, fuse_rk=False, dual_store=True):
        self.d, self.dim, self.a, self.b, self.mats, self.world = d, dim, a, b, mats, world
        lev, sup = d["level"], d["suppt"]
        self.parts = [D.FibrePartition(lev, sup, world, r) if world > 1 else None for r in range(world)]
        self.plans = [S.StagePlan(dim, a, b, dim, part=self.parts[r], fuse_rk=fuse_rk, dual_store=dual_store) for r in range(world)]
        self.lay = [S.SlabLayout(self.plans[r], lev.shape[0]) for r in range(world)]
        self.base = [(r + 1) << 44 for r in range(world)]                     # fake byte addresses, far apart
        self.slab = [np.zeros(int(self.lay[r].total[r])) for r in range(world)]
        self.maps = [self.lay[r].maps(self.base)[0] if world > 1 else {} for r in range(world)]
        self.rows = [{L: (self.parts[r].local[L] if world > 1 else np.arange(lev.shape[0])) for L in ("X", "V")} for r in range(world)]
        self.rel_cache = {}

    def view(self, r, name):
        plan = self.plans[r]
        key = plan.alias.get(name, name)
        bb = plan.bufs[key]
        n = len(self.rows[r][bb.layout])
        o = self.lay[r].offset(key)
        return self.slab[r][o:o + n * bb.width].reshape(n, bb.width)

    def store_mapped(self, r, dst, src_layout, values):
        """values[i] -> block of local row i in the owner's copy of dst, through the destination map"""
        plan = self.plans[r]
        w = plan.bufs[dst].width
        mp = self.maps[r][(dst, src_layout)]
        start = self.base[r] // 8 + self.lay[r].offset(dst)                    # "address" (in doubles) of this rank's copy
        for i in range(values.shape[0]):
            addr = start + int(mp[i])
            rr = (addr * 8 >> 44) - 1
            off = addr - self.base[rr] // 8
            assert 0 <= rr < self.world and 0 <= off and off + w <= self.slab[rr].size and mp[i] % 2 == 0 or w % 2
            self.slab[rr][off:off + w] = values[i]

    def rels(self, r, lay, t, kind):
        key = (r, lay, t, kind)
        if key not in self.rel_cache:
            rows = self.rows[r][lay]
            self.rel_cache[key] = O.relations(self.d["level"][rows], self.d["suppt"][rows], t, kind)
        return self.rel_cache[key]

    def run(self):
        d = self.d
        n_ops = len(self.plans[0].ops)
        assert all(len(p.ops) == n_ops for p in self.plans)
        for i in range(n_ops):
            for r in range(self.world):
                o = self.plans[r].ops[i]
                assert o[0] == self.plans[0].ops[i][0]
                if o[0] == "sweep":
                    _, lay, opn, rel, lu, t, jobs = o[:7]
                    rows = self.rows[r][lay]
                    if not len(rows):
                        continue
                    mat, kf, kt = self.mats[opn]
                    for j in jobs:
                        coef = self.coef(j["coef"])
                        out, _ = O.transform_1d(self.view(r, j["src"]), j["sizes"], mat, ("L", "U", "full")[lu], self.rels(r, lay, t, ("vol", "flx")[rel]),
                                                d["level"][rows], d["order_elem"][rows], t, kf - 1, kt - 1, coef=coef)
                        if j["acc"]:
                            out = out + self.view(r, j["acc_from"] if j.get("acc_from") else j["dst"])
                        if j.get("push") and self.world > 1:
                            self.store_mapped(r, j["dst"], lay, out)
                        else:
                            self.view(r, j["dst"])[:] = out
                        if j.get("dst2"):
                            self.store_mapped(r, j["dst2"], lay, out)
                elif o[0] == "scatter":
                    _, src, dst = o
                    lay = self.plans[r].bufs[src].layout
                    self.store_mapped(r, dst, lay, self.view(r, src))
                elif o[0] == "pointwise":
                    _, lay, up, fps = o
                    for c, f in enumerate(fps):
                        self.view(r, f)[:] = flux(c, self.view(r, up))
                elif o[0] == "lincomb":
                    _, dst, parts, beta = o[:4]
                    cf = [1.0] * len(parts) if len(o) < 5 else [self.coef(x) for x in o[4]]
                    self.view(r, dst)[:] = sum(c * self.view(r, p) for c, p in zip(cf, parts)) + (beta * self.view(r, dst) if beta else 0.0)
                elif o[0] == "rk":
                    _, u_tn, u, rhs = o
                    self.view(r, u)[:] = 0.75 * self.view(r, u_tn) + 0.25 * (self.view(r, u) + DT * self.view(r, rhs))

    def gather(self, name):
        """global array of a buffer (by the layout it lives in)"""
        plan = self.plans[0]
        bb = plan.bufs[plan.alias.get(name, name)]
        out = np.zeros((self.d["level"].shape[0], bb.width))
        for r in range(self.world):
            out[self.rows[r][bb.layout]] = self.view(r, name)
        return out


@pytest.mark.parametrize("name,worlds", [("cfg5_vlasov_d6_k1_n2", (1, 2, 8)), ("cfg4_burgers_lagr_d2_k2_n4", (1, 2, 3)), ("cfg2_rt_d4_k3_n3", (2,)),
                                         ("variants_lagr_d3_k1_n3", (1, 2, 5))])
def test_partitioned_stage_equals_single_grid(name, worlds):
    d, dim, nmax, a, b, mats = setup(name)
    rng = np.random.default_rng(5)
    u = rng.uniform(-1, 1, size=(d["level"].shape[0], a ** dim)) * np.ldexp(1.0, -d["level"].sum(axis=1))[:, None]
    up_ref, fuc_ref, rhs_ref, u_ref = reference(d, dim, a, b, mats, u)
    rel = lambda x, y: np.linalg.norm(x - y) / np.linalg.norm(y)
    for world, fuse in [(w, f) for w in worlds for f in (False, True)]:
        E = Emulator(d, dim, a, b, mats, world, fuse_rk=fuse, dual_store=fuse)          # the older form (row scatters) rides with the unfused plan
        for r in range(world):
            E.view(r, "u")[:] = u[E.rows[r]["X"]]
            E.view(r, "u_tn")[:] = u[E.rows[r]["X"]]
        E.run()
        p = E.plans[0]
        assert rel(E.gather(p.up), up_ref) < 1e-13
        for c in range(dim):
            assert rel(E.gather(p.fuc[c]), fuc_ref[c]) < 1e-13
        if fuse:
            # the RK combination rides in the sweep epilogues: no rhs array, no "rk" operation, fewer launches than the unfused plan
            assert p.rhs is None and not any(o[0] == "rk" for o in p.ops)
            assert p.launches() < S.StagePlan(dim, a, b, dim, part=E.parts[0], dual_store=False).launches()
            if world > 1 and dim > 2:
                # second destinations replace the row scatters of the down-pass buffers and of the hierarchised fluxes
                n_sc = lambda q: sum(1 for o in q.ops if o[0] == "scatter")
                assert n_sc(p) < n_sc(S.StagePlan(dim, a, b, dim, part=E.parts[0], fuse_rk=True, dual_store=False))
        else:
            assert rel(E.gather(p.rhs), rhs_ref) < 1e-12
        assert rel(E.gather(p.result), u_ref) < 1e-13
        if world > 1:
            assert p.n_barrier == (6 if dim > 2 else 6) or dim <= 2
            # every rank's maps hit every remote row of a pushed buffer exactly once (checked through the final values above) and are even
            for r in range(world):
                for (dst, lay), mp in E.maps[r].items():
                    if p.bufs[dst].width % 2 == 0:
                        assert (mp % 2 == 0).all()


def test_plan_summary_cfg5():
    """the 6-D plan: launches per stage and exchange volume per element (what bench.py reports)"""
    lev, sup = A.sparse_grid(6, 4)
    part = D.FibrePartition(lev, sup, 8, 3)
    p1 = S.StagePlan(6, 2, 3, 6)
    p8 = S.StagePlan(6, 2, 3, 6, part=part)
    assert p1.n_barrier == 0 and p8.n_barrier == 6
    assert sum(1 for o in p8.ops if o[0] == "sweep") < 70
    # exchange volume per element and stage: interpolation 1000 + 3375 + u 64 + pen 64 + up 729 + hierarchisation 6*729 + rhs 6*(3375 + 1000)
    vol = 64 + 1000 + 3375 + 64 + 729 + 6 * 729 + 6 * (3375 + 1000)
    assert S.StagePlan(6, 2, 3, 6, part=part, dual_store=False).push_bytes == vol
    # second destinations move the same blocks from the epilogues of their producers; u@V serves the interpolation as well (one scatter of u less)
    assert p8.push_bytes == vol - 64
    assert sum(1 for o in p8.ops if o[0] == "scatter") == 1 and sum(1 for o in S.StagePlan(6, 2, 3, 6, part=part, dual_store=False).ops if o[0] == "scatter") == 29


def test_rk_coefficients_of_the_fused_plan_match_every_scheme():
    """stage.rk_coefficients (the a, b, c of u_new = a u_tn + b u + c rhs that the fused plan folds into the sweep epilogues) against the oracle's
    ExplicitRK::step_stage restatement for every scheme and stage (source/ODESolver.cpp:209-330)"""
    rng = np.random.default_rng(11)
    u_tn, u, rhs = rng.standard_normal((3, 50))
    dt = 0.0137
    for scheme, name, stages in ((A.RK_EULER, "euler", 1), (A.RK_RK2SSP, "rk2ssp", 2), (A.RK_RK2MID, "rk2mid", 2), (A.RK_RK3SSP, "rk3ssp", 3), (A.RK_RK3HEUN, "rk3heun", 3)):
        for stage in range(stages):
            a, b, c = S.rk_coefficients(scheme, stage, dt)
            uu = u_tn if (stage == 0) else u                      # at stage 0 the current state is u_tn in every scheme
            want = O.rk_stage(name, stage, u_tn, uu, rhs, dt)
            got = a * u_tn + b * uu + c * rhs
            assert np.allclose(got, want, rtol=1e-14, atol=1e-15), (name, stage)
