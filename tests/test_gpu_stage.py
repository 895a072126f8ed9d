"""GPU tests of the pieces the batched / partitioned stage is made of (adaptive-multiresolution-dg_b200/stage.py): mapped destinations and
accumulate-from in the sweep kernels, row scatter, point-wise expressions, linear combination, the device-side barrier, and the whole stage
program on one GPU against the reference's dumps."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch

from conftest import load_golden
from pipeline import DevCase, rel

pytestmark = pytest.mark.gpu
TOL = 1e-12
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("kernel", [5, 8])
@pytest.mark.parametrize("name", ["cfg5_vlasov_d6_k1_n2", "adapt_d2_k2_n6", "cfg2_rt_d4_k3_n3"])
def test_mapped_destination_and_accumulate_from(name, kernel):
    """a sweep with a destination map writes every element block where the map says (here: a permuted, padded array); with acc_from the old
    values come from a second array in the plain layout"""
    d = load_golden(name)
    c = DevCase(d, kernel=kernel)
    A = c.amdg
    g = torch.Generator(device="cuda").manual_seed(3)
    for t in sorted({0, c.dim - 1}):
        for lu in (A.LU_L, A.LU_U, A.LU_FULL):
            sizes = [c.b if q < t else c.a for q in range(c.dim)]
            s_from = int(np.prod(sizes)); s_to = s_from // c.a * c.b
            x = torch.rand(c.ne, s_from, dtype=torch.float64, device="cuda", generator=g)
            base = torch.rand(c.ne, s_to, dtype=torch.float64, device="cuda", generator=g)
            plain = base.clone()
            c.ctx.sweep1d(c.op_pt, A.REL_VOL, lu, t, sizes, x, plain, coef=0.5, accumulate=True)
            perm = torch.randperm(c.ne, device="cuda", generator=g)
            pitch = s_to + (2 if s_to % 2 == 0 else 1)                       # even offsets whenever the block size is even
            out = torch.full((c.ne + 3, pitch), -7.0, dtype=torch.float64, device="cuda")
            dmap = (perm * pitch).to(torch.int64).contiguous()
            c.ctx.sweep1d_batch_mapped(c.op_pt, A.REL_VOL, lu, t, [sizes], [x], [out], coefs=[0.5], accumulates=[1], dst_maps=[dmap], acc_froms=[base])
            c.ctx.sync()
            got = out[perm][:, :s_to]
            assert torch.allclose(got, plain, rtol=0, atol=1e-14 * float(plain.abs().max()))
            assert float(out[:, s_to:].min()) == -7.0 and float(out[:, s_to:].max()) == -7.0      # nothing written outside the mapped blocks
            # without accumulate
            plain2 = torch.empty_like(base)
            c.ctx.sweep1d(c.op_pt, A.REL_VOL, lu, t, sizes, x, plain2)
            c.ctx.sweep1d_batch_mapped(c.op_pt, A.REL_VOL, lu, t, [sizes], [x], [out], dst_maps=[dmap])
            c.ctx.sync()
            assert torch.equal(out[perm][:, :s_to], plain2) or torch.allclose(out[perm][:, :s_to], plain2, rtol=0, atol=1e-14 * float(plain2.abs().max()))
            # second destination (amdg_sweep1d_batch_dual): the plain local result AND a mapped copy from the same launch (column kernel: two stores in
            # the epilogue; other kernels: a row scatter issued by the library), with and without accumulation
            for acc in (0, 1):
                loc = base.clone()
                out.fill_(-7.0)
                c.ctx.sweep1d_batch_mapped(c.op_pt, A.REL_VOL, lu, t, [sizes], [x], [loc], coefs=[0.5], accumulates=[acc], dst2s=[out], dst2_maps=[dmap])
                c.ctx.sync()
                want = plain if acc else 0.5 * plain2
                assert torch.allclose(loc, want, rtol=0, atol=1e-14 * float(plain.abs().max()))
                assert torch.equal(out[perm][:, :s_to], loc)
                assert float(out[:, s_to:].min()) == -7.0 and float(out[:, s_to:].max()) == -7.0
    c.close()


def test_scatter_rows_lincomb_barrier():
    d = load_golden("cfg4_burgers_lagr_d2_k2_n4")
    c = DevCase(d)
    A = c.amdg
    n, w = 37, 24
    src = torch.rand(n, w, dtype=torch.float64, device="cuda")
    perm = torch.randperm(n, device="cuda")
    dst = torch.zeros(n + 2, w + 4, dtype=torch.float64, device="cuda")
    mp = (perm * (w + 4)).to(torch.int64).contiguous()
    c.ctx.scatter_rows(src, n, w, dst, mp)
    xs = [torch.rand(1000, dtype=torch.float64, device="cuda") for _ in range(5)]
    y = torch.rand(1000, dtype=torch.float64, device="cuda")
    y0 = y.clone()
    c.ctx.lincomb([1.0, -2.0, 0.5, 3.0, 1.0], xs, y, beta=0.25)
    # a barrier of a single rank with itself (flags, epoch and error words in one small buffer), three times, also from a CUDA graph
    st = torch.zeros(64, dtype=torch.int32, device="cuda")
    flags, epoch, error = st.data_ptr(), st.data_ptr() + 128, st.data_ptr() + 136
    for _ in range(3):
        c.ctx.peer_barrier([flags], 0, epoch, error)
    c.ctx.sync()
    assert torch.equal(dst[perm][:, :w], src) and float(dst[:, w:].abs().max()) == 0.0
    assert torch.allclose(y, 0.25 * y0 + xs[0] - 2 * xs[1] + 0.5 * xs[2] + 3 * xs[3] + xs[4], rtol=1e-14, atol=1e-14)
    assert int(st[0]) == 3 and int(st[32]) == 3 and int(st[34]) == 0
    c.close()


def test_pointwise_expressions_match_enumerated_and_oracle():
    """amdg_pointwise_expr: the enumerated fluxes as stack programs; coordinates from the 1D table; a two-variable flux; a field through an element map"""
    sys.path.insert(0, ROOT)
    import bench
    d = load_golden("cfg5_vlasov_d6_k1_n2")
    c = DevCase(d)
    A = c.amdg
    P = A.PW
    c.ctx.points_set(d["lagr.intep_pt"])
    up = c.to_dev(d["up_intp"][:, 0, :])
    npts = c.b ** c.dim
    # 1. the Vlasov products of the fixture (v_t f, E_t(x) f with the harness' analytic field): program vs the reference's fp_intp
    prog, ptr, consts = bench.vlasov_program(A, c.dim)
    outs = [torch.zeros(c.ne, npts, dtype=torch.float64, device="cuda") for _ in range(c.dim)]
    c.ctx.pointwise_expr([up], [], None, outs, prog, ptr, consts)
    for t in range(c.dim):
        assert rel(c.to_host(outs[t]), d["fp_intp"][:, 0, t, :]) < 1e-14
    # 2. two unknowns: the coupled Schroedinger source of example/06_schrodinger_02_coupled_adapt.cpp:183-194, component 0: -(u0^2 + u1^2) u1
    u1 = torch.rand_like(up)
    o = torch.zeros_like(up)
    c.ctx.pointwise_expr([up, u1], [], None, [o], [(P["VAR"], 0), (P["SQR"], 0), (P["VAR"], 1), (P["SQR"], 0), (P["ADD"], 0), (P["VAR"], 1), (P["MUL"], 0), (P["NEG"], 0)], [0, 8])
    assert torch.allclose(o, -(up * up + u1 * u1) * u1, rtol=1e-15, atol=0)
    # 3. a field living on another grid's elements, gathered through an element map (DGSolution::copy_up_intp_to_f): E * f
    nE = 11
    E = torch.rand(nE, npts, dtype=torch.float64, device="cuda")
    emap = torch.randint(0, nE, (c.ne,), dtype=torch.int32, device="cuda")
    c.ctx.pointwise_expr([up], [E], emap, [o], [(P["OTHER"], 0), (P["VAR"], 0), (P["MUL"], 0)], [0, 3])
    assert torch.equal(o, E[emap.long()] * up)
    # 4. malformed programs are rejected, nothing is launched
    with pytest.raises(A.AmdgError):
        c.ctx.pointwise_expr([up], [], None, [o], [(P["VAR"], 0), (P["ADD"], 0)], [0, 2])
    with pytest.raises(A.AmdgError):
        c.ctx.pointwise_expr([up], [], None, [o], [(P["VAR"], 3)], [0, 1])
    c.close()


@pytest.mark.parametrize("name", ["pw_d2_k2_n4_v2", "pw_vlasov_d4_k1_n3_v2"])
def test_pointwise_bodies_vs_reference(name):
    """amdg_pointwise_expr against the reference's own point-wise kernels (fixtures from the compiled reference, run "pw" of oracle/ref_harness.cpp):
    a coupled two-variable flux through LagrInterpolation::eval_fp_Lag (all VEC_NUM unknowns reach the flux, source/Interplation.cpp:271-286), a
    coefficient of position through var_coeff_u_Lagr_fast / eval_coe_u_Lag (:648-698, 4199-4213), and interp_Vlasov_2D2V (:4508-4580) with the
    field values of a second solution reaching f through the element map of DGSolution::copy_up_intp_to_f (source/DGSolution.cpp:1024-1065)"""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import amdg_oracle as O
    d = load_golden(name)
    c = DevCase(d)
    A = c.amdg
    P = A.PW
    dim, B = c.dim, c.b ** c.dim
    c.ctx.points_set(d["lagr.intep_pt"])
    up = [c.eval_up(c.to_dev(d["ucoe_alpt.in"][:, v, :])) for v in range(2)]
    for v in range(2):
        assert rel(c.to_host(up[v]), d["pw.sys.up_intp"][:, v, :]) < TOL
    new = lambda n: [torch.zeros(c.ne, B, dtype=torch.float64, device="cuda") for _ in range(n)]
    # (A) f_{0,t} = (t+1) u0 u1, f_{1,t} = u1^2 / 2 - (t+1) u0; then hierarchisation of every component
    consts = [0.5] + [t + 1.0 for t in range(dim)]
    for i in range(2):
        prog, ptr = [], [0]
        for t in range(dim):
            if i == 0:
                prog += [(P["CONST"], 1 + t), (P["VAR"], 0), (P["MUL"], 0), (P["VAR"], 1), (P["MUL"], 0)]
            else:
                prog += [(P["CONST"], 0), (P["VAR"], 1), (P["SQR"], 0), (P["MUL"], 0), (P["CONST"], 1 + t), (P["VAR"], 0), (P["MUL"], 0), (P["SUB"], 0)]
            ptr.append(len(prog))
        outs = new(dim)
        c.ctx.pointwise_expr(up, [], None, outs, prog, ptr, consts)
        for t in range(dim):
            assert rel(c.to_host(outs[t]), d["pw.sys.fp_intp"][:, i, t, :]) < 1e-14
            assert rel(c.to_host(c.hier(outs[t])), d["pw.sys.fucoe_intp"][:, i, t, :]) < TOL
    # (B) coe(x, t) u_v with coe(x, t) = 0.3 (t+1) + sum_q w_q sin(2 pi (x_q + q/10)), w_t = 1, else 1/4: coordinates from the 1D point table
    consts = [2.0 * np.pi, 0.25] + [0.3 * (t + 1.0) for t in range(dim)] + [0.1 * q for q in range(dim)]
    for v in range(2):
        prog, ptr = [], [0]
        for t in range(dim):
            prog += [(P["CONST"], 2 + t)]
            for q in range(dim):
                prog += [(P["X"], q), (P["CONST"], 2 + dim + q), (P["ADD"], 0), (P["CONST"], 0), (P["MUL"], 0), (P["SIN"], 0)]
                if q != t:
                    prog += [(P["CONST"], 1), (P["MUL"], 0)]
                prog += [(P["ADD"], 0)]
            prog += [(P["VAR"], v), (P["MUL"], 0)]
            ptr.append(len(prog))
        outs = new(dim)
        c.ctx.pointwise_expr(up, [], None, outs, prog, ptr, consts)
        for t in range(dim):
            assert rel(c.to_host(outs[t]), d["pw.coe.fp_intp"][:, v, t, :]) < 1e-13
            assert rel(c.to_host(c.hier(outs[t])), d["pw.coe.fucoe_intp"][:, v, t, :]) < TOL
    # (C) interp_Vlasov_2D2V: E = (E1, E2) lives on its own grid (full in x, level 0 in v); its point values reach f through the element map
    if "pw.vl.fp_intp" in d:
        le, se = d["pw.vl.E.level"], d["pw.vl.E.suppt"]
        cE = A.Context(dim, c.nmax, c.pa, c.pl, device=0)
        cE.set_stream(torch.cuda.current_stream().cuda_stream)
        cE.grid_set(le, se)
        op_pt = cE.op_generate_points(A.BASIS_LAGRANGE, c.pl)
        Eup = []
        for v in range(2):
            uE = torch.from_numpy(np.ascontiguousarray(d["pw.vl.E.ucoe_alpt"][:, v, :])).cuda()
            o = torch.zeros(le.shape[0], B, dtype=torch.float64, device="cuda")
            cE.apply_tensor([op_pt] * dim, [A.REL_VOL] * dim, uE, o)
            Eup.append(o)
        rows = O.field_rows_of(d["level"][c.perm], d["suppt"][c.perm], le, se, (2, 3))
        emap = torch.from_numpy(rows).cuda()
        prog = [(P["X"], 2), (P["VAR"], 0), (P["MUL"], 0), (P["X"], 3), (P["VAR"], 0), (P["MUL"], 0),
                (P["OTHER"], 0), (P["VAR"], 0), (P["MUL"], 0), (P["OTHER"], 1), (P["VAR"], 0), (P["MUL"], 0)]
        outs = new(4)
        c.ctx.pointwise_expr([up[0]], Eup, emap, outs, prog, [0, 3, 6, 9, 12])
        for t in range(4):
            assert rel(c.to_host(outs[t]), d["pw.vl.fp_intp"][:, 0, t, :]) < TOL
            assert rel(c.to_host(c.hier(outs[t])), d["pw.vl.fucoe_intp"][:, 0, t, :]) < TOL
        cE.close()
    c.close()


@pytest.mark.parametrize("generated", [False, True])
def test_vlasov_ampere_2d2v_rk3_step_on_device(generated):
    """one complete RK3SSP step of the coupled 2D2V Vlasov-Ampere system on the device, no host round trip between the stages: f by the field broadcast
    through the element map (DGSolution::copy_up_intp_to_f), the Vlasov products (interp_Vlasov_2D2V), hierarchisation, vol + flx + penalty sweeps;
    E_t = -J by the velocity moments (compute_moment_2D2V); both by amdg_rk_stage.  Every stage against the compiled reference running
    example/07_vlasov_ampere_02_2D2V_accuracy.cpp:255-318 on the same grids (fixture vlasov_ampere_d4_k1_n3_v2); with `generated` no table of the
    reference is used either"""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import amdg_oracle as O
    d = load_golden("vlasov_ampere_d4_k1_n3_v2")
    c = DevCase(d, generated=generated)
    A = c.amdg
    P = A.PW
    dim, B, dt = 4, c.b ** 4, 0.002
    if generated:
        c.ctx.points_generate(A.BASIS_LAGRANGE, c.pl)
    else:
        c.ctx.points_set(d["lagr.intep_pt"])
    le, se = d["va.E.level"], d["va.E.suppt"]
    nE = le.shape[0]
    cE = A.Context(dim, c.nmax, c.pa, c.pl, device=0)
    cE.set_stream(torch.cuda.current_stream().cuda_stream)
    cE.grid_set(le, se)
    op_pt_E = cE.op_generate_points(A.BASIS_LAGRANGE, c.pl) if generated else cE.op_register(d["Lag_pt_Alpt_1D"].T.copy(), c.a, c.b)
    lev_f, sup_f = d["level"][c.perm], d["suppt"][c.perm]
    emap = torch.from_numpy(O.field_rows_of(lev_f, sup_f, le, se, (2, 3))).cuda()                 # f element -> field element (broadcast)
    pmap = torch.from_numpy(O.field_partner(lev_f, sup_f, le, se)).cuda()                          # field element -> f element (moments)
    f = c.to_dev(d["ucoe_alpt.in"][:, 0, :]); f_tn = f.clone()
    E = torch.from_numpy(np.ascontiguousarray(d["va.E.ucoe_alpt.in"].transpose(1, 0, 2))).cuda(); E_tn = E.clone()      # [2][nE][a^4]
    prog = [(P["X"], 2), (P["VAR"], 0), (P["MUL"], 0), (P["X"], 3), (P["VAR"], 0), (P["MUL"], 0),
            (P["OTHER"], 0), (P["VAR"], 0), (P["MUL"], 0), (P["OTHER"], 1), (P["VAR"], 0), (P["MUL"], 0)]
    launches0 = c.ctx.launch_count + cE.launch_count
    for stage in range(3):
        up = c.eval_up(f)
        Eup = torch.zeros(2, nE, B, dtype=torch.float64, device="cuda")
        cE.apply_tensor([op_pt_E] * dim, [A.REL_VOL] * dim, E, Eup, n_comp=2)
        fp = [torch.zeros(c.ne, B, dtype=torch.float64, device="cuda") for _ in range(4)]
        c.ctx.pointwise_expr([up], [Eup[0], Eup[1]], emap, fp, prog, [0, 3, 6, 9, 12])
        rhs = c.zeros(c.a)
        c.rhs_vol_flx([c.hier(x) for x in fp], rhs)
        c.penalty(f, rhs, 1.2)
        if stage == 0:
            assert rel(c.to_host(rhs), d["va.stage0.rhs_f"][:, 0, :]) < TOL
        rhs_E = torch.zeros_like(E)
        c.ctx.moment(pmap, 2, (1, 0), -1.0, f, rhs_E[0])                # the moments read f at the start of the stage, as the reference does
        c.ctx.moment(pmap, 2, (0, 1), -1.0, f, rhs_E[1])
        c.ctx.rk_stage(A.RK_RK3SSP, stage, dt, f_tn, f, rhs)
        cE.rk_stage(A.RK_RK3SSP, stage, dt, E_tn, E, rhs_E)
        assert rel(c.to_host(f), d["va.stage%d.f" % stage][:, 0, :]) < TOL
        assert rel(E.cpu().numpy().transpose(1, 0, 2), d["va.stage%d.E" % stage]) < TOL
    assert c.ctx.launch_count + cE.launch_count > launches0
    cE.close()
    c.close()


@pytest.mark.parametrize("kernel", [0, 5, 8])
def test_stage_program_single_gpu_vs_reference(kernel):
    """the whole batched stage program (stage.py) on one GPU: right-hand side and RK stage 0 of the d=6 Vlasov fixture against the reference"""
    sys.path.insert(0, ROOT)
    import bench
    A = importlib.import_module("adaptive-multiresolution-dg_b200")
    S = importlib.import_module("adaptive-multiresolution-dg_b200.stage")
    D = importlib.import_module("adaptive-multiresolution-dg_b200.dist")
    err, berr = bench.parity_check(A, S, D, 1, 0, 0, torch.cuda.current_stream(), kernel)
    assert err < TOL and berr == 0


def test_stage_program_burgers_2d_vs_reference():
    """the same program in 2D (cfg4, Burgers, one flux component per dimension) against the reference's nonlinear right-hand side"""
    sys.path.insert(0, ROOT)
    import bench
    A = importlib.import_module("adaptive-multiresolution-dg_b200")
    S = importlib.import_module("adaptive-multiresolution-dg_b200.stage")
    D = importlib.import_module("adaptive-multiresolution-dg_b200.dist")
    d = load_golden("cfg1_adv_d2_k2_n4")           # Burgers flux in both dimensions, RK3SSP stages dumped
    dim, nmax, n0, sparse, pa, pl = [int(x) for x in d["config"][:6]]
    stream = torch.cuda.current_stream()
    st, plan, part = bench.make_stage(A, S, D, dim, nmax, pa, pl, d["level"], d["suppt"], d, "burgers", 1, 0, 0, stream.cuda_stream, 0, (A.RK_RK3SSP, 0, 0.002), dense=True)
    u0 = torch.from_numpy(np.ascontiguousarray(d["ucoe_alpt.in"][:, 0, :])).cuda()
    st.view("u").copy_(u0); st.view("u_tn").copy_(u0)
    st.run()
    torch.cuda.synchronize()
    assert rel(st.view("rhs").cpu().numpy(), d["rhs_all"][:, 0, :]) < TOL
    assert rel(st.view("u").cpu().numpy(), d["stage0.ucoe_alpt"][:, 0, :]) < TOL
    n_unfused = plan.launches()
    st.close()
    # the fused plan: all three RK3SSP stages in a row (the accumulator of one stage is the input of the next), each against the reference's stage dump
    u_tn = u0.clone()
    cur = u0
    for stage in range(3):
        st, plan, part = bench.make_stage(A, S, D, dim, nmax, pa, pl, d["level"], d["suppt"], d, "burgers", 1, 0, 0, stream.cuda_stream, 0, (A.RK_RK3SSP, stage, 0.002),
                                          dense=True, fuse_rk=True)
        assert plan.launches() < n_unfused and plan.rhs is None
        st.view("u").copy_(cur); st.view("u_tn").copy_(u_tn)
        st.run()
        torch.cuda.synchronize()
        cur = st.view(plan.result).clone()
        assert rel(cur.cpu().numpy(), d["stage%d.ucoe_alpt" % stage][:, 0, :]) < TOL
        st.swap_result()                                         # the accumulator becomes "u" of the next run of this program (pointer swap)
        assert torch.equal(st.view("u"), cur)
        st.close()


def _n_gpus():
    return torch.cuda.device_count()


@pytest.mark.skipif(_n_gpus() < 2, reason="needs 2 GPUs (the driver's 1-GPU box runs the same check inside `bench.py --gpus 2`)")
def test_two_gpu_stage_parity():
    import subprocess
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29671",
                        os.path.join(ROOT, "tests", "dist_check.py")], capture_output=True, text=True, timeout=900)
    assert "DIST_CHECK OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_live_reference_dropin_burgers_adapt():
    """examples/live_burgers_adapt: the reference's own DGAdapt / OperatorMatrix1D / HermInterpolation objects (oracle/_ref/libsgdg_ref.a, unmodified)
    with every fast class replaced by the mirror over the C ABI, ten adaptive Burgers steps with refine() and coarsen() each step
    (example/02_hyperbolic_05_burgers_adapt.cpp:223-470, -imex 0 -v 0) in lockstep with the stock reference run: same element sets, coefficients
    within 1e-10, identical L2 error against the exact solution"""
    import subprocess
    exe = os.path.join(ROOT, "examples", "live_burgers_adapt")
    if not os.path.exists(exe):
        pytest.skip("examples/live_burgers_adapt is built only where the reference sources exist (__graft_entry__.build())")
    r = subprocess.run([exe, "-NM", "6", "-N0", "2", "-steps", "10"], capture_output=True, text=True, timeout=900)
    assert "LIVE OK" in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]
    # -gen 1: the device arm takes no table from the reference objects (library-generated Hermite / Alpert tables, point table, stencils): the
    # adaptive run must still make the same refine / coarsen decisions and stay within the full-run bound
    r = subprocess.run([exe, "-NM", "6", "-N0", "2", "-steps", "10", "-gen", "1"], capture_output=True, text=True, timeout=900)
    assert "LIVE OK" in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]


def test_live_reference_full_runs_static_burgers():
    """full runs to t = 0.05 on static sparse grids of level 3..6 (2-D Burgers, k = 2, Hermite flux interpolation, RK3SSP, the time step of the
    example: 4 .. 32 steps): after every step the device coefficients stay within 1e-10 of the stock reference run (north_star's full-run bound) and the
    L1 / L2 / Linf errors against ExactSolution are the same for both arms"""
    import re
    import subprocess
    exe = os.path.join(ROOT, "examples", "live_burgers_adapt")
    if not os.path.exists(exe):
        pytest.skip("examples/live_burgers_adapt is built only where the reference sources exist (__graft_entry__.build())")
    for n in (3, 4, 5, 6):
        r = subprocess.run([exe, "-NM", str(n), "-N0", str(n), "-static", "1", "-tf", "0.05"], capture_output=True, text=True, timeout=900)
        assert "LIVE OK" in r.stdout and "static-grid" in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]
        m = re.search(r"reference (\S+) (\S+) (\S+) \| device (\S+) (\S+) (\S+)", r.stdout)
        assert abs(float(m.group(2)) - float(m.group(5))) <= 1e-10 * max(1.0, float(m.group(2)))


def test_live_reference_advection_convergence_orders():
    """examples/live_advection_convergence: the reference's linear advection example (example/02_hyperbolic_01_scalar_const_coefficient.cpp, k = 2,
    sparse grids N = 3..6, RK3SSP to t = 0.1) run to the end by the stock assembled-matrix path and by the device sweeps: coefficients within 1e-10 after
    the full run, identical L2 errors against the exact solution, and so identical observed orders of convergence (about k + 1/2 .. k + 1 on sparse grids)"""
    import re
    import subprocess
    exe = os.path.join(ROOT, "examples", "live_advection_convergence")
    if not os.path.exists(exe):
        pytest.skip("examples/live_advection_convergence is built only where the reference sources exist (__graft_entry__.build())")
    r = subprocess.run([exe, "-Nmin", "3", "-Nmax", "6"], capture_output=True, text=True, timeout=1500)
    assert "CONVERGENCE OK" in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]
    rows = re.findall(r"^\| (\d) \| (\S+) \| (\S+) \| (\S+) \| (\S+) \|$", r.stdout, flags=re.M)
    orders = [float(x[4]) for x in rows if x[4] != "-"]
    print(r.stdout[-900:])
    assert len(orders) == 3 and min(orders) > 2.0          # third order scheme on sparse grids (log factors): observed 2.3 .. 3.1 in the reference as well


@pytest.mark.parametrize("name", ["moment_d3_k2_n4", "moment_d4_k1_n3"])
def test_velocity_moments_device(name):
    """amdg_moment against DGAdapt::compute_moment_1D2V / _2D2V of the compiled reference (three accumulated calls), and the mixed first-order
    moment (product form, not offered by the reference) against the numpy restatement"""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import amdg_oracle as O
    A = importlib.import_module("adaptive-multiresolution-dg_b200")
    d = load_golden(name)
    dim, nmax, n0, sparse, pa, pl = [int(x) for x in d["config"][:6]]
    a = pa + 1
    ctx = A.Context(dim, nmax, pa, pl, device=0)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    ctx.grid_set(d["level"], d["suppt"])
    f = torch.from_numpy(np.ascontiguousarray(d["ucoe_alpt.in"][:, 0, :])).cuda()
    partner = O.field_partner(d["level"], d["suppt"], d["moment.level"], d["moment.suppt"])
    pmap = torch.from_numpy(partner).cuda()
    rhs = torch.zeros(d["moment.rhs"].shape, dtype=torch.float64, device="cuda")
    for order, w in (((0, 0), 1.25), ((1, 0), -0.5), ((0, 1), 2.0)):
        ctx.moment(pmap, 2, order, w, f, rhs)
    assert rel(rhs.cpu().numpy(), d["moment.rhs"]) < TOL
    mixed = torch.zeros_like(rhs)
    ctx.moment(pmap, 2, (1, 1), 3.0, f, mixed)
    ref = O.moments(d["ucoe_alpt.in"][:, 0, :], partner, a, dim, 2, (1, 1), 3.0, np.zeros_like(d["moment.rhs"]))
    assert rel(mixed.cpu().numpy(), ref) < TOL
    ctx.close()
