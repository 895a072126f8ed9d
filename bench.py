#!/usr/bin/env python
"""Benchmark of the fast sparse-grid transform path (BASELINE.json metric: sparse-grid DoF-stage updates/s, FP64).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg5|cfg2|cfg4] [--kernel V]

Default workload (every N): **cfg5**, example/07_vlasov_maxwell_sparse scaled to 3D3V -- d=6, k=1, m=2, NMAX=7 full sparse grid (16 172 elements,
1.04 M DoF, 11.8 M interpolation points) -- ONE complete nonlinear RK3SSP stage: FastLagrIntp::eval_up_Lagr -> Vlasov point-wise products ->
eval_fp_to_coe_D_Lag -> rhs_vol + rhs_flx of all six dimensions -> LxF penalty sweeps -> ExplicitRK::step_stage, as the batched program of
adaptive-multiresolution-dg_b200/stage.py.  N = 1 runs it on one context; N > 1 (torchrun, one rank per GPU) partitions the grid by fibre
ownership: the layout switches are stores into CUDA-IPC-mapped peer memory from the epilogues of the producing sweeps plus row scatters, separated
by device-side barriers -- STRONG scaling, value = global DoF / max-over-ranks stage time.  Before timing, the same program is run on the reference's
d=6 fixture (tests/golden/cfg5_vlasov_d6_k1_n2) at the same N and compared with the reference's dump: `config.parity_rel_l2`.
At N = 1 the line also carries `secondary`: cfg2 (example/01_interp_01_high_dim, Lagrange round trip d=4, k=3, m=3, NMAX=8), the transform-only
workload, with the roofline of its <4,4> sweep kernel.

Timing: CUDA events on the launch stream around CUDA-graph replays of the step, an L2 flush (256 MiB write) between timed steps, max over ranks.
`e2e`: the same step from pinned host buffers (H2D of the coefficients, the stage, D2H of the result inside the timed region).  The reference arm
(--impl reference) and `cpu_baseline` time the compiled, unmodified reference (oracle/_ref/ref_harness) on the host cores.
"""
import argparse
import ctypes
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # cpu_nmax: bounded sample for the cpu_baseline leg; ref_nmax: the --impl reference arm (same config where the reference finishes in minutes)
    "cfg2": dict(kind="roundtrip", dim=4, k=3, m=3, nmax=8, cpu_nmax=7, ref_nmax=8, desc="example/01_interp_01_high_dim: Lagrange interpolation round trip d=4 k=3 m=3 NMAX=8 (full sparse grid)"),
    "cfg5": dict(kind="stage", flux="vlasov", dim=6, k=1, m=2, nmax=7, cpu_nmax=5, ref_nmax=6, ref_single=True, desc="example/07_vlasov_maxwell_sparse scaled to 3D3V: d=6 k=1 m=2 NMAX=7 full sparse grid, one nonlinear RK3SSP stage (interpolate, Vlasov products with a prescribed smooth field, hierarchise, vol+flx+penalty, RK)"),
    "cfg1": dict(kind="linear", op="advection", dim=2, k=2, m=3, nmax=7, cpu_nmax=7, ref_nmax=7, stages=3, desc="example/02_hyperbolic_01_scalar_const_coefficient: 2D linear advection, Alpert k=2, NMAX=7 full sparse grid, one RK3SSP stage (operator as 1D sweeps: u_vx + upwind flux per dimension)"),
    "cfg3": dict(kind="linear", op="wave", dim=3, k=2, m=3, nmax=7, cpu_nmax=7, ref_nmax=7, stages=4, desc="example/03_wave_01_const_coeff_periodic: 3D second-order wave, k=2, NMAX=7, IPDG (sigma=20), one RK4ODE2nd stage (operator as 1D sweeps: four terms per dimension merged into one operator)"),
    "cfg4": dict(kind="stage", flux="burgers", dim=2, k=2, m=3, nmax=9, cpu_nmax=9, ref_nmax=9, desc="example/02_hyperbolic_05_burgers_adapt at BASELINE's NMAX=9 on the static upper-bound grid (full sparse grid, 2 816 elements; Lagrange flux): one nonlinear RK3SSP stage.  The adaptive run itself (refine + coarsen every step, live reference DGAdapt) is examples/live_burgers_adapt"),
}
LXF_ALPHA, DT = 1.2, 1e-4


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel_key):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the roofline kernel, from the profile summary the profiling script wrote
    (profiles/r02_traffic.json, keyed by kernel); None when there is no capture of this kernel"""
    p = os.path.join(ROOT, "profiles", "r02_traffic.json")
    if not os.path.exists(p):
        return None
    try:
        return json.load(open(p)).get(kernel_key, {}).get("dram_bytes_per_launch")
    except Exception:
        return None


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled during the timed region"""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag, self.proc = index, [], False, None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(",")])
                if self.stop_flag:
                    break
        except Exception:
            pass

    def finish(self):
        self.stop_flag = True
        if self.proc:
            self.proc.terminate()
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            for nm, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx[0] if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


def load_tables(A, w):
    """the 1D operator / point / stencil tables of the workload, generated by the library itself in compact form (csrc/tables.hpp; the bundles of
    reference outputs that earlier rounds shipped are now only the fixture of tests/test_tables.py)"""
    return A.generate_tables(w["nmax"], w["k"], w["m"])


def synthetic_field(level, block, n_comp, seed):
    """i.i.d. U(-1,1) * 2^-(n_1+...+n_d): the decay keeps the hierarchical sums well conditioned"""
    rng = np.random.default_rng(seed)
    scale = np.ldexp(1.0, -level.sum(axis=1).astype(np.int64))
    return (rng.uniform(-1.0, 1.0, size=(n_comp, level.shape[0], block)) * scale[None, :, None])


def run_reference(args, w, n_threads=None, as_baseline=False):
    """the compiled, unmodified reference on the host cores"""
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_harness")
    if not os.path.exists(exe):
        return None
    nthr = n_threads or os.cpu_count()
    nmax = w["cpu_nmax"] if as_baseline else w["ref_nmax"]
    reps = 2 if as_baseline else max(1, min(args.steps, 2 if nmax >= w["nmax"] else 5))
    warm = 1 if (as_baseline or nmax >= w["nmax"]) else min(args.warmup, 2)
    if w.get("ref_single") and not as_baseline:
        reps, warm = 1, 0          # one stage of the reference at this size is > 1 minute of host time (plus ~40 s of neighbour set-up): a single timed stage
    cmd = [exe, "--dim", str(w["dim"]), "--nmax", str(nmax), "--pa", str(w["k"]), "--pl", str(w["m"]), "--time", str(reps + warm), "--threads", str(nthr)]
    if w["kind"] == "roundtrip":
        cmd += ["--run", "roundtrip"]
        phases = ("intp", "hier", "init")
        what = "same round trip"
    elif w["kind"] == "linear":
        # the shipped path of these examples: assembled SpMV inside ExplicitRK::step_rk (all stages of one step are timed together)
        cmd += ["--run", w["op"], "--dt", "1e-4"]
        phases = ("adv_step_rk",) if w["op"] == "advection" else ("wave_step_rk",)
        what = "one %s::step_rk on the assembled operator (the shipped path), divided by its %d stages" % ("RK3SSP" if w["op"] == "advection" else "RK4ODE2nd", w["stages"])
    else:
        cmd += ["--run", "rhs", "--flux", w["flux"]]
        phases = ("intp", "pointwise", "hier", "rhs_vol", "rhs_flx", "rhs_penalty")
        what = "same nonlinear stage (interpolate, point-wise, hierarchise, vol + flx + penalty sweeps; RK axpy not counted)"
    t0 = time.time()
    out = subprocess.run(cmd, check=True, stdout=subprocess.PIPE, text=True).stdout.strip().splitlines()[-1]
    wall = time.time() - t0
    r = json.loads(out)
    t_step = sum(r[p] for p in phases) / (w["stages"] if w["kind"] == "linear" else 1)          # medians over the repetitions
    dof = r["dof"]
    same = nmax == w["nmax"]
    return {"value": dof / t_step, "unit": "DoF-stage/s", "cores": r["threads"], "kind": "reference", "same_config": same,
            "sample": "%s at NMAX=%d (%d elements, %d DoF)%s: median of %d repetitions, %.3f s per step; reference built from /root/reference/source with -O3 -fopenmp (oracle/Makefile)"
                      % (what, nmax, r["n_elem"], dof, " -- the benchmark's own configuration" if same else " -- a bounded sample of the NMAX=%d workload, DoF-normalised" % w["nmax"], r["reps"], t_step),
            "ms_per_step": t_step * 1e3, "wall_s": wall}


# ---------------------------------------------------------------------------------------------------------------------------------------
# point-wise programs (amdg_pointwise_expr)
def vlasov_program(A, dim):
    """fp_t = v_t f for the x dims, E_t(x) f for the v dims, E_t(x) = sum_s sin(2 pi (x_s + (t - d/2 + 1)/8)): the generalised
    interp_Vlasov_2D2V body (reference source/Interplation.cpp:4508-4580) with the prescribed smooth field of oracle/ref_harness.cpp"""
    P = A.PW
    hd = dim // 2
    prog, ptr, consts = [], [0], [2.0 * np.pi]
    for t in range(dim):
        if t < hd:
            prog += [(P["X"], hd + t), (P["VAR"], 0), (P["MUL"], 0)]
        else:
            consts.append(0.125 * (t - hd + 1))
            ci = len(consts) - 1
            for s in range(hd):
                prog += [(P["X"], s), (P["CONST"], ci), (P["ADD"], 0), (P["CONST"], 0), (P["MUL"], 0), (P["SIN"], 0)]
                if s:
                    prog.append((P["ADD"], 0))
            prog += [(P["VAR"], 0), (P["MUL"], 0)]
        ptr.append(len(prog))
    return prog, ptr, consts


def burgers_program(A, dim):
    P = A.PW
    prog, ptr = [], [0]
    for t in range(dim):
        prog += [(P["VAR"], 0), (P["SQR"], 0), (P["CONST"], 0), (P["MUL"], 0)]
        ptr.append(len(prog))
    return prog, ptr, [0.5]


def make_stage(A, S, D, dim, nmax, k, m, lev, sup, tables, flux, world, rank, device, stream_ptr, kernel, rk, dense=False, fuse_rk=False, dual_store=True):
    """DeviceStage of one rank for the grid (lev, sup); tables: compact bundles (bench) or dense dump tables (fixture).  fuse_rk: the RK combination
    rides in the epilogues of the right-hand-side sweeps (stage.StagePlan)"""
    a, b = k + 1, m + 1
    part = D.FibrePartition(lev, sup, world, rank) if world > 1 else None
    plan = S.StagePlan(dim, a, b, dim, part=part, fuse_rk=fuse_rk, dual_store=dual_store)

    def make_ops(c):
        if dense:
            reg = lambda nm, kf, kt: c.op_register(tables[nm], kf, kt)
            pt = c.op_register(tables["Lag_pt_Alpt_1D"].T.copy(), a, b)
            hier = c.op_register_hier(tables["lagr.pw_anc"], tables["lagr.pw_wt"])
        else:
            reg = lambda nm, kf, kt: c.op_register_compact(tables[nm])
            pt = c.op_register_compact(tables["pt"])
            hier = c.op_register_compact(tables["hier"], hier=True)
        uv, uvx = reg("lagr.u_v", b, a), reg("lagr.u_vx", b, a)
        uave = c.op_combine(reg("lagr.ulft_vjp", b, a), 1.0, reg("lagr.urgt_vjp", b, a), 1.0)      # ulft_vjp + urgt_vjp, source/FastMultiplyLU.cpp:1165
        c.points_set(tables["lagr.intep_pt"])
        return {"pt": pt, "uv": uv, "volflx": c.op_combine(uvx, 1.0, uave, 0.5), "pen": reg("alpt.ujp_vjp", a, a), "hier": hier}
    holder = {}                                    # id(context) -> its local element rows (filled in once the stage exists)
    if flux == "vlasov":
        # The field E_t(x) lives on the elements with level 0 in the velocity dimensions (the reference's aux-dimension DGSolution).  Every stage
        # it is evaluated at their interpolation points by one small point-wise launch, and the flux kernel reads it through the element map
        # DGSolution::copy_up_intp_to_f builds (reference source/DGSolution.cpp:1024-1065) -- the broadcast itself is never materialised.
        import torch
        P = A.PW
        hd = dim // 2
        frows = np.nonzero((lev[:, hd:] == 0).all(axis=1))[0]
        fkey = {tuple(lev[i, :hd]) + tuple(sup[i, :hd]): n for n, i in enumerate(frows)}
        fmap_all = np.array([fkey[tuple(lev[i, :hd]) + tuple(sup[i, :hd])] for i in range(lev.shape[0])], dtype=np.int32)
        fctx = A.Context(dim, nmax, k, m, device=device)
        if stream_ptr is not None:
            fctx.set_stream(stream_ptr)
        fctx.grid_set(lev[frows], sup[frows])
        fctx.points_set(tables["lagr.intep_pt"])
        E = torch.zeros(hd, len(frows), b ** dim, dtype=torch.float64, device="cuda")
        progE, ptrE, constsE = [], [0], [2.0 * np.pi]
        for t in range(hd):
            constsE.append(0.125 * (t + 1))
            for s_ in range(hd):
                progE += [(P["X"], s_), (P["CONST"], len(constsE) - 1), (P["ADD"], 0), (P["CONST"], 0), (P["MUL"], 0), (P["SIN"], 0)]
                if s_:
                    progE.append((P["ADD"], 0))
            ptrE.append(len(progE))
        prog, ptr = [], [0]
        for t in range(dim):
            prog += [(P["X"], hd + t), (P["VAR"], 0), (P["MUL"], 0)] if t < hd else [(P["OTHER"], t - hd), (P["VAR"], 0), (P["MUL"], 0)]
            ptr.append(len(prog))
        maps = {}

        def pointwise(c, up_ptr, fp_ptrs):
            if id(c) not in maps:
                maps[id(c)] = torch.from_numpy(np.ascontiguousarray(fmap_all[holder[id(c)]])).cuda()
            fctx.pointwise_expr([], [], None, [E[t] for t in range(hd)], progE, ptrE, constsE)
            c.pointwise_expr([up_ptr], [E[t] for t in range(hd)], maps[id(c)], fp_ptrs, prog, ptr, [])
        holder["close"] = fctx.close
    else:
        prog, ptr, consts = burgers_program(A, dim)

        def pointwise(c, up_ptr, fp_ptrs):
            c.pointwise_expr([up_ptr], [], None, fp_ptrs, prog, ptr, consts)

    def exchange(obj):
        import torch.distributed as dist
        out = [None] * world
        dist.all_gather_object(out, obj)
        return out
    st = S.DeviceStage(A, plan, lev, sup, nmax, k, m, device, make_ops, pointwise, -LXF_ALPHA / 2.0, rk, exchange=exchange, stream_ptr=stream_ptr, kernel=kernel)
    for L, c in st.ctx.items():
        holder[id(c)] = st.rows[L]
    if "close" in holder:
        close_stage, close_field = st.close, holder["close"]
        st.close = lambda: (close_stage(), close_field())
    return st, plan, part


def parity_check(A, S, D, world, rank, device, stream, kernel):
    """the stage program at this N on the reference's d=6 fixture, as the unfused plan (rhs array + RK kernel) and as the fused plan bench.py times
    (RK combination in the sweep epilogues): max over (rhs, stage update unfused, stage update fused) of the relative L2 error against the dump"""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import refdump
    d = refdump.load(os.path.join(ROOT, "tests", "golden", "cfg5_vlasov_d6_k1_n2.dump.xz"))
    dim, nmax, n0, sparse, pa, pl = [int(x) for x in d["config"][:6]]
    tables = d
    num = np.zeros(3)
    berr = 0
    for fuse in (False, True):
        with torch.cuda.stream(stream):
            st, plan, part = make_stage(A, S, D, dim, nmax, pa, pl, d["level"], d["suppt"], tables, "vlasov", world, rank, device, stream.cuda_stream, kernel,
                                        (A.RK_RK3SSP, 0, 0.001), dense=True, fuse_rk=fuse, dual_store=fuse)
            rows = part.local["X"] if part is not None else np.arange(d["level"].shape[0])
            u0 = torch.from_numpy(np.ascontiguousarray(d["ucoe_alpt.in"][:, 0, :][rows])).cuda()
            if len(rows):
                st.view("u").copy_(u0); st.view("u_tn").copy_(u0)
            stream.synchronize()
            if world > 1:
                import torch.distributed as dist
                dist.barrier()
            st.run()
        stream.synchronize()
        if len(rows):
            if not fuse:
                num[0] = np.linalg.norm(st.view("rhs").cpu().numpy() - d["rhs_all"][:, 0, :][rows]) ** 2
            num[2 if fuse else 1] = np.linalg.norm(st.view(plan.result).cpu().numpy() - d["stage0.ucoe_alpt"][:, 0, :][rows]) ** 2
        berr = max(berr, st.barrier_error())
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        st.close()
    t = torch.from_numpy(num).cuda()
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(t)
    num = t.cpu().numpy()
    nu = np.linalg.norm(d["stage0.ucoe_alpt"][:, 0, :])
    err = max(float(np.sqrt(num[0]) / np.linalg.norm(d["rhs_all"][:, 0, :])), float(np.sqrt(num[1]) / nu), float(np.sqrt(num[2]) / nu))
    return err, berr


def sweep_roofline(A, ctx, stream, op, sizes_of_t, kf, kt, dim, ne, flush, label, kernel_key):
    """device time per launch of single-job full sweeps along each dimension in turn over rotating buffers larger than L2 (graph replay)"""
    import torch
    peak, peak_src = peaks()
    nbuf = dim * ((8 + dim - 1) // dim)           # buffer i is always swept along dimension i % dim (the block shapes differ per dimension)
    with torch.cuda.stream(stream):
        bufs, dsts = [], []
        for i in range(nbuf):
            sz = sizes_of_t(i % dim)
            bufs.append(torch.rand(ne, int(np.prod(sz)), dtype=torch.float64, device="cuda"))
            dsts.append(torch.empty(ne, int(np.prod(sz)) // kf * kt, dtype=torch.float64, device="cuda"))
            ctx.sweep1d(op, A.REL_VOL, A.LU_FULL, i % dim, sz, bufs[i], dsts[i])
    stream.synchronize()
    nrep, n_replay = 4 * nbuf, 8
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=stream):
        for i in range(nrep):
            ctx.sweep1d(op, A.REL_VOL, A.LU_FULL, i % dim, sizes_of_t(i % dim), bufs[i % nbuf], dsts[i % nbuf])
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        g.replay()
        flush.fill_(0.0)
        e0.record(stream)
        for _ in range(n_replay):
            g.replay()
        e1.record(stream)
    stream.synchronize()
    t_launch = e0.elapsed_time(e1) / (nrep * n_replay) * 1e-3
    bytes_launch = float(np.mean([8.0 * ne * (np.prod(sizes_of_t(t)) + np.prod(sizes_of_t(t)) // kf * kt) for t in range(dim)]))    # B_sweep = 8 N_e (S_from + S_to)
    achieved = bytes_launch / t_launch / 1e9
    return {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": ncu_traffic(kernel_key),
            "kernel": label, "bytes_per_launch": bytes_launch, "us_per_launch": t_launch * 1e6, "peak_source": peak_src}


def stage_sweep_roofline(torch, S, st, plan, stream, kf, kt, kernel_label, kernel_key):
    """The dominant kernel of the stage in the mix the stage runs it: every kf -> kt sweep launch of the plan (the L sweeps of the down pass, the
    full sweeps along the last dimension and the accumulating U sweeps of the up pass, batched as in the stage), replayed from a CUDA graph;
    achieved = sum over jobs of B_sweep = 8 N_e (S_from + S_to) (read once, written once; the re-read of an accumulating sweep is not
    counted) / device time.  The launches touch far more than L2 (tens of GB per pass)."""
    peak, peak_src = peaks()
    ops = [o for o in plan.ops if o[0] == "sweep" and o[2] in ("uv", "volflx")]          # the kf -> kt operators of the right-hand side
    byts, n_jobs = 0.0, 0
    for o in ops:
        ne_loc = len(st.rows[o[1]])
        for j in o[6]:
            s_from = int(np.prod(j["sizes"]))
            byts += 8.0 * ne_loc * (s_from + s_from // kf * kt)
            n_jobs += 1
    if not ops or byts == 0:
        return None
    l0 = st.launch_count()
    with torch.cuda.stream(stream):
        for o in ops:
            st._run_op(o)
    stream.synchronize()
    n_launch = st.launch_count() - l0
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=stream):
        for o in ops:
            st._run_op(o)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n_replay = 4
    with torch.cuda.stream(stream):
        g.replay()
        e0.record(stream)
        for _ in range(n_replay):
            g.replay()
        e1.record(stream)
    stream.synchronize()
    t_pass = e0.elapsed_time(e1) / n_replay * 1e-3
    achieved = byts / t_pass / 1e9
    return {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": ncu_traffic(kernel_key),
            "kernel": kernel_label, "launches": int(n_launch), "sweep_jobs": int(n_jobs), "bytes_per_launch": byts / max(n_launch, 1),
            "us_per_launch": t_pass * 1e6 / max(n_launch, 1), "pass_ms": t_pass * 1e3, "peak_source": peak_src}


KERNEL_NAMES = {0: "sweep_tc_kernel", 5: "sweep_tc_kernel", 4: "sweep_mma_kernel", 8: "sweep_col_kernel"}


def timed_replays(torch, stream, flush, run_step, steps, barrier):
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    barrier()
    with torch.cuda.stream(stream):
        for s in range(steps):
            flush.fill_(0.0)                  # L2 flush between timed steps (outside the event pair)
            ev[s][0].record(stream)
            run_step()
            ev[s][1].record(stream)
    barrier()
    return float(np.sum([e0.elapsed_time(e1) for e0, e1 in ev])) / steps


def roundtrip_record(args, A, torch, stream, flush, local_rank, w, with_cpu):
    """cfg2: Alpert -> point values -> hierarchical coefficients -> Alpert on one GPU"""
    dim, k, m, nmax = w["dim"], w["k"], w["m"], w["nmax"]
    a, b = k + 1, m + 1
    lev, sup = A.sparse_grid(dim, nmax)
    keys = np.array([A.hash_key(l, s) for l, s in zip(lev, sup)])
    o = np.argsort(keys, kind="stable")
    lev, sup = lev[o], sup[o]
    ne = lev.shape[0]
    dof = ne * a ** dim
    ctx = A.Context(dim, nmax, k, m, device=local_rank)
    ctx.set_stream(stream.cuda_stream)
    ctx.set_schedule(args.schedule)
    ctx.set_kernel(args.kernel)
    ctx.grid_set(lev, sup)
    tb = load_tables(A, w)
    op_pt, op_uv, op_hier = ctx.op_register_compact(tb["pt"]), ctx.op_register_compact(tb["lagr.u_v"]), ctx.op_register_compact(tb["hier"], hier=True)
    vol = [A.REL_VOL] * dim
    host_in = torch.from_numpy(synthetic_field(lev, a ** dim, 1, 20240901)).pin_memory()
    host_out = torch.empty_like(host_in).pin_memory()
    hin, hout = host_in.numpy(), host_out.numpy()
    with torch.cuda.stream(stream):
        u = host_in.to("cuda", non_blocking=True)
        up = torch.zeros(1, ne, b ** dim, dtype=torch.float64, device="cuda")
        out = torch.zeros(1, ne, a ** dim, dtype=torch.float64, device="cuda")

    def step():
        ctx.apply_tensor([op_pt] * dim, vol, u, up)                # Alpert -> point values
        ctx.hierarchize(op_hier, up, up)                            # point values -> hierarchical coefficients
        ctx.apply_tensor([op_uv] * dim, vol, up, out)               # -> Alpert
    sync = lambda: torch.cuda.synchronize()
    with torch.cuda.stream(stream):
        for _ in range(max(args.warmup, 3)):
            step()
    sync()
    l_before = ctx.launch_count
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=stream):
        step()
    launches_per_step = ctx.launch_count - l_before
    with torch.cuda.stream(stream):
        graph.replay()
    sync()
    t_ms = timed_replays(torch, stream, flush, graph.replay, args.steps, sync)
    for _ in range(2):
        ctx.host_roundtrip(op_pt, op_hier, op_uv, hin, out=hout)
    sync()
    t0 = time.perf_counter()
    n_e2e = max(3, min(args.steps, 10))
    for _ in range(n_e2e):
        ctx.host_roundtrip(op_pt, op_hier, op_uv, hin, out=hout)
    sync()
    t_e2e = (time.perf_counter() - t0) / n_e2e
    kname = KERNEL_NAMES.get(args.kernel, "sweep kernel %d" % args.kernel)
    roof = sweep_roofline(A, ctx, stream, op_pt, lambda t: [a] * dim, a, b, dim, ne, flush,
                          "%s<%d,%d> (one 1D sweep along each dimension in turn, single job; FP64 DMMA m8n8k4)" % (kname, a, b), "%s<%d,%d>" % (kname, a, b))
    chain = sum((a ** (dim - i) * b ** i + a ** (dim - i - 1) * b ** (i + 1)) for i in range(dim))
    b_alg = 8.0 * ne * (2 * 2 ** (dim - 1) * chain + dim * 2 * b ** dim)
    peak = roof["peak"]
    roof["step"] = {"b_alg_bytes": b_alg, "gbs": b_alg / (t_ms * 1e-3) / 1e9, "frac": b_alg / (t_ms * 1e-3) / 1e9 / peak,
                    "note": "reference sweep list bytes / measured step time; the shared-prefix schedule runs %d instead of %d sweeps per transform" % (3 * 2 ** (dim - 1) - 2, dim * 2 ** (dim - 1))}
    rec = {"workload": w["desc"], "metric": "sparse-grid DoF-stage updates/sec (FP64)", "value": dof / (t_ms * 1e-3), "unit": "DoF-stage/s", "ms_per_step": t_ms,
           "n_elem": int(ne), "dof": int(dof), "gpu_launches": int(launches_per_step * args.steps), "kernel": args.kernel,
           "e2e": {"value": dof / t_e2e, "unit": "DoF-stage/s", "h2d_bytes_per_step": int(hin.nbytes), "d2h_bytes_per_step": int(hout.nbytes), "ms_per_step": t_e2e * 1e3,
                   "api": "amdg_host_roundtrip (pinned host buffers)"},
           "roofline": roof}
    if with_cpu:
        cb = run_reference(args, w, as_baseline=True)
        if cb:
            rec["cpu_baseline"] = {kk: cb[kk] for kk in ("value", "unit", "cores", "kind", "sample")}
    ctx.close()
    return rec


def linear_record(args, A, torch, stream, flush, local_rank, w, with_cpu):
    """cfg1 / cfg3: one RK stage of a linear operator applied as single 1D sweeps (FastRHS::transform_ucoe_alpt_to_rhs, reference
    source/FastMultiplyLU.cpp:46-59; the shipped examples assemble the same operator as a sparse matrix, source/BilinearForm.cpp:691-710, 877-929)"""
    dim, k, nmax = w["dim"], w["k"], w["nmax"]
    a = k + 1
    lev, sup = A.sparse_grid(dim, nmax)
    keys = np.array([A.hash_key(l, s_) for l, s_ in zip(lev, sup)])
    o = np.argsort(keys, kind="stable")
    lev, sup = lev[o], sup[o]
    ne = lev.shape[0]
    dof = ne * a ** dim
    ctx = A.Context(dim, nmax, k, w["m"], device=local_rank)
    ctx.set_stream(stream.cuda_stream)
    ctx.set_kernel(args.kernel)
    ctx.grid_set(lev, sup)
    tb = load_tables(A, w)
    reg = lambda nm: ctx.op_register_compact(tb["alpt." + nm])
    if w["op"] == "advection":
        op = ctx.op_combine(reg("u_vx"), 1.0, reg("ulft_vjp"), 1.0)                     # volume + upwind flux (c >= 0: left trace), one operator under the flx relation
    else:
        sigma_dx = 20.0 * 2 ** nmax
        op = ctx.op_combine(ctx.op_combine(reg("ux_vx"), -1.0, reg("uxave_vjp"), -1.0), 1.0, ctx.op_combine(reg("ujp_vxave"), -1.0, reg("ujp_vjp"), -sigma_dx), 1.0)
    host_in = torch.from_numpy(synthetic_field(lev, a ** dim, 1, 20240901)).pin_memory()
    host_out = torch.empty_like(host_in).pin_memory()
    with torch.cuda.stream(stream):
        u = host_in.to("cuda", non_blocking=True)
        u_tn, v, v_tn, rhs = u.clone(), u.clone(), u.clone(), torch.zeros_like(u)
        ku, kv = torch.zeros(4, *u.shape, dtype=torch.float64, device="cuda"), torch.zeros(4, *u.shape, dtype=torch.float64, device="cuda")

    def step():
        for t in range(dim):
            ctx.sweep1d(op, A.REL_FLX, A.LU_FULL, t, [a] * dim, u, rhs, accumulate=t > 0)
        if w["op"] == "advection":
            ctx.rk_stage(A.RK_RK3SSP, 1, DT, u_tn, u, rhs)
        else:
            ctx.rk4_ode2nd_stage(1, DT, u_tn, v_tn, u, v, rhs, ku, kv)
    sync = lambda: torch.cuda.synchronize()
    with torch.cuda.stream(stream):
        for _ in range(max(args.warmup, 3)):
            step()
    sync()
    l0 = ctx.launch_count
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=stream):
        step()
    launches = ctx.launch_count - l0
    with torch.cuda.stream(stream):
        graph.replay()
    sync()
    t_ms = timed_replays(torch, stream, flush, graph.replay, args.steps, sync)

    def e2e_call():
        with torch.cuda.stream(stream):
            u.copy_(host_in, non_blocking=True)
            graph.replay()
            host_out.copy_(u, non_blocking=True)
        stream.synchronize()
    for _ in range(2):
        e2e_call()
    t0 = time.perf_counter()
    n_e2e = max(3, min(args.steps, 10))
    for _ in range(n_e2e):
        e2e_call()
    t_e2e = (time.perf_counter() - t0) / n_e2e
    peak, peak_src = peaks()
    byts = 8.0 * ne * 2 * a ** dim
    # the sweeps alone (graph of dim launches)
    g2 = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g2, stream=stream):
        for t in range(dim):
            ctx.sweep1d(op, A.REL_FLX, A.LU_FULL, t, [a] * dim, u, rhs, accumulate=t > 0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        g2.replay(); e0.record(stream)
        for _ in range(20):
            g2.replay()
        e1.record(stream)
    sync()
    t_launch = e0.elapsed_time(e1) / (20 * dim) * 1e-3
    kname = "sweep_mma_kernel" if args.kernel == 0 else KERNEL_NAMES.get(args.kernel, "sweep kernel %d" % args.kernel)
    roof = {"bound": "hbm", "achieved": byts / t_launch / 1e9, "peak": peak, "unit": "GB/s", "frac": byts / t_launch / 1e9 / peak, "traffic": None,
            "kernel": "%s<%d,%d> (one 1D sweep, %.2f MB vector: L2 resident, the launch is latency bound -- SURVEY.md 8d caveat)" % (kname, a, a, byts / 2e6),
            "bytes_per_launch": byts, "us_per_launch": t_launch * 1e6, "peak_source": peak_src}
    rec = {"metric": "sparse-grid DoF-stage updates/sec (FP64)", "value": dof / (t_ms * 1e-3), "unit": "DoF-stage/s", "ms_per_step": t_ms, "n_elem": int(ne), "dof": int(dof),
           "gpu_launches": int(launches * args.steps), "launches_per_stage": int(launches),
           "e2e": {"value": dof / t_e2e, "unit": "DoF-stage/s", "h2d_bytes_per_step": int(host_in.numel() * 8), "d2h_bytes_per_step": int(host_out.numel() * 8), "ms_per_step": t_e2e * 1e3,
                   "api": "pinned host buffer -> device, the stage's C-ABI calls (CUDA-graph replay), device -> pinned host buffer"},
           "roofline": roof}
    if with_cpu:
        cb = run_reference(args, w, as_baseline=True)
        if cb:
            rec["cpu_baseline"] = {kk: cb[kk] for kk in ("value", "unit", "cores", "kind", "sample")}
    ctx.close()
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--workload", default="cfg5")
    ap.add_argument("--kernel", type=int, default=0)
    ap.add_argument("--schedule", type=int, default=1)
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-dual-store", action="store_true", help="partitioned runs: move the buffers that both layouts consume by row scatters instead of second destinations of their producing sweeps")
    ap.add_argument("--no-fuse-rk", action="store_true", help="separate rhs array, per-application sums and RK kernel instead of the RK combination in the sweep epilogues")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (quick kernel comparisons)")
    ap.add_argument("--no-secondary", action="store_true", help="skip the cfg2 secondary record at N = 1")
    ap.add_argument("--breakdown", action="store_true", help="also report device time per kind of operation of the stage (eager launches, CUDA events; max over ranks)")
    args = ap.parse_args()
    w = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    # the ONE JSON line goes to the real stdout; everything else a library prints there (NCCL's version banner, ...) is sent to stderr
    out_fd = os.dup(1)
    sys.stdout.flush()
    os.dup2(2, 1)

    def emit(obj):
        os.write(out_fd, (json.dumps(obj) + "\n").encode())

    if args.impl == "reference":
        if rank != 0:
            return 0
        r = run_reference(args, w)
        line = {"impl": "reference", "metric": "sparse-grid DoF-stage updates/sec (FP64)", "value": r["value"], "unit": "DoF-stage/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "strong" if w["kind"] == "stage" else "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": w["desc"], "sample": r["sample"], "same_config": r["same_config"]},
                "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": r["value"], "unit": "DoF-stage/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        emit(line)
        return 0

    import torch
    import torch.distributed as dist
    assert torch.cuda.is_available(), "bench.py needs a CUDA device: there is no CPU path"
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    A = importlib.import_module("adaptive-multiresolution-dg_b200")
    S = importlib.import_module("adaptive-multiresolution-dg_b200.stage")
    D = importlib.import_module("adaptive-multiresolution-dg_b200.dist")
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device="cuda")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if w["kind"] in ("roundtrip", "linear"):
        if world > 1:
            raise SystemExit("this workload has no exchange step: run it on one GPU (the multi-GPU benchmark is the cfg5 stage)")
        sampler = ClockSampler(local_rank)
        sampler.start()
        rec = (roundtrip_record if w["kind"] == "roundtrip" else linear_record)(args, A, torch, stream, flush, local_rank, w, with_cpu=not args.no_cpu)
        clocks = sampler.finish()
        line = {"metric": rec["metric"], "value": rec["value"], "unit": rec["unit"], "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": rec["ms_per_step"],
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": w["desc"], "n_elem": rec["n_elem"], "dof": rec["dof"], "kernel": args.kernel, "cuda_graph": True,
                           "l2": "flushed (256 MiB write) between timed steps; per-step CUDA event pairs on the launch stream"},
                "clocks": clocks, "e2e": rec["e2e"], "gpu_launches": rec["gpu_launches"], "roofline": rec["roofline"]}
        if "cpu_baseline" in rec:
            line["cpu_baseline"] = rec["cpu_baseline"]
        emit(line)
        return 0

    # ---------------------------------------------------------------- the stage workloads (cfg5 default, cfg4)
    dim, k, m, nmax = w["dim"], w["k"], w["m"], w["nmax"]
    a, b = k + 1, m + 1
    parity = parity_check(A, S, D, world, rank, local_rank, stream, args.kernel) if dim == 6 else None
    lev, sup = A.sparse_grid(dim, nmax)
    keys = np.array([A.hash_key(l, s) for l, s in zip(lev, sup)])
    o = np.argsort(keys, kind="stable")
    lev, sup = lev[o], sup[o]
    ne = lev.shape[0]
    dof = ne * a ** dim
    tb = load_tables(A, w)
    with torch.cuda.stream(stream):
        st, plan, part = make_stage(A, S, D, dim, nmax, k, m, lev, sup, tb, w["flux"], world, rank, local_rank, stream.cuda_stream, args.kernel, (A.RK_RK3SSP, 1, DT),
                                    fuse_rk=not args.no_fuse_rk, dual_store=not args.no_dual_store)
    rows = part.local["X"] if part is not None else np.arange(ne)
    u_all = synthetic_field(lev, a ** dim, 1, 20240901)[0]
    host_in = torch.from_numpy(np.ascontiguousarray(u_all[rows])).pin_memory()
    host_out = torch.empty_like(host_in).pin_memory()
    hin, hout = host_in.numpy(), host_out.numpy()
    d_u = ctypes.c_void_p(st.local_ptr("u"))
    d_res = ctypes.c_void_p(st.local_ptr(plan.result))          # the stage's result: "u" (in place) or the RK accumulator of the fused plan
    c0 = st.ctx["X"]
    dp = ctypes.POINTER(ctypes.c_double)
    with torch.cuda.stream(stream):
        if len(rows):
            st.view("u").copy_(host_in, non_blocking=True)
            st.view("u_tn").copy_(host_in, non_blocking=True)
    stream.synchronize()
    barrier()
    with torch.cuda.stream(stream):
        for _ in range(max(args.warmup, 3)):
            st.run()
    barrier()
    graph, launches_per_step = None, None
    l_before = st.launch_count()
    if not args.no_graph:
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=stream):
            st.run()
        launches_per_step = st.launch_count() - l_before
        with torch.cuda.stream(stream):
            graph.replay()
        barrier()
    else:
        with torch.cuda.stream(stream):
            st.run()
        launches_per_step = st.launch_count() - l_before
        barrier()
    run_step = (lambda: graph.replay()) if graph is not None else st.run
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.3)
    t_step_ms = timed_replays(torch, stream, flush, run_step, args.steps, barrier)

    # ---- e2e: host buffers -> device, the stage, device -> host, every step
    def e2e_call():
        if len(rows):
            A.lib.amdg_dev_upload(c0._h, d_u, hin.ctypes.data_as(dp), hin.size)
        with torch.cuda.stream(stream):
            run_step()
        if len(rows):
            A.lib.amdg_dev_download(c0._h, hout.ctypes.data_as(dp), d_res, hout.size)
        c0.sync()
    for _ in range(2):
        e2e_call()
    barrier()
    t0 = time.perf_counter()
    n_e2e = max(3, min(args.steps, 10))
    for _ in range(n_e2e):
        e2e_call()
    barrier()
    t_e2e = (time.perf_counter() - t0) / n_e2e
    clocks = sampler.finish()
    berr = st.barrier_error()
    breakdown = None
    if args.breakdown:
        acc = {}
        for _ in range(5):
            prof = {}
            barrier()
            with torch.cuda.stream(stream):
                st.run(prof)
            stream.synchronize()
            for kk, vv in S.DeviceStage.profile_summary(prof).items():
                acc[kk] = acc.get(kk, 0.0) + vv / 5.0
        keys = sorted(acc)
        bt = torch.tensor([acc[kk] for kk in keys], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(bt, op=dist.ReduceOp.MAX)          # the plan (and so the key list) is the same on every rank
        breakdown = {kk: float(v) for kk, v in zip(keys, bt.cpu().numpy())}

    tt = torch.tensor([t_step_ms, t_e2e, float(len(rows)), float(len(part.local["V"]) if part is not None else ne), float(st.sent_bytes), float(berr)], dtype=torch.float64, device="cuda")
    mx, sm = tt.clone(), tt.clone()
    if world > 1:
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
    t_step_ms, t_e2e = float(mx[0]), float(mx[1])
    value = dof / (t_step_ms * 1e-3)

    # roofline of the dominant kernel of the stage: the b -> a sweeps of the right-hand-side applications (d of the d+1 tensor applications), in the
    # mix of L / full / accumulating-U launches the stage runs (every rank replays its own launches; rank 0 reports its GPU)
    col = args.kernel in (0, 8) and b * a <= 9 and dim >= 3
    kname = "sweep_col_kernel" if col else KERNEL_NAMES.get(args.kernel, "sweep kernel %d" % args.kernel)
    roof = stage_sweep_roofline(torch, S, st, plan, stream, b, a,
                                "%s<%d,%d,*>: all %d -> %d sweep launches of one stage (L sweeps, full sweeps along the last dimension, accumulating U sweeps of the shared-prefix schedule; FP64 %s)"
                                % (kname, b, a, b, a, "DFMA, one thread per column" if col else "DMMA m8n8k4"), "%s<%d,%d>" % (kname, b, a))
    barrier()
    if rank == 0:
        if roof is None:
            peak, peak_src = peaks()
            roof = {"bound": "hbm", "achieved": None, "peak": peak, "unit": "GB/s", "frac": None, "traffic": None, "kernel": kname, "peak_source": peak_src}
        C = lambda p, q: sum((p ** (dim - i) * q ** i + p ** (dim - i - 1) * q ** (i + 1)) for i in range(dim))
        n_ch, nf = 2 ** (dim - 1), dim
        b_alg = 8.0 * ne * (n_ch * C(a, b) + (1 + nf) * b ** dim + nf * dim * 2 * b ** dim + 2 * nf * n_ch * C(b, a) + dim * 2 * a ** dim + 4 * a ** dim)
        roof["step"] = {"b_alg_bytes": b_alg, "gbs": b_alg / (t_step_ms * 1e-3) / 1e9, "frac_of_n_gpus": b_alg / (t_step_ms * 1e-3) / 1e9 / (roof["peak"] * world),
                        "note": "reference sweep list bytes / measured stage time / (N x measured HBM peak); every tensor application runs %d instead of %d sweeps (shared-prefix schedule), vol + flx of a dimension are one application" % (3 * n_ch - 2, dim * n_ch)}
        line = {
            "metric": "sparse-grid DoF-stage updates/sec (FP64)", "value": value, "unit": "DoF-stage/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": t_step_ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": w["desc"], "n_elem": int(ne), "dof": int(dof), "points": int(ne * b ** dim), "kernel": args.kernel, "cuda_graph": graph is not None,
                       "l2": "flushed (256 MiB write) between timed steps; per-step CUDA event pairs on the launch stream",
                       "multi_gpu": ("fibre-partitioned: %d ranks, layout switches = stores into IPC-mapped peer memory from the sweep epilogues + row scatters, %d device-side barriers per stage"
                                     % (world, plan.n_barrier)) if world > 1 else "single GPU (same batched program, one layout)",
                       "launches_per_stage": int(launches_per_step), "barriers_per_stage": int(plan.n_barrier),
                       "row_scatters_per_stage": int(sum(1 for o in plan.ops if o[0] == "scatter")),
                       "rk_update": "in the epilogues of the right-hand-side sweeps (no rhs array, no RK kernel)" if plan.fuse_rk else "separate kernel",
                       "exchange_bytes_per_stage_all_ranks": float(sm[4]), "exchange_doubles_per_element": int(plan.push_bytes) if world > 1 else 0,
                       "max_local_elements": [int(mx[2]), int(mx[3])], "ideal_local_elements": ne / world,
                       "parity_rel_l2": (parity[0] if parity else None), "parity_fixture": "tests/golden/cfg5_vlasov_d6_k1_n2 (reference dump): rhs and RK stage after one stage at this N, unfused and fused plan" if parity else None,
                       "barrier_timeouts": int(mx[5]) + (parity[1] if parity else 0)},
            "clocks": clocks,
            "e2e": {"value": dof / t_e2e, "unit": "DoF-stage/s", "h2d_bytes_per_step": int(dof * 8), "d2h_bytes_per_step": int(dof * 8), "ms_per_step": t_e2e * 1e3,
                    "api": "amdg_dev_upload + the stage's C-ABI calls (CUDA-graph replay) + amdg_dev_download, pinned host buffers, per rank its own rows"},
            "gpu_launches": int(launches_per_step * args.steps),
            "roofline": roof,
        }
        if breakdown is not None:
            line["config"]["breakdown_ms_eager_max_over_ranks"] = breakdown
    st.close()
    if rank == 0:
        if world == 1 and not args.no_cpu:
            cb = run_reference(args, w, as_baseline=True)
            if cb:
                line["cpu_baseline"] = {kk: cb[kk] for kk in ("value", "unit", "cores", "kind", "sample")}
        if world == 1 and not args.no_secondary and args.workload == "cfg5":
            line["secondary"] = roundtrip_record(args, A, torch, stream, flush, local_rank, WORKLOADS["cfg2"], with_cpu=not args.no_cpu)
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
