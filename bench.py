#!/usr/bin/env python
"""Benchmark of the fast sparse-grid transform path (BASELINE.json metric: sparse-grid DoF-stage updates/s, FP64).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2|cfg4|cfg5]

One "step" = one pass of the hot path over one batch of synthetic input:
  cfg2 (default, the config the metric is quoted on that fits one GPU): the Lagrange interpolation round trip
       d=4, k=3, m=3, NMAX=8 full sparse grid -- FastLagrIntp::eval_up_Lagr -> eval_up_to_coe_D_Lag ->
       FastLagrInit::eval_ucoe_Alpt_Lagr (reference source/FastMultiplyLU.cpp:1362-1365, 1617-1620,
       source/Interplation.cpp:891-1048) -- one DoF-stage update = one DoF through one forward+inverse transform.
Multi-GPU (torchrun, one rank per GPU): the round trip has no exchange step, so the ranks run independent
grids-worth of components (weak scaling, no data-path collective); the barrier + max-over-ranks timing uses NCCL.

Timing: CUDA events on the stream the kernels are launched on, every timed step bracketed by its own event
pair with an L2 flush (256 MiB write) between steps, max over ranks.  `e2e` goes through the host-buffer C-ABI
entry point (amdg_host_roundtrip) with pinned host buffers: H2D + kernels + D2H inside the timed region.
The reference arm (--impl reference) and `cpu_baseline` time the compiled, unmodified reference
(oracle/_ref/ref_harness) on the host cores on a bounded sample (smaller NMAX) of the same workload.
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
    # name: dim, k, m, nmax, sample nmax for the CPU arm
    "cfg2": dict(kind="roundtrip", dim=4, k=3, m=3, nmax=8, cpu_nmax=7, ref_nmax=6, desc="example/01_interp_01_high_dim: Lagrange interpolation round trip d=4 k=3 m=3 NMAX=8 (full sparse grid)"),
    # one nonlinear RK stage (interpolate -> point-wise products -> hierarchise -> vol + flx + penalty sweeps -> RK3SSP stage)
    "cfg5": dict(kind="stage", flux="vlasov", dim=6, k=1, m=2, nmax=7, cpu_nmax=4, ref_nmax=3, desc="example/07_vlasov_maxwell_sparse scaled to 3D3V: d=6 k=1 m=2 NMAX=7, one nonlinear RK3SSP stage with a prescribed smooth field"),
    "cfg4": dict(kind="stage", flux="burgers", dim=2, k=2, m=3, nmax=7, cpu_nmax=7, ref_nmax=7, desc="example/02_hyperbolic_05_burgers_adapt (static upper-bound grid NMAX=7, Lagrange flux): one nonlinear RK3SSP stage"),
}


# dram__bytes_read.sum + dram__bytes_write.sum of one single-job full sweep launch (ncu --set full, profiles/r01_sweep_tc_ncu.md,
# launch 1); the written half of the compulsory bytes is still in L2 when the kernel ends, so only the reads show up
TRAFFIC_NCU = 23.1e6


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


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
    path = os.path.join(ROOT, "adaptive-multiresolution-dg_b200", "data", "tables_k%d_m%d_n%d.npz" % (w["k"], w["m"], w["nmax"]))
    return np.load(path)


def synthetic_field(level, block, n_comp, seed):
    """i.i.d. U(-1,1) * 2^-(n_1+...+n_d): the decay keeps the hierarchical sums well conditioned"""
    rng = np.random.default_rng(seed)
    scale = np.ldexp(1.0, -level.sum(axis=1).astype(np.int64))
    return (rng.uniform(-1.0, 1.0, size=(n_comp, level.shape[0], block)) * scale[None, :, None])


def run_reference(args, w, n_threads=None, as_baseline=False):
    """the compiled, unmodified reference on the host cores, bounded sample"""
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_harness")
    if not os.path.exists(exe):
        return None
    nthr = n_threads or os.cpu_count()
    reps = max(1, args.steps if not as_baseline else 2)
    nmax = w["cpu_nmax"] if as_baseline else w["ref_nmax"]
    cmd = [exe, "--dim", str(w["dim"]), "--nmax", str(nmax), "--pa", str(w["k"]), "--pl", str(w["m"]),
           "--time", str(reps + (args.warmup if not as_baseline else 1)), "--threads", str(nthr)]
    if w["kind"] == "roundtrip":
        cmd += ["--run", "roundtrip"]
        phases = ("intp", "hier", "init")
        what = "same round trip"
    else:
        cmd += ["--run", "rhs", "--flux", w["flux"]]
        phases = ("intp", "pointwise", "hier", "rhs_vol", "rhs_flx", "rhs_penalty")
        what = "same nonlinear stage (interpolate, point-wise, hierarchise, vol + flx + penalty sweeps; RK axpy not counted)"
    t0 = time.time()
    out = subprocess.run(cmd, check=True, stdout=subprocess.PIPE, text=True).stdout.strip().splitlines()[-1]
    wall = time.time() - t0
    r = json.loads(out)
    t_step = sum(r[p] for p in phases)          # medians over the repetitions
    dof = r["dof"]
    return {"value": dof / t_step, "unit": "DoF-stage/s", "cores": r["threads"], "kind": "reference",
            "sample": "%s at NMAX=%d (%d elements, %d DoF): median of %d repetitions, %.3f s per step; reference built from /root/reference/source with -O3 -fopenmp (oracle/Makefile)"
                      % (what, nmax, r["n_elem"], dof, r["reps"], t_step),
            "ms_per_step": t_step * 1e3, "wall_s": wall}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--kernel", type=int, default=0)
    ap.add_argument("--schedule", type=int, default=1)
    ap.add_argument("--ncomp", type=int, default=1)
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (quick kernel comparisons)")
    ap.add_argument("--literal-rhs", action="store_true", help="stage workloads: rhs_vol and rhs_flx as separate tensor applications")
    args = ap.parse_args()
    w = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank != 0:
            return 0
        r = run_reference(args, w)
        line = {"impl": "reference", "metric": "sparse-grid DoF-stage updates/sec (FP64)", "value": r["value"], "unit": "DoF-stage/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": w["desc"], "sample": r["sample"]},
                "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": r["value"], "unit": "DoF-stage/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    import torch
    import torch.distributed as dist
    assert torch.cuda.is_available(), "bench.py needs a CUDA device: there is no CPU path"
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    A = importlib.import_module("adaptive-multiresolution-dg_b200")

    dim, k, m, nmax = w["dim"], w["k"], w["m"], w["nmax"]
    a, b = k + 1, m + 1
    lev, sup = A.sparse_grid(dim, nmax)
    keys = np.array([A.hash_key(l, s) for l, s in zip(lev, sup)])
    o = np.argsort(keys, kind="stable")
    lev, sup = lev[o], sup[o]
    ne = lev.shape[0]
    ncomp = args.ncomp
    dof = ne * a ** dim * ncomp

    stream = torch.cuda.Stream()
    ctx = A.Context(dim, nmax, k, m, device=local_rank)
    ctx.set_stream(stream.cuda_stream)
    ctx.set_schedule(args.schedule)
    ctx.set_kernel(args.kernel)
    ctx.grid_set(lev, sup)
    tb = load_tables(A, w)
    op_pt = ctx.op_register_compact(tb["pt"])
    op_uv = ctx.op_register_compact(tb["lagr.u_v"])
    op_hier = ctx.op_register_compact(tb["hier"], hier=True)
    vol = [A.REL_VOL] * dim

    host_in = torch.from_numpy(synthetic_field(lev, a ** dim, ncomp, 20240901 + rank)).pin_memory()
    host_out = torch.empty_like(host_in).pin_memory()
    hin, hout = host_in.numpy(), host_out.numpy()
    with torch.cuda.stream(stream):
        u = host_in.to("cuda", non_blocking=True)
        flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device="cuda")

    if w["kind"] == "roundtrip":
        with torch.cuda.stream(stream):
            up = torch.zeros(ncomp, ne, b ** dim, dtype=torch.float64, device="cuda")
            out = torch.zeros(ncomp, ne, a ** dim, dtype=torch.float64, device="cuda")
        ops_f, ops_i = [op_pt] * dim, [op_uv] * dim

        def step():
            ctx.apply_tensor(ops_f, vol, u, up, n_comp=ncomp)          # Alpert -> point values
            ctx.hierarchize(op_hier, up, up, n_comp=ncomp)              # point values -> hierarchical coefficients
            ctx.apply_tensor(ops_i, vol, up, out, n_comp=ncomp)         # -> Alpert

        def e2e_call():
            ctx.host_roundtrip(op_pt, op_hier, op_uv, hin, n_comp=ncomp, out=hout)
        e2e_api = "amdg_host_roundtrip (pinned host buffers)"
        # algorithmic bytes of the reference's sweep list: 2*2^(d-1)*C(d,a,b) + d*2*b^d doubles per element
        chain = sum((a ** (dim - i) * b ** i + a ** (dim - i - 1) * b ** (i + 1)) for i in range(dim))
        b_alg = 8.0 * ne * ncomp * (2 * 2 ** (dim - 1) * chain + dim * 2 * b ** dim)
        step_note = "reference sweep list bytes / measured step time; the shared-prefix schedule runs %d instead of %d sweeps per transform" % (3 * 2 ** (dim - 1) - 2, dim * 2 ** (dim - 1))
    else:
        assert ncomp == 1
        # one nonlinear RK3SSP stage of a scalar conservation law, Lagrange flux interpolation
        # (the loop of example/07_vlasov_ampere_02_2D2V_accuracy.cpp:247-303)
        op_uvx = ctx.op_register_compact(tb["lagr.u_vx"])
        op_ul = ctx.op_register_compact(tb["lagr.ulft_vjp"])
        op_ur = ctx.op_register_compact(tb["lagr.urgt_vjp"])
        op_uave = ctx.op_combine(op_ul, 1.0, op_ur, 1.0)                 # ulft_vjp + urgt_vjp, source/FastMultiplyLU.cpp:1165
        op_volflx = ctx.op_combine(op_uvx, 1.0, op_uave, 0.5)            # u_vx + (ulft_vjp + urgt_vjp)/2 under the flx relation
        op_pen = ctx.op_register_compact(tb["alpt.ujp_vjp"])
        nf = dim
        flux_ids = [A.FLUX_VLASOV_SMOOTH_E if w["flux"] == "vlasov" else A.FLUX_BURGERS] * nf
        prm = [[t, 0, 0, 0] for t in range(nf)]
        lxf_alpha, dt = 1.2, 1e-4
        with torch.cuda.stream(stream):
            up = torch.zeros(ne, b ** dim, dtype=torch.float64, device="cuda")
            pts = torch.zeros(ne, b ** dim, dim, dtype=torch.float64, device="cuda")
            fp = torch.zeros(nf, ne, b ** dim, dtype=torch.float64, device="cuda")
            fuc = torch.zeros_like(fp)
            rhs = torch.zeros(ne, a ** dim, dtype=torch.float64, device="cuda")
            u_tn = u.clone()
            ctx.point_coords(tb["lagr.intep_pt"], pts)
        merged = not args.literal_rhs

        def step():
            ctx.apply_tensor([op_pt] * dim, vol, u, up)                                  # FastLagrIntp::eval_up_Lagr
            ctx.pointwise(flux_ids, prm, up, fp, pts)                                    # eval_fp_Lag / Vlasov products
            ctx.hierarchize(op_hier, fp, fuc, n_comp=nf)                                 # eval_fp_to_coe_D_Lag
            for t in range(dim):
                if merged:      # rhs_vol_scalar + rhs_flx_intp_scalar of dim t as ONE tensor application
                    ops = [op_volflx if s == t else op_uv for s in range(dim)]
                    rels = [A.REL_FLX if s == t else A.REL_VOL for s in range(dim)]
                    ctx.apply_tensor(ops, rels, fuc[t], rhs, accumulate=t > 0)           # t == 0 overwrites: DGSolution::set_rhs_zero
                else:
                    ctx.apply_tensor([op_uvx if s == t else op_uv for s in range(dim)], vol, fuc[t], rhs, accumulate=t > 0)
                    ctx.apply_tensor([op_uave if s == t else op_uv for s in range(dim)], [A.REL_FLX if s == t else A.REL_VOL for s in range(dim)],
                                     fuc[t], rhs, coef=0.5, accumulate=True)
            for t in range(dim):                                                         # HyperbolicAlptRHS::rhs_flx_penalty_scalar
                ctx.sweep1d(op_pen, A.REL_FLX, A.LU_FULL, t, [a] * dim, u, rhs, coef=-lxf_alpha / 2.0, accumulate=True)
            ctx.rk_stage(A.RK_RK3SSP, 1, dt, u_tn, u, rhs)                               # RK3SSP::step_stage(1): u <- 3/4 u_n + 1/4 (u + dt rhs)

        d_u = ctypes.c_void_p(u.data_ptr())

        def e2e_call():
            A.lib.amdg_dev_upload(ctx._h, d_u, hin.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), hin.size)
            step()
            A.lib.amdg_dev_download(ctx._h, hout.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), d_u, hout.size)
            ctx.sync()
        e2e_api = "amdg_dev_upload + the stage's C-ABI calls + amdg_dev_download (pinned host buffers)"
        C = lambda p, q: sum((p ** (dim - i) * q ** i + p ** (dim - i - 1) * q ** (i + 1)) for i in range(dim))
        n_ch = 2 ** (dim - 1)
        # interpolation + point-wise + hierarchisation + vol and flx tensor applications + penalty sweeps + 4 axpy vectors
        b_alg = 8.0 * ne * (n_ch * C(a, b) + (1 + nf) * b ** dim + nf * dim * 2 * b ** dim + 2 * nf * n_ch * C(b, a) + dim * 2 * a ** dim + 4 * a ** dim)
        step_note = ("reference sweep list bytes / measured step time; here every tensor application runs %d instead of %d sweeps (shared-prefix schedule)%s"
                     % (3 * n_ch - 2, dim * n_ch, " and vol + flx of a dimension are one application (u_vx + (ulft_vjp+urgt_vjp)/2 under the flx relation)" if merged else ""))
    stream.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.cuda.stream(stream):
        for _ in range(max(args.warmup, 3)):
            step()
    barrier()
    # the step is a fixed sequence of ~36 kernel launches: capture it once into a CUDA graph and replay it
    graph = None
    launches_per_step = None
    if not args.no_graph:
        l_before = ctx.launch_count
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=stream):
            step()
        launches_per_step = ctx.launch_count - l_before
        with torch.cuda.stream(stream):
            graph.replay()
        barrier()
    run_step = (lambda: graph.replay()) if graph is not None else step
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.3)
    l0 = ctx.launch_count
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    with torch.cuda.stream(stream):
        for s in range(args.steps):
            flush.fill_(0.0)                  # L2 flush between timed steps (outside the event pair)
            ev[s][0].record(stream)
            run_step()
            ev[s][1].record(stream)
    barrier()
    launches = (launches_per_step * args.steps) if graph is not None else (ctx.launch_count - l0)
    times = np.array([e0.elapsed_time(e1) for e0, e1 in ev])          # ms
    t_total = float(times.sum())

    # ---- e2e through the host-buffer entry point, pinned host memory, copies inside the timed region
    for _ in range(2):
        e2e_call()
    barrier()
    t0 = time.perf_counter()
    n_e2e = max(3, min(args.steps, 10))
    for _ in range(n_e2e):
        e2e_call()
    torch.cuda.synchronize()
    t_e2e = (time.perf_counter() - t0) / n_e2e
    clocks = sampler.finish()

    # ---- roofline of the dominant kernel (the 1D sweep): single-job launches over a rotating set of buffers larger
    # than L2, so that every launch streams its source from HBM
    peak, peak_src = peaks()
    nbuf = 8
    with torch.cuda.stream(stream):
        bufs = [torch.rand(ne, a ** dim, dtype=torch.float64, device="cuda") for _ in range(nbuf)]
        dsts = [torch.empty(ne, a ** (dim - 1) * b, dtype=torch.float64, device="cuda") for _ in range(nbuf)]
        for i in range(nbuf):
            ctx.sweep1d(op_pt, A.REL_VOL, A.LU_FULL, i % dim, [a] * dim, bufs[i], dsts[i])
    stream.synchronize()
    # the launches are replayed from a CUDA graph (as in the step), so the figure is device time per launch, not host enqueue rate
    nrep, n_replay = 4 * nbuf, 8
    l_roof0 = ctx.launch_count
    g_roof = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g_roof, stream=stream):
        for i in range(nrep):
            ctx.sweep1d(op_pt, A.REL_VOL, A.LU_FULL, i % dim, [a] * dim, bufs[i % nbuf], dsts[i % nbuf])
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        g_roof.replay()
        flush.fill_(0.0)
        e0.record(stream)
        for _ in range(n_replay):
            g_roof.replay()
        e1.record(stream)
    stream.synchronize()
    t_launch = e0.elapsed_time(e1) / (nrep * n_replay) * 1e-3
    bytes_launch = 8.0 * ne * (a ** dim + a ** (dim - 1) * b)           # B_sweep = 8 N_e (S_from + S_to), SURVEY.md 8(d): one dimension goes from edge a to edge b
    achieved = bytes_launch / t_launch / 1e9

    # per-rank step time -> max over ranks
    t_step_ms = t_total / args.steps
    if world > 1:
        tt = torch.tensor([t_step_ms, t_e2e], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_step_ms, t_e2e = float(tt[0]), float(tt[1])
    value = dof * world / (t_step_ms * 1e-3)

    if rank == 0:
        line = {
            "metric": "sparse-grid DoF-stage updates/sec (FP64)", "value": value, "unit": "DoF-stage/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": t_step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": w["desc"], "n_elem": int(ne), "dof_per_gpu": int(dof), "components": ncomp, "schedule": "shared-prefix" if args.schedule else "literal",
                       "kernel": args.kernel, "cuda_graph": graph is not None, "l2": "flushed (256 MiB write) between timed steps; per-step CUDA event pairs on the launch stream",
                       "multi_gpu": "independent replicas per rank (no exchange step in this workload)" if world > 1 else "single GPU"},
            "clocks": clocks,
            "e2e": {"value": dof * world / t_e2e, "unit": "DoF-stage/s", "h2d_bytes_per_step": int(hin.nbytes), "d2h_bytes_per_step": int(hout.nbytes),
                    "ms_per_step": t_e2e * 1e3, "api": e2e_api},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": TRAFFIC_NCU if (args.kernel in (0, 5) and args.workload == "cfg2") else None,
                         "kernel": ("sweep_tc_kernel<%d,%d> (one 1D sweep along each dimension in turn, single job; FP64 DMMA m8n8k4)" % (a, b) if args.kernel in (0, 5) else
                                    "sweep_mma_kernel<%d,%d> (one 1D sweep, single job; FP64 DMMA m8n8k4)" % (a, b) if args.kernel == 4 else "sweep kernel variant %d (one 1D sweep, single job)" % args.kernel), "bytes_per_launch": bytes_launch, "us_per_launch": t_launch * 1e6,
                         "peak_source": peak_src,
                         "step": {"b_alg_bytes": b_alg, "gbs": b_alg / (t_step_ms * 1e-3) / 1e9, "frac": b_alg / (t_step_ms * 1e-3) / 1e9 / peak,
                                  "note": step_note}},
        }
        if world == 1 and not args.no_cpu:
            cb = run_reference(args, w, as_baseline=True)
            if cb:
                line["cpu_baseline"] = {kk: cb[kk] for kk in ("value", "unit", "cores", "kind", "sample")}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    ctx.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
