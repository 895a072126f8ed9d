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
    "cfg2": dict(dim=4, k=3, m=3, nmax=8, cpu_nmax=7, ref_nmax=6, desc="example/01_interp_01_high_dim: Lagrange interpolation round trip d=4 k=3 m=3 NMAX=8 (full sparse grid)"),
}


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
    cmd = [exe, "--dim", str(w["dim"]), "--nmax", str(nmax), "--pa", str(w["k"]), "--pl", str(w["m"]), "--run", "roundtrip",
           "--time", str(reps + (args.warmup if not as_baseline else 1)), "--threads", str(nthr)]
    t0 = time.time()
    out = subprocess.run(cmd, check=True, stdout=subprocess.PIPE, text=True).stdout.strip().splitlines()[-1]
    wall = time.time() - t0
    r = json.loads(out)
    t_step = r["intp"] + r["hier"] + r["init"]          # medians over the repetitions
    dof = r["dof"]
    return {"value": dof / t_step, "unit": "DoF-stage/s", "cores": r["threads"], "kind": "reference",
            "sample": "same round trip at NMAX=%d (%d elements, %d DoF): median of %d repetitions, %.3f s per step; reference built from /root/reference/source with -O3 -fopenmp (oracle/Makefile)"
                      % (nmax, r["n_elem"], dof, r["reps"], t_step),
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

    host_in = torch.from_numpy(synthetic_field(lev, a ** dim, ncomp, 20240901 + rank)).pin_memory()
    host_out = torch.empty_like(host_in).pin_memory()
    with torch.cuda.stream(stream):
        u = host_in.to("cuda", non_blocking=True)
        up = torch.zeros(ncomp, ne, b ** dim, dtype=torch.float64, device="cuda")
        out = torch.zeros(ncomp, ne, a ** dim, dtype=torch.float64, device="cuda")
        flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device="cuda")
    stream.synchronize()
    ops_f, ops_i, rels = [op_pt] * dim, [op_uv] * dim, [A.REL_VOL] * dim

    def step():
        ctx.apply_tensor(ops_f, rels, u, up, n_comp=ncomp)          # Alpert -> point values
        ctx.hierarchize(op_hier, up, up, n_comp=ncomp)               # point values -> hierarchical coefficients
        ctx.apply_tensor(ops_i, rels, up, out, n_comp=ncomp)         # -> Alpert

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.cuda.stream(stream):
        for _ in range(max(args.warmup, 3)):
            step()
    barrier()
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
            step()
            ev[s][1].record(stream)
    barrier()
    launches = ctx.launch_count - l0
    times = np.array([e0.elapsed_time(e1) for e0, e1 in ev])          # ms
    t_total = float(times.sum())

    # ---- e2e through the host-buffer entry point, pinned host memory, copies inside the timed region
    hin, hout = host_in.numpy(), host_out.numpy()
    for _ in range(2):
        ctx.host_roundtrip(op_pt, op_hier, op_uv, hin, n_comp=ncomp, out=hout)
    barrier()
    t0 = time.perf_counter()
    n_e2e = max(3, min(args.steps, 10))
    for _ in range(n_e2e):
        ctx.host_roundtrip(op_pt, op_hier, op_uv, hin, n_comp=ncomp, out=hout)
    torch.cuda.synchronize()
    t_e2e = (time.perf_counter() - t0) / n_e2e
    clocks = sampler.finish()

    # ---- roofline of the dominant kernel (the 1D sweep): single-job launches over a rotating set of buffers larger
    # than L2, so that every launch streams its source from HBM
    peak, peak_src = peaks()
    nbuf = 8
    with torch.cuda.stream(stream):
        bufs = [torch.rand(ne, a ** dim, dtype=torch.float64, device="cuda") for _ in range(nbuf)]
        dsts = [torch.empty(ne, b ** dim, dtype=torch.float64, device="cuda") for _ in range(nbuf)]
        for i in range(nbuf):
            ctx.sweep1d(op_pt, A.REL_VOL, A.LU_FULL, i % dim, [a] * dim, bufs[i], dsts[i])
    stream.synchronize()
    nrep = 4 * nbuf
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        flush.fill_(0.0)
        e0.record(stream)
        for i in range(nrep):
            ctx.sweep1d(op_pt, A.REL_VOL, A.LU_FULL, i % dim, [a] * dim, bufs[i % nbuf], dsts[i % nbuf])
        e1.record(stream)
    stream.synchronize()
    t_launch = e0.elapsed_time(e1) / nrep * 1e-3
    bytes_launch = 8.0 * ne * (a ** dim + b ** dim)                     # B_sweep = 8 N_e (S_from + S_to), SURVEY.md 8(d)
    achieved = bytes_launch / t_launch / 1e9

    # per-rank step time -> max over ranks
    t_step_ms = t_total / args.steps
    if world > 1:
        tt = torch.tensor([t_step_ms, t_e2e], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_step_ms, t_e2e = float(tt[0]), float(tt[1])
    value = dof * world / (t_step_ms * 1e-3)

    # algorithmic bytes of the reference's sweep list for the step: 2*2^(d-1)*C(d,a,b) + d*2*b^d doubles per element
    chain = sum((a ** (dim - i) * b ** i + a ** (dim - i - 1) * b ** (i + 1)) for i in range(dim))
    b_alg = 8.0 * ne * ncomp * (2 * 2 ** (dim - 1) * chain + dim * 2 * b ** dim)

    if rank == 0:
        line = {
            "metric": "sparse-grid DoF-stage updates/sec (FP64)", "value": value, "unit": "DoF-stage/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": t_step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": w["desc"], "n_elem": int(ne), "dof_per_gpu": int(dof), "components": ncomp, "schedule": "shared-prefix" if args.schedule else "literal",
                       "kernel": args.kernel, "l2": "flushed (256 MiB write) between timed steps; per-step CUDA event pairs on the launch stream",
                       "multi_gpu": "independent replicas per rank (no exchange step in this workload)" if world > 1 else "single GPU"},
            "clocks": clocks,
            "e2e": {"value": dof * world / t_e2e, "unit": "DoF-stage/s", "h2d_bytes_per_step": int(hin.nbytes), "d2h_bytes_per_step": int(hout.nbytes),
                    "ms_per_step": t_e2e * 1e3, "api": "amdg_host_roundtrip (pinned host buffers)"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                         "kernel": "sweep_gather_kernel<4,4> (one 1D sweep, single job)", "bytes_per_launch": bytes_launch, "us_per_launch": t_launch * 1e6,
                         "peak_source": peak_src,
                         "step": {"b_alg_bytes": b_alg, "gbs": b_alg / (t_step_ms * 1e-3) / 1e9, "frac": b_alg / (t_step_ms * 1e-3) / 1e9 / peak,
                                  "note": "reference sweep list bytes / measured step time; the shared-prefix schedule runs 22 instead of 32 sweeps per transform"}},
        }
        if world == 1:
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
