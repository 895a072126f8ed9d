"""Device time of single 1D sweeps (graph replay over rotating buffers larger than L2), per dimension and L/U/full part.
    python tools/sweep_time.py --workload cfg2|cfg5 --kernel 5|6 [--kf 4 --kt 4] [--lus 2]
Prints one JSON line per (kf, kt, t, lu): us per launch, GB/s on B_sweep = 8 N_e (S_from + S_to), fraction of the measured HBM peak."""
import argparse, importlib, json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--kernel", type=int, default=6)
    ap.add_argument("--lus", default="0,1,2")
    ap.add_argument("--dims", default="")
    ap.add_argument("--shapes", default="")     # e.g. "a>b,b>a,b>b,a>a"
    ap.add_argument("--acc", type=int, default=0)
    ap.add_argument("--tag", default="")
    args = ap.parse_args()
    w = bench.WORKLOADS[args.workload]
    A = importlib.import_module("adaptive-multiresolution-dg_b200")
    dim, k, m, nmax = w["dim"], w["k"], w["m"], w["nmax"]
    a, b = k + 1, m + 1
    lev, sup = A.sparse_grid(dim, nmax)
    ne = lev.shape[0]
    stream = torch.cuda.Stream()
    ctx = A.Context(dim, nmax, k, m, device=0)
    ctx.set_stream(stream.cuda_stream)
    ctx.set_kernel(args.kernel)
    ctx.grid_set(lev, sup)
    tb = bench.load_tables(A, w)
    ops = {"a>b": (ctx.op_register_compact(tb["pt"]), a, b), "b>a": (ctx.op_register_compact(tb["lagr.u_v"]), b, a),
           "b>b": (ctx.op_register_compact(tb["hier"], hier=True), b, b), "a>a": (ctx.op_register_compact(tb["alpt.ujp_vjp"]), a, a)}
    peak, _ = bench.peaks()
    shapes = args.shapes.split(",") if args.shapes else (["a>b"] if args.workload == "cfg2" else ["a>b", "b>a", "b>b", "a>a"])
    dims = [int(x) for x in args.dims.split(",")] if args.dims else list(range(dim))
    for sh in shapes:
        op, kf, kt = ops[sh]
        for t in dims:
            # block shape as in the chain of a tensor application: dims before t already have the target edge
            sizes = [kt if q < t else kf for q in range(dim)]
            s_from = int(np.prod(sizes)); s_to = s_from // kf * kt
            nbuf = max(2, int(np.ceil(400e6 / (8.0 * ne * (s_from + s_to)))))
            with torch.cuda.stream(stream):
                srcs = [torch.rand(ne, s_from, dtype=torch.float64, device="cuda") for _ in range(nbuf)]
                dsts = [torch.zeros(ne, s_to, dtype=torch.float64, device="cuda") for _ in range(nbuf)]
            for lu in [int(x) for x in args.lus.split(",")]:
                with torch.cuda.stream(stream):
                    for i in range(nbuf):
                        ctx.sweep1d(op, A.REL_VOL, lu, t, sizes, srcs[i], dsts[i], accumulate=bool(args.acc))
                stream.synchronize()
                g = torch.cuda.CUDAGraph()
                nrep = 2 * nbuf
                with torch.cuda.graph(g, stream=stream):
                    for i in range(nrep):
                        ctx.sweep1d(op, A.REL_VOL, lu, t, sizes, srcs[i % nbuf], dsts[i % nbuf], accumulate=bool(args.acc))
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                with torch.cuda.stream(stream):
                    g.replay()
                    e0.record(stream)
                    for _ in range(5):
                        g.replay()
                    e1.record(stream)
                stream.synchronize()
                us = e0.elapsed_time(e1) / (5 * nrep) * 1e3
                byts = 8.0 * ne * (s_from + s_to) + (8.0 * ne * s_to if args.acc else 0)
                print(json.dumps({"tag": args.tag, "workload": args.workload, "kernel": args.kernel, "shape": "%d>%d" % (kf, kt), "t": t, "lu": "LUF"[lu], "acc": args.acc,
                                  "us": round(us, 2), "gbs": round(byts / us / 1e3, 1), "frac": round(byts / us / 1e3 / peak, 3)}), flush=True)
            del srcs, dsts
    ctx.close()


if __name__ == "__main__":
    main()
