"""Tuning sweep of the tensor-core sweep kernel's work-list knobs: python tools/tune_mma.py DIM NMAX K M "cap:items[:ent],cap:items,..."
(AMDG_LIB selects a library built with another AMDG_MMA_THREADS)."""
import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
A = importlib.import_module("adaptive-multiresolution-dg_b200")
dim, nmax, k, m = [int(x) for x in sys.argv[1:5]]
a, b = k + 1, m + 1
lev, sup = A.sparse_grid(dim, nmax)
ne = lev.shape[0]
st = torch.cuda.Stream()
nbuf = max(2, int(400e6 // (ne * (a ** dim + b ** dim) * 8)) + 1)
with torch.cuda.stream(st):
    us = [torch.rand(1, ne, a ** dim, dtype=torch.float64, device="cuda") for _ in range(nbuf)]
    vs = [torch.zeros(1, ne, a ** (dim - 1) * b, dtype=torch.float64, device="cuda") for _ in range(nbuf)]
byt = 8.0 * ne * (a ** dim + a ** (dim - 1) * b)
rng = np.random.default_rng(0)
for cfg in sys.argv[5].split(","):
    f = cfg.split(":")
    os.environ["AMDG_MMA_CAP"], os.environ["AMDG_MMA_ITEMS"] = f[0], f[1]
    os.environ["AMDG_TC_CAP"], os.environ["AMDG_TC_ITEMS"] = f[0], f[1]
    if len(f) > 2: os.environ["AMDG_MMA_ENT"] = os.environ["AMDG_TC_ENT"] = f[2]
    if len(f) > 3: os.environ["AMDG_MMA_STAGE_A"] = os.environ["AMDG_TC_STAGE_A"] = f[3]
    ctx = A.Context(dim, nmax, k, m, device=0)
    ctx.set_stream(st.cuda_stream)
    ctx.grid_set(lev, sup)
    src_, tgt_, vol_ = ctx.pairs()
    op = ctx.op_register_compact(rng.standard_normal((len(src_), a, b)))
    res = []
    for lu, nm in ((A.LU_FULL, "full"), (A.LU_L, "L"), (A.LU_U, "U")):
        for t in sorted(set([0, dim - 2, dim - 1])):
            with torch.cuda.stream(st):
                for i in range(3):
                    ctx.sweep1d(op, A.REL_VOL, lu, t, [a] * dim, us[i % nbuf], vs[i % nbuf])
            st.synchronize()
            nrep = 5 * nbuf
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=st):
                for i in range(nrep):
                    ctx.sweep1d(op, A.REL_VOL, lu, t, [a] * dim, us[i % nbuf], vs[i % nbuf])
            with torch.cuda.stream(st):
                g.replay()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(st)
                for _ in range(4):
                    g.replay()
                e1.record(st)
            st.synchronize()
            res.append(e0.elapsed_time(e1) / (4 * nrep) * 1e3)
            del g
    print("%s cfg %-16s mean %6.1f us (%6.0f GB/s)  " % (os.path.basename(A.LIB_PATH), cfg, np.mean(res), byt / np.mean(res) / 1e3) + " ".join("%.1f" % r for r in res), flush=True)
    ctx.close()
