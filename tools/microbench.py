"""Micro-benchmark of single sweeps: python tools/microbench.py DIM NMAX K M NCOMP  (random operator blocks)"""
import importlib, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
A = importlib.import_module("adaptive-multiresolution-dg_b200")
dim, nmax, k, m, ncomp = [int(x) for x in sys.argv[1:6]]
a, b = k + 1, m + 1
lev, sup = A.sparse_grid(dim, nmax)
ctx = A.Context(dim, nmax, k, m, device=0)
st = torch.cuda.Stream()
ctx.set_stream(st.cuda_stream)
ctx.grid_set(lev, sup)
src_, tgt_, vol_ = ctx.pairs()
rng = np.random.default_rng(0)
op = ctx.op_register_compact(rng.standard_normal((len(src_), a, b)))
ne = lev.shape[0]
nbuf = max(2, int(400e6 // (ne * ncomp * (a ** dim + b ** dim) * 8)) + 1)
with torch.cuda.stream(st):
    us = [torch.rand(ncomp, ne, a ** dim, dtype=torch.float64, device="cuda") for _ in range(nbuf)]
    vs = [torch.zeros(ncomp, ne, a ** (dim - 1) * b, dtype=torch.float64, device="cuda") for _ in range(nbuf)]
st.synchronize()
byt = 8.0 * ne * ncomp * (a ** dim + a ** (dim - 1) * b)
print("n_elem %d ncomp %d bytes/sweep %.1f MB nbuf %d" % (ne, ncomp, byt / 1e6, nbuf))
for lu, nm in ((A.LU_FULL, "full"), (A.LU_L, "L"), (A.LU_U, "U")):
    for t in sorted(set([0, dim // 2, dim - 1])):
        with torch.cuda.stream(st):
            for i in range(3):
                ctx.sweep1d(op, A.REL_VOL, lu, t, [a] * dim, us[i % nbuf], vs[i % nbuf], n_comp=ncomp)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            nrep = 20
            for i in range(nrep):
                ctx.sweep1d(op, A.REL_VOL, lu, t, [a] * dim, us[i % nbuf], vs[i % nbuf], n_comp=ncomp)
            e1.record(st)
        st.synchronize()
        us_ = e0.elapsed_time(e1) / nrep * 1e3
        print("  %-4s t=%d  %8.1f us  %7.1f GB/s" % (nm, t, us_, byt / us_ / 1e3))
