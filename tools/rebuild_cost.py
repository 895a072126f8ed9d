"""Cost of a grid change (what DGAdapt::refine/coarsen triggers): amdg_grid_set (host tables, O(N log N), + upload) and the first
application afterwards (work lists and operator fragments are built lazily), against a steady-state application.
    python tools/rebuild_cost.py DIM NMAX K M"""
import importlib, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
A = importlib.import_module("adaptive-multiresolution-dg_b200")
dim, nmax, k, m = [int(x) for x in sys.argv[1:5]]
a, b = k + 1, m + 1
lev, sup = A.sparse_grid(dim, nmax)
ne = lev.shape[0]
ctx = A.Context(dim, nmax, k, m, device=0)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
ctx.grid_set(lev, sup)
src_, tgt_, vol_ = ctx.pairs()
op = ctx.op_register_compact(np.random.default_rng(0).standard_normal((len(src_), a, b)))
u = torch.rand(ne, a ** dim, dtype=torch.float64, device="cuda")
v = torch.zeros(ne, b ** dim, dtype=torch.float64, device="cuda")
res = {}
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ctx.grid_set(lev, sup)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    ctx.apply_tensor([op] * dim, [0] * dim, u, v)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    for _ in range(5):
        ctx.apply_tensor([op] * dim, [0] * dim, u, v)
    torch.cuda.synchronize(); t3 = time.perf_counter()
    res = dict(grid_set_ms=(t1 - t0) * 1e3, first_apply_ms=(t2 - t1) * 1e3, steady_apply_ms=(t3 - t2) / 5 * 1e3)
print("d=%d NMAX=%d k=%d m=%d: %d elements | amdg_grid_set %.2f ms | first tensor application after it %.2f ms | steady state %.3f ms (eager)"
      % (dim, nmax, k, m, ne, res["grid_set_ms"], res["first_apply_ms"], res["steady_apply_ms"]))
