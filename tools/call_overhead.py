"""Host cost of one call through the C ABI on a tiny grid (launch-bound regime of the adaptive runs): wall time per call over a long loop, no
synchronisation inside.    python tools/call_overhead.py"""
import importlib, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
A = importlib.import_module("adaptive-multiresolution-dg_b200")
dim, nmax, k, m = 2, 5, 2, 3
a, b = k + 1, m + 1
lev, sup = A.sparse_grid(dim, nmax)
ne = lev.shape[0]
for kernel in (1, 4, 0):
    ctx = A.Context(dim, nmax, k, m, device=0)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    ctx.set_kernel(kernel)
    ctx.grid_set(lev, sup)
    op = ctx.op_generate_points(A.BASIS_LAGRANGE, m)
    opa = ctx.op_generate(A.BASIS_ALPERT, k, "ujp_vjp")
    u = torch.rand(ne, a ** dim, dtype=torch.float64, device="cuda")
    v = torch.zeros(ne, b ** dim, dtype=torch.float64, device="cuda")
    w = torch.zeros(ne, a ** dim, dtype=torch.float64, device="cuda")
    res = {}
    def timed(name, fn, n=2000):
        for _ in range(20): fn()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(n): fn()
        t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
        res[name] = ((t1 - t0) / n * 1e6, (t2 - t0) / n * 1e6)
    timed("sweep1d a->a", lambda: ctx.sweep1d(opa, A.REL_FLX, A.LU_FULL, 0, [a] * dim, u, w))
    timed("apply_tensor a->b (3 launches)", lambda: ctx.apply_tensor([op] * dim, [A.REL_VOL] * dim, u, v), 1000)
    timed("apply_tensor accumulate", lambda: ctx.apply_tensor([op] * dim, [A.REL_VOL] * dim, u, v, accumulate=True), 1000)
    timed("rk_stage", lambda: ctx.rk_stage(A.RK_RK3SSP, 1, 1e-3, u, w, w))
    timed("axpby", lambda: A._check(A.lib.amdg_axpby(ctx._h, u.numel(), 1.0, A._ptr(u), 1.0, A._ptr(w))))
    print("kernel %d (%d elements): " % (kernel, ne) + "; ".join("%s: %.1f us issue / %.1f us incl. drain" % (n, x[0], x[1]) for n, x in res.items()))
    ctx.close()
