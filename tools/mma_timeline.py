"""Per-CTA phase timing of the tensor-core sweep kernel by fibre length."""
import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
A = importlib.import_module("adaptive-multiresolution-dg_b200")
dim, nmax, k, m = 4, 8, 3, 3
lev, sup = A.sparse_grid(dim, nmax)
ctx = A.Context(dim, nmax, k, m, device=0)
ctx.set_kernel(4)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
ctx.grid_set(lev, sup)
src_, tgt_, vol_ = ctx.pairs()
op = ctx.op_register_compact(np.random.default_rng(0).standard_normal((len(src_), 4, 4)))
ne = lev.shape[0]
us = [torch.rand(ne, 256, dtype=torch.float64, device="cuda") for _ in range(8)]
vs = [torch.zeros(ne, 256, dtype=torch.float64, device="cuda") for _ in range(8)]
for nm in sys.argv[1:] or ["U", "full"]:
    lu = {"full": A.LU_FULL, "L": A.LU_L, "U": A.LU_U}[nm]
    for i in range(4):
        ctx.sweep1d(op, A.REL_VOL, lu, 0, [4] * dim, us[i], vs[i])
    dbg = torch.zeros(4096 * 8, dtype=torch.int64, device="cuda")
    ctx.set_debug_buffer(dbg)
    ctx.sweep1d(op, A.REL_VOL, lu, 0, [4] * dim, us[5], vs[5])
    torch.cuda.synchronize()
    ctx.set_debug_buffer(None)
    d = dbg.cpu().numpy().reshape(-1, 8)
    d = d[d[:, 0] > 0]
    print(nm, "CTAs", len(d))
    for mm in sorted(set(d[:, 6].tolist())):
        x = d[d[:, 6] == mm]
        f = lambda a, b: (x[:, b] - x[:, a]).mean()
        print("  m=%4d items %4d (nfib*1000+cols %s): item %6.0f issue %6.0f wait %6.0f compute %7.0f total %7.0f (max %7.0f) cycles" % (
            mm, len(x), sorted(set(x[:, 7].tolist()))[:3], f(0, 1), f(1, 2), f(2, 3), f(3, 4), f(0, 4), (x[:, 4] - x[:, 0]).max()))
