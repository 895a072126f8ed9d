"""Per-CTA phase timing of one sweep (clock64 stamps written by the kernel)."""
import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
A = importlib.import_module("adaptive-multiresolution-dg_b200")
dim, nmax, k, m = 4, 8, 3, 3
lev, sup = A.sparse_grid(dim, nmax)
ctx = A.Context(dim, nmax, k, m, device=0)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
ctx.grid_set(lev, sup)
src_, tgt_, vol_ = ctx.pairs()
op = ctx.op_register_compact(np.random.default_rng(0).standard_normal((len(src_), 4, 4)))
ne = lev.shape[0]
us = [torch.rand(ne, 256, dtype=torch.float64, device="cuda") for _ in range(8)]
vs = [torch.zeros(ne, 256, dtype=torch.float64, device="cuda") for _ in range(8)]
lu = {"full": A.LU_FULL, "L": A.LU_L, "U": A.LU_U}[sys.argv[1] if len(sys.argv) > 1 else "U"]
for i in range(4):
    ctx.sweep1d(op, A.REL_VOL, lu, 0, [4] * dim, us[i], vs[i])
dbg = torch.zeros(4096 * 8, dtype=torch.int64, device="cuda")
ctx.set_debug_buffer(dbg)
ctx.sweep1d(op, A.REL_VOL, lu, 0, [4] * dim, us[5], vs[5])
torch.cuda.synchronize()
d = dbg.cpu().numpy().reshape(-1, 8)
d = d[d[:, 0] > 0]
t0 = d[:, 0].min()
pk = d[d[:, 6] > 0]
print("CTAs", len(d), "packed", len(pk), "kernel span cycles", (d[:, [0, 1, 2, 3, 4]].max() - t0))
for nm, a, b in (("item fetch", 0, 1), ("issue loads", 1, 2), ("wait+sync", 2, 3), ("compute+store", 3, 4), ("total", 0, 4)):
    x = pk[:, b] - pk[:, a]
    print("  packed %-14s mean %7.0f  p50 %7.0f  p90 %7.0f  max %7.0f cycles" % (nm, x.mean(), np.median(x), np.percentile(x, 90), x.max()))
print("  packed rows staged: mean %.1f" % pk[:, 7].mean())
# start time distribution: when do CTAs start
st = np.sort(d[:, 0] - t0)
print("  CTA start times (cycles): p10 %d p50 %d p90 %d max %d" % (st[len(st)//10], st[len(st)//2], st[len(st)*9//10], st[-1]))
sm = d[:, 5]
print("  CTAs per SM: min %d max %d" % (np.bincount(sm.astype(int)).min(), np.bincount(sm.astype(int)).max()))
