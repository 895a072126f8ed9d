"""Per-CTA phase timing of the tensor-core sweep kernel by fibre length."""
import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
A = importlib.import_module("adaptive-multiresolution-dg_b200")
dim, nmax, k, m = 4, 8, 3, 3
lev, sup = A.sparse_grid(dim, nmax)
ctx = A.Context(dim, nmax, k, m, device=0)
ctx.set_kernel(int(os.environ.get("KERNEL", "4")))
SLOTS = 444 if os.environ.get("KERNEL", "4") == "4" else 148 * 6
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
ctx.grid_set(lev, sup)
src_, tgt_, vol_ = ctx.pairs()
op = ctx.op_register_compact(np.random.default_rng(0).standard_normal((len(src_), 4, 4)))
ne = lev.shape[0]
us = [torch.rand(ne, 256, dtype=torch.float64, device="cuda") for _ in range(8)]
vs = [torch.zeros(ne, 256, dtype=torch.float64, device="cuda") for _ in range(8)]
T = int(os.environ.get("T", "0"))
for nm in sys.argv[1:] or ["U", "full"]:
    lu = {"full": A.LU_FULL, "L": A.LU_L, "U": A.LU_U}[nm]
    for i in range(4):
        ctx.sweep1d(op, A.REL_VOL, lu, T, [4] * dim, us[i], vs[i])
    dbg = torch.zeros(4096 * 8, dtype=torch.int64, device="cuda")
    ctx.set_debug_buffer(dbg)
    ctx.sweep1d(op, A.REL_VOL, lu, T, [4] * dim, us[5], vs[5])
    torch.cuda.synchronize()
    ctx.set_debug_buffer(None)
    d = dbg.cpu().numpy().reshape(-1, 8)
    d = d[d[:, 0] > 0]
    t0 = d[:, 5].min()
    start = (d[:, 5] - t0) / 1e3
    end = start + (d[:, 4] - d[:, 0]) / 1965.0
    print(nm, "t=%d" % T, "CTAs", len(d), "kernel span %.1f us; CTA lifetime sum %.0f us (/%d slots = %.1f us)" % (end.max(), (end - start).sum(), SLOTS, (end - start).sum() / SLOTS))
    for q in (0.5, 0.9, 0.99, 1.0):
        print("   %3.0f%% of CTAs finished by %.1f us" % (q * 100, np.quantile(end, q)))
    o = np.argsort(-end)[:6]
    for i in o:
        print("   late CTA: m=%d code=%d start %.1f end %.1f us (item %d issue %d wait %d compute %d cycles)" % (d[i, 6], d[i, 7], start[i], end[i], d[i,1]-d[i,0], d[i,2]-d[i,1], d[i,3]-d[i,2], d[i,4]-d[i,3]))
    for mm in sorted(set(d[:, 6].tolist())):
        x = d[d[:, 6] == mm]
        f = lambda a, b: (x[:, b] - x[:, a]).mean()
        print("  m=%4d items %4d (nfib*1e6+n_ent*100+cols %s): item %6.0f issue %6.0f wait %6.0f compute %7.0f total %7.0f (max %7.0f) cycles" % (
            mm, len(x), sorted(set(x[:, 7].tolist()))[:3], f(0, 1), f(1, 2), f(2, 3), f(3, 4), f(0, 4), (x[:, 4] - x[:, 0]).max()))
