"""Per-CTA time split of the pipelined sweep kernel (clock64 sums written by the kernel)."""
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
for nm in sys.argv[1:] or ["U", "L", "full"]:
    lu = {"full": A.LU_FULL, "L": A.LU_L, "U": A.LU_U}[nm]
    for i in range(4):
        ctx.sweep1d(op, A.REL_VOL, lu, 0, [4] * dim, us[i], vs[i])
    dbg = torch.zeros(4096 * 8, dtype=torch.int64, device="cuda")
    ctx.set_debug_buffer(dbg)
    ctx.sweep1d(op, A.REL_VOL, lu, 0, [4] * dim, us[5], vs[5])
    torch.cuda.synchronize()
    ctx.set_debug_buffer(None)
    d = dbg.cpu().numpy().reshape(-1, 8)
    d = d[d[:, 5] == 1]
    tot = d[:, 0] + d[:, 1] + d[:, 2] + d[:, 3]
    print("%s: CTAs %d, items/CTA mean %.1f; cycles per CTA: wait %.0f issue %.0f compute %.0f final %.0f total mean %.0f max %.0f" % (
        nm, len(d), d[:, 4].mean(), d[:, 0].mean(), d[:, 1].mean(), d[:, 2].mean(), d[:, 3].mean(), tot.mean(), tot.max()))
    print("    per item: wait %.0f issue %.0f compute %.0f" % ((d[:, 0] / d[:, 4]).mean(), (d[:, 1] / d[:, 4]).mean(), (d[:, 2] / d[:, 4]).mean()))
