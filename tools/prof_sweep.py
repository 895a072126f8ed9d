"""Profiling driver: a few single sweeps and one round trip at the cfg2 size (run under ncu)."""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
A = importlib.import_module("adaptive-multiresolution-dg_b200")

dim, k, m, nmax = 4, 3, 3, int(os.environ.get("NMAX", "8"))
mode = sys.argv[1] if len(sys.argv) > 1 else "sweeps"
lev, sup = A.sparse_grid(dim, nmax)
ctx = A.Context(dim, nmax, k, m, device=0)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
ctx.grid_set(lev, sup)
tb = A.generate_tables(8, 3, 3)
assert nmax == 8
op_pt = ctx.op_register_compact(tb["pt"])
op_uv = ctx.op_register_compact(tb["lagr.u_v"])
op_hier = ctx.op_register_compact(tb["hier"], hier=True)
ne = lev.shape[0]
u = torch.rand(ne, 256, dtype=torch.float64, device="cuda")
v = torch.zeros_like(u)
w = torch.zeros_like(u)
if mode == "sweeps":
    for rep in range(2):
        for t in range(dim):
            for lu in (A.LU_FULL, A.LU_L, A.LU_U):
                ctx.sweep1d(op_pt, A.REL_VOL, lu, t, [4] * dim, u, v)
else:
    for rep in range(2):
        ctx.apply_tensor([op_pt] * dim, [0] * dim, u, v)
        ctx.hierarchize(op_hier, v, v)
        ctx.apply_tensor([op_uv] * dim, [0] * dim, v, w)
torch.cuda.synchronize()
print("done", ctx.launch_count)
