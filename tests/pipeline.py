"""Test helper: drive the CUDA path (through the C ABI) over a golden dump of the reference, phase by phase.
torch only carries device buffers."""
import importlib

import numpy as np
import torch


def rel(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


class DevCase:
    def __init__(self, d, schedule=None, kernel=None, perm=None, generated=False, msh_case=1):
        amdg = self.amdg = importlib.import_module("adaptive-multiresolution-dg_b200")
        self.d = d
        (self.dim, self.nmax, self.n0, self.sparse, self.pa, self.pl, self.ph, self.vecnum, self.herm, self.ne) = [int(x) for x in d["config"]]
        self.a = self.pa + 1
        self.b = (self.ph if self.herm else self.pl) + 1
        self.perm = np.arange(self.ne) if perm is None else perm
        self.ctx = amdg.Context(self.dim, self.nmax, self.pa, self.ph if self.herm else self.pl, device=0)
        self.ctx.set_stream(torch.cuda.current_stream().cuda_stream)    # share torch's stream: no cross-stream races
        if schedule is not None:
            self.ctx.set_schedule(schedule)
        if kernel is not None:
            self.ctx.set_kernel(kernel)
        self.ctx.grid_set(d["level"][self.perm], d["suppt"][self.perm])
        pre = "herm" if self.herm else "lagr"
        c = self.ctx
        if generated:
            # every table from the library's own generator (amdg_op_generate*): nothing but the grid and the coefficients comes from the dump
            basis, P = (amdg.BASIS_HERMITE, self.ph) if self.herm else (amdg.BASIS_LAGRANGE, self.pl)
            self.op_pt = c.op_generate_points(basis, P, msh_case)
            self.op_uv, self.op_uvx = c.op_generate(basis, P, "u_v", msh_case), c.op_generate(basis, P, "u_vx", msh_case)
            self.op_ul, self.op_ur = c.op_generate(basis, P, "ulft_vjp", msh_case), c.op_generate(basis, P, "urgt_vjp", msh_case)
            self.op_uave = c.op_combine(self.op_ul, 1.0, self.op_ur, 1.0)
            self.op_hier = c.op_generate_hier(basis, P, msh_case)
            self.alpt = {k: c.op_generate(amdg.BASIS_ALPERT, self.pa, k) for k in ("u_v", "u_vx", "ulft_vjp", "urgt_vjp", "ujp_vjp", "ux_vx", "uxave_vjp", "ujp_vxave")}
            if not self.herm:
                self.op_pt_d1 = c.op_generate_points(basis, P, msh_case, 1)
            return
        pt = (d["Her_pt_Alpt_1D"] if self.herm else d["Lag_pt_Alpt_1D"]).T.copy()     # FastLagrIntp ctor transpose
        self.op_pt = c.op_register(pt, self.a, self.b)
        self.op_uv = c.op_register(d[pre + ".u_v"], self.b, self.a)
        self.op_uvx = c.op_register(d[pre + ".u_vx"], self.b, self.a)
        self.op_ul = c.op_register(d[pre + ".ulft_vjp"], self.b, self.a)
        self.op_ur = c.op_register(d[pre + ".urgt_vjp"], self.b, self.a)
        self.op_uave = c.op_combine(self.op_ul, 1.0, self.op_ur, 1.0)
        self.op_hier = c.op_register_hier(d[pre + ".pw_anc"], d[pre + ".pw_wt"])
        self.alpt = {k: c.op_register(d["alpt." + k], self.a, self.a) for k in ("u_v", "u_vx", "ulft_vjp", "urgt_vjp", "ujp_vjp", "ux_vx", "uxave_vjp", "ujp_vxave")}
        if not self.herm and "Lag_pt_Alpt_1D_d1" in d:
            self.op_pt_d1 = c.op_register(d["Lag_pt_Alpt_1D_d1"].T.copy(), self.a, self.b)

    def close(self):
        self.ctx.close()

    # element arrays: dump order <-> device order
    def to_dev(self, arr):
        """arr [ne, block] (dump order) -> device tensor in the context's element order"""
        return torch.from_numpy(np.ascontiguousarray(arr[self.perm])).cuda()

    def to_host(self, t):
        out = np.empty_like(t.cpu().numpy())
        out[self.perm] = t.cpu().numpy()
        return out

    def zeros(self, edge):
        return torch.zeros(self.ne, edge ** self.dim, dtype=torch.float64, device="cuda")

    def eval_up(self, u_dev, op=None, per_dim_ops=None):
        up = self.zeros(self.b)
        ops = per_dim_ops if per_dim_ops is not None else [self.op_pt if op is None else op] * self.dim
        self.ctx.apply_tensor(ops, [self.amdg.REL_VOL] * self.dim, u_dev, up)
        return up

    def hier(self, v_dev):
        out = torch.empty_like(v_dev)
        self.ctx.hierarchize(self.op_hier, v_dev, out)
        return out

    def to_alpt(self, c_dev):
        out = self.zeros(self.a)
        self.ctx.apply_tensor([self.op_uv] * self.dim, [self.amdg.REL_VOL] * self.dim, c_dev, out)
        return out

    def rhs_vol_flx(self, fucoe_dev_list, rhs):
        """HyperbolicLagrRHS::rhs_vol_scalar + rhs_flx_intp_scalar: rhs += ..."""
        A = self.amdg
        for t in range(self.dim):
            ops = [self.op_uvx if s == t else self.op_uv for s in range(self.dim)]
            self.ctx.apply_tensor(ops, [A.REL_VOL] * self.dim, fucoe_dev_list[t], rhs, accumulate=True)
        vol = rhs.clone()
        for t in range(self.dim):
            ops = [self.op_uave if s == t else self.op_uv for s in range(self.dim)]
            rels = [A.REL_FLX if s == t else A.REL_VOL for s in range(self.dim)]
            self.ctx.apply_tensor(ops, rels, fucoe_dev_list[t], rhs, coef=0.5, accumulate=True)
        return vol

    def penalty(self, u_dev, rhs, alpha):
        """HyperbolicAlptRHS::rhs_flx_penalty_scalar"""
        A = self.amdg
        if self.dim == 1:
            return    # reference quirk, see tests/test_oracle.py
        for t in range(self.dim):
            self.ctx.sweep1d(self.alpt["ujp_vjp"], A.REL_FLX, A.LU_FULL, t, [self.a] * self.dim, u_dev, rhs, coef=-alpha / 2.0, accumulate=True)
