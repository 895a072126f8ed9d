"""TEST FIXTURE generator: the reference's 1D tables at the benchmark sizes, for tests/test_tables.py.

The dense OperatorMatrix1D / Lag_pt_Alpt_1D tables are produced by the compiled reference (oracle/_ref/ref_harness --dump-tables) and
stored in the compact per-pair block form of include/amdg.h (amdg_pairs order).  Rounds 1-2 shipped these bundles as inputs of bench.py;
the library now generates the tables itself (csrc/tables.hpp, amdg_op_generate*) and the bundles only pin that generator.  Run in the
build container:

    python tests/golden/tables/make_tables.py K M N [msh_case]
"""
import importlib
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    k, m, n = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    msh = int(sys.argv[4]) if len(sys.argv) > 4 else 1
    import refdump
    amdg = importlib.import_module("adaptive-multiresolution-dg_b200")
    tmp = "/tmp/amdg_tables_%d_%d_%d.dump" % (k, m, n)
    subprocess.run([os.path.join(ROOT, "oracle", "_ref", "ref_harness"), "--dim", "1", "--nmax", str(n), "--pa", str(k), "--pl", str(m),
                    "--msh-lagr", str(msh), "--run", "grid", "--dump-tables", "1", "--out", tmp], check=True)
    d = refdump.load(tmp)
    os.remove(tmp)
    ctx = amdg.Context(1, n, k, m, device=-1)
    a, b = k + 1, m + 1
    out = {"meta": np.array([k, m, n, msh], dtype=np.int32)}
    def compact(dense, kf, kt):
        return ctx.op_blocks(ctx.op_register(dense, kf, kt), kf, kt)
    out["pt"] = compact(d["Lag_pt_Alpt_1D"].T.copy(), a, b)
    out["pt_d1"] = compact(d["Lag_pt_Alpt_1D_d1"].T.copy(), a, b)
    for nm in ("u_v", "u_vx", "ulft_vjp", "urgt_vjp"):
        out["lagr." + nm] = compact(d["lagr." + nm], b, a)
    for nm in ("u_vx", "ulft_vjp", "urgt_vjp", "ujp_vjp", "ux_vx", "uxave_vjp", "ujp_vxave"):
        out["alpt." + nm] = compact(d["alpt." + nm], a, a)
    out["hier"] = ctx.op_blocks(ctx.op_register_hier(d["lagr.pw_anc"], d["lagr.pw_wt"]), b, b)
    out["lagr.pw_anc"] = d["lagr.pw_anc"]
    out["lagr.pw_wt"] = d["lagr.pw_wt"]
    out["lagr.intep_pt"] = d["lagr.intep_pt"]
    path = os.path.join(HERE, "tables_k%d_m%d_n%d.npz" % (k, m, n))
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path))


if __name__ == "__main__":
    main()
