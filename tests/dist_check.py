"""Multi-GPU parity check (run under torchrun, one rank per GPU): the fibre-partitioned tensor application
(dist.DistTensorApply, NCCL all-to-all layout switches) reproduces the compiled reference's eval_up_Lagr and
eval_ucoe_Alpt_Lagr on the d=6 and d=4 fixtures."""
import importlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import refdump


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    A = importlib.import_module("adaptive-multiresolution-dg_b200")
    D = importlib.import_module("adaptive-multiresolution-dg_b200.dist")
    worst = 0.0
    for name in ("cfg5_vlasov_d6_k1_n2", "cfg2_rt_d4_k3_n3"):
        d = refdump.load(os.path.join(ROOT, "tests", "golden", name + ".dump.xz"))
        dim, nmax, n0, sparse, pa, pl = [int(x) for x in d["config"][:6]]
        a, b = pa + 1, pl + 1
        part = D.FibrePartition(d["level"], d["suppt"], world, rank)

        def register(c):
            return {"pt": c.op_register(d["Lag_pt_Alpt_1D"].T.copy(), a, b), "u_v": c.op_register(d["lagr.u_v"], b, a)}
        T = D.DistTensorApply(A, part, dim, nmax, pa, pl, local, register)
        u = torch.from_numpy(np.ascontiguousarray(d["ucoe_alpt.in"][:, 0, :][part.local["X"]])).cuda()
        up = T.apply(["pt"] * dim, [A.REL_VOL] * dim, u, a, b)
        key = "up_intp" if "up_intp" in d else "rt.up_intp"
        ref = d[key][:, 0, :][part.local["X"]]
        e = float(np.linalg.norm(up.cpu().numpy() - ref) / np.linalg.norm(ref))
        worst = max(worst, e)
        if "rt.ucoe_intp" in d:
            c_in = torch.from_numpy(np.ascontiguousarray(d["rt.ucoe_intp"][:, 0, :][part.local["X"]])).cuda()
            ua = T.apply(["u_v"] * dim, [A.REL_VOL] * dim, c_in, b, a)
            ref = d["rt.ucoe_alpt"][:, 0, :][part.local["X"]]
            worst = max(worst, float(np.linalg.norm(ua.cpu().numpy() - ref) / np.linalg.norm(ref)))
        if rank == 0:
            print("%s: rel-L2 %.3e, switches %d (%.2f MB sent by rank 0)" % (name, e, T.switches, T.switch_bytes / 1e6))
        T.close()
    # the whole nonlinear stage program at this world size against the reference's dump (bench.py runs the same check before timing)
    import bench
    S = importlib.import_module("adaptive-multiresolution-dg_b200.stage")
    stream = torch.cuda.Stream()
    err, berr = bench.parity_check(A, S, D, world, rank, local, stream, 0)
    worst = max(worst, err, float(berr))
    if rank == 0:
        print("stage program (cfg5 fixture, %d ranks): rel-L2 %.3e, barrier time-outs %d" % (world, err, berr))
    w = torch.tensor([worst], device="cuda", dtype=torch.float64)
    dist.all_reduce(w, op=dist.ReduceOp.MAX)
    if rank == 0:
        print("DIST_CHECK", "OK" if float(w) < 1e-12 else "FAIL", float(w))
    dist.destroy_process_group()
    return 0 if float(w) < 1e-12 else 1


if __name__ == "__main__":
    sys.exit(main())
