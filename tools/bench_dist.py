"""Strong-scaling measurement of the fibre-partitioned tensor application: the cfg5 interpolation transform
FastLagrIntp::eval_up_Lagr on the d=6, k=1, m=2, NMAX=7 sparse grid, element blocks owned by fibre (adaptive-multiresolution-dg_b200/dist.py),
two NCCL all-to-all layout switches per application.  Run under torchrun, one rank per GPU:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29655 tools/bench_dist.py
Prints one JSON line (rank 0): DoF-transforms/s over all ranks (max over ranks of the device time)."""
import importlib, json, os, sys
import numpy as np, torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    dim, k, m, nmax = [int(x) for x in (sys.argv[1:5] if len(sys.argv) > 4 else (6, 1, 2, 7))]
    steps = int(os.environ.get("STEPS", "10"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    A = importlib.import_module("adaptive-multiresolution-dg_b200")
    D = importlib.import_module("adaptive-multiresolution-dg_b200.dist")
    a, b = k + 1, m + 1
    lev, sup = A.sparse_grid(dim, nmax)
    tb = A.generate_tables(nmax, k, m)
    part = D.FibrePartition(lev, sup, world, rank)
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        T = D.DistTensorApply(A, part, dim, nmax, k, m, local, lambda c: {"pt": c.op_register_compact(tb["pt"])})
        rng = np.random.default_rng(1 + rank)
        u = torch.from_numpy(rng.uniform(-1, 1, size=(len(part.local["X"]), a ** dim))).cuda()
        ops, rels = ["pt"] * dim, [A.REL_VOL] * dim
        for _ in range(3):
            out = T.apply(ops, rels, u, a, b)
    stream.synchronize()
    sw0, by0 = T.switches, T.switch_bytes
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    with torch.cuda.stream(stream):
        for s in range(steps):
            ev[s][0].record(stream)
            out = T.apply(ops, rels, u, a, b)
            ev[s][1].record(stream)
    torch.cuda.synchronize()
    t = np.median([e0.elapsed_time(e1) for e0, e1 in ev])
    tt = torch.tensor([t, float(len(part.local["X"])), float(len(part.local["V"]))], dtype=torch.float64, device="cuda")
    mx = tt.clone()
    if world > 1:
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    if rank == 0:
        ne = lev.shape[0]
        print(json.dumps({"workload": "fibre-partitioned eval_up_Lagr d=%d k=%d m=%d NMAX=%d" % (dim, k, m, nmax), "n_gpus": world, "n_elem": int(ne),
                          "ms_per_application": float(mx[0]), "dof_transforms_per_s": ne * a ** dim / (float(mx[0]) * 1e-3),
                          "max_local_elements": [int(mx[1]), int(mx[2])], "ideal_local_elements": ne / world,
                          "switches_per_application": (T.switches - sw0) // steps, "mb_sent_per_application_rank0": (T.switch_bytes - by0) / steps / 1e6,
                          "launch": "eager (python enqueue, one sweep1d call per sweep)", "scaling": "strong"}))
    T.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
