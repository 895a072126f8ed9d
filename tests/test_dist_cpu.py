"""World-size-2 gloo tests (CPU) of the fibre-partition index logic behind the multi-GPU path
(adaptive-multiresolution-dg_b200/dist.py): ownership keeps fibres whole, and the all-to-all layout switch is a
permutation that lands every element block on its new owner."""
import importlib
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, dim, nmax, q):
    import sys
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        A = importlib.import_module("adaptive-multiresolution-dg_b200")
        D = importlib.import_module("adaptive-multiresolution-dg_b200.dist")
        lev, sup = A.sparse_grid(dim, nmax)
        part = D.FibrePartition(lev, sup, world, rank)
        n = lev.shape[0]
        # every element has exactly one owner per layout, loads are balanced within 2x
        for k in ("X", "V"):
            cnt = np.bincount(part.owner[k], minlength=world)
            assert cnt.sum() == n and cnt.max() <= 2 * max(1, cnt.min()) + 8, cnt
        # fibres along the local dims are complete: all related elements (reference relation tables) are local
        ctx = A.Context(dim, nmax, 1, 2, device=-1)
        ctx.grid_set(lev, sup)
        for layout, dims in (("X", part.dims_x), ("V", part.dims_v)):
            mine = set(part.local[layout].tolist())
            for t in dims:
                ptr, idx = ctx.grid_relation(t, A.REL_FLX)
                for e in part.local[layout]:
                    assert set(idx[ptr[e]:ptr[e + 1]].tolist()) <= mine
        ctx.close()
        # layout switch: blocks labelled by (global id, column)
        blk = 5
        x = torch.tensor(part.local["X"][:, None] * 100 + np.arange(blk)[None, :], dtype=torch.float64)
        v = part.switch(x, "X", "V")
        assert torch.equal(v, torch.tensor(part.local["V"][:, None] * 100 + np.arange(blk)[None, :], dtype=torch.float64))
        back = part.switch(v, "V", "X")
        assert torch.equal(back, x)
        # several live buffers of different block sizes change layout in ONE all-to-all (switch_many), as the schedule's levels do
        xs = [torch.tensor(part.local["X"][:, None] * 100 + 7 * i + np.arange(w)[None, :], dtype=torch.float64) for i, w in enumerate((3, 8, 1))]
        vs = part.switch_many(xs, "X", "V")
        for i, w in enumerate((3, 8, 1)):
            assert vs[i].is_contiguous() and torch.equal(vs[i], torch.tensor(part.local["V"][:, None] * 100 + 7 * i + np.arange(w)[None, :], dtype=torch.float64))
        q.put((rank, "ok"))
    except Exception as e:      # pragma: no cover
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("dim,nmax", [(4, 3), (6, 2), (2, 4)])
def test_partition_and_switch_world2(dim, nmax):
    world = 2
    ctxm = mp.get_context("spawn")
    q = ctxm.Queue()
    port = _free_port()
    procs = [ctxm.Process(target=_worker, args=(r, world, port, dim, nmax, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for r, msg in res:
        assert msg == "ok", msg


def test_partition_with_more_ranks_than_groups():
    """a layout with fewer ownership groups than ranks (coarse grid, many GPUs) leaves ranks without elements; the partition still covers every
    element exactly once in both layouts, empty ranks get empty row lists, and the stage plan of an empty rank has the same operation list as the
    others (it must take part in every barrier)"""
    A = importlib.import_module("adaptive-multiresolution-dg_b200")
    D = importlib.import_module("adaptive-multiresolution-dg_b200.dist")
    S = importlib.import_module("adaptive-multiresolution-dg_b200.stage")
    lev, sup = A.sparse_grid(4, 1)                    # 4-D level 1: 5 elements, at most 3 ownership groups per layout
    world = 8
    parts = [D.FibrePartition(lev, sup, world, r) for r in range(world)]
    for k in ("X", "V"):
        rows = np.concatenate([p.local[k] for p in parts])
        assert sorted(rows.tolist()) == list(range(lev.shape[0]))
        assert sum(1 for p in parts if len(p.local[k]) == 0) >= world - lev.shape[0]
    plans = [S.StagePlan(4, 2, 3, 4, part=p, fuse_rk=True) for p in parts]
    kinds = [[o[0] for o in pl.ops] for pl in plans]
    assert all(k == kinds[0] for k in kinds) and all(pl.n_barrier == plans[0].n_barrier for pl in plans)
