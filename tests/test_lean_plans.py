"""Host-side invariants of the default sweep kernel's work plans (csrc/capi.cu: lean_plan), checked without a GPU through the
diagnostic entry point amdg_lean_plan_check: every row tile of a fibre shape's tile program is in exactly one piece with all its
entries and pairs, staged pieces index their own source rows and fit in shared memory, streamed (coarse) pieces keep fibre-local
sources and at most 8 columns, the rectangles of a piece tile the column plane exactly once."""
import importlib

import numpy as np
import pytest

A = importlib.import_module("adaptive-multiresolution-dg_b200")


def random_adaptive_grid(dim, nmax, seed, keep=0.55):
    """a random downward-closed subset of the sparse grid (leaves removed at random), as DGAdapt::coarsen produces"""
    lev, sup = A.sparse_grid(dim, nmax)
    elems = {tuple(l) + tuple(s) for l, s in zip(lev.tolist(), sup.tolist())}
    rng = np.random.default_rng(seed)

    def children(e):
        out = []
        for d in range(dim):
            n, j = e[d], e[dim + d]
            if n >= nmax:
                continue
            for cj in ([1] if n == 0 else [2 * j - 1, 2 * j + 1]):
                c = list(e); c[d] = n + 1; c[dim + d] = cj
                out.append(tuple(c))
        return out
    target = int(keep * len(elems))
    while len(elems) > target:
        leaves = [e for e in elems if sum(e[:dim]) > 0 and not any(c in elems for c in children(e))]
        rng.shuffle(leaves)
        for e in leaves[:max(1, len(leaves) // 3)]:
            elems.discard(e)
    arr = np.array(sorted(elems), dtype=np.int32)
    return np.ascontiguousarray(arr[:, :dim]), np.ascontiguousarray(arr[:, dim:])


def check_all(ctx, dim, a, b):
    tot = dict(pieces=0, coarse_pieces=0)
    for t in range(dim):
        for sizes, kf, kt in (([a] * dim, a, b), ([b if q < t else a for q in range(dim)], a, b), ([b] * dim, b, a)):
            for rel in (A.REL_VOL, A.REL_FLX):
                for lu in (A.LU_L, A.LU_U, A.LU_FULL):
                    r = ctx.lean_plan_check(t, sizes, kf, kt, rel, lu)
                    assert r["shapes"] > 0 and r["pieces"] >= r["shapes"] and r["max_smem_doubles"] <= 4608
                    tot["pieces"] += r["pieces"]; tot["coarse_pieces"] += r["coarse_pieces"]
    return tot


@pytest.mark.parametrize("dim,nmax,k,m", [(4, 8, 3, 3), (6, 5, 1, 2), (2, 9, 2, 3), (3, 7, 2, 2)])
def test_plans_on_sparse_grids(dim, nmax, k, m):
    lev, sup = A.sparse_grid(dim, nmax)
    ctx = A.Context(dim, nmax, k, m, device=-1)
    ctx.grid_set(lev, sup)
    tot = check_all(ctx, dim, k + 1, m + 1)
    if nmax >= 8:
        assert tot["coarse_pieces"] > 0          # long fibres: the coarse targets are streamed
    ctx.close()


@pytest.mark.parametrize("dim,nmax,k,m,seed", [(2, 8, 2, 3, 1), (3, 6, 1, 2, 2), (4, 5, 3, 3, 3)])
def test_plans_on_random_adaptive_grids(dim, nmax, k, m, seed):
    lev, sup = random_adaptive_grid(dim, nmax, seed)
    ctx = A.Context(dim, nmax, k, m, device=-1)
    ctx.grid_set(lev, sup)
    check_all(ctx, dim, k + 1, m + 1)
    # a grid change keeps the plans of the shapes that survive: the same check passes again after re-setting a coarser grid
    lev2, sup2 = random_adaptive_grid(dim, nmax, seed + 10, keep=0.4)
    ctx.grid_set(lev2, sup2)
    check_all(ctx, dim, k + 1, m + 1)
    ctx.close()


def test_shape_cache_eviction(monkeypatch, capfd):
    """a long adaptive run keeps meeting new fibre shapes: once the known shapes exceed the eviction threshold, plans of shapes the
    current grid no longer has are dropped (csrc/capi.cu: evict_shape_caches) and the plans of the surviving / returning shapes stay valid"""
    monkeypatch.setenv("AMDG_CACHE_SHAPES", "0")
    monkeypatch.setenv("AMDG_VERBOSE", "1")
    dim, nmax, k, m = 2, 8, 2, 3
    ctx = A.Context(dim, nmax, k, m, device=-1)
    evictions = 0
    for seed in range(6):
        lev, sup = random_adaptive_grid(dim, nmax, 100 + seed, keep=0.3 + 0.1 * (seed % 3))
        ctx.grid_set(lev, sup)
        check_all(ctx, dim, k + 1, m + 1)
        evictions += capfd.readouterr().err.count("cache eviction")
    assert evictions >= 1
    ctx.close()
