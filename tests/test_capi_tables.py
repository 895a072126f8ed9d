"""CPU-side checks of the C-ABI library: it loads without a GPU, exports every symbol include/amdg.h declares,
and its host index tables (hash keys, 1D orders, fibres, vol/flx relation lists) are bit-exact with the
reference's (tests/golden/, dumped from the compiled reference)."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT, golden_names, load_golden


def test_library_exports_every_declared_symbol(amdg):
    hdr = open(os.path.join(ROOT, "include", "amdg.h")).read()
    declared = set(re.findall(r"\b(amdg_[a-z0-9_]+)\s*\(", hdr))
    declared.discard("amdg_ctx")
    assert declared == set(amdg.SYMBOLS), declared ^ set(amdg.SYMBOLS)
    for name in declared:
        assert getattr(amdg.lib, name) is not None
    assert b"sm_100a" in amdg.lib.amdg_version()


def test_no_cpu_compute_path(amdg):
    ctx = amdg.Context(2, 3, 2, 3, device=-1)
    lev, sup = amdg.sparse_grid(2, 3)
    ctx.grid_set(lev, sup)
    op = ctx.op_register(np.zeros((8 * 3, 8 * 3)), 3, 3)
    with pytest.raises(amdg.AmdgError, match="no CPU compute path"):
        ctx.host_sweep1d(op, amdg.REL_VOL, amdg.LU_FULL, 0, [3, 3], np.zeros(lev.shape[0] * 9), 3)
    ctx.close()


def test_new_entry_points_refuse_to_run_without_a_device(amdg):
    """the entry points added in round 2 keep the contract of the header: a context created without a device builds tables only, every compute call
    fails with AMDG_ENODEVICE and a message (no CPU path exists), table generation works"""
    import ctypes
    ctx = amdg.Context(2, 3, 2, 3, device=-1)
    lev, sup = amdg.sparse_grid(2, 3)
    ctx.grid_set(lev, sup)
    op = ctx.op_generate_points(amdg.BASIS_LAGRANGE, 3)
    assert ctx.op_blocks(op, 3, 4).shape[1:] == (3, 4)
    null = ctypes.c_void_p(0)
    one = (ctypes.c_void_p * 1)(ctypes.c_void_p(8))
    ops = (ctypes.c_int * 2)(op, op); rels = (ctypes.c_int * 2)(0, 0)
    assert amdg.lib.amdg_apply_tensor_coarse(ctx._h, ops, rels, ctypes.c_void_p(8), ctypes.c_void_p(16), 1, 1.0, 0, 2) == -2
    assert b"no CPU compute path" in amdg.lib.amdg_last_error()
    assert amdg.lib.amdg_indicator_norm(ctx._h, 1, one, ctypes.c_void_p(16)) == -2
    sizes = (ctypes.c_int * 2)(3, 3)
    assert amdg.lib.amdg_sweep1d_batch_dual(ctx._h, op, 0, 2, 0, sizes, one, one, None, None, None, None, None, None, 1) == -2
    info = (ctypes.c_int * 5)()
    assert amdg.lib.amdg_ctx_info(ctx._h, info) == 0 and list(info) == [2, 3, 2, 3, -1]
    ctx.close()


@pytest.mark.parametrize("name,key", [("moment_d3_k2_n4", "moment"), ("moment_d4_k1_n3", "moment"), ("vlasov_ampere_d4_k1_n3_v2", "va.E"), ("pw2_vm_d3_k2_n3_v3", "pw2.vm.BE")])
def test_auxiliary_dimension_grid(amdg, name, key):
    """amdg_aux_grid against the reference's DGSolution constructor with auxiliary dimensions (the grids of E / B in the Vlasov examples), joined
    through the bit-exact hash key"""
    d = load_golden(name)
    dim, nmax = int(d["config"][0]), int(d["config"][1])
    lev, sup = amdg.aux_grid(dim, nmax, 2)
    keys = np.array([amdg.hash_key(l, s) for l, s in zip(lev, sup)])
    o = np.argsort(keys, kind="stable")
    assert np.array_equal(keys[o], d[key + ".hash_key"])
    assert np.array_equal(lev[o], d[key + ".level"]) and np.array_equal(sup[o], d[key + ".suppt"])
    assert (lev[:, dim - 2:] == 0).all() and lev[:, :dim - 2].max() == nmax


def test_duplicate_element_in_a_large_grid_is_rejected(amdg):
    """the radix-sorted path of amdg_grid_set (>= 256 elements, threads from 2 048) still finds a duplicated element, and the context keeps its old grid"""
    for nmax in (6, 9):
        lev, sup = amdg.sparse_grid(2, nmax)
        ctx = amdg.Context(2, nmax, 1, 2, device=-1)
        ctx.grid_set(lev, sup)
        n_ok = ctx.n_elem
        bad_l, bad_s = np.vstack([lev, lev[lev.shape[0] // 2:lev.shape[0] // 2 + 1]]), np.vstack([sup, sup[sup.shape[0] // 2:sup.shape[0] // 2 + 1]])
        with pytest.raises(amdg.AmdgError, match="duplicate"):
            ctx.grid_set(bad_l, bad_s)
        assert amdg.lib.amdg_grid_size(ctx._h) == n_ok
        ctx.grid_set(lev[::-1].copy(), sup[::-1].copy())              # and it still takes a valid grid afterwards
        ctx.close()


def test_invalid_arguments(amdg):
    with pytest.raises(amdg.AmdgError):
        amdg.Context(0, 3, 2, 3, device=-1)
    ctx = amdg.Context(2, 3, 2, 3, device=-1)
    with pytest.raises(amdg.AmdgError):
        ctx.grid_set(np.array([[0, 0], [0, 0]]), np.array([[1, 1], [1, 1]]))      # duplicate element
    with pytest.raises(amdg.AmdgError):
        ctx.grid_set(np.array([[4, 0]]), np.array([[1, 1]]))                        # level > nmax
    with pytest.raises(amdg.AmdgError):
        ctx.grid_set(np.array([[2, 0]]), np.array([[5, 1]]))                        # support index out of range
    with pytest.raises(amdg.AmdgError):
        ctx.op_register(np.zeros((5, 5)), 3, 3)                                     # wrong table shape
    ctx.close()


@pytest.mark.parametrize("name", golden_names())
def test_tables_bit_exact(amdg, name):
    d = load_golden(name)
    dim, nmax, n0, sparse, pa, pl, ph, vecnum, herm, ne = [int(x) for x in d["config"]]
    if not name.startswith("adapt_"):
        # the library's own grid generator, then sorted by its own hash key == the reference's sorted order
        lev, sup = amdg.sparse_grid(dim, n0, sparse == 1)
        keys = np.array([amdg.hash_key(l, s) for l, s in zip(lev, sup)])
        o = np.argsort(keys, kind="stable")
        assert np.array_equal(keys[o], d["hash_key"])
        assert np.array_equal(lev[o], d["level"]) and np.array_equal(sup[o], d["suppt"])
    ctx = amdg.Context(dim, nmax, pa, ph if herm else pl, device=-1)
    ctx.grid_set(d["level"], d["suppt"])
    hk, od = ctx.grid_keys()
    assert np.array_equal(hk, d["hash_key"]) and np.array_equal(od, d["order_elem"])
    for t in range(dim):
        for rel, nm in ((amdg.REL_VOL, "vol"), (amdg.REL_FLX, "flx")):
            ptr, idx = ctx.grid_relation(t, rel)
            assert np.array_equal(ptr, d["%s_d%d_ptr" % (nm, t)].astype(np.int64))
            assert np.array_equal(idx, d["%s_d%d_idx" % (nm, t)])
        fptr, fel = ctx.grid_fibres(t)
        assert sorted(fel.tolist()) == list(range(ne))
        others = [k for k in range(dim) if k != t]
        for f in range(len(fptr) - 1):
            rows = fel[fptr[f]:fptr[f + 1]]
            assert (d["order_elem"][rows][:, others] == d["order_elem"][rows[0]][others]).all()
            assert (np.diff(d["order_elem"][rows][:, t]) > 0).all()
    ctx.close()


def test_tables_in_permuted_element_order(amdg):
    """the caller's element order is arbitrary (the glue passes unordered_map iteration order)"""
    d = load_golden("cfg3_wave_d3_k2_n3")
    dim, nmax = int(d["config"][0]), int(d["config"][1])
    ne = d["level"].shape[0]
    perm = np.random.default_rng(1).permutation(ne)
    inv = np.argsort(perm)
    ctx = amdg.Context(dim, nmax, 2, 3, device=-1)
    ctx.grid_set(d["level"][perm], d["suppt"][perm])
    for t in range(dim):
        ptr, idx = ctx.grid_relation(t, amdg.REL_FLX)
        rptr, ridx = d["flx_d%d_ptr" % t], d["flx_d%d_idx" % t]
        for e_new in range(ne):
            e_old = perm[e_new]
            got = sorted(perm[idx[ptr[e_new]:ptr[e_new + 1]]].tolist())
            assert got == ridx[rptr[e_old]:rptr[e_old + 1]].tolist()
    ctx.close()


def _brute_relation(lev, sup, t, flx):
    """Element::is_vol_alpt / is_flx_alpt (source/Element.cpp:265-299, 337-386) by the O(N^2) scan of source/DGSolution.cpp:675-728"""
    n, dim = lev.shape
    def supp(l, j):
        return (0.0, 1.0) if l <= 1 else (2.0 ** (1 - l) * (j - 1) / 2, 2.0 ** (1 - l) * (j + 1) / 2)
    rows = []
    others = [k for k in range(dim) if k != t]
    for e in range(n):
        same = np.all(lev[:, others] == lev[e, others], axis=1) & np.all(sup[:, others] == sup[e, others], axis=1)
        u0, u1 = supp(lev[e, t], sup[e, t])
        out = []
        for f in np.nonzero(same)[0]:
            v0, v1 = supp(lev[f, t], sup[f, t])
            hit = not (u0 >= v1 or u1 <= v0)
            if flx:
                hit = hit or abs(u0 - v1) < 1e-13 or abs(u1 - v0) < 1e-13 or (abs(u0) < 1e-13 and abs(v1 - 1) < 1e-13) or (abs(v0) < 1e-13 and abs(u1 - 1) < 1e-13)
            if hit:
                out.append(int(f))
        rows.append(out)
    return rows


@pytest.mark.parametrize("dim,nmax", [(2, 6), (3, 4)])
def test_grid_change_sequence_matches_fresh_build(amdg, dim, nmax):
    """a context that goes through a sequence of refine / coarsen-like grid changes (neighbour lists of known fibre shapes come from its
    per-shape cache, tables are rebuilt into the previous grid's storage) has exactly the tables of a fresh context, in any row order,
    and they are the brute-force relations of the reference"""
    from test_lean_plans import random_adaptive_grid
    ctx = amdg.Context(dim, nmax, 1, 2, device=-1)
    rng = np.random.default_rng(7)
    for step, keep in enumerate([1.0, 0.7, 0.5, 0.8, 0.35, 1.0]):
        if keep == 1.0:
            lev, sup = amdg.sparse_grid(dim, nmax)
        else:
            lev, sup = random_adaptive_grid(dim, nmax, seed=step, keep=keep)
        o = rng.permutation(lev.shape[0])                     # DGSolution::dg iteration order changes with every rehash
        lev, sup = np.ascontiguousarray(lev[o]), np.ascontiguousarray(sup[o])
        ctx.grid_set(lev, sup)
        fresh = amdg.Context(dim, nmax, 1, 2, device=-1)
        fresh.grid_set(lev, sup)
        assert np.array_equal(ctx.grid_keys()[0], fresh.grid_keys()[0]) and np.array_equal(ctx.grid_keys()[1], fresh.grid_keys()[1])
        for t in range(dim):
            pa, ea = ctx.grid_fibres(t); pb, eb = fresh.grid_fibres(t)
            assert np.array_equal(pa, pb) and np.array_equal(ea, eb)
            for rel in (amdg.REL_VOL, amdg.REL_FLX):
                p1, i1 = ctx.grid_relation(t, rel); p2, i2 = fresh.grid_relation(t, rel)
                assert np.array_equal(p1, p2) and np.array_equal(i1, i2)
                if step in (1, 4):
                    brute = _brute_relation(lev, sup, t, rel == amdg.REL_FLX)
                    for e in range(lev.shape[0]):
                        assert sorted(i1[p1[e]:p1[e + 1]].tolist()) == brute[e]
        fresh.close()
    ctx.close()
