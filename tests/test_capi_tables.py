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
