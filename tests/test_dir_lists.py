"""The work lists of the register-direct sweep kernel (csrc/dir_items.hpp, kernels_dir.cu), replayed on the CPU and compared with the oracle.

The replay below does, unit by unit and column tile by column tile, exactly what a warp of sweep_dir_kernel does with the exported tables
(B fragments through tab_b, one 8x4 operator fragment per (source, row tile) of the source's mask, C fragments stored through tab_c), in
numpy.  It is test infrastructure: it pins the host-built plans, tables and operator fragments without a device; the device kernel itself is
compared with the reference dumps in test_gpu_parity.py."""
import importlib

import numpy as np
import pytest

import amdg_oracle as O
from conftest import load_golden

A = importlib.import_module("adaptive-multiresolution-dg_b200")


def replay(L, src, n_elem, s_from, s_to, kf, kt, inner, coef=1.0, old=None):
    ktp = 1 if kt <= 1 else (2 if kt <= 2 else (4 if kt <= 4 else 8))
    tg = 8 // ktp
    lanes = np.arange(32)
    kk, nn, row8 = lanes & 3, lanes >> 2, lanes >> 2
    dkp = (np.minimum(4 + kk, kf - 1) - np.minimum(kk, kf - 1)) * inner
    dst = np.full((n_elem, s_to), np.nan) if old is None else old.copy()
    written = np.zeros((n_elem, s_to), dtype=np.int32)
    for u in L["units"]:
        pool_ofs, fib_ofs, nfib, m, ct0, nct, n_src, prog, variant, n_rt = [int(x) for x in u[:10]]
        rt_id = [int(x) for x in u[10:14]]
        assert n_rt <= (1, 2, 4, 1)[variant]
        codes = L["pool"][pool_ofs:pool_ofs + n_src]
        masks = L["pool"][pool_ofs + n_src:pool_ofs + 2 * n_src]
        Ap = L["A"][L["prog_ent_ptr"][prog]:L["prog_ent_ptr"][prog + 1]]
        assert int(sum(bin(int(x)).count("1") for x in masks)) == Ap.shape[0]
        for b in range(nfib):
            fo = fib_ofs + b * m
            for ct in range(ct0, ct0 + nct):
                acc = np.zeros((n_rt, 8, 8))
                if old is not None:
                    for r in range(n_rt):
                        for lane in range(32):
                            tl = rt_id[r] * tg + (row8[lane] // ktp)
                            if tl >= m:
                                continue
                            e = L["elem_pool"][fo + tl]
                            for h in range(2):
                                off = L["tab_c"][ct, lane, h]
                                if off >= 0:
                                    acc[r, row8[lane], 2 * (lane & 3) + h] = old[e, off]
                abase = 0
                for s in range(n_src):
                    row = L["elem_pool"][fo + (codes[s] >> 1)]
                    boff = L["tab_b"][ct] + np.where(codes[s] & 1, dkp, 0)
                    B = np.zeros((4, 8))
                    B[kk, nn] = src[row, boff] * coef
                    mk = int(masks[s])
                    for r in range(n_rt):
                        if (mk >> r) & 1:
                            Am = np.zeros((8, 4))
                            Am[row8, kk] = Ap[abase + bin(mk & ((1 << r) - 1)).count("1")]
                            acc[r] += Am @ B
                    abase += bin(mk).count("1")
                for r in range(n_rt):
                    for lane in range(32):
                        tl = rt_id[r] * tg + (row8[lane] // ktp)
                        if tl >= m:
                            continue
                        e = L["elem_pool"][fo + tl]
                        for h in range(2):
                            off = L["tab_c"][ct, lane, h]
                            if off >= 0:
                                dst[e, off] = acc[r, row8[lane], 2 * (lane & 3) + h]
                                written[e, off] += 1
    assert (written == 1).all(), "every output is stored exactly once"
    return dst


CASES = [("cfg1_adv_d2_k2_n4", "alpt"), ("cfg2_rt_d4_k3_n3", "pt"), ("cfg5_vlasov_d6_k1_n2", "pt"), ("adapt_d2_k2_n6", "pt"), ("line_d1_k2_n5", "pt")]


@pytest.mark.parametrize("name,which", CASES)
def test_dir_list_replay(name, which):
    d = load_golden(name)
    dim, nmax, n0, sparse, pa, pl = [int(x) for x in d["config"][:6]]
    a, b = pa + 1, pl + 1
    lev, sup, ord1d = d["level"], d["suppt"], d["order_elem"]
    n = lev.shape[0]
    ctx = A.Context(dim, nmax, pa, pl, device=-1)
    ctx.grid_set(lev, sup)
    rng = np.random.default_rng(7)
    if which == "pt" and "Lag_pt_Alpt_1D" in d:
        mat, kf, kt = d["Lag_pt_Alpt_1D"].T.copy(), a, b
    else:
        key = [k for k in d if k.endswith("ujp_vjp") or k.endswith("u_vx")][0]
        mat, kf, kt = d[key], a, a
    op = ctx.op_register(mat, kf, kt)
    ts = sorted({0, dim - 1, dim // 2})
    for t in ts:
        for relname, rel in (("vol", A.REL_VOL), ("flx", A.REL_FLX)):
            rels = O.relations(lev, sup, t, relname)
            for luname, lu in (("L", A.LU_L), ("U", A.LU_U), ("full", A.LU_FULL)):
                sizes = [kt if q < t else kf for q in range(dim)]           # dims already swept have the target edge
                outer = int(np.prod(sizes[:t])) if t else 1
                inner = int(np.prod(sizes[t + 1:])) if t < dim - 1 else 1
                src = rng.uniform(-1, 1, size=(n, outer * kf * inner))
                ref, _ = O.transform_1d(src, sizes, mat, luname, rels, lev, ord1d, t, kf - 1, kt - 1, coef=0.7)
                L = ctx.dir_list_export(op, rel, lu, t, sizes)
                got = replay(L, src, n, outer * kf * inner, outer * kt * inner, kf, kt, inner, coef=0.7)
                assert not np.isnan(got).any()
                err = np.linalg.norm(got - ref) / max(np.linalg.norm(ref), 1e-300)
                assert err < 1e-13, (name, t, relname, luname, err)
    # accumulate: old values initialise the accumulators
    t = 0
    sizes = [kf] * dim
    inner = int(np.prod(sizes[1:])) if dim > 1 else 1
    src = rng.uniform(-1, 1, size=(n, kf * inner)); old = rng.uniform(-1, 1, size=(n, kt * inner))
    ref, _ = O.transform_1d(src, sizes, mat, "full", O.relations(lev, sup, 0, "vol"), lev, ord1d, 0, kf - 1, kt - 1)
    L = ctx.dir_list_export(op, A.REL_VOL, A.LU_FULL, 0, sizes)
    got = replay(L, src, n, kf * inner, kt * inner, kf, kt, inner, old=old)
    assert np.linalg.norm(got - (ref + old)) / np.linalg.norm(ref + old) < 1e-13
    ctx.close()


def test_dir_list_long_fibres_cover():
    """full-size cfg2 grid (fibres up to 256 elements): heavy row tiles become narrow pieces; every (fibre, row tile, column tile) is owned once"""
    lev, sup = A.sparse_grid(4, 8)
    ctx = A.Context(4, 8, 3, 3, device=-1)
    ctx.grid_set(lev, sup)
    src, tgt, vol = ctx.pairs()
    op = ctx.op_register_compact(np.ones((len(src), 4, 4)))
    for lu in (A.LU_L, A.LU_U, A.LU_FULL):
        L = ctx.dir_list_export(op, A.REL_VOL, lu, 1, [4, 4, 4, 4])
        u = L["units"]
        assert (u[:, 8] == 3).any() or lu == A.LU_U          # narrow pieces exist where coarse targets read long source lists
        assert L["vec_ok"] and L["nct"] == 8
        # tiles owned: sum over units of fibres * row tiles * column tiles == sum over fibres of (row tiles of the fibre) * 8
        fptr, fel = ctx.grid_fibres(1)
        n_rt_total = sum((int(fptr[i + 1] - fptr[i]) + 1) // 2 for i in range(len(fptr) - 1))
        assert int((u[:, 2] * u[:, 9] * u[:, 5]).sum()) == n_rt_total * 8
    ctx.close()
