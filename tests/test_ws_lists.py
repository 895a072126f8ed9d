"""The work lists of the warp-specialised streaming sweep kernel (csrc/ws_items.hpp, kernels_ws.cu), replayed on the CPU and compared with
the oracle: per item the producer's staged slots (runs of the source blocks in the rectangle's compact layout), per sub-unit the consumer's
entry walk (B fragments out of the slots through tab_b, operator fragments in entry order, C fragments through tab_c), heavy items from
global memory.  Test infrastructure: pins the host-built plans, rectangles, tables and operator fragments without a device."""
import importlib

import numpy as np
import pytest

import amdg_oracle as O
from conftest import load_golden

A = importlib.import_module("adaptive-multiresolution-dg_b200")
F = ("prog", "pool_ofs", "fib_ofs", "nfib", "m", "n_rt", "n_src", "n_ent", "tab", "nct", "src_origin", "dst_origin", "nrun", "run_len", "gstride", "slot",
     "kstride", "heavy", "vec", "rows_ofs")


def replay(L, src, n_elem, s_to, kf, kt, coef=1.0, old=None, stage_cap=4096):
    ktp = 1 if kt <= 1 else (2 if kt <= 2 else (4 if kt <= 4 else 8))
    tg = 8 // ktp
    lanes = np.arange(32)
    kk, nn, row8 = lanes & 3, lanes >> 2, lanes >> 2
    dk = np.minimum(4 + kk, kf - 1) - np.minimum(kk, kf - 1)
    dst = np.full((n_elem, s_to), np.nan)
    written = np.zeros((n_elem, s_to), dtype=np.int32)
    pool, ep = L["pool"], L["elem_pool"]
    assert L["cta_ptr"][0] == 0 and L["cta_ptr"][-1] == len(L["items"]) and (np.diff(L["cta_ptr"]) >= 0).all()
    cta_of = np.searchsorted(L["cta_ptr"], np.arange(len(L["items"])), side="right") - 1
    for ii, it in enumerate(L["items"]):
        h = dict(zip(F, [int(x) for x in it[:20]]))
        rows = L["rows"][L["rows_ptr"][cta_of[ii]] + h["rows_ofs"]:]            # the kernel reads sources and targets from here (shared memory copy)
        rt_ptr = pool[h["pool_ofs"]:h["pool_ofs"] + h["n_rt"] + 1]
        rt_id = pool[h["pool_ofs"] + h["n_rt"] + 1:h["pool_ofs"] + 2 * h["n_rt"] + 1]
        ent = pool[h["pool_ofs"] + 2 * h["n_rt"] + 1:h["pool_ofs"] + 2 * h["n_rt"] + 1 + h["n_ent"]]
        src_local = pool[h["pool_ofs"] + 2 * h["n_rt"] + 1 + h["n_ent"]:h["pool_ofs"] + 2 * h["n_rt"] + 1 + h["n_ent"] + h["n_src"]]
        Ap = L["A"][L["prog_ent_ptr"][h["prog"]]:L["prog_ent_ptr"][h["prog"] + 1]]
        assert Ap.shape[0] == h["n_ent"] == rt_ptr[-1]
        if not h["heavy"]:
            assert h["nfib"] * h["n_src"] * h["slot"] <= stage_cap and h["slot"] == h["nrun"] * h["run_len"]
        else:
            assert h["nct"] == 1 and h["n_rt"] == 1
        for b in range(h["nfib"]):
            fo = h["fib_ofs"] + b * h["m"]
            slots = []
            if not h["heavy"]:
                for s in range(h["n_src"]):
                    row = rows[b * h["n_src"] + s]
                    assert row == ep[fo + src_local[s]]
                    slots.append(np.concatenate([src[row, h["src_origin"] + r * h["gstride"]:h["src_origin"] + r * h["gstride"] + h["run_len"]] for r in range(h["nrun"])]))
            for ri in range(h["n_rt"]):
                rt = int(rt_id[ri])
                for ct in range(h["nct"]):
                    tile = h["tab"] + ct
                    acc = np.zeros((8, 8))
                    for p in range(rt_ptr[ri], rt_ptr[ri + 1]):
                        code = int(ent[p])
                        boff = L["tab_b"][tile] + (code & 1) * dk * h["kstride"]
                        B = np.zeros((4, 8))
                        if h["heavy"]:
                            B[kk, nn] = src[rows[p], h["src_origin"] + boff]
                        else:
                            B[kk, nn] = slots[code >> 1][boff]
                        Am = np.zeros((8, 4)); Am[row8, kk] = Ap[p]
                        acc += Am @ B
                    for lane in range(32):
                        tl = rt * tg + (row8[lane] // ktp)
                        e = rows[(h["n_ent"] if h["heavy"] else h["nfib"] * h["n_src"] + (b * h["n_rt"] + ri) * tg) + (row8[lane] // ktp)]
                        assert (e < 0) == (tl >= h["m"])
                        if e < 0:
                            continue
                        assert e == ep[fo + tl]
                        for hh in range(2):
                            off = L["tab_c"][tile, lane, hh]
                            if off >= 0:
                                v = coef * acc[row8[lane], 2 * (lane & 3) + hh]
                                if old is not None:
                                    v += old[e, h["dst_origin"] + off]
                                dst[e, h["dst_origin"] + off] = v
                                written[e, h["dst_origin"] + off] += 1
    assert (written == 1).all(), "every output is stored exactly once"
    return dst


CASES = [("cfg1_adv_d2_k2_n4", "alpt"), ("cfg2_rt_d4_k3_n3", "pt"), ("cfg5_vlasov_d6_k1_n2", "pt"), ("adapt_d2_k2_n6", "pt"), ("line_d1_k2_n5", "pt")]


@pytest.mark.parametrize("name,which", CASES)
def test_ws_list_replay(name, which):
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
    for t in sorted({0, dim - 1, dim // 2}):
        for relname, rel in (("vol", A.REL_VOL), ("flx", A.REL_FLX)):
            rels = O.relations(lev, sup, t, relname)
            for luname, lu in (("L", A.LU_L), ("U", A.LU_U), ("full", A.LU_FULL)):
                sizes = [kt if q < t else kf for q in range(dim)]
                outer = int(np.prod(sizes[:t])) if t else 1
                inner = int(np.prod(sizes[t + 1:])) if t < dim - 1 else 1
                src = rng.uniform(-1, 1, size=(n, outer * kf * inner))
                ref, _ = O.transform_1d(src, sizes, mat, luname, rels, lev, ord1d, t, kf - 1, kt - 1, coef=0.7)
                L = ctx.ws_list_export(op, rel, lu, t, sizes, n_cta=7)
                got = replay(L, src, n, outer * kt * inner, kf, kt, coef=0.7)
                assert not np.isnan(got).any()
                err = np.linalg.norm(got - ref) / max(np.linalg.norm(ref), 1e-300)
                assert err < 1e-13, (name, t, relname, luname, err)
    ctx.close()


@pytest.mark.parametrize("dim,nmax,k,m,cap", [(2, 7, 2, 3, 1200), (3, 5, 1, 2, 1200), (2, 8, 3, 3, 4096), (3, 4, 2, 5, 1400)])
def test_ws_list_small_stage_and_long_fibres(dim, nmax, k, m, cap, monkeypatch):
    """long fibres (heavy items, target-cut pieces) and, through a small stage (AMDG_WS_CAP), the o-slice and i-strip rectangles; block edge 6
    (two k-parts per source): replay == oracle"""
    monkeypatch.setenv("AMDG_WS_CAP", str(cap))
    lev, sup = A.sparse_grid(dim, nmax)
    n = lev.shape[0]
    ord1d = np.array([[0 if l == 0 else 2 ** (l - 1) + (j - 1) // 2 for l, j in zip(ll, jj)] for ll, jj in zip(lev, sup)])
    ctx = A.Context(dim, nmax, k, m, device=-1)
    ctx.grid_set(lev, sup)
    a, b = k + 1, m + 1
    T = 2 ** nmax
    rng = np.random.default_rng(3)
    src_, tgt_, vol_ = ctx.pairs()
    for kf, kt in ((b, a), (a, b)):
        blocks = rng.standard_normal((len(src_), kf, kt))
        mat = np.zeros((T * kf, T * kt))
        for p in range(len(src_)):
            mat[src_[p] * kf:(src_[p] + 1) * kf, tgt_[p] * kt:(tgt_[p] + 1) * kt] = blocks[p]
        op = ctx.op_register_compact(blocks)
        for t in range(dim):
            sizes = [kt if q < t else kf for q in range(dim)]
            outer = int(np.prod(sizes[:t])) if t else 1
            inner = int(np.prod(sizes[t + 1:])) if t < dim - 1 else 1
            src = rng.uniform(-1, 1, size=(n, outer * kf * inner))
            for relname, rel, luname, lu in (("vol", A.REL_VOL, "full", A.LU_FULL), ("flx", A.REL_FLX, "L", A.LU_L)):
                ref, _ = O.transform_1d(src, sizes, mat, luname, O.relations(lev, sup, t, relname), lev, ord1d, t, kf - 1, kt - 1)
                L = ctx.ws_list_export(op, rel, lu, t, sizes, n_cta=5)
                got = replay(L, src, n, outer * kt * inner, kf, kt, stage_cap=cap)
                assert np.linalg.norm(got - ref) / np.linalg.norm(ref) < 1e-13
                assert (L["items"][:, 17] == 1).any() or nmax < 7
                if cap < 1500 and t == 0 and dim > 1 and inner * kf * 8 > cap:
                    assert (L["items"][:, 12] > 1).any()          # i-strips: kf runs per source
                if cap < 1500 and t == dim - 1 and dim > 1 and outer * kf * 8 > cap:
                    assert (L["items"][:, 10] > 0).any()          # o-slices: rectangles that do not start at column 0
    ctx.close()
