"""B200-native fast sparse-grid transform path of AdaM-DG (FP64 CUDA for sm_100a behind a C ABI).

This package is a thin ctypes binding of libamdg_b200.so (include/amdg.h).  There is no CPU fallback: if the
shared library is missing the import fails, and every compute call needs a CUDA device.  torch is used only
as plumbing (device buffers, streams, torch.distributed) -- no torch op is on the compute path.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("AMDG_LIB", os.path.join(_HERE, "libamdg_b200.so"))

if not os.path.exists(LIB_PATH):
    raise ImportError(
        "libamdg_b200.so is not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
        "(nvcc -gencode arch=compute_100a,code=sm_100a).  There is no fallback path.")

lib = ctypes.CDLL(LIB_PATH)

REL_VOL, REL_FLX = 0, 1
BASIS_ALPERT, BASIS_LAGRANGE, BASIS_HERMITE = 0, 1, 2
TABLES = {"u_v": 0, "u_vx": 1, "ulft_vjp": 2, "urgt_vjp": 3, "ujp_vjp": 4, "uave_vjp": 5, "ujp_vxlft": 6, "ujp_vxrgt": 7, "ux_vx": 8, "uxave_vjp": 9,
          "ujp_vxave": 10, "ux_v": 11}
LU_L, LU_U, LU_FULL = 0, 1, 2
SCHED_LITERAL, SCHED_SHARED = 0, 1
PW = dict(VAR=1, X=2, OTHER=3, CONST=4, ADD=5, SUB=6, MUL=7, DIV=8, NEG=9, SIN=10, COS=11, SQR=12, EXP=13, SQRT=14, ABS=15, POW=16, TANH=17, MIN=18, MAX=19)
FLUX_LINEAR, FLUX_BURGERS, FLUX_SIN, FLUX_COS, FLUX_BUCKLEY_X, FLUX_BUCKLEY_Y, FLUX_VLASOV_SMOOTH_E = range(7)
RK_EULER, RK_RK2SSP, RK_RK2MID, RK_RK3SSP, RK_RK3HEUN = range(5)

_i, _i64, _d, _p = ctypes.c_int, ctypes.c_int64, ctypes.c_double, ctypes.c_void_p
_ip = ctypes.POINTER(ctypes.c_int)
_lp = ctypes.POINTER(ctypes.c_int64)
_dp = ctypes.POINTER(ctypes.c_double)

# every symbol include/amdg.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "amdg_version": (ctypes.c_char_p, []),
    "amdg_last_error": (ctypes.c_char_p, []),
    "amdg_ctx_create": (_i, [_i, _i, _i, _i, _i, ctypes.POINTER(_p)]),
    "amdg_ctx_destroy": (_i, [_p]),
    "amdg_ctx_set_stream": (_i, [_p, _p]),
    "amdg_ctx_sync": (_i, [_p]),
    "amdg_ctx_set_schedule": (_i, [_p, _i]),
    "amdg_ctx_set_kernel": (_i, [_p, _i]),
    "amdg_ctx_launch_count": (_i64, [_p]),
    "amdg_ctx_set_debug_buffer": (_i, [_p, _p]),
    "amdg_lean_plan_check": (_i, [_p, _i, _ip, _i, _i, _i, _i, _lp]),
    "amdg_hash_key": (_i, [_i, _ip, _ip]),
    "amdg_order_elem": (_i, [_i, _i]),
    "amdg_sparse_grid": (_i64, [_i, _i, _i, _ip, _ip]),
    "amdg_aux_grid": (_i64, [_i, _i, _i, _ip, _ip]),
    "amdg_grid_set": (_i, [_p, _i64, _ip, _ip]),
    "amdg_grid_size": (_i64, [_p]),
    "amdg_grid_keys": (_i, [_p, _ip, _ip]),
    "amdg_grid_relation": (_i64, [_p, _i, _i, _lp, _ip]),
    "amdg_grid_fibres": (_i64, [_p, _i, _lp, _ip]),
    "amdg_op_register": (_i, [_p, _dp, _i, _i, _i, _i, _ip]),
    "amdg_pairs": (_i64, [_p, _ip, _ip, _ip]),
    "amdg_op_register_compact": (_i, [_p, _dp, _i64, _i, _i, _i, _ip]),
    "amdg_op_blocks": (_i, [_p, _i, _dp]),
    "amdg_op_register_hier": (_i, [_p, _ip, _dp, _i, _ip]),
    "amdg_op_combine": (_i, [_p, _i, _d, _i, _d, _ip]),
    "amdg_sweep1d": (_i, [_p, _i, _i, _i, _i, _ip, _p, _p, _i, _d, _i]),
    "amdg_sweep1d_batch": (_i, [_p, _i, _i, _i, _i, _ip, _p, _p, _dp, _ip, _i, _i]),
    "amdg_ctx_info": (_i, [_p, _p]),
    "amdg_op_generate": (_i, [_p, _i, _i, _i, _i, _p]),
    "amdg_op_generate_bc": (_i, [_p, _i, _i, _i, _i, _i, _p]),
    "amdg_op_generate_points": (_i, [_p, _i, _i, _i, _i, _p]),
    "amdg_op_generate_hier": (_i, [_p, _i, _i, _i, _p]),
    "amdg_points_generate": (_i, [_p, _i, _i, _i, _p]),
    "amdg_sweep1d_batch_dual": (_i, [_p, _i, _i, _i, _i, _ip, _p, _p, _dp, _ip, _p, _p, _p, _p, _i]),
    "amdg_apply_tensor": (_i, [_p, _ip, _ip, _p, _p, _i, _d, _i]),
    "amdg_apply_tensor_coarse": (_i, [_p, _ip, _ip, _p, _p, _i, _d, _i, _i]),
    "amdg_hierarchize": (_i, [_p, _i, _p, _p, _i]),
    "amdg_pointwise": (_i, [_p, _i, _ip, _dp, _p, _p, _p]),
    "amdg_pointwise_hermite2d": (_i, [_p, _i, _ip, _dp, _p, _p]),
    "amdg_point_coords": (_i, [_p, _dp, _p]),
    "amdg_rk_stage": (_i, [_p, _i, _i, _d, _p, _p, _p, _i64]),
    "amdg_rk4_ode2nd_stage": (_i, [_p, _i, _d, _p, _p, _p, _p, _p, _p, _p, _i64]),
    "amdg_axpby": (_i, [_p, _i64, _d, _p, _d, _p]),
    "amdg_lincomb": (_i, [_p, _i64, _i, _dp, _p, _d, _p]),
    "amdg_moment": (_i, [_p, _i64, _p, _i, _ip, _d, _p, _p]),
    "amdg_indicator_norm": (_i, [_p, _i, _p, _p]),
    "amdg_sweep1d_batch_mapped": (_i, [_p, _i, _i, _i, _i, _ip, _p, _p, _dp, _ip, _p, _p, _i]),
    "amdg_points_set": (_i, [_p, _dp]),
    "amdg_pointwise_expr": (_i, [_p, _i, _p, _i, _p, _p, _i, _p, _ip, _i, _ip, _dp, _i]),
    "amdg_peer_export": (_i, [_p, _p, _p]),
    "amdg_peer_open": (_i, [_p, _p, ctypes.POINTER(_p)]),
    "amdg_peer_close": (_i, [_p, _p]),
    "amdg_peer_barrier": (_i, [_p, _p, _i, _i, _p, _p]),
    "amdg_scatter_rows": (_i, [_p, _p, _i64, _i, _p, _p]),
    "amdg_host_apply_tensor": (_i, [_p, _ip, _ip, _dp, _dp, _i, _d, _i]),
    "amdg_host_sweep1d": (_i, [_p, _i, _i, _i, _i, _ip, _dp, _dp, _i, _d, _i]),
    "amdg_host_hierarchize": (_i, [_p, _i, _dp, _dp, _i]),
    "amdg_host_roundtrip": (_i, [_p, _i, _i, _i, _dp, _dp, _i]),
    "amdg_dev_alloc": (_i, [_p, _i64, ctypes.POINTER(_p)]),
    "amdg_dev_free": (_i, [_p, _p]),
    "amdg_dev_upload": (_i, [_p, _p, _dp, _i64]),
    "amdg_dev_download": (_i, [_p, _dp, _p, _i64]),
    "amdg_dev_zero": (_i, [_p, _p, _i64]),
}
for _name, (_res, _args) in SYMBOLS.items():
    _f = getattr(lib, _name)
    _f.restype = _res
    _f.argtypes = _args


class AmdgError(RuntimeError):
    pass


def _check(rc):
    if rc < 0:
        raise AmdgError("amdg error %d: %s" % (rc, lib.amdg_last_error().decode()))
    return rc


def _ints(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(_ip)


def _dbls(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(_dp)


def _ptr(t):
    """device pointer of a torch tensor (contiguous, on the context's device) or a raw int"""
    if isinstance(t, (int, np.integer)):
        return ctypes.c_void_p(int(t))
    assert t.is_cuda and t.is_contiguous(), "need a contiguous CUDA tensor"
    return ctypes.c_void_p(t.data_ptr())


def hash_key(level, suppt):
    l, lp = _ints(level)
    j, jp = _ints(suppt)
    return lib.amdg_hash_key(len(l), lp, jp)


def sparse_grid(dim, level_init, sparse=True):
    """Initial grid of DGSolution (reference source/DGSolution.cpp:10-57), construction order."""
    n = _check(lib.amdg_sparse_grid(dim, level_init, int(sparse), None, None))
    lev = np.zeros((n, dim), dtype=np.int32)
    sup = np.zeros((n, dim), dtype=np.int32)
    _check(lib.amdg_sparse_grid(dim, level_init, int(sparse), lev.ctypes.data_as(_ip), sup.ctypes.data_as(_ip)))
    return lev, sup


def aux_grid(dim, level_init, aux_dim):
    """Grid of a field solution with auxiliary dimensions (reference source/DGSolution.cpp:59-116): full grid in the first dim - aux_dim dimensions,
    level 0 in the rest."""
    n = _check(lib.amdg_aux_grid(dim, level_init, aux_dim, None, None))
    lev = np.zeros((n, dim), dtype=np.int32)
    sup = np.zeros((n, dim), dtype=np.int32)
    _check(lib.amdg_aux_grid(dim, level_init, aux_dim, lev.ctypes.data_as(_ip), sup.ctypes.data_as(_ip)))
    return lev, sup


class Context:
    """One problem configuration on one device (dim, NMAX, Alpert and interpolation degrees)."""

    def __init__(self, dim, nmax, pmax_alpt, pmax_intp, device=0):
        self.dim, self.nmax, self.a, self.b = dim, nmax, pmax_alpt + 1, pmax_intp + 1
        self.device = device
        h = _p()
        _check(lib.amdg_ctx_create(dim, nmax, pmax_alpt, pmax_intp, device, ctypes.byref(h)))
        self._h = h
        self.n_elem = 0

    def close(self):
        if getattr(self, "_h", None):
            lib.amdg_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- configuration
    def set_stream(self, cuda_stream_ptr):
        _check(lib.amdg_ctx_set_stream(self._h, ctypes.c_void_p(cuda_stream_ptr)))

    def sync(self):
        _check(lib.amdg_ctx_sync(self._h))

    def set_schedule(self, sched):
        _check(lib.amdg_ctx_set_schedule(self._h, sched))

    def set_kernel(self, variant):
        _check(lib.amdg_ctx_set_kernel(self._h, variant))

    def set_debug_buffer(self, t):
        _check(lib.amdg_ctx_set_debug_buffer(self._h, ctypes.c_void_p(t.data_ptr()) if t is not None else None))

    @property
    def launch_count(self):
        return lib.amdg_ctx_launch_count(self._h)

    # ---- grid
    def grid_set(self, level, suppt):
        l, lp = _ints(level)
        j, jp = _ints(suppt)
        assert l.shape == j.shape and l.ndim == 2 and l.shape[1] == self.dim
        _check(lib.amdg_grid_set(self._h, l.shape[0], lp, jp))
        self.n_elem = l.shape[0]

    def grid_keys(self):
        hk = np.zeros(self.n_elem, dtype=np.int32)
        od = np.zeros((self.n_elem, self.dim), dtype=np.int32)
        _check(lib.amdg_grid_keys(self._h, hk.ctypes.data_as(_ip), od.ctypes.data_as(_ip)))
        return hk, od

    def grid_relation(self, t, rel):
        nnz = _check(lib.amdg_grid_relation(self._h, t, rel, None, None))
        ptr = np.zeros(self.n_elem + 1, dtype=np.int64)
        idx = np.zeros(max(nnz, 1), dtype=np.int32)
        _check(lib.amdg_grid_relation(self._h, t, rel, ptr.ctypes.data_as(_lp), idx.ctypes.data_as(_ip)))
        return ptr, idx[:nnz]

    def grid_fibres(self, t):
        nf = _check(lib.amdg_grid_fibres(self._h, t, None, None))
        ptr = np.zeros(nf + 1, dtype=np.int64)
        el = np.zeros(self.n_elem, dtype=np.int32)
        _check(lib.amdg_grid_fibres(self._h, t, ptr.ctypes.data_as(_lp), el.ctypes.data_as(_ip)))
        return ptr, el

    # ---- operators
    def op_register(self, dense, edge_from, edge_to):
        m, mp = _dbls(dense)
        out = _i()
        _check(lib.amdg_op_register(self._h, mp, m.shape[0], m.shape[1], edge_from, edge_to, ctypes.byref(out)))
        return out.value

    def op_register_hier(self, anc, wt):
        a, ap = _ints(anc)
        w, wp = _dbls(wt)
        out = _i()
        _check(lib.amdg_op_register_hier(self._h, ap, wp, w.shape[-1], ctypes.byref(out)))
        return out.value

    # ---- tables generated by the library (no reference run): compact blocks per related 1D pair
    def op_generate(self, basis_u, pmax_u, table, msh_case=1, boundary="period"):
        out = _i()
        bc = {"period": 0, "zero": 1, "inside": 2}[boundary]
        _check(lib.amdg_op_generate_bc(self._h, basis_u, pmax_u, msh_case, TABLES[table] if isinstance(table, str) else table, bc, ctypes.byref(out)))
        return out.value

    def op_generate_points(self, basis, pmax, msh_case=1, derivative=0):
        out = _i()
        _check(lib.amdg_op_generate_points(self._h, basis, pmax, msh_case, derivative, ctypes.byref(out)))
        return out.value

    def op_generate_hier(self, basis, pmax, msh_case=1):
        out = _i()
        _check(lib.amdg_op_generate_hier(self._h, basis, pmax, msh_case, ctypes.byref(out)))
        return out.value

    def points_generate(self, basis, pmax, msh_case=1):
        pts = np.zeros((1 << self.nmax) * (pmax + 1), dtype=np.float64)
        _check(lib.amdg_points_generate(self._h, basis, pmax, msh_case, pts.ctypes.data_as(ctypes.c_void_p)))
        return pts

    def pairs(self):
        n = _check(lib.amdg_pairs(self._h, None, None, None))
        src, tgt, vol = (np.zeros(n, dtype=np.int32) for _ in range(3))
        _check(lib.amdg_pairs(self._h, src.ctypes.data_as(_ip), tgt.ctypes.data_as(_ip), vol.ctypes.data_as(_ip)))
        return src, tgt, vol

    def op_register_compact(self, blocks, hier=False):
        b, bp = _dbls(blocks)
        out = _i()
        _check(lib.amdg_op_register_compact(self._h, bp, b.shape[0], b.shape[1], b.shape[2], int(hier), ctypes.byref(out)))
        return out.value

    def op_blocks(self, op, edge_from, edge_to):
        n = _check(lib.amdg_pairs(self._h, None, None, None))
        out = np.zeros((n, edge_from, edge_to))
        _check(lib.amdg_op_blocks(self._h, op, out.ctypes.data_as(_dp)))
        return out

    def op_combine(self, op_a, alpha, op_b, beta):
        out = _i()
        _check(lib.amdg_op_combine(self._h, op_a, alpha, op_b, beta, ctypes.byref(out)))
        return out.value

    # ---- device compute (torch tensors are only carriers of device pointers)
    def sweep1d(self, op, rel, lu, t, sizes_from, src, dst, n_comp=1, coef=1.0, accumulate=False):
        s, sp = _ints(sizes_from)
        _check(lib.amdg_sweep1d(self._h, op, rel, lu, t, sp, _ptr(src), _ptr(dst), n_comp, coef, int(accumulate)))

    def lean_plan_check(self, t, sizes_from, kf, kt, rel, lu):
        """host-only self check of the default sweep kernel's work plans (amdg_lean_plan_check); returns a dict of counts"""
        s, sp = _ints(sizes_from)
        out = np.zeros(6, dtype=np.int64)
        _check(lib.amdg_lean_plan_check(self._h, t, sp, kf, kt, rel, lu, out.ctypes.data_as(_lp)))
        return dict(zip(("shapes", "pieces", "coarse_pieces", "entries", "max_staged_rows", "max_smem_doubles"), out.tolist()))

    def sweep1d_batch(self, op, rel, lu, t, sizes_from, srcs, dsts, coefs=None, accumulates=None, n_comp=1):
        """one launch for several (src, dst) pairs that share operator, relation, L/U part and dimension (amdg_sweep1d_batch)"""
        n = len(srcs)
        s, sp = _ints(np.asarray(sizes_from).reshape(n, self.dim))
        ps = (ctypes.c_void_p * n)(*[_ptr(x) for x in srcs])
        pd = (ctypes.c_void_p * n)(*[_ptr(x) for x in dsts])
        cf, cp = _dbls(np.ones(n) if coefs is None else coefs)
        ac, ap = _ints(np.zeros(n) if accumulates is None else accumulates)
        _check(lib.amdg_sweep1d_batch(self._h, op, rel, lu, t, sp, ps, pd, cp, ap, n, n_comp))

    def apply_tensor(self, ops, rels, src, dst, n_comp=1, coef=1.0, accumulate=False):
        o, op = _ints(ops)
        r, rp = _ints(rels)
        _check(lib.amdg_apply_tensor(self._h, op, rp, _ptr(src), _ptr(dst), n_comp, coef, int(accumulate)))

    def apply_tensor_coarse(self, ops, rels, src, dst, mesh_nmax, n_comp=1, coef=1.0, accumulate=False):
        """the *_coarse_grid transforms: only elements with sum of levels <= mesh_nmax take part"""
        o, op = _ints(ops)
        r, rp = _ints(rels)
        _check(lib.amdg_apply_tensor_coarse(self._h, op, rp, _ptr(src), _ptr(dst), n_comp, coef, int(accumulate), int(mesh_nmax)))

    def hierarchize(self, hier_op, src, dst, n_comp=1):
        _check(lib.amdg_hierarchize(self._h, hier_op, _ptr(src), _ptr(dst), n_comp))

    def pointwise(self, flux_ids, params, up, fp, pts=None):
        f, fp_ = _ints(flux_ids)
        prm = np.zeros((len(f), 4)) if params is None else np.asarray(params, dtype=np.float64).reshape(len(f), 4)
        prm, pp = _dbls(prm)
        _check(lib.amdg_pointwise(self._h, len(f), fp_, pp, _ptr(up), _ptr(fp), _ptr(pts) if pts is not None else None))

    def pointwise_hermite2d(self, flux_ids, params, up, fp):
        f, fp_ = _ints(flux_ids)
        prm = np.zeros((len(f), 4)) if params is None else np.asarray(params, dtype=np.float64).reshape(len(f), 4)
        prm, pp = _dbls(prm)
        _check(lib.amdg_pointwise_hermite2d(self._h, len(f), fp_, pp, _ptr(up), _ptr(fp)))

    def point_coords(self, pts1d, dev_pts):
        p, pp = _dbls(pts1d)
        _check(lib.amdg_point_coords(self._h, pp, _ptr(dev_pts)))

    def rk_stage(self, scheme, stage, dt, u_tn, u, rhs):
        _check(lib.amdg_rk_stage(self._h, scheme, stage, dt, _ptr(u_tn), _ptr(u), _ptr(rhs), u.numel()))

    def rk4_ode2nd_stage(self, stage, dt, u_tn, v_tn, u, v, rhs, ku, kv):
        _check(lib.amdg_rk4_ode2nd_stage(self._h, stage, dt, _ptr(u_tn), _ptr(v_tn), _ptr(u), _ptr(v), _ptr(rhs), _ptr(ku), _ptr(kv), u.numel()))

    def lincomb(self, coefs, xs, y, beta=0.0):
        cf, cp = _dbls(coefs)
        px = (ctypes.c_void_p * len(xs))(*[_ptr(x) for x in xs])
        n = y.numel() if hasattr(y, "numel") else None
        _check(lib.amdg_lincomb(self._h, n, len(xs), cp, px, beta, _ptr(y)))

    def sweep1d_batch_mapped(self, op, rel, lu, t, sizes_from, srcs, dsts, coefs=None, accumulates=None, dst_maps=None, acc_froms=None, dst2s=None, dst2_maps=None):
        """amdg_sweep1d_batch_mapped / _dual; srcs/dsts/dst_maps/acc_froms/dst2s/dst2_maps are raw device addresses (ints) or torch tensors, None entries
        allowed in the optional ones"""
        n = len(srcs)
        if dst2s is not None and any(x is not None for x in dst2s):
            s, sp = _ints(np.asarray(sizes_from).reshape(n, self.dim))
            arr = lambda xs: (ctypes.c_void_p * n)(*[(_ptr(x) if x is not None else None) for x in (xs or [None] * n)])
            cf, cp = _dbls(np.ones(n) if coefs is None else coefs)
            ac, ap = _ints(np.zeros(n) if accumulates is None else accumulates)
            _check(lib.amdg_sweep1d_batch_dual(self._h, op, rel, lu, t, sp, arr(srcs), arr(dsts), cp, ap, arr(dst_maps), arr(acc_froms), arr(dst2s), arr(dst2_maps), n))
            return
        s, sp = _ints(np.asarray(sizes_from).reshape(n, self.dim))
        ps = (ctypes.c_void_p * n)(*[_ptr(x) for x in srcs])
        pd = (ctypes.c_void_p * n)(*[_ptr(x) for x in dsts])
        pm = (ctypes.c_void_p * n)(*[(_ptr(x) if x is not None else None) for x in (dst_maps or [None] * n)])
        pa = (ctypes.c_void_p * n)(*[(_ptr(x) if x is not None else None) for x in (acc_froms or [None] * n)])
        cf, cp = _dbls(np.ones(n) if coefs is None else coefs)
        ac, ap = _ints(np.zeros(n) if accumulates is None else accumulates)
        _check(lib.amdg_sweep1d_batch_mapped(self._h, op, rel, lu, t, sp, ps, pd, cp, ap, pm, pa, n))

    def points_set(self, pts1d):
        p, pp = _dbls(pts1d)
        _check(lib.amdg_points_set(self._h, pp))

    def pointwise_expr(self, ups, others, other_map, outs, prog, out_ptr, consts=()):
        """amdg_pointwise_expr: prog = [(op, arg), ...] for all outputs back to back, out_ptr[c] = first op of output c"""
        pu = (ctypes.c_void_p * max(len(ups), 1))(*[_ptr(x) for x in ups])
        po = (ctypes.c_void_p * max(len(others), 1))(*[_ptr(x) for x in others])
        pout = (ctypes.c_void_p * len(outs))(*[_ptr(x) for x in outs])
        pr, prp = _ints(np.asarray(prog, dtype=np.int32).reshape(-1, 2))
        op_, opp = _ints(out_ptr)
        cs, csp = _dbls(np.asarray(consts if len(consts) else [0.0]))
        _check(lib.amdg_pointwise_expr(self._h, len(ups), pu, len(others), po, _ptr(other_map) if other_map is not None else None, len(outs), pout,
                                       prp, pr.shape[0], opp, csp, len(consts)))

    def moment(self, field_map, n_vdim, order, weight, f, rhs_field):
        """amdg_moment: rhs_field[e][x, v = 0] += weight * velocity moment of f at the partner element field_map[e] (int32 device tensor, -1 = none)"""
        od, odp = _ints(order)
        _check(lib.amdg_moment(self._h, int(field_map.numel()), _ptr(field_map), n_vdim, odp, weight, _ptr(f), _ptr(rhs_field)))

    def indicator_norm(self, us, norm):
        """amdg_indicator_norm: norm[e] = sum_v ||us[v][e]||_2 (DGAdapt::indicator_norm)"""
        pu = (ctypes.c_void_p * len(us))(*[_ptr(x) for x in us])
        _check(lib.amdg_indicator_norm(self._h, len(us), pu, _ptr(norm)))

    def scatter_rows(self, src, n_rows, width, dst_base, dst_map):
        _check(lib.amdg_scatter_rows(self._h, _ptr(src), n_rows, width, _ptr(dst_base), _ptr(dst_map)))

    def dev_alloc(self, n_doubles):
        out = _p()
        _check(lib.amdg_dev_alloc(self._h, n_doubles, ctypes.byref(out)))
        return out.value

    def dev_free(self, ptr):
        _check(lib.amdg_dev_free(self._h, ctypes.c_void_p(ptr)))

    def peer_export(self, ptr):
        buf = ctypes.create_string_buffer(64)
        _check(lib.amdg_peer_export(self._h, ctypes.c_void_p(ptr), buf))
        return buf.raw

    def peer_open(self, handle):
        out = _p()
        _check(lib.amdg_peer_open(self._h, ctypes.c_char_p(handle), ctypes.byref(out)))
        return out.value

    def peer_close(self, ptr):
        _check(lib.amdg_peer_close(self._h, ctypes.c_void_p(ptr)))

    def peer_barrier(self, flag_ptrs, rank, epoch_ptr, error_ptr):
        fp = (ctypes.c_void_p * len(flag_ptrs))(*[ctypes.c_void_p(x) for x in flag_ptrs])
        _check(lib.amdg_peer_barrier(self._h, fp, len(flag_ptrs), rank, ctypes.c_void_p(epoch_ptr), ctypes.c_void_p(error_ptr)))

    def axpby(self, alpha, x, beta, y):
        _check(lib.amdg_axpby(self._h, y.numel(), alpha, _ptr(x), beta, _ptr(y)))

    # ---- host-buffer entry points (numpy in, numpy out; copies inside)
    def host_apply_tensor(self, ops, rels, src, edge_to, n_comp=1, coef=1.0, out=None):
        o, op = _ints(ops)
        r, rp = _ints(rels)
        s, sp = _dbls(src)
        acc = out is not None
        if out is None:
            out = np.empty(n_comp * self.n_elem * edge_to ** self.dim)
        _check(lib.amdg_host_apply_tensor(self._h, op, rp, sp, out.ctypes.data_as(_dp), n_comp, coef, int(acc)))
        return out

    def host_sweep1d(self, op, rel, lu, t, sizes_from, src, edge_to, n_comp=1, coef=1.0, out=None):
        z, zp = _ints(sizes_from)
        s, sp = _dbls(src)
        acc = out is not None
        if out is None:
            blk = int(np.prod(z)) // int(z[t]) * edge_to
            out = np.empty(n_comp * self.n_elem * blk)
        _check(lib.amdg_host_sweep1d(self._h, op, rel, lu, t, zp, sp, out.ctypes.data_as(_dp), n_comp, coef, int(acc)))
        return out

    def host_hierarchize(self, hier_op, src, n_comp=1):
        s, sp = _dbls(src)
        out = np.empty_like(s)
        _check(lib.amdg_host_hierarchize(self._h, hier_op, sp, out.ctypes.data_as(_dp), n_comp))
        return out

    def host_roundtrip(self, op_fwd, hier_op, op_inv, ucoe, n_comp=1, out=None):
        s, sp = _dbls(ucoe)
        if out is None:
            out = np.empty_like(s)
        _check(lib.amdg_host_roundtrip(self._h, op_fwd, hier_op, op_inv, sp, out.ctypes.data_as(_dp), n_comp))
        return out


def generate_tables(nmax, k, m, msh_case=1):
    """Every 1D table the path uses for Alpert degree k and Lagrange interpolation degree m, generated by the library in compact form
    (blocks per related 1D pair, amdg_pairs order) -- the replacement of a run of the reference's OperatorMatrix1D / LagrInterpolation
    constructors.  Keys as in the reference: "alpt.<table>", "lagr.<table>", "pt" / "pt_d1" (transposed Lag_pt_Alpt_1D / _d1), "hier"
    (I + W of the hierarchisation stencils), "lagr.intep_pt"."""
    ctx = Context(1, nmax, k, m, device=-1)
    a, b = k + 1, m + 1
    out = {}
    for t in ("u_v", "u_vx", "ulft_vjp", "urgt_vjp", "ujp_vjp", "ux_vx", "uxave_vjp", "ujp_vxave"):
        out["alpt." + t] = ctx.op_blocks(ctx.op_generate(BASIS_ALPERT, k, t), a, a)
    for t in ("u_v", "u_vx", "ulft_vjp", "urgt_vjp", "uave_vjp", "ujp_vxlft", "ujp_vxrgt"):
        out["lagr." + t] = ctx.op_blocks(ctx.op_generate(BASIS_LAGRANGE, m, t, msh_case), b, a)
    out["pt"] = ctx.op_blocks(ctx.op_generate_points(BASIS_LAGRANGE, m, msh_case, 0), a, b)
    out["pt_d1"] = ctx.op_blocks(ctx.op_generate_points(BASIS_LAGRANGE, m, msh_case, 1), a, b)
    out["hier"] = ctx.op_blocks(ctx.op_generate_hier(BASIS_LAGRANGE, m, msh_case), b, b)
    out["lagr.intep_pt"] = ctx.points_generate(BASIS_LAGRANGE, m, msh_case)
    ctx.close()
    return out
