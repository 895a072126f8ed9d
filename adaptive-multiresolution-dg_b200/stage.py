"""One nonlinear Runge-Kutta stage of a scalar conservation law on a (possibly fibre-partitioned) sparse grid, as ONE batched program.

The stage is the loop body of the reference's nonlinear examples (example/02_hyperbolic_05_burgers_adapt.cpp:223-470,
example/07_vlasov_ampere_02_2D2V_accuracy.cpp:247-303):

    FastLagrIntp::eval_up_Lagr -> point-wise flux -> eval_fp_to_coe_D_Lag -> HyperbolicLagrRHS::rhs_vol_scalar + rhs_flx_intp_scalar (all dims)
    -> HyperbolicAlptRHS::rhs_flx_penalty_scalar -> ExplicitRK::step_stage

Here it is planned once as a list of OPS (batched sweeps, row scatters, barriers, point-wise, linear combination, RK) over named buffers:

  * every tensor application uses the shared-prefix schedule (csrc/capi.cu: apply_tensor_shared), and the d right-hand-side applications advance
    in lockstep: the sweeps of one schedule level of ALL applications go out as one or two batched launches (amdg_sweep1d_batch_mapped);
  * on N > 1 GPUs the grid is partitioned by fibre ownership (dist.FibrePartition): layout X owns an element by its sub-index in the V dims
    (second half of the dims), so fibres along the X dims are local, layout V the converse.  A sweep whose output is consumed in the other
    layout stores every element block straight into the memory of the rank that owns it there (destination maps into CUDA-IPC-mapped peer
    memory, SweepJob::dst_map; accumulating sweeps read their old value locally, SweepJob::acc_from) -- the transfer of a layout switch
    happens in the epilogue of the sweeps that produce the data, tile by tile, over NVLink; buffers that no sweep produces at the right moment
    are moved by a row scatter.  A device-side barrier over peer-mapped flags (amdg_peer_barrier) separates producers from consumers:
    six barriers per stage, no host synchronisation, the whole stage is one CUDA graph per rank.
  * N == 1 runs the same program with both layouts equal (no maps, no scatters, no barriers).

The plan is pure index logic (numpy); `DeviceStage` executes it through the C ABI; tests/test_stage_plan.py executes it with the numpy oracle
for every rank in one process and compares with the single-grid oracle.
"""
import ctypes

import numpy as np

REL_VOL, REL_FLX = 0, 1
LU_L, LU_U, LU_FULL = 0, 1, 2


class Buf:
    __slots__ = ("name", "layout", "width")

    def __init__(self, name, layout, width):
        self.name, self.layout, self.width = name, layout, width


class StagePlan:
    """ops for one rank; `part` is a dist.FibrePartition (or None for a single GPU)"""

    def __init__(self, dim, a, b, n_flux, part=None, n_x_dims=None, fuse_rk=False, dual_store=True):
        self.dim, self.a, self.b, self.nf = dim, a, b, n_flux
        # dual_store (partitioned plans): a buffer that is consumed locally in layout X and, after the switch, in layout V (the outputs of the early
        # down-pass levels, the hierarchised fluxes) is stored to both places by the sweep that produces it (SweepJob::dst2) instead of being moved by a
        # row scatter before the switch; twin[x buffer] = its copy in layout V
        self.dual_store = dual_store
        self.twin = {}
        # fuse_rk: the Runge-Kutta combination u_new = a u_tn + b u + c dt rhs (ExplicitRK::step_stage, source/ODESolver.cpp:209-301) rides in the
        # epilogues of the sweeps that produce the right-hand side: the accumulator starts as a u_tn + b u, every right-hand-side chain carries the
        # factor c dt and its last sweeps accumulate straight into it -- no rhs array, no per-application sums, no separate RK pass
        self.fuse_rk = fuse_rk
        self.dist = part is not None and part.world > 1
        self.part = part
        self.h = (dim // 2 if n_x_dims is None else n_x_dims) if self.dist else dim      # dims < h are swept in layout X, the rest in layout V
        self.bufs, self.alias, self.ops = {}, {}, []
        self.n_barrier = 0
        self.n_group = 0
        self.max_jobs = 32           # jobs per launch (MAX_JOBS of the kernels)
        self.push_bytes = 0          # doubles per element pushed / scattered to the other layout (exchange volume)
        self._build()

    # ---- buffers
    def buf(self, name, layout, width):
        if not self.dist:
            layout = "X"
        key = self.alias.get(name, name)
        if key not in self.bufs:
            self.bufs[key] = Buf(key, layout, width)
        bb = self.bufs[key]
        assert bb.width == width and bb.layout == layout, (name, bb.layout, layout, bb.width, width)
        return key

    def other(self, layout):
        return "V" if layout == "X" else "X"

    def layout_of(self, k):
        return "X" if (k < self.h or not self.dist) else "V"

    def barrier(self):
        if self.dist:
            self.ops.append(("barrier",))
            self.n_barrier += 1

    def move(self, src, dst_name, dst_layout):
        """buffer `src` is needed under the name dst_name in dst_layout: a row scatter to the peers, or an alias on one GPU"""
        w = self.bufs[src].width
        if not self.dist:
            self.alias[dst_name] = src
            return src
        d = self.buf(dst_name, dst_layout, w)
        self.ops.append(("scatter", src, d))
        self.push_bytes += w
        return d

    # ---- batched tensor applications (shared-prefix schedule)
    def apply_batch(self, tag, apps, kf, kt, final_layout="X", acc_into=None):
        """apps: list of dict(src=buffer in layout X, ops=[name per dim], rels=[per dim], coef).  Returns the result buffer of every application,
        living in `final_layout` (X: the natural end of the schedule; V: the last sweeps push it).  acc_into (single GPU): an existing buffer the
        result is ADDED to -- the chain of accumulating sweeps of the up pass ends in the output of the full sweep of the all-L chain, so that one
        job accumulates into acc_into instead of writing a scratch buffer and the sum over applications needs no extra pass."""
        d, h = self.dim, self.h
        assert acc_into is None or not self.dist
        edge = lambda S, k: kt if (S >> k) & 1 else kf
        width = lambda sizes: int(np.prod(sizes))
        if d == 1:
            outs = []
            for i, ap in enumerate(apps):
                o = self.buf("%s.res%d" % (tag, i), "X", kt)
                if acc_into is not None:
                    o = acc_into
                self.ops.append(("sweep", "X", ap["ops"][0], ap["rels"][0], LU_FULL, 0, [dict(sizes=[kf], src=ap["src"], dst=o, coef=ap.get("coef", 1.0), acc=acc_into is not None)]))
                outs.append(o)
            return outs
        X = [{0: ap["src"]} for ap in apps]

        def emit(layout, k, lu, jobs_by_app):
            # the launches of one schedule level are independent of each other (they read and write different buffers): they carry a common
            # group id, and DeviceStage runs a group on parallel streams
            groups = {}
            for i, jobs in enumerate(jobs_by_app):
                for j in jobs:
                    groups.setdefault((apps[i]["ops"][k], apps[i]["rels"][k]), []).append(j)
            self.n_group += 1
            for (opn, rel), jobs in groups.items():
                for c0 in range(0, len(jobs), self.max_jobs):
                    self.ops.append(("sweep", layout, opn, rel, lu, k, jobs[c0:c0 + self.max_jobs], self.n_group))

        def switch_down(k_next):
            """all live X buffers of the down pass move from layout X to layout V before level k_next"""
            for i in range(len(apps)):
                for S in sorted(X[i]):
                    name = "%s.x%d.%d@V" % (tag, i, S)
                    if self.bufs[X[i][S]].layout == "X" and self.dist:
                        X[i][S] = self.twin[X[i][S]] if X[i][S] in self.twin else self.move(X[i][S], name, "V")
            self.barrier()

        # down pass: L_k applied to X_S for every S subset of {0..k-1}
        for k in range(d - 1):
            lay = self.layout_of(k)
            if self.dist and k == h:
                switch_down(k)
            push = self.dist and k == h - 1                           # the outputs of the last X level are consumed in layout V
            jobs_by_app = []
            for i in range(len(apps)):
                jobs = []
                for S in sorted(s for s in X[i] if s < (1 << k)):
                    sizes = [edge(S, q) if q < k else kf for q in range(d)]
                    osz = list(sizes); osz[k] = kt
                    dl = self.other(lay) if push else lay
                    dst = self.buf("%s.x%d.%d%s" % (tag, i, S | (1 << k), "@V" if (push or lay == "V") and self.dist else ""), dl, width(osz))
                    job = dict(sizes=sizes, src=X[i][S], dst=dst, coef=1.0, acc=False, push=push)
                    if self.dist and self.dual_store and lay == "X" and not push and k < h:
                        # consumed by the later X levels here and by the V levels after the switch: stored to both places
                        job["dst2"] = self.buf("%s.x%d.%d@V" % (tag, i, S | (1 << k)), "V", width(osz))
                        self.twin[dst] = job["dst2"]
                        self.push_bytes += width(osz)
                    jobs.append(job)
                    if push:
                        self.push_bytes += width(osz)
                    X[i][S | (1 << k)] = dst
                jobs_by_app.append(jobs)
            emit(lay, k, LU_L, jobs_by_app)
        if self.dist and h == d - 1:
            switch_down(d - 1)
        # full sweep along d-1
        lay = self.layout_of(d - 1)
        R = [dict() for _ in apps]
        jobs_by_app = []
        to_x_after_full = self.dist and lay == "V" and h == d - 1     # no U level in layout V: the full sweep itself pushes
        for i, ap in enumerate(apps):
            jobs = []
            for S in sorted(X[i]):
                sizes = [edge(S, q) for q in range(d - 1)] + [kf]
                osz = sizes[:-1] + [kt]
                dl = "X" if to_x_after_full else lay
                into = acc_into is not None and S == (1 << (d - 1)) - 1
                dst = acc_into if into else self.buf("%s.y%d.%d%s" % (tag, i, S, "@V" if dl == "V" and self.dist else ""), dl, width(osz))
                jobs.append(dict(sizes=sizes, src=X[i][S], dst=dst, coef=ap.get("coef", 1.0), acc=into, push=to_x_after_full))
                if to_x_after_full:
                    self.push_bytes += width(osz)
                R[i][S] = dst
            jobs_by_app.append(jobs)
        emit(lay, d - 1, LU_FULL, jobs_by_app)
        if to_x_after_full:
            self.barrier()
        # up pass: R_k(S) = U_k R_{k+1}(S) + R_{k+1}(S + {k}), accumulated into the buffer of S + {k}
        for k in range(d - 2, -1, -1):
            lay = self.layout_of(k)
            to_x = self.dist and k == h and h >= 1                   # last V level: results are consumed in layout X
            to_v_final = self.dist and k == 0 and final_layout == "V"
            jobs_by_app = []
            for i in range(len(apps)):
                jobs, newR = [], {}
                for S in sorted(s for s in R[i] if not (s >> k) & 1 and s < (1 << k)):
                    sizes = [edge(S, q) if q <= k else kt for q in range(d)]
                    osz = list(sizes); osz[k] = kt
                    hi = R[i][S | (1 << k)]
                    if to_x or to_v_final:
                        dl = "X" if to_x else "V"
                        dst = self.buf("%s.r%d.%d.%d@%s" % (tag, i, k, S, dl), dl, width(osz))
                        jobs.append(dict(sizes=sizes, src=R[i][S], dst=dst, coef=1.0, acc=True, push=True, acc_from=hi))
                        self.push_bytes += width(osz)
                        newR[S] = dst
                    else:
                        jobs.append(dict(sizes=sizes, src=R[i][S], dst=hi, coef=1.0, acc=True, push=False))
                        newR[S] = hi
                R[i] = newR
                jobs_by_app.append(jobs)
            emit(lay, k, LU_U, jobs_by_app)
            if to_x or to_v_final:
                self.barrier()
        return [R[i][0] for i in range(len(apps))]

    # ---- the stage
    def _build(self):
        d, a, b, nf = self.dim, self.a, self.b, self.nf
        A, B = a ** d, b ** d
        u = self.buf("u", "X", A)
        self.buf("u_tn", "X", A)
        fuse = self.fuse_rk
        acc = None
        if fuse and not self.dist:
            acc = self.buf("u_new", "X", A)
            self.ops.append(("lincomb", acc, ["u_tn", u], 0.0, ["rk_a", "rk_b"]))
        # penalty in the V dims needs u in layout V (rides on the first barrier of the interpolation)
        u_v = self.move(u, "u@V", "V") if self.dist else u
        if self.dist and self.dual_store:
            self.twin[u] = u_v                      # the interpolation's X_0 in layout V is this copy: no second scatter
        # 1. Alpert coefficients -> point values; the last sweeps deliver them in layout V (where the point-wise products and the V-dim
        #    hierarchisation run)
        pw_layout = "V" if self.dist else "X"
        up = self.apply_batch("intp", [dict(src=u, ops=["pt"] * d, rels=[REL_VOL] * d)], a, b, final_layout=pw_layout)[0]
        # penalty sweeps of the V dims (u@V arrived with the first barrier): pen_v = sum_t coef * P_t u, last sweep pushes it to layout X
        pen_parts = []
        v_dims = [t for t in range(d) if self.layout_of(t) == "V"] if self.dist else []
        x_dims = [t for t in range(d) if t not in v_dims]
        if d > 1:
            if v_dims:
                pv = self.buf("pen@V", "V", A)
                px = self.buf("pen.fromV", "X", A)
                for n_, t in enumerate(v_dims):
                    last = n_ == len(v_dims) - 1
                    job = dict(sizes=[a] * d, src=u_v, dst=px if last else pv, coef="pen", acc=n_ > 0, push=last)
                    if last and n_ > 0:
                        job["acc_from"] = pv
                    self.ops.append(("sweep", "V", "pen", REL_FLX, LU_FULL, t, [job]))
                self.push_bytes += A
                pen_parts.append(px)
        # 2. point-wise flux
        fp = [self.buf("fp%d" % c, pw_layout, B) for c in range(nf)]
        self.ops.append(("pointwise", pw_layout, up, fp))
        # 3. hierarchisation: the V dims first (layout V), then the X dims; one job per flux component, ping-pong buffers
        cur = fp
        order = (v_dims + x_dims) if self.dist else list(range(d))
        for n_, t in enumerate(order):
            lay = self.layout_of(t)
            push = self.dist and lay == "V" and (n_ + 1 == len(order) or self.layout_of(order[n_ + 1]) == "X")
            dl = "X" if push else lay
            nxt = [self.buf("h%d.%d%s" % (n_ % 2, c, "@X" if dl == "X" and self.dist else ""), dl, B) for c in range(nf)]
            sizes = [b] * d
            hjobs = [dict(sizes=sizes, src=cur[c], dst=nxt[c], coef=1.0, acc=False, push=push) for c in range(nf)]
            if self.dist and self.dual_store and n_ + 1 == len(order) and dl == "X" and self.h < d:
                # the hierarchised fluxes are the X_0 of the right-hand-side applications: needed in layout X (first L sweeps) and in layout V (full sweep)
                for c in range(nf):
                    hjobs[c]["dst2"] = self.buf("rhs.x%d.0@V" % c, "V", B)
                    self.twin[nxt[c]] = hjobs[c]["dst2"]
                    self.push_bytes += B
            self.ops.append(("sweep", lay, "hier", REL_VOL, LU_U, t, hjobs))
            if push:
                self.push_bytes += nf * B
                self.barrier()
            cur = nxt
        fuc = cur
        # 4. right-hand side: rhs_vol + rhs_flx of dimension t as one application (u_vx + (ulft_vjp + urgt_vjp)/2 under the flx relation in dim t)
        apps = [dict(src=fuc[t], ops=["volflx" if s == t else "uv" for s in range(d)], rels=[REL_FLX if s == t else REL_VOL for s in range(d)]) for t in range(nf)]
        if acc is not None:
            # single GPU, fused: every application accumulates into the RK accumulator, the penalty sweeps as well; nothing is left to combine
            for t in range(nf):
                self.apply_batch("rhs", [dict(apps[t], coef="rk_c")], b, a, final_layout="X", acc_into=acc)
            if d > 1:
                for t in x_dims:
                    self.ops.append(("sweep", "X", "pen", REL_FLX, LU_FULL, t, [dict(sizes=[a] * d, src=u, dst=acc, coef="pen*rk_c", acc=True, push=False)]))
            self.result, self.rhs, self.up, self.fuc = acc, None, up, fuc
            return
        rhs = self.buf("rhs", "X", A)
        first = True
        if self.dist:
            # all applications in lockstep: two barriers for the whole right-hand side
            res = self.apply_batch("rhs", apps, b, a, final_layout="X")
        else:
            # one GPU: one application at a time over the same scratch buffers (the working set of a schedule level stays near the L2 size),
            # each result added to rhs before the next application overwrites it
            res = []
            for t in range(nf):
                r1 = self.apply_batch("rhs", [apps[t]], b, a, final_layout="X")
                self.ops.append(("lincomb", rhs, r1, 0.0 if first else 1.0))
                first = False
        # 5. penalty sweeps of the X dims, all parts joined, RK stage
        if d > 1:
            px2 = self.buf("pen@X", "X", A)
            for n_, t in enumerate(x_dims):
                self.ops.append(("sweep", "X", "pen", REL_FLX, LU_FULL, t, [dict(sizes=[a] * d, src=u, dst=px2, coef="pen", acc=n_ > 0, push=False)]))
            if x_dims:
                pen_parts.append(px2)
        if fuse:
            # partitioned, fused: the pushed partial results, the penalty parts and the RK combination in ONE pass (in place: element-wise)
            parts = res + pen_parts
            self.ops.append(("lincomb", u, ["u_tn", u] + parts, 0.0, ["rk_a", "rk_b"] + ["rk_c"] * len(parts)))
            self.result, self.rhs, self.up, self.fuc = u, None, up, fuc
            return
        if res + pen_parts:
            self.ops.append(("lincomb", rhs, res + pen_parts, 0.0 if first else 1.0))
        self.ops.append(("rk", "u_tn", u, rhs))
        self.result, self.rhs, self.up, self.fuc = u, rhs, up, fuc

    # ---- summaries
    def launches(self):
        return sum(1 for o in self.ops if o[0] != "barrier") + self.n_barrier


class SlabLayout:
    """Where every buffer of a plan lives in every rank's slab (pure index logic, identical on all ranks), and the destination maps of the
    pushed buffers given the addresses at which this rank sees the slabs."""

    def __init__(self, plan, n_elem):
        self.plan = plan
        part = plan.part
        self.world = part.world if plan.dist else 1
        self.rank = part.rank if plan.dist else 0
        layouts = ("X", "V") if plan.dist else ("X",)
        self.n_loc = {L: ([len(part.local_of(L, r)) for r in range(self.world)] if plan.dist else [n_elem]) for L in layouts}
        names = list(plan.bufs)
        self.index = {nm: i for i, nm in enumerate(names)}
        self.off = np.zeros((self.world, len(names)), dtype=np.int64)          # doubles, multiples of 32 (256 bytes)
        tot = np.zeros(self.world, dtype=np.int64)
        for i, nm in enumerate(names):
            bb = plan.bufs[nm]
            for r in range(self.world):
                self.off[r, i] = tot[r]
                tot[r] += (self.n_loc[bb.layout][r] * bb.width + 31) // 32 * 32
        self.total = tot                                                         # doubles per rank (the barrier flags follow)

    def offset(self, name, r=None):
        return int(self.off[self.rank if r is None else r, self.index[self.plan.alias.get(name, name)]])

    def maps(self, base):
        """base[r] = address (bytes) at which this rank sees rank r's slab.  Returns {(dst buffer, source layout): int64[n_local(source layout)]}:
        offset in doubles, relative to this rank's copy of dst, of the block of every local row in the copy of the rank that owns the element in
        dst's layout; and the bytes stored into other ranks' memory per stage"""
        plan, part = self.plan, self.plan.part
        need = set()
        for o in plan.ops:
            if o[0] == "sweep":
                for j in o[6]:
                    if j.get("push"):
                        need.add((j["dst"], o[1]))
                    if j.get("dst2"):
                        need.add((j["dst2"], o[1]))
            elif o[0] == "scatter":
                need.add((o[2], plan.bufs[o[1]].layout))
        out, sent = {}, 0
        for (dst, src_layout) in sorted(need):
            bb = plan.bufs[dst]
            i = self.index[dst]
            mine = part.local[src_layout]
            owner = part.owner[bb.layout][mine]
            row = part.row_in[bb.layout][mine]
            rel = np.array([(int(base[r]) - int(base[self.rank])) // 8 + self.off[r, i] - self.off[self.rank, i] for r in range(self.world)], dtype=np.int64)
            out[(dst, src_layout)] = rel[owner] + row.astype(np.int64) * bb.width
            sent += int((owner != self.rank).sum()) * bb.width * 8
        return out, sent


class DeviceStage:
    """Executes a StagePlan through the C ABI on this rank's GPU.  `make_ops(ctx)` registers the operators on a context and returns
    {"pt", "uv", "volflx", "pen", "hier"} -> handle; `exchange(obj)` all-gathers a python object over the ranks (torch.distributed.all_gather_object)."""

    def __init__(self, amdg, plan, level, suppt, nmax, k, m, device, make_ops, pointwise, pen_coef, rk, exchange=None, stream_ptr=None, kernel=0):
        self.A, self.plan = amdg, plan
        part = plan.part
        self.world = part.world if plan.dist else 1
        self.rank = part.rank if plan.dist else 0
        self.pointwise, self.pen_coef, self.rk = pointwise, pen_coef, rk
        import os
        self.n_side, self._side = int(os.environ.get("AMDG_STAGE_STREAMS", "3")), None      # parallel streams for the launches of one schedule level
        self.ctx, self.ops, self.rows = {}, {}, {}
        layouts = ("X", "V") if plan.dist else ("X",)
        for L in layouts:
            rows = part.local[L] if plan.dist else np.arange(level.shape[0])
            c = amdg.Context(plan.dim, nmax, k, m, device=device)
            if stream_ptr is not None:
                c.set_stream(stream_ptr)
            c.set_kernel(kernel)
            if len(rows):
                c.grid_set(level[rows], suppt[rows])
            self.ctx[L], self.rows[L] = c, rows
            self.ops[L] = make_ops(c) if len(rows) else {}
        c0 = self.ctx["X"]
        # ---- slab: every buffer of the plan at a 256-byte aligned offset; the same walk gives the offsets on every rank
        self.layout = SlabLayout(plan, level.shape[0])
        tot = self.layout.total
        flag_doubles = 64
        self.slab_doubles = int(tot[self.rank]) + flag_doubles
        self.slab = c0.dev_alloc(self.slab_doubles)
        amdg.lib.amdg_dev_zero(c0._h, ctypes.c_void_p(self.slab), self.slab_doubles)
        c0.sync()
        self.base = [self.slab]
        self.sent_bytes = 0
        self.flag_ofs = [int(tot[r]) for r in range(self.world)]
        self.maps = {}
        if plan.dist:
            import torch
            handles = exchange(c0.peer_export(self.slab))
            self.base = [self.slab if r == self.rank else c0.peer_open(handles[r]) for r in range(self.world)]
            mp, self.sent_bytes = self.layout.maps(self.base)
            self.maps = {k: torch.from_numpy(v).cuda() for k, v in mp.items()}
        # barrier state: flags (unsigned[world]) after the buffers of every slab, epoch and error words on this device
        self.flags = [self.base[r] + 8 * self.flag_ofs[r] for r in range(self.world)]
        self.epoch = self.slab + 8 * self.flag_ofs[self.rank] + 256
        self.error = self.epoch + 8

    def local_ptr(self, name):
        return self.slab + 8 * self.layout.offset(name)

    def swap_result(self):
        """single-GPU fused plans leave the stage's result in the RK accumulator ("u_new"), not in "u": exchange the slab offsets of the two buffers so
        that the next run() of the same (eager) program reads the new state -- a pointer swap, no copy.  CUDA graphs bake addresses in: a time
        loop that replays graphs captures one graph per parity instead."""
        r = self.plan.alias.get(self.plan.result, self.plan.result)
        if r == "u":
            return
        i, j = self.layout.index["u"], self.layout.index[r]
        self.layout.off[:, [i, j]] = self.layout.off[:, [j, i]]

    def close(self):
        c0 = self.ctx["X"]
        c0.sync()
        for r in range(self.world):
            if r != self.rank:
                c0.peer_close(self.base[r])
        c0.dev_free(self.slab)
        for c in self.ctx.values():
            c.close()

    def view(self, name):
        """torch view of a local buffer (tests, initial data)"""
        import torch
        key = self.plan.alias.get(name, name)
        bb = self.plan.bufs[key]
        n = len(self.rows[bb.layout])
        return _as_tensor(self.local_ptr(key), n * bb.width).view(n, bb.width)

    def run(self, profile=None):
        """executes the plan on the context's stream; `profile` (a dict) collects device time per kind of operation through CUDA events
        (eager launches only: the caller synchronises and calls profile_summary)"""
        A, plan = self.A, self.plan
        ops = plan.ops
        if profile is None and self.n_side > 0:
            import torch
            main = torch.cuda.current_stream()
            if self._side is None:
                self._side = [torch.cuda.Stream() for _ in range(self.n_side)]
            i = 0
            while i < len(ops):
                o = ops[i]
                j = i + 1
                if o[0] == "sweep" and len(o) > 7:
                    while j < len(ops) and ops[j][0] == "sweep" and len(ops[j]) > 7 and ops[j][7] == o[7]:
                        j += 1
                if j - i == 1:
                    self._run_op(o)
                else:
                    # fork: the launches of one schedule level on parallel streams; join before the next level
                    fork = torch.cuda.Event()
                    fork.record(main)
                    used = []
                    for n, op in enumerate(ops[i:j]):
                        st = main if n % (self.n_side + 1) == 0 else self._side[n % (self.n_side + 1) - 1]
                        if st is not main and st not in used:
                            st.wait_event(fork)
                            used.append(st)
                        self.ctx[op[1]].set_stream(st.cuda_stream)
                        self._run_op(op)
                        self.ctx[op[1]].set_stream(main.cuda_stream)
                    for st in used:
                        e = torch.cuda.Event()
                        e.record(st)
                        main.wait_event(e)
                i = j
            return
        for o in ops:
            kind = o[0]
            if profile is not None:
                import torch
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                key = kind if kind != "sweep" else "sweep:%s" % o[2]
                profile.setdefault("_events", []).append((key, e0, e1))
                e0.record()
                self._run_op(o)
                e1.record()
            else:
                self._run_op(o)

    @staticmethod
    def profile_summary(profile):
        """milliseconds per kind of operation from the events of run(profile) (call after a synchronise)"""
        out = {}
        for key, e0, e1 in profile.pop("_events", []):
            out[key] = out.get(key, 0.0) + e0.elapsed_time(e1)
        return out

    def _run_op(self, o):
        A, plan = self.A, self.plan
        if True:
            kind = o[0]
            if kind == "sweep":
                _, lay, opn, rel, lu, t, jobs = o[:7]
                c = self.ctx[lay]
                if not len(self.rows[lay]):
                    return
                srcs = [self.local_ptr(j["src"]) for j in jobs]
                dsts = [self.local_ptr(j["dst"]) for j in jobs]
                coefs = [self.coef(j["coef"]) for j in jobs]
                maps = [self.maps[(j["dst"], lay)] if (plan.dist and j.get("push")) else None for j in jobs]
                accf = [self.local_ptr(j["acc_from"]) if j.get("acc_from") else None for j in jobs]
                d2 = [self.local_ptr(j["dst2"]) if j.get("dst2") else None for j in jobs]
                m2 = [self.maps[(j["dst2"], lay)] if j.get("dst2") else None for j in jobs]
                c.sweep1d_batch_mapped(self.ops[lay][opn], rel, lu, t, [j["sizes"] for j in jobs], srcs, dsts, coefs=coefs,
                                       accumulates=[int(j["acc"]) for j in jobs], dst_maps=maps, acc_froms=accf, dst2s=d2, dst2_maps=m2)
            elif kind == "scatter":
                _, src, dst = o
                lay = plan.bufs[src].layout
                n = len(self.rows[lay])
                if n:
                    self.ctx[lay].scatter_rows(self.local_ptr(src), n, plan.bufs[src].width, self.local_ptr(dst), self.maps[(dst, lay)])
            elif kind == "barrier":
                self.ctx["X"].peer_barrier(self.flags, self.rank, self.epoch, self.error)
            elif kind == "pointwise":
                _, lay, up, fps = o
                if len(self.rows[lay]):
                    self.pointwise(self.ctx[lay], self.local_ptr(up), [self.local_ptr(f) for f in fps])
            elif kind == "lincomb":
                _, dst, parts, beta = o[:4]
                n = len(self.rows["X"]) * plan.bufs[plan.alias.get(dst, dst)].width
                if n:
                    c = self.ctx["X"]
                    cf = np.ones(len(parts)) if len(o) < 5 else np.array([self.coef(x) for x in o[4]], dtype=np.float64)
                    px = (ctypes.c_void_p * len(parts))(*[ctypes.c_void_p(self.local_ptr(p)) for p in parts])
                    A._check(A.lib.amdg_lincomb(c._h, n, len(parts), cf.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), px, beta, ctypes.c_void_p(self.local_ptr(dst))))
            elif kind == "rk":
                _, u_tn, u, rhs = o
                n = len(self.rows["X"]) * plan.bufs[u].width
                if n:
                    scheme, stage, dt = self.rk
                    A._check(A.lib.amdg_rk_stage(self.ctx["X"]._h, scheme, stage, dt, ctypes.c_void_p(self.local_ptr(u_tn)), ctypes.c_void_p(self.local_ptr(u)),
                                                 ctypes.c_void_p(self.local_ptr(rhs)), n))

    def coef(self, c):
        """numbers, or the symbols of the plan: "pen" (penalty coefficient), "rk_a" / "rk_b" / "rk_c" (u_new = rk_a u_tn + rk_b u + rk_c rhs), products "x*y"""
        if isinstance(c, str):
            v = 1.0
            for f in c.split("*"):
                v *= self.pen_coef if f == "pen" else self._rk_abc()[("rk_a", "rk_b", "rk_c").index(f)]
            return v
        return c

    def _rk_abc(self):
        return rk_coefficients(*self.rk)

    def launch_count(self):
        return sum(c.launch_count for c in self.ctx.values())

    def barrier_error(self):
        """non-zero when a barrier gave up waiting for a peer"""
        out = np.zeros(1)
        c = self.ctx["X"]
        self.A._check(self.A.lib.amdg_dev_download(c._h, out.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), ctypes.c_void_p(self.error), 1))
        return int(out.view(np.uint32)[0])


def rk_coefficients(scheme, stage, dt):
    """(a, b, c) of u_new = a u_tn + b u + c rhs for ExplicitRK::step_stage (source/ODESolver.cpp:209-330), as amdg_rk_stage applies it
    (csrc/kernels.cu: launch_rk_stage).  Schemes: 0 ForwardEuler, 1 RK2SSP, 2 RK2Midpoint, 3 RK3SSP, 4 RK3HeunLinear."""
    c_tn, c_u, c_rhs = 1.0, 0.0, dt
    if scheme == 1 and stage == 1:
        c_tn, c_u = 0.5, 0.5
    elif scheme == 2 and stage == 0:
        c_rhs = 0.5 * dt
    elif scheme == 3 and stage == 1:
        c_tn, c_u = 3.0 / 4.0, 1.0 / 4.0
    elif scheme == 3 and stage == 2:
        c_tn, c_u = 1.0 / 3.0, 2.0 / 3.0
    elif scheme == 4 and stage in (0, 1):
        c_rhs = (1.0 / 3.0 if stage == 0 else 1.0 / 2.0) * dt
    return (c_tn, c_u, c_u * c_rhs) if c_u != 0.0 else (c_tn, 0.0, c_rhs)


def _as_tensor(ptr, n_doubles):
    """torch float64 tensor over raw device memory (plumbing only)"""
    import torch

    class _Holder:
        pass
    h = _Holder()
    h.__cuda_array_interface__ = {"shape": (int(n_doubles),), "typestr": "<f8", "data": (int(ptr), False), "version": 2}
    return torch.as_tensor(h, device="cuda")
