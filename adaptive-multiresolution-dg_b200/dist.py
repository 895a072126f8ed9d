"""Fibre-partitioned multi-GPU form of the tensor application (one process per GPU, torch.distributed for plumbing).

A sweep along dimension t couples only elements that agree in (level, suppt) in every other dimension
(reference source/Element.cpp:265-299), so fibres along t are independent.  Two layouts:

  layout X: an element is owned by the rank chosen for its sub-index in the V dims (second half of the dims)
            -> every fibre along an X dim (first half) is complete on one rank;
  layout V: owned by the sub-index in the X dims -> fibres along V dims are local.

The shared-prefix schedule (csrc/capi.cu) applies, in this order, the L sweeps of dims 0..d-2,
the full sweep of dim d-1 and the U sweeps of dims d-2..0.  With X dims = 0..h-1 and V dims = h..d-1 that is
[X-local] -> switch -> [V-local] -> switch -> [X-local]: exactly two layout switches per tensor application, each an
all-to-all of element blocks (NCCL over NVLink on GPUs; gloo in the CPU tests of the index logic).
"""
import numpy as np
import torch
import torch.distributed as dist


def _group_owner(level, suppt, dims_key, world):
    """owner rank of every element such that elements equal in (level, suppt) on dims_key share the owner;
    groups are spread by longest-processing-time-first on the group size"""
    key = np.concatenate([level[:, dims_key], suppt[:, dims_key]], axis=1)
    _, inv, counts = np.unique(key, axis=0, return_inverse=True, return_counts=True)
    inv = inv.reshape(-1)
    order = np.argsort(-counts, kind="stable")
    load = np.zeros(world, dtype=np.int64)
    owner_of_group = np.zeros(len(counts), dtype=np.int64)
    for g in order:
        r = int(np.argmin(load))
        owner_of_group[g] = r
        load[r] += counts[g]
    return owner_of_group[inv]


class FibrePartition:
    """Ownership tables for the two layouts and the exchange plan between them (pure index logic, no device code)."""

    def __init__(self, level, suppt, world, rank, n_x_dims=None):
        level = np.asarray(level)
        suppt = np.asarray(suppt)
        self.n, self.dim = level.shape
        self.world, self.rank = world, rank
        self.h = self.dim // 2 if n_x_dims is None else n_x_dims
        self.dims_x = list(range(self.h))
        self.dims_v = list(range(self.h, self.dim))
        self.owner = {"X": _group_owner(level, suppt, self.dims_v, world), "V": _group_owner(level, suppt, self.dims_x, world)}
        # local element lists (ascending global id) per layout
        self.local = {k: np.nonzero(self.owner[k] == rank)[0] for k in ("X", "V")}
        # row of every element in the local list of the rank that owns it, per layout (ascending global id inside a rank)
        self.row_in = {}
        for k in ("X", "V"):
            ri = np.zeros(self.n, dtype=np.int64)
            for r in range(world):
                idx = np.nonzero(self.owner[k] == r)[0]
                ri[idx] = np.arange(len(idx))
            self.row_in[k] = ri
        self.level, self.suppt = level, suppt

    def local_of(self, layout, r):
        """global ids (ascending) of the elements rank r owns in `layout`"""
        return np.nonzero(self.owner[layout] == r)[0]

    def plan(self, src, dst):
        """exchange plan src layout -> dst layout for this rank: (send_index [local src rows, grouped by destination rank],
        send_counts, recv_counts, recv_index [position in the local dst list of every received row])"""
        mine = self.local[src]
        dest = self.owner[dst][mine]
        send_order = np.argsort(dest, kind="stable")                      # by destination, ascending global id inside
        send_counts = np.bincount(dest, minlength=self.world)
        theirs = self.local[dst]                                           # rows I own after the switch
        srcrank = self.owner[src][theirs]
        recv_order = np.argsort(srcrank, kind="stable")                    # arrival order: by source rank, ascending global id
        recv_counts = np.bincount(srcrank, minlength=self.world)
        recv_index = np.empty(len(theirs), dtype=np.int64)
        recv_index[np.arange(len(theirs))] = recv_order                    # received row i goes to local position recv_order[i]
        return send_order, send_counts, recv_counts, recv_order

    def _device_plan(self, src, dst, dev):
        """the exchange plan as device index tensors and python count lists (cached: the layout switches of every application reuse it)"""
        key = (src, dst, str(dev))
        if not hasattr(self, "_plans"):
            self._plans = {}
        if key not in self._plans:
            send_order, send_counts, recv_counts, recv_order = self.plan(src, dst)
            self._plans[key] = (torch.as_tensor(send_order, device=dev), [int(c) for c in send_counts], [int(c) for c in recv_counts],
                                torch.as_tensor(recv_order, device=dev))
        return self._plans[key]

    def switch(self, x, src, dst, group=None):
        """x: [n_local(src), block] tensor in layout src -> [n_local(dst), block] in layout dst"""
        return self.switch_many([x], src, dst, group)[0]

    def switch_many(self, xs, src, dst, group=None):
        """all tensors of xs ([n_local(src), block_i]) change layout in ONE all-to-all: element rows are concatenated along the block axis"""
        dev = xs[0].device
        send_idx, send_counts, recv_counts, recv_idx = self._device_plan(src, dst, dev)
        widths = [int(x.shape[1]) for x in xs]
        cat = xs[0] if len(xs) == 1 else torch.cat(xs, dim=1)
        sbuf = cat.index_select(0, send_idx)
        rbuf = torch.empty(sum(recv_counts), cat.shape[1], dtype=cat.dtype, device=dev)
        if self.world == 1:
            rbuf.copy_(sbuf)
        else:
            dist.all_to_all_single(rbuf, sbuf, recv_counts, send_counts, group=group)
        out = torch.empty_like(rbuf)
        out.index_copy_(0, recv_idx, rbuf)
        return [o.contiguous() for o in torch.split(out, widths, dim=1)] if len(xs) > 1 else [out]


class DistTensorApply:
    """amdg_apply_tensor over a fibre-partitioned grid.  `make_ctx(level, suppt)` builds a Context on the local device for
    a local element list; operators are registered per layout through `register(ctx)` -> dict name -> handle."""

    def __init__(self, amdg, part, dim, nmax, pmax_alpt, pmax_intp, device, register):
        self.A, self.part, self.dim = amdg, part, dim
        self.ctx, self.ops = {}, {}
        for k in ("X", "V"):
            rows = part.local[k]
            c = amdg.Context(dim, nmax, pmax_alpt, pmax_intp, device=device)
            c.set_stream(torch.cuda.current_stream().cuda_stream)
            # a layout with fewer ownership groups than ranks leaves some rank without elements: it keeps the context for the plumbing only
            # (amdg_grid_set rejects an empty grid) and still takes part in every all-to-all with zero rows
            if len(rows):
                c.grid_set(part.level[rows], part.suppt[rows])
            self.ctx[k] = c
            self.ops[k] = register(c) if len(rows) else {}
        self.switches = 0
        self.switch_bytes = 0

    def close(self):
        for c in self.ctx.values():
            c.close()

    def _switch_all(self, bufs, src, dst):
        """one all-to-all for all live buffers of the schedule"""
        keys = list(bufs.keys())
        outs = self.part.switch_many([bufs[S] for S in keys], src, dst)
        self.switches += 1
        self.switch_bytes += sum(bufs[S].numel() for S in keys) * 8
        return dict(zip(keys, outs))

    def apply(self, op_names, rels, src_x, kf, kt, coef=1.0):
        """src_x: [n_local(X), kf^dim] in layout X -> returns [n_local(X), kt^dim] in layout X.
        op_names[t]: operator name of dimension t (looked up per layout).  The sweeps of one level of the shared-prefix
        schedule go out as one batched launch (amdg_sweep1d_batch); every layout switch is one all-to-all."""
        A, d, h = self.A, self.dim, self.part.h
        edge = lambda S, k: kt if (S >> k) & 1 else kf

        def alloc(layout, sizes):
            return torch.empty(len(self.part.local[layout]), int(np.prod(sizes)), dtype=torch.float64, device=src_x.device)

        def batch(layout, lu, k, jobs, coef=1.0):
            # jobs: (sizes_from, src, dst, accumulate)
            if not jobs or len(self.part.local[layout]) == 0:
                return
            c = self.ctx[layout]
            c.sweep1d_batch(self.ops[layout][op_names[k]], rels[k], lu, k, [j[0] for j in jobs], [j[1] for j in jobs], [j[2] for j in jobs],
                            coefs=[coef] * len(jobs), accumulates=[int(j[3]) for j in jobs])

        # down pass: X_S for S subset of {0..d-2}; sizes of X_S: dims in S have kt
        X = {0: src_x}
        layout = "X"
        for k in range(d - 1):
            if k == h:
                X = self._switch_all(X, "X", "V")
                layout = "V"
            jobs = []
            for S in list(X.keys()):
                sizes = [edge(S, q) if q < k else kf for q in range(d)]
                out_sizes = list(sizes)
                out_sizes[k] = kt
                y = alloc(layout, out_sizes)
                jobs.append((sizes, X[S], y, False))
                X[S | (1 << k)] = y
            batch(layout, A.LU_L, k, jobs)
        if d - 1 >= h and layout == "X":
            X = self._switch_all(X, "X", "V")
            layout = "V"
        # full sweep along d-1
        R, jobs = {}, []
        for S, x in X.items():
            sizes = [edge(S, q) for q in range(d - 1)] + [kf]
            y = alloc(layout, sizes[:-1] + [kt])
            jobs.append((sizes, x, y, False))
            R[S] = y
        batch(layout, A.LU_FULL, d - 1, jobs, coef=coef)
        # up pass: R_k(S) = U_k R_{k+1}(S) + R_{k+1}(S + {k})
        for k in range(d - 2, -1, -1):
            if k == h - 1 and layout == "V":
                R = self._switch_all(R, "V", "X")
                layout = "X"
            newR, jobs = {}, []
            for S in [s for s in R if not (s >> k) & 1]:
                sizes = [edge(S, q) if q <= k else kt for q in range(d)]
                hi = R[S | (1 << k)]
                jobs.append((sizes, R[S], hi, True))
                newR[S] = hi
            batch(layout, A.LU_U, k, jobs)
            R = newR
        if layout == "V":
            R = self._switch_all(R, "V", "X")
        return R[0]
