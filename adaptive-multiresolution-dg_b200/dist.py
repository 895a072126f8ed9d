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
        self.level, self.suppt = level, suppt

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

    def switch(self, x, src, dst, group=None):
        """x: [n_local(src), block] tensor in layout src -> [n_local(dst), block] in layout dst"""
        send_order, send_counts, recv_counts, recv_order = self.plan(src, dst)
        blk = x.shape[1]
        dev = x.device
        sbuf = x.index_select(0, torch.as_tensor(send_order, device=dev)).contiguous()
        rbuf = torch.empty(int(recv_counts.sum()), blk, dtype=x.dtype, device=dev)
        if self.world == 1:
            rbuf.copy_(sbuf)
        else:
            dist.all_to_all_single(rbuf, sbuf, [int(c) for c in recv_counts], [int(c) for c in send_counts], group=group)
        out = torch.empty_like(rbuf)
        out.index_copy_(0, torch.as_tensor(recv_order, device=dev), rbuf)
        return out


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
            c.grid_set(part.level[rows], part.suppt[rows])
            self.ctx[k] = c
            self.ops[k] = register(c)
        self.switches = 0
        self.switch_bytes = 0

    def close(self):
        for c in self.ctx.values():
            c.close()

    def _switch_all(self, bufs, src, dst):
        out = {}
        for S, x in bufs.items():
            out[S] = self.part.switch(x, src, dst)
            self.switches += 1
            self.switch_bytes += x.numel() * 8
        return out

    def apply(self, op_names, rels, src_x, kf, kt, coef=1.0):
        """src_x: [n_local(X), kf^dim] in layout X -> returns [n_local(X), kt^dim] in layout X.
        op_names[t]: operator name of dimension t (looked up per layout)."""
        A, d, h = self.A, self.dim, self.part.h
        edge = lambda S, k: kt if (S >> k) & 1 else kf

        def sweep(layout, lu, k, sizes, x, out, coef=1.0, accumulate=False):
            c = self.ctx[layout]
            c.sweep1d(self.ops[layout][op_names[k]], rels[k], lu, k, sizes, x, out, coef=coef, accumulate=accumulate)

        def alloc(layout, sizes):
            return torch.zeros(len(self.part.local[layout]), int(np.prod(sizes)), dtype=torch.float64, device=src_x.device)

        # down pass: X_S for S subset of {0..d-2}; sizes of X_S: dims in S have kt
        X = {0: src_x}
        layout = "X"
        for k in range(d - 1):
            if k == h:
                X = self._switch_all(X, "X", "V")
                layout = "V"
            for S in list(X.keys()):
                sizes = [edge(S, q) if q < k else kf for q in range(d)]
                out_sizes = list(sizes)
                out_sizes[k] = kt
                y = alloc(layout, out_sizes)
                sweep(layout, A.LU_L, k, sizes, X[S], y)
                X[S | (1 << k)] = y
        if d - 1 >= h and layout == "X":
            X = self._switch_all(X, "X", "V")
            layout = "V"
        # full sweep along d-1
        R = {}
        for S, x in X.items():
            sizes = [edge(S, q) for q in range(d - 1)] + [kf]
            out_sizes = sizes[:-1] + [kt]
            y = alloc(layout, out_sizes)
            sweep(layout, A.LU_FULL, d - 1, sizes, x, y, coef=coef)
            R[S] = y
        # up pass: R_k(S) = U_k R_{k+1}(S) + R_{k+1}(S + {k})
        for k in range(d - 2, -1, -1):
            if k == h - 1 and layout == "V":
                R = self._switch_all(R, "V", "X")
                layout = "X"
            newR = {}
            for S in [s for s in R if not (s >> k) & 1]:
                sizes = [edge(S, q) if q <= k else kt for q in range(d)]
                hi = R[S | (1 << k)]
                sweep(layout, A.LU_U, k, sizes, R[S], hi, accumulate=True)
                newR[S] = hi
            R = newR
        if layout == "V":
            R = self._switch_all(R, "V", "X")
        return R[0]
