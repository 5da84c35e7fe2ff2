"""Row-sharded multi-GPU operation: one process per GPU, `torch.distributed` for the plumbing.

The operator shards by spatial dof rows: rank p owns the rows I_p of every K_m and of all vectors, for ALL
modes, so the G-coupling is local.  One exchange per application (halo rows of X from the neighbouring ranks,
`batch_isend_irecv` = grouped NCCL send/recv over NVLink), scalars of the Krylov loop by all-reduce.  The
reference has no distributed path at all (dead `using Distributed`, src/ExtendableASGFEM.jl:3).

Everything here is host-side bookkeeping (integer partitions, halo lists, the Krylov recurrence on scalars); all
arithmetic on vectors happens in the per-rank `Context` (libasgfem_cuda.so).  The backend object is injectable so
that the exchange/partition logic is tested with gloo on CPU (tests/test_distributed_cpu.py).

Since round 2 the production path is INSIDE the library (csrc/dist.cu: asgfem_comm_init / asgfem_set_halo /
asgfem_precond_setup_global - NCCL halo exchange, all-reduced inner products, exact mean preconditioner); what remains
here is the partitioner (`LocalProblem`, `strip_shard`), the context setup helpers and `DistributedOperator` / `pcg`
as the host-driven reference implementation of the same algorithm (rank-local block-Jacobi preconditioner) that the
gloo tests exercise without a GPU.
"""
from __future__ import annotations

import numpy as np


# ---- partitioning ----------------------------------------------------------------------------------
def partition_rows(n, nparts, coords=None):
    """owner[i] in [0, nparts): recursive coordinate bisection of the dof coordinates (METIS is not available);
    contiguous index blocks if no coordinates are given."""
    owner = np.zeros(n, dtype=np.int64)
    if coords is None:
        bounds = np.linspace(0, n, nparts + 1).astype(np.int64)
        for p in range(nparts):
            owner[bounds[p]:bounds[p + 1]] = p
        return owner

    def rcb(idx, p0, np_):
        if np_ == 1:
            owner[idx] = p0
            return
        c = coords[idx]
        axis = int(np.argmax(c.max(axis=0) - c.min(axis=0)))
        left = np_ // 2
        order = np.argsort(c[:, axis], kind="stable")
        cut = len(idx) * left // np_
        rcb(idx[order[:cut]], p0, left)
        rcb(idx[order[cut:]], p0 + left, np_ - left)

    rcb(np.arange(n), 0, nparts)
    return owner


class LocalProblem:
    """Local numbering of rank `rank`: owned dofs first (rows without halo columns, then rows with halo columns, each
    ascending in the global id), then halo dofs grouped by owner."""

    def __init__(self, rank, owner, indptr, indices):
        self.rank = rank
        owner = np.asarray(owner)
        indptr = np.asarray(indptr, dtype=np.int64)
        indices = np.asarray(indices, dtype=np.int64)
        n = len(owner)
        row_of = np.repeat(np.arange(n, dtype=np.int64), np.diff(indptr))  # row of every stored entry
        mine_row = owner[row_of] == rank
        foreign_col = owner[indices] != rank
        # interior rows (all columns owned) first, rows that reference a halo column last: the operator on the first
        # n_interior rows does not need the halo exchange and runs while it is in flight
        touches = np.zeros(n, dtype=bool)
        touches[row_of[mine_row & foreign_col]] = True
        owned = np.where(owner == rank)[0]
        th = touches[owned]
        self.owned = np.concatenate([owned[~th], owned[th]])
        self.n_owned = len(self.owned)
        self.n_interior = int(np.count_nonzero(~th))
        halo = np.unique(indices[mine_row & foreign_col])
        halo = halo[np.lexsort((halo, owner[halo]))]
        self.halo = halo
        self.local_to_global = np.concatenate([self.owned, halo])
        self.n_local = len(self.local_to_global)
        g2l = -np.ones(n, dtype=np.int64)
        g2l[self.local_to_global] = np.arange(self.n_local)
        self.global_to_local = g2l
        # receive lists: local halo positions per neighbour rank (ascending global id)
        self.recv = {int(q): g2l[halo[owner[halo] == q]] for q in np.unique(owner[halo])}
        # send lists: my dofs that the rows of rank q reference = the columns of q's rows that I own (no symmetry of the
        # pattern assumed: this is exactly the halo of q restricted to my dofs, in the same ascending global order)
        self.send = {}
        col_mine = owner[indices] == rank
        for q in np.unique(owner[row_of[col_mine & ~mine_row]]) if n else []:
            sel = col_mine & (owner[row_of] == q)
            self.send[int(q)] = g2l[np.unique(indices[sel])]
        # local CSR: owned rows with local column ids (sorted), halo rows empty
        cnt = np.diff(indptr)[self.owned]
        lp = np.zeros(self.n_local + 1, dtype=np.int64)
        lp[1:self.n_owned + 1] = np.cumsum(cnt)
        lp[self.n_owned + 1:] = lp[self.n_owned]
        # entries of the owned rows in the new row order
        starts = indptr[self.owned]
        src = np.repeat(starts - lp[:self.n_owned], cnt) + np.arange(lp[self.n_owned], dtype=np.int64)
        lrow = np.repeat(np.arange(self.n_owned, dtype=np.int64), cnt)
        lcol = g2l[indices[src]]
        order = np.lexsort((lcol, lrow))
        self.indptr = lp
        self.indices = lcol[order]
        self._src = src[order]  # position in the global CSR of every local entry

    def local_values(self, indptr, indices, vals):
        """Values of the owned rows in the order of the local CSR (columns sorted by local id)."""
        return np.asarray(vals)[self._src]


# ---- halo exchange + Krylov loop ---------------------------------------------------------------------
class HaloExchange:
    """Grouped send/recv of halo rows.  `backend.pack_rows(slot, rows0)` returns a torch tensor (device of the
    process group), `backend.unpack_rows(slot, rows0, tensor)` scatters it back."""

    def __init__(self, local: LocalProblem, backend, dist):
        self.local, self.backend, self.dist = local, backend, dist

    def start(self, slot):
        """Packs and posts the sends / receives; returns the handle for `finish` (None: nothing to exchange)."""
        d = self.dist
        if d is None or d.get_world_size() == 1:
            return None
        ops, recv_bufs, send_bufs = [], {}, []
        for q, rows in self.local.send.items():
            send_bufs.append(self.backend.pack_rows(slot, rows))
            ops.append(d.P2POp(d.isend, send_bufs[-1], q))
        for q, rows in self.local.recv.items():
            recv_bufs[q] = self.backend.empty_rows(len(rows))
            ops.append(d.P2POp(d.irecv, recv_bufs[q], q))
        works = d.batch_isend_irecv(ops) if ops else []
        return works, recv_bufs, send_bufs

    def finish(self, slot, handle):
        if handle is None:
            return
        works, recv_bufs, _send_bufs = handle
        for w in works:
            w.wait()
        self.backend.sync()
        for q, rows in self.local.recv.items():
            self.backend.unpack_rows(slot, rows, recv_bufs[q])

    def __call__(self, slot):
        self.finish(slot, self.start(slot))


class DistributedOperator:
    def __init__(self, local: LocalProblem, backend, dist):
        self.local, self.backend, self.dist = local, backend, dist
        self.exchange = HaloExchange(local, backend, dist)

    def apply(self, sx, sy):
        """Y = A X on the owned rows: the rows without halo columns are applied while the halo rows of X travel."""
        if not hasattr(self.backend, "apply_rows"):
            self.exchange(sx)
            self.backend.apply(sx, sy)
            return
        handle = self.exchange.start(sx)
        self.backend.apply_rows(sx, sy, 0, self.local.n_interior)
        self.exchange.finish(sx, handle)
        self.backend.apply_rows(sx, sy, self.local.n_interior, self.local.n_owned)

    def dot(self, a, b):
        import torch
        v = torch.tensor([self.backend.dot_owned(a, b)], dtype=torch.float64, device=self.backend.device)
        if self.dist is not None and self.dist.get_world_size() > 1:
            self.dist.all_reduce(v)
        return float(v.item())


def pcg(op: DistributedOperator, slots, atol=1e-14, rtol=1e-14, itmax=1000):
    """Preconditioned CG with the rank-local mean preconditioner (block-Jacobi over the row partition: every rank
    factorises its own block of K_0; same converged solution as the global mean preconditioner, more iterations -
    SURVEY.md §7 'Multi-GPU preconditioner', option 2).  slots = dict(x, b, r, z, p, q) of vector slot ids;
    x holds the warm start, b the right-hand side (boundary rows zeroed)."""
    be = op.backend
    x, b, r, z, p, q = (slots[k] for k in ("x", "b", "r", "z", "p", "q"))
    op.apply(x, q)
    be.copy(b, r)
    be.axpy(-1.0, q, r)
    be.precond_apply(r, z)
    be.copy(z, p)
    rz = op.dot(r, z)
    rz0 = rz
    eps = atol + rtol * np.sqrt(max(rz0, 0.0))
    k = 0
    hist = [np.sqrt(max(rz, 0.0))]
    while k < itmax and np.sqrt(max(rz, 0.0)) > eps:
        op.apply(p, q)
        alpha = rz / op.dot(p, q)
        be.axpy(alpha, p, x)
        be.axpy(-alpha, q, r)
        be.precond_apply(r, z)
        rz_new = op.dot(r, z)
        be.xpay(z, rz_new / rz, p)
        rz = rz_new
        k += 1
        hist.append(np.sqrt(max(rz, 0.0)))
    return dict(niter=k, solved=np.sqrt(max(rz, 0.0)) <= eps, residuals=hist)


class ContextBackend:
    """Backend on top of a libasgfem_cuda Context (one GPU)."""

    def __init__(self, ctx, N):
        import torch
        self.ctx, self.N, self.torch = ctx, N, torch
        self.device = torch.device("cuda", torch.cuda.current_device())

    def empty_rows(self, nrows):
        return self.torch.empty(nrows * self.N, dtype=self.torch.float64, device=self.device)

    def pack_rows(self, slot, rows0):
        buf = self.empty_rows(len(rows0))
        self.ctx.pack_rows(slot, np.asarray(rows0) + 1, buf.data_ptr())
        return buf

    def unpack_rows(self, slot, rows0, buf):
        self.ctx.unpack_rows(slot, np.asarray(rows0) + 1, buf.data_ptr())

    def sync(self):
        self.torch.cuda.synchronize()

    def apply(self, sx, sy):
        self.ctx.apply(sx, sy)

    def apply_rows(self, sx, sy, row0, row1):
        if row1 > row0:
            self.ctx.apply_rows(sx, sy, row0, row1)

    def dot_owned(self, a, b):
        return self.ctx.vec_dot_owned(a, b)

    def axpy(self, alpha, x, y):
        self.ctx.vec_axpy(alpha, x, y)

    def xpay(self, x, beta, y):
        self.ctx.vec_xpay(x, beta, y)

    def copy(self, src, dst):
        self.ctx.vec_copy(src, dst)

    def precond_apply(self, r, z):
        self.ctx.precond_apply(r, z)


def setup_context(ctx, local: LocalProblem, mats, bdofs0, family, multi_indices):
    """Loads the rows owned by this rank (local numbering) into a Context.  mats = [K_0, K_1, ..., K_M] as global
    scipy CSR matrices; bdofs0 = global Dirichlet dofs (0-based).  Halo rows are empty and flagged as excluded, so
    the rank-local mean preconditioner factorises the owned interior block only (block-Jacobi)."""
    import scipy.sparse as sp
    n = local.n_local
    ctx.set_multiindices(family, np.asarray(multi_indices, dtype=np.int64))
    pat = sp.csr_matrix((np.ones(len(local.indices)), local.indices, local.indptr), shape=(n, n)).tocsc()
    pat.sort_indices()
    ctx.set_pattern_csc(n, pat.indptr.astype(np.int64) + 1, pat.indices.astype(np.int64) + 1)
    ctx.set_num_stiffness(len(mats) - 1)
    for m, K in enumerate(mats):
        K = sp.csr_matrix(K)
        vals = local.local_values(K.indptr, K.indices, K.data)
        Kl = sp.csr_matrix((vals, local.indices, local.indptr), shape=(n, n)).tocsc()
        Kl.sort_indices()
        ctx.set_stiffness_csc(m, Kl.indptr.astype(np.int64) + 1, Kl.indices.astype(np.int64) + 1, Kl.data)
    g2l = local.global_to_local
    bl = g2l[np.asarray(bdofs0)]
    bl = bl[bl >= 0]
    excluded = np.unique(np.concatenate([bl, np.arange(local.n_owned, n)]))
    ctx.set_bdofs(excluded + 1)
    ctx.set_owned_rows(local.n_owned)
    return ctx


# ---- strips of a structured mesh: dof ownership by coordinates (benchmark configs[4], sharded estimator) -----------------
class StripShard:
    """Row shard of a structured P1 / P2 mesh of ncx x (ncy_per_rank * world) squares cut into horizontal strips.

    Rank r owns the dofs with y-index in [r, r + 1) * ncy_per_rank (for P2 edge dofs: the half row above an owned node
    row; the last rank also owns the top row).  Its mesh = the cells that touch an owned dof plus `upper_halo_rows - 1`
    more cell rows above (the estimator needs the cell across every face of an owned cell: upper_halo_rows = 2).  Local
    numbering: owned dofs sorted by (y, x) - the rows along the lower cut first, along the upper cut last, so that the
    rows without halo columns are one contiguous range `interior` -, then the lower halo, then the upper halo.
    send / recv: dict neighbour rank -> local dof ids (0-based), both sides in (y, x) order.  cell_owned: the cells whose
    lower node row is owned (every cell of the global mesh is owned by exactly one rank)."""


def strip_shard(rank, world, order, ncx, ncy_per_rank, upper_halo_rows=1):
    from . import grids as _g
    S = StripShard()
    nx = ncx + 1
    ny_tot = ncy_per_rank * world + 1
    y_lo, y_hi = rank * ncy_per_rank, (rank + 1) * ncy_per_rank
    if rank == world - 1:
        y_hi = ny_tot
    ext_lo, ext_hi = max(y_lo - 1, 0), min(y_hi + upper_halo_rows, ny_tot)
    hx, hy = 1.0 / ncx, 1.0 / (ny_tot - 1)
    g = _g.structured_unitsquare(nx, ext_hi - ext_lo, 0.0, 1.0, ext_lo * hy, (ext_hi - 1) * hy)
    fes = _g.FESpace(g, order)
    ix = np.rint(g.coords[:, 0] / hx).astype(np.int64)
    iy = np.rint(g.coords[:, 1] / hy).astype(np.int64)
    px, py = 2 * ix, 2 * iy  # dof positions in half mesh widths
    if order == 2:
        fa, fb = g.facenodes[:, 0], g.facenodes[:, 1]
        px = np.concatenate([px, ix[fa] + ix[fb]])
        py = np.concatenate([py, iy[fa] + iy[fb]])
    owned = (py >= 2 * y_lo) & (py < 2 * y_hi)
    lower, upper = py < 2 * y_lo, py >= 2 * y_hi
    key = py * (4 * nx) + px

    def order_of(mask):
        idx = np.where(mask)[0]
        return idx[np.argsort(key[mask], kind="stable")]

    o_own, o_lo, o_up = order_of(owned), order_of(lower), order_of(upper)
    perm = np.concatenate([o_own, o_lo, o_up])  # new -> old
    new_of_old = np.empty(fes.ndofs, dtype=np.int64)
    new_of_old[perm] = np.arange(fes.ndofs)
    S.grid, S.fes, S.order = g, fes, order
    S.n_owned, S.n_local = len(o_own), fes.ndofs
    S.celldofs = new_of_old[fes.celldofs]
    S.perm, S.new_of_old = perm, new_of_old
    pxn, pyn = px[perm], py[perm]
    S.px, S.py = pxn, pyn
    S.bdofs = np.where((pxn == 0) | (pxn == 2 * ncx) | (pyn == 0) | (pyn == 2 * (ny_tot - 1)))[0]
    S.send, S.recv = {}, {}
    if y_lo > 0:  # lower neighbour: its upper halo = my dofs with y in [y_lo, y_lo + upper_halo_rows - 1]
        S.recv[rank - 1] = S.n_owned + np.arange(len(o_lo))
        S.send[rank - 1] = np.where(pyn[:S.n_owned] <= 2 * (y_lo + upper_halo_rows - 1))[0]
    if y_hi < ny_tot:  # upper neighbour: its lower halo = my dofs with y in [y_hi - 1, y_hi)
        S.recv[rank + 1] = S.n_owned + len(o_lo) + np.arange(len(o_up))
        S.send[rank + 1] = np.where(pyn[:S.n_owned] >= 2 * (y_hi - 1))[0]
    i0 = int(np.sum(pyn[:S.n_owned] == 2 * y_lo)) if y_lo > 0 else 0
    i1 = int(np.sum(pyn[:S.n_owned] < 2 * (y_hi - 1))) if y_hi < ny_tot else S.n_owned
    S.interior = (i0, i1)
    cy = iy[g.cellnodes].min(axis=1)  # lower node row of a cell
    S.cell_owned = ((cy >= y_lo) & (cy < y_hi)).astype(np.uint8)
    S.global_dof_key = key[perm]  # (y, x) key: identical for the same dof on every rank
    S.rank, S.world = rank, world
    return S


def setup_strip_context(ctx, S, coefficient, M, quadrature):
    """Mesh, renumbered space, coefficient tables and device assembly of K_0..K_M for a strip shard; halo rows are flagged
    like Dirichlet rows (they are not part of this rank's system)."""
    g = S.grid
    ctx.set_mesh(g.coords, g.cellnodes + 1)
    ctx.set_space(S.order, S.n_local, S.celldofs + 1)
    ctx.set_coefficient_cosinus(coefficient.mean_value, coefficient.decay_factors, coefficient.b1, coefficient.b2)
    xref, w = quadrature
    ctx.assemble_stiffness(M, xref, w)
    if S.world > 1:
        ctx.set_bdofs(np.union1d(S.bdofs, np.arange(S.n_owned, S.n_local)) + 1)
        ctx.set_owned_rows(S.n_owned)
    else:
        ctx.set_bdofs(S.bdofs + 1)
