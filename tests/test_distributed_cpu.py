"""World-size-2 gloo tests (CPU) of the multi-GPU host logic: row partition, local numbering, halo lists, the
grouped send/recv exchange and the distributed PCG recurrence.  The per-rank compute backend is a numpy/scipy
stand-in built from the oracle (tests only); on a GPU box the same classes run on ContextBackend."""
import os

import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import problem as oproblem
from oracle import solver as osolver

import asgfem_b200  # noqa: F401
from asgfem_b200 import distributed as D


class NumpyBackend:
    """Local vectors: dict slot -> (n_local, N) array; local K_m = owned rows with local columns."""

    def __init__(self, L, P):
        self.L, self.N = L, P.N
        self.device = torch.device("cpu")
        self.v = {}
        n_local = L.n_local
        self.K = []
        for A in [P.A0] + P.Am:
            A = sp.csr_matrix(A)
            vals = L.local_values(A.indptr, A.indices, A.data)
            self.K.append(sp.csr_matrix((vals, L.indices, L.indptr), shape=(n_local, n_local)))
        self.G = P.G.tocsr()
        self.Mlen = P.M
        bmask = np.zeros(n_local, dtype=bool)
        bmask[L.global_to_local[P.bdofs][L.global_to_local[P.bdofs] >= 0]] = True
        self.bmask = bmask
        own_int = np.where(~bmask[:L.n_owned])[0]
        self.own_int = own_int
        self.lu = spla.splu(sp.csc_matrix(self.K[0][own_int][:, own_int]))

    def vec(self, s):
        return self.v.setdefault(s, np.zeros((self.L.n_local, self.N)))

    def empty_rows(self, nrows):
        return torch.empty(nrows * self.N, dtype=torch.float64)

    def pack_rows(self, slot, rows0):
        return torch.from_numpy(self.vec(slot)[rows0].reshape(-1).copy())

    def unpack_rows(self, slot, rows0, buf):
        self.vec(slot)[rows0] = buf.numpy().reshape(len(rows0), self.N)

    def sync(self):
        pass

    def apply(self, sx, sy):
        X, N = self.vec(sx), self.N
        Y = np.zeros_like(X)
        for mu in range(N):
            Y[:, mu] += self.K[0] @ X[:, mu]
            for e in range(self.Mlen):
                row = e * N + mu
                for p in range(self.G.indptr[row], self.G.indptr[row + 1]):
                    Y[:, mu] += self.G.data[p] * (self.K[e + 1] @ X[:, self.G.indices[p]])
        Y[self.bmask] = 0
        Y[self.L.n_owned:] = self.vec(sy)[self.L.n_owned:]  # halo rows are not written
        self.v[sy] = Y

    def apply_rows(self, sx, sy, row0, row1):
        """Rows [row0, row1) only, the others keep their content (asgfem_apply_rows)."""
        keep = self.vec(sy).copy()
        self.apply(sx, sy)
        keep[row0:row1] = self.v[sy][row0:row1]
        self.v[sy] = keep

    def dot_owned(self, a, b):
        n = self.L.n_owned
        return float(np.sum(self.vec(a)[:n] * self.vec(b)[:n]))

    def axpy(self, alpha, x, y):
        self.vec(y)[:] += alpha * self.vec(x)

    def xpay(self, x, beta, y):
        self.v[y] = self.vec(x) + beta * self.vec(y)

    def copy(self, src, dst):
        self.v[dst] = self.vec(src).copy()

    def precond_apply(self, r, z):
        R = self.vec(r)
        Z = np.zeros_like(R)
        Z[self.own_int] = self.lu.solve(R[self.own_int])
        self.v[z] = Z


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        P = oproblem.poisson_simple(nrefs=2, order=1)
        A0 = sp.csr_matrix(P.A0)
        owner = D.partition_rows(P.n, world, P.mesh.coords)
        L = D.LocalProblem(rank, owner, A0.indptr, A0.indices)
        be = NumpyBackend(L, P)
        op = D.DistributedOperator(L, be, dist)
        # --- operator: global oracle vs sharded apply ---------------------------------------------------
        xg = np.random.default_rng(0).standard_normal(P.n * P.N)
        S = osolver.SystemPrimal(P.A0, P.Am, P.G, P.bdofs, P.N)
        ref = S.mul(xg).reshape(P.N, P.n).T
        X = xg.reshape(P.N, P.n).T
        be.vec(0)[:L.n_owned] = X[L.owned]          # halo rows intentionally left at zero: exchange must fill them
        assert 0 < L.n_interior < L.n_owned  # some rows run behind the exchange, some need the halo
        op.apply(0, 1)
        err_apply = np.abs(be.vec(1)[:L.n_owned] - ref[L.owned]).max() / np.abs(ref).max()
        nrm = op.dot(0, 0)
        # --- PCG with the rank-local mean preconditioner vs the oracle solution -----------------------------
        refsol = np.zeros(P.n * P.N)
        osolver.solve_primal(refsol, P.A0, P.Am, P.b0, P.G, P.N, P.bdofs)
        refsol = refsol.reshape(P.N, P.n).T
        b = np.zeros((L.n_local, P.N))
        b[:L.n_owned, 0] = P.b0[L.owned]
        b[be.bmask] = 0
        be.v = {0: np.zeros((L.n_local, P.N)), 1: b}
        st = D.pcg(op, dict(x=0, b=1, r=2, z=3, p=4, q=5), atol=1e-14, rtol=1e-13, itmax=500)
        err_sol = np.abs(be.vec(0)[:L.n_owned] - refsol[L.owned]).max() / np.abs(refsol).max()
        q.put((rank, err_apply, abs(nrm - float(xg @ xg)) / float(xg @ xg), st["niter"], bool(st["solved"]), err_sol,
               L.n_owned, sorted(L.recv.keys())))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_operator_and_pcg_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500) + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sum(r[6] for r in res) == oproblem.poisson_simple(nrefs=2, order=1).n
    for rank, err_apply, err_dot, niter, solved, err_sol, n_owned, nbrs in res:
        assert err_apply < 1e-13, (rank, err_apply)
        assert err_dot < 1e-13
        assert solved and niter < 200
        assert err_sol < 1e-10
        assert len(nbrs) >= 1


def test_partition_is_balanced_and_complete():
    P = oproblem.poisson_simple(nrefs=3, order=1)
    for parts in (2, 4, 8):
        owner = D.partition_rows(P.n, parts, P.mesh.coords)
        counts = np.bincount(owner, minlength=parts)
        assert counts.sum() == P.n and counts.max() - counts.min() <= parts
    owner = D.partition_rows(10, 3)
    assert list(owner) == [0, 0, 0, 1, 1, 1, 2, 2, 2, 2]


@pytest.mark.parametrize("order", [1, 2])
@pytest.mark.parametrize("upper", [1, 2])
def test_strip_shards_partition_dofs_and_cells(order, upper):
    """Strips of a structured mesh (bench configs[4], sharded estimator): every dof and every cell has exactly one owner,
    and the send list of a rank names the same dofs in the same order as the receive list of its neighbour."""
    from asgfem_b200 import distributed as D
    world, ncx, ncy = 4, 16, 4
    shards = [D.strip_shard(r, world, order, ncx, ncy, upper) for r in range(world)]
    n_glob = (ncx + 1) * (ncy * world + 1) if order == 1 else (2 * ncx + 1) * (2 * ncy * world + 1)
    assert sum(S.n_owned for S in shards) == n_glob
    assert sum(int(S.cell_owned.sum()) for S in shards) == 2 * ncx * ncy * world
    keys = np.concatenate([S.global_dof_key[:S.n_owned] for S in shards])
    assert len(np.unique(keys)) == n_glob
    for r in range(world - 1):
        lo, up = shards[r], shards[r + 1]
        assert np.array_equal(lo.global_dof_key[lo.send[r + 1]], up.global_dof_key[up.recv[r]])
        assert np.array_equal(up.global_dof_key[up.send[r]], lo.global_dof_key[lo.recv[r + 1]])
    for S in shards:
        i0, i1 = S.interior
        assert 0 <= i0 <= i1 <= S.n_owned
        # rows [i0, i1) touch no halo dof: no cell contains both such a dof and a halo dof
        touches = np.zeros(S.n_local, dtype=bool)
        halo_cell = (S.celldofs >= S.n_owned).any(axis=1)
        touches[np.unique(S.celldofs[halo_cell])] = True
        assert not touches[i0:i1].any()
