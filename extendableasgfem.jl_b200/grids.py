"""Host-side stand-ins for the ExtendableGrids / ExtendableFEMBase objects a Julia caller already owns
(grid, FESpace, QuadratureRule): enough to drive the library from Python and to generate the synthetic
benchmark problems of BASELINE.json.  Index arrays are kept 0-based here and converted at the C ABI.

Conventions follow SURVEY.md Appendix B (ExtendableGrids: grid_unitsquare / grid_lshape / uniform_refine,
faces numbered by first appearance; ExtendableFEMBase: H1Pk dof layout [nodes; faces], quadrature rules).
"""
from __future__ import annotations

import numpy as np
from scipy.special import roots_jacobi

_LF = np.array([[0, 1], [1, 2], [2, 0]])


class Grid:
    def __init__(self, coords, cellnodes, bfacenodes):
        self.coords = np.ascontiguousarray(coords, dtype=np.float64)
        self.cellnodes = np.ascontiguousarray(cellnodes, dtype=np.int64)
        self.bfacenodes = np.ascontiguousarray(bfacenodes, dtype=np.int64)
        nn = len(self.coords)
        a = self.cellnodes[:, _LF[:, 0]].reshape(-1)
        b = self.cellnodes[:, _LF[:, 1]].reshape(-1)
        key = np.minimum(a, b) * nn + np.maximum(a, b)
        uniq, first, inv = np.unique(key, return_index=True, return_inverse=True)
        rank = np.empty(len(uniq), dtype=np.int64)
        rank[np.argsort(first, kind="stable")] = np.arange(len(uniq))
        self.cellfaces = rank[inv].reshape(-1, 3)
        self.nfaces = len(uniq)
        fn = np.empty((self.nfaces, 2), dtype=np.int64)
        fid = rank[inv]
        fn[fid[::-1], 0] = a[::-1]
        fn[fid[::-1], 1] = b[::-1]
        self.facenodes = fn
        ba, bb = self.bfacenodes[:, 0], self.bfacenodes[:, 1]
        pos = np.searchsorted(uniq, np.minimum(ba, bb) * nn + np.maximum(ba, bb))
        self.bfacefaces = rank[pos]
        x = self.coords
        c = self.cellnodes
        e1, e2 = x[c[:, 1]] - x[c[:, 0]], x[c[:, 2]] - x[c[:, 0]]
        self.cellvolumes = 0.5 * np.abs(e1[:, 0] * e2[:, 1] - e1[:, 1] * e2[:, 0])

    nnodes = property(lambda self: len(self.coords))
    ncells = property(lambda self: len(self.cellnodes))


def grid_unitsquare():
    return Grid([[0, 0], [1, 0], [1, 1], [0, 1], [0.5, 0.5]],
                np.array([[1, 2, 5], [2, 3, 5], [3, 4, 5], [4, 1, 5]]) - 1,
                np.array([[1, 2], [2, 3], [3, 4], [4, 1]]) - 1)


def grid_lshape():
    return Grid([[0, 0], [1, 0], [1, 1], [0, 1], [-1, 1], [-1, 0], [-1, -1], [0, -1]],
                np.array([[1, 2, 3], [1, 3, 4], [1, 4, 5], [5, 6, 1], [1, 6, 7], [1, 7, 8]]) - 1,
                np.array([[1, 2], [2, 3], [3, 4], [4, 5], [5, 6], [6, 7], [7, 8], [8, 1]]) - 1)


def uniform_refine(g: Grid, nrefs=1):
    for _ in range(nrefs):
        nn = g.nnodes
        coords = np.vstack([g.coords, 0.5 * (g.coords[g.facenodes[:, 0]] + g.coords[g.facenodes[:, 1]])])
        loc = np.hstack([g.cellnodes, g.cellfaces + nn])
        cells = loc[:, np.array([[1, 4, 6], [4, 2, 5], [6, 5, 3], [5, 6, 4]]) - 1].reshape(-1, 3)
        bm = g.bfacefaces + nn
        bf = np.stack([g.bfacenodes[:, 0], bm, bm, g.bfacenodes[:, 1]], axis=1).reshape(-1, 2)
        g = Grid(coords, cells, bf)
    return g


def structured_unitsquare(nx, ny=None, x0=0.0, x1=1.0, y0=0.0, y1=1.0):
    """nx x ny nodes, row-major numbering, every square cut by the same diagonal (synthetic benchmark mesh)."""
    ny = ny or nx
    X, Y = np.meshgrid(np.linspace(x0, x1, nx), np.linspace(y0, y1, ny), indexing="xy")
    coords = np.stack([X.reshape(-1), Y.reshape(-1)], axis=1)
    i, j = np.meshgrid(np.arange(nx - 1), np.arange(ny - 1), indexing="xy")
    n00 = (i + nx * j).reshape(-1)
    n10, n01 = n00 + 1, n00 + nx
    n11 = n01 + 1
    cells = np.stack([np.stack([n00, n10, n11], 1), np.stack([n00, n11, n01], 1)], axis=1).reshape(-1, 3)
    bot = np.stack([np.arange(nx - 1), np.arange(1, nx)], 1)
    right = np.stack([nx - 1 + nx * np.arange(ny - 1), nx - 1 + nx * np.arange(1, ny)], 1)
    top = np.stack([nx * (ny - 1) + np.arange(nx - 1, 0, -1), nx * (ny - 1) + np.arange(nx - 2, -1, -1)], 1)
    left = np.stack([nx * np.arange(ny - 1, 0, -1), nx * np.arange(ny - 2, -1, -1)], 1)
    return Grid(coords, cells, np.vstack([bot, right, top, left]))


class FESpace:
    """H1Pk{1,2,order}, order 1 or 2: celldofs (FES[CellDofs]), bdofs in first-occurrence order of the
    BFaceDofs sweep (solvers_poisson_primal.jl:136-142)."""

    def __init__(self, grid: Grid, order: int):
        assert order in (1, 2)
        self.grid, self.order = grid, order
        if order == 1:
            self.ndofs = grid.nnodes
            self.celldofs = grid.cellnodes.copy()
            bfd = grid.bfacenodes
        else:
            self.ndofs = grid.nnodes + grid.nfaces
            self.celldofs = np.hstack([grid.cellnodes, grid.cellfaces + grid.nnodes])
            bfd = np.hstack([grid.bfacenodes, (grid.bfacefaces + grid.nnodes)[:, None]])
        flat = bfd.reshape(-1)
        _, first = np.unique(flat, return_index=True)
        self.bdofs = flat[np.sort(first)]  # 0-based

    def rhs(self, f=None, bonus_quadorder=0):
        """b = (f, phi_i), LinearOperator(rhs, [id(1)]; bonus_quadorder) (poisson_primal.jl:66-68); host-side
        like in the reference (the rhs is a user closure)."""
        g = self.grid
        xref, w = quadrature_rule(self.order + bonus_quadorder)
        lam = np.stack([1 - xref[:, 0] - xref[:, 1], xref[:, 0], xref[:, 1]], axis=1)
        if self.order == 1:
            phi = lam
        else:
            phi = np.hstack([lam * (2 * lam - 1), 4 * lam[:, [0, 1, 2]] * lam[:, [1, 2, 0]]])
        c = g.cellnodes
        x1, x2, x3 = g.coords[c[:, 0]], g.coords[c[:, 1]], g.coords[c[:, 2]]
        xq = x1[:, None, :] + xref[None, :, 0:1] * (x2 - x1)[:, None, :] + xref[None, :, 1:2] * (x3 - x1)[:, None, :]
        fv = np.ones(xq.shape[:2]) if f is None else f(xq[:, :, 0], xq[:, :, 1])
        loc = np.einsum("c,q,cq,qi->ci", g.cellvolumes, w, fv, phi)
        b = np.zeros(self.ndofs)
        np.add.at(b, self.celldofs.reshape(-1), loc.reshape(-1))
        return b


def quadrature_rule(order):
    """QuadratureRule{Float64,Triangle2D}(order) stand-in: (xref (nq,2), w) with sum(w) = 1."""
    if order <= 1:
        return np.array([[1 / 3, 1 / 3]]), np.array([1.0])
    if order == 2:
        return np.array([[0.5, 0.0], [0.5, 0.5], [0.0, 0.5]]), np.full(3, 1 / 3)
    n = order // 2 + 1
    r, a = np.polynomial.legendre.leggauss(n)
    s, b = roots_jacobi(n, 1.0, 0.0)
    r, s = 0.5 * r + 0.5, 0.5 * s + 0.5
    pts = np.array([[s[j], r[i] * (1 - s[j])] for j in range(n) for i in range(n)])
    w = np.array([a[i] * b[j] for j in range(n) for i in range(n)])
    return pts, w / w.sum()


def quadrature_rule_1d(order):
    """QuadratureRule{Float64,Edge1D}(order) stand-in on [0,1]."""
    if order <= 1:
        return np.array([0.5]), np.array([1.0])
    x, w = np.polynomial.legendre.leggauss(order // 2 + 1)
    return 0.5 * x + 0.5, 0.5 * w
