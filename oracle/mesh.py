"""Oracle: triangular meshes with ExtendableGrids-style adjacencies.

Test infrastructure only (see oracle/__init__.py).  The reference delegates all of this to the
un-vendored package ExtendableGrids >= 1.13 (call sites: scripts/poisson.jl:206-209,
scripts/poisson_simple.jl:39-45, src/estimate.jl:371-373,401-413).  What is restated here are the
ASSUMED semantics of SURVEY.md Appendix B.1 (parity unpinned, each isolated in one function):

* grid_unitsquare(Triangle2D): 5 nodes, 4 cells fanned around the centre node
* grid_lshape(Triangle2D): 8 nodes, 6 cells fanned around the re-entrant corner
* uniform_refine: red refinement, midpoint of face f becomes node nnodes+f, children
  [1 4 6; 4 2 5; 6 5 3; 5 6 4], boundary faces split in two
* faces are numbered by first appearance in a cell-major sweep over local faces (1,2),(2,3),(3,1)

All indices are 0-based inside Python; anything crossing the C ABI is converted to the 1-based
Int32/Int64 arrays Julia holds.
"""
from __future__ import annotations

import numpy as np

LOCAL_FACES = np.array([[0, 1], [1, 2], [2, 0]])


class Mesh:
    def __init__(self, coords, cellnodes, bfacenodes, bfaceregions=None):
        self.coords = np.ascontiguousarray(coords, dtype=np.float64)  # (nnodes, 2)
        self.cellnodes = np.ascontiguousarray(cellnodes, dtype=np.int64)  # (ncells, 3)
        self.bfacenodes = np.ascontiguousarray(bfacenodes, dtype=np.int64)  # (nbfaces, 2)
        self.bfaceregions = (
            np.ones(len(self.bfacenodes), dtype=np.int64) if bfaceregions is None else np.asarray(bfaceregions)
        )
        self._faces()
        self._volumes()

    @property
    def nnodes(self):
        return self.coords.shape[0]

    @property
    def ncells(self):
        return self.cellnodes.shape[0]

    @property
    def nfaces(self):
        return self.facenodes.shape[0]

    def _faces(self):
        cn = self.cellnodes
        nn = self.nnodes
        a = cn[:, LOCAL_FACES[:, 0]].reshape(-1)  # cell-major, local face minor
        b = cn[:, LOCAL_FACES[:, 1]].reshape(-1)
        key = np.minimum(a, b) * nn + np.maximum(a, b)
        uniq, first, inv = np.unique(key, return_index=True, return_inverse=True)
        order = np.argsort(first, kind="stable")  # face id = rank of first appearance
        rank = np.empty_like(order)
        rank[order] = np.arange(len(order))
        faceid = rank[inv]
        self.cellfaces = faceid.reshape(-1, 3)
        nfaces = len(uniq)
        fn = np.empty((nfaces, 2), dtype=np.int64)
        fn[faceid[::-1], 0] = a[::-1]  # reversed write => first appearance wins
        fn[faceid[::-1], 1] = b[::-1]
        self.facenodes = fn
        cellofslot = np.repeat(np.arange(self.ncells), 3)
        fc = -np.ones((nfaces, 2), dtype=np.int64)
        firstslot = first[order]
        fc[:, 0] = cellofslot[firstslot]
        # second cell: the other slot with the same face id (if any)
        slot_sorted = np.argsort(faceid, kind="stable")
        counts = np.bincount(faceid, minlength=nfaces)
        starts = np.concatenate([[0], np.cumsum(counts)[:-1]])
        has2 = counts == 2
        fc[has2, 1] = cellofslot[slot_sorted[starts[has2] + 1]]
        self.facecells = fc
        # boundary faces -> face ids
        ba, bb = self.bfacenodes[:, 0], self.bfacenodes[:, 1]
        bkey = np.minimum(ba, bb) * nn + np.maximum(ba, bb)
        pos = np.searchsorted(uniq, bkey)
        assert np.all(uniq[pos] == bkey), "boundary face not found among faces"
        self.bfacefaces = rank[pos]

    def _volumes(self):
        x = self.coords
        c = self.cellnodes
        e1 = x[c[:, 1]] - x[c[:, 0]]
        e2 = x[c[:, 2]] - x[c[:, 0]]
        self.cellvolumes = 0.5 * np.abs(e1[:, 0] * e2[:, 1] - e1[:, 1] * e2[:, 0])
        d = x[self.facenodes[:, 1]] - x[self.facenodes[:, 0]]
        self.facevolumes = np.sqrt(d[:, 0] ** 2 + d[:, 1] ** 2)


def grid_unitsquare():
    coords = [[0, 0], [1, 0], [1, 1], [0, 1], [0.5, 0.5]]
    cells = np.array([[1, 2, 5], [2, 3, 5], [3, 4, 5], [4, 1, 5]]) - 1
    bfaces = np.array([[1, 2], [2, 3], [3, 4], [4, 1]]) - 1
    return Mesh(coords, cells, bfaces, [1, 2, 3, 4])


def grid_lshape():
    coords = [[0, 0], [1, 0], [1, 1], [0, 1], [-1, 1], [-1, 0], [-1, -1], [0, -1]]
    cells = np.array([[1, 2, 3], [1, 3, 4], [1, 4, 5], [5, 6, 1], [1, 6, 7], [1, 7, 8]]) - 1
    bfaces = np.array([[1, 2], [2, 3], [3, 4], [4, 5], [5, 6], [6, 7], [7, 8], [8, 1]]) - 1
    return Mesh(coords, cells, bfaces, [1, 2, 3, 4, 5, 6, 7, 8])


def uniform_refine(mesh: Mesh, nrefs: int = 1):
    for _ in range(nrefs):
        nn = mesh.nnodes
        mid = 0.5 * (mesh.coords[mesh.facenodes[:, 0]] + mesh.coords[mesh.facenodes[:, 1]])
        coords = np.vstack([mesh.coords, mid])
        c = mesh.cellnodes
        f = mesh.cellfaces + nn
        loc = np.stack([c[:, 0], c[:, 1], c[:, 2], f[:, 0], f[:, 1], f[:, 2]], axis=1)  # local 1..6
        rule = np.array([[1, 4, 6], [4, 2, 5], [6, 5, 3], [5, 6, 4]]) - 1
        cells = loc[:, rule].reshape(-1, 3)
        bmid = mesh.bfacefaces + nn
        b = mesh.bfacenodes
        bfaces = np.stack([b[:, 0], bmid, bmid, b[:, 1]], axis=1).reshape(-1, 2)
        breg = np.repeat(mesh.bfaceregions, 2)
        mesh = Mesh(coords, cells, bfaces, breg)
    return mesh


def structured_unitsquare(nx: int, ny: int | None = None):
    """Synthetic benchmark mesh of SURVEY.md §8(d): nx x ny nodes, row-major node numbering, every
    square split by the same diagonal.  Defines the bench inputs; not part of the reference."""
    if ny is None:
        ny = nx
    xs = np.linspace(0.0, 1.0, nx)
    ys = np.linspace(0.0, 1.0, ny)
    X, Y = np.meshgrid(xs, ys, indexing="xy")
    coords = np.stack([X.reshape(-1), Y.reshape(-1)], axis=1)
    i, j = np.meshgrid(np.arange(nx - 1), np.arange(ny - 1), indexing="xy")
    n00 = (i + nx * j).reshape(-1)
    n10 = n00 + 1
    n01 = n00 + nx
    n11 = n01 + 1
    cells = np.stack([np.stack([n00, n10, n11], 1), np.stack([n00, n11, n01], 1)], axis=1).reshape(-1, 3)
    bot = np.stack([np.arange(nx - 1), np.arange(1, nx)], 1)
    right = np.stack([nx - 1 + nx * np.arange(ny - 1), nx - 1 + nx * np.arange(1, ny)], 1)
    top = np.stack([nx * (ny - 1) + np.arange(nx - 1, 0, -1), nx * (ny - 1) + np.arange(nx - 2, -1, -1)], 1)
    left = np.stack([nx * np.arange(ny - 1, 0, -1), nx * np.arange(ny - 2, -1, -1)], 1)
    bfaces = np.vstack([bot, right, top, left])
    breg = np.concatenate([np.full(len(bot), 1), np.full(len(right), 2), np.full(len(top), 3), np.full(len(left), 4)])
    return Mesh(coords, cells, bfaces, breg)
