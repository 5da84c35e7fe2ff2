"""Oracle: P1/P2 Lagrange spaces, quadrature and assembly of K_0..K_M and the rhs.

Test infrastructure only (see oracle/__init__.py).  Reference call sites:
src/modelproblems/poisson_primal.jl:56-63 (K_m = BilinearOperator(get_am_x(m,C),[grad(1)],[grad(1)];
bonus_quadorder)), 66-68 (rhs LinearOperator(rhs,[id(1)])), src/modelproblems/
solvers_poisson_primal.jl:136-142 (bdofs).  The arithmetic lives in the un-vendored packages
ExtendableFEM >= 1.10 / ExtendableFEMBase >= 1.5.1 -> PARITY UNPINNED; the assumed semantics
 are isolated here:

* quadrature_rule: order<=1 centroid; order 2 the three edge midpoints (1/3 each); higher orders the
  generic Stroud conical product rule with div(order,2)+1 points per direction; weights sum to 1
* H1Pk{1,2,1}: dof = node.  H1Pk{1,2,2}: dofs = [nodes; faces], basis l_i(2l_i-1), 4 l_i l_j;
  CellDofs = 3 node dofs then 3 face dofs in local face order; BFaceDofs = 2 nodes (+ face dof)
* assembly: A[test,ansatz] += |T| w_q a_m(x_q) grad(phi_ansatz).grad(phi_test); operator quadrature
  order = sum(poly order - derivative order) + bonus_quadorder (P1: 2, P2: 4 at bonus 2)
* one shared sorted CSR/CSC pattern (all dof pairs sharing a cell) for every m
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
from scipy.special import roots_jacobi


def quadrature_rule(order: int):
    """Triangle rule on the reference triangle; returns (xref (nq,2), w (nq,)) with sum(w)=1.
    A point xref maps to x1 + xref[0]*(x2-x1) + xref[1]*(x3-x1)."""
    if order <= 1:
        return np.array([[1.0 / 3.0, 1.0 / 3.0]]), np.array([1.0])
    if order == 2:
        return np.array([[0.5, 0.0], [0.5, 0.5], [0.0, 0.5]]), np.array([1.0 / 3.0] * 3)
    n = order // 2 + 1
    r, a = np.polynomial.legendre.leggauss(n)
    s, b = roots_jacobi(n, 1.0, 0.0)
    r = 0.5 * r + 0.5
    s = 0.5 * s + 0.5
    a = 0.5 * a
    b = 0.25 * b  # weight (1-s) on [0,1]
    pts, wts = [], []
    for j in range(n):
        for i in range(n):
            pts.append([s[j], r[i] * (1.0 - s[j])])
            wts.append(a[i] * b[j])
    w = np.array(wts)
    return np.array(pts), w / w.sum()


def quadrature_rule_1d(order: int):
    """Rule on the reference face [0,1]; sum(w)=1.  order<=1 midpoint, else Gauss-Legendre with
    div(order,2)+1 points (assumed semantics of QuadratureRule{T,Edge1D})."""
    if order <= 1:
        return np.array([0.5]), np.array([1.0])
    n = order // 2 + 1
    x, w = np.polynomial.legendre.leggauss(n)
    return 0.5 * x + 0.5, 0.5 * w


class FESpace:
    """H1Pk{1,2,order} for order 1 or 2."""

    def __init__(self, mesh, order: int):
        assert order in (1, 2)
        self.mesh = mesh
        self.order = order
        if order == 1:
            self.ndofs = mesh.nnodes
            self.celldofs = mesh.cellnodes.copy()
            bfd = mesh.bfacenodes
        else:
            self.ndofs = mesh.nnodes + mesh.nfaces
            self.celldofs = np.hstack([mesh.cellnodes, mesh.cellfaces + mesh.nnodes])
            bfd = np.hstack([mesh.bfacenodes, (mesh.bfacefaces + mesh.nnodes)[:, None]])
        self.ndofs4cell = self.celldofs.shape[1]
        self.bfacedofs = bfd
        # bdofs = unique(all BFaceDofs entries), order of first occurrence (solvers_poisson_primal.jl:136-142)
        flat = bfd.reshape(-1)
        _, first = np.unique(flat, return_index=True)
        self.bdofs = flat[np.sort(first)]

    # reference basis ------------------------------------------------------------------------
    def basis(self, xref):
        """phi (nq, nd), dphi/dlambda (nq, nd, 3) at reference points xref (nq,2);
        barycentrics l1 = 1-x-y, l2 = x, l3 = y."""
        xref = np.atleast_2d(xref)
        l = np.stack([1 - xref[:, 0] - xref[:, 1], xref[:, 0], xref[:, 1]], axis=1)
        nq = l.shape[0]
        if self.order == 1:
            phi = l.copy()
            dphi = np.tile(np.eye(3)[None], (nq, 1, 1))
            return phi, dphi
        phi = np.zeros((nq, 6))
        dphi = np.zeros((nq, 6, 3))
        for i in range(3):
            phi[:, i] = l[:, i] * (2 * l[:, i] - 1)
            dphi[:, i, i] = 4 * l[:, i] - 1
        for f, (i, j) in enumerate([(0, 1), (1, 2), (2, 0)]):
            phi[:, 3 + f] = 4 * l[:, i] * l[:, j]
            dphi[:, 3 + f, i] = 4 * l[:, j]
            dphi[:, 3 + f, j] = 4 * l[:, i]
        return phi, dphi

    def lambda_gradients(self):
        """grad(lambda_i) per cell: (ncells, 3, 2)."""
        x = self.mesh.coords
        c = self.mesh.cellnodes
        x1, x2, x3 = x[c[:, 0]], x[c[:, 1]], x[c[:, 2]]
        det = (x2[:, 0] - x1[:, 0]) * (x3[:, 1] - x1[:, 1]) - (x2[:, 1] - x1[:, 1]) * (x3[:, 0] - x1[:, 0])
        g = np.empty((len(c), 3, 2))
        g[:, 0, 0] = (x2[:, 1] - x3[:, 1]) / det
        g[:, 0, 1] = (x3[:, 0] - x2[:, 0]) / det
        g[:, 1, 0] = (x3[:, 1] - x1[:, 1]) / det
        g[:, 1, 1] = (x1[:, 0] - x3[:, 0]) / det
        g[:, 2, 0] = (x1[:, 1] - x2[:, 1]) / det
        g[:, 2, 1] = (x2[:, 0] - x1[:, 0]) / det
        return g

    def laplacians(self):
        """Laplacian of the basis functions per cell (ncells, nd); zero for P1, constant for P2."""
        if self.order == 1:
            return np.zeros((self.mesh.ncells, 3))
        g = self.lambda_gradients()
        out = np.zeros((self.mesh.ncells, 6))
        for i in range(3):
            out[:, i] = 4 * np.einsum("cd,cd->c", g[:, i], g[:, i])
        for f, (i, j) in enumerate([(0, 1), (1, 2), (2, 0)]):
            out[:, 3 + f] = 8 * np.einsum("cd,cd->c", g[:, i], g[:, j])
        return out

    def physical_points(self, xref):
        x = self.mesh.coords
        c = self.mesh.cellnodes
        x1, x2, x3 = x[c[:, 0]], x[c[:, 1]], x[c[:, 2]]
        xref = np.atleast_2d(xref)
        # (ncells, nq, 2)
        return x1[:, None, :] + xref[None, :, 0:1] * (x2 - x1)[:, None, :] + xref[None, :, 1:2] * (x3 - x1)[:, None, :]


def pattern(space: FESpace):
    """Shared sorted CSR pattern (indptr, indices) of all dof pairs sharing a cell, plus for every
    (cell, i, j) the position in the value array (ncells, nd, nd)."""
    cd = space.celldofs
    nd = space.ndofs4cell
    rows = np.repeat(cd, nd, axis=1).reshape(-1)
    cols = np.tile(cd, (1, nd)).reshape(-1)
    n = space.ndofs
    key = rows * n + cols
    uniq, inv = np.unique(key, return_inverse=True)
    r = uniq // n
    c = uniq % n
    indptr = np.zeros(n + 1, dtype=np.int64)
    np.add.at(indptr, r + 1, 1)
    indptr = np.cumsum(indptr)
    return indptr, c.astype(np.int64), inv.reshape(len(cd), nd, nd)


def assemble_stiffness(space: FESpace, coeff, M: int, bonus_quadorder: int = 2, modes=None):
    """Returns (indptr, indices, vals) with vals of shape (M+1, nnz): K_m[i,j] = int a_m grad phi_j . grad phi_i.
    Values are accumulated cell by cell in cell order (np.add.at) like a serial assembly loop."""
    mesh = space.mesh
    qorder = 2 * (space.order - 1) + bonus_quadorder
    xref, w = quadrature_rule(qorder)
    indptr, indices, pos = pattern(space)
    nnz = len(indices)
    g = space.lambda_gradients()  # (nc,3,2)
    _, dphi = space.basis(xref)  # (nq, nd, 3)
    gradphi = np.einsum("qdl,clx->cqdx", dphi, g)  # (nc, nq, nd, 2)
    stiff_q = np.einsum("cqix,cqjx->cqij", gradphi, gradphi)  # test i, ansatz j
    xq = space.physical_points(xref)  # (nc, nq, 2)
    vol = mesh.cellvolumes
    ms = range(M + 1) if modes is None else modes
    vals = np.zeros((M + 1, nnz))
    for m in ms:
        am = coeff.am(m, xq[:, :, 0], xq[:, :, 1])  # (nc, nq)
        loc = np.einsum("c,q,cq,cqij->cij", vol, w, am, stiff_q)
        np.add.at(vals[m], pos.reshape(-1), loc.reshape(-1))
    return indptr, indices, vals


def assemble_rhs(space: FESpace, f=None, bonus_quadorder: int = 0):
    """b_i = int f phi_i with quadrature order = poly order + bonus (poisson_primal.jl:66-68)."""
    mesh = space.mesh
    xref, w = quadrature_rule(space.order + bonus_quadorder)
    phi, _ = space.basis(xref)
    xq = space.physical_points(xref)
    fv = np.ones(xq.shape[:2]) if f is None else f(xq[:, :, 0], xq[:, :, 1])
    loc = np.einsum("c,q,cq,qi->ci", mesh.cellvolumes, w, fv, phi)
    b = np.zeros(space.ndofs)
    np.add.at(b, space.celldofs.reshape(-1), loc.reshape(-1))
    return b


def csr(indptr, indices, vals, n):
    return sp.csr_matrix((vals, indices, indptr), shape=(n, n))


# ---- log-transformed primal problem (src/modelproblems/logpoisson_primal.jl:95-128) ---------------------------------------
def assemble_logprimal_matrices(space: FESpace, coeff, M: int, bonus_quadorder: int = 2):
    """Returns (indptr, indices, vals) with vals of shape (M+1, nnz) on the pattern of assemble_stiffness:
    vals[0] = A[i,j] = int grad phi_j . grad phi_i                      (:96-97, BilinearOperator([grad(1)]))
    vals[m] = N_m[i,j] = - int (grad a_m . grad phi_j) phi_i            (:100-105, kernel get_gradam_x_sigma,
                                                                         src/coefficients/coefficients.jl:169-176; factor -1)
    One quadrature rule of order (order - 1) + order + bonus for all planes (exact for A)."""
    mesh = space.mesh
    xref, w = quadrature_rule(2 * space.order - 1 + bonus_quadorder)
    indptr, indices, pos = pattern(space)
    nnz = len(indices)
    g = space.lambda_gradients()
    phi, dphi = space.basis(xref)  # (nq, nd), (nq, nd, 3)
    gradphi = np.einsum("qdl,clx->cqdx", dphi, g)  # (nc, nq, nd, 2)
    xq = space.physical_points(xref)
    vol = mesh.cellvolumes
    vals = np.zeros((M + 1, nnz))
    loc = np.einsum("c,q,cqix,cqjx->cij", vol, w, gradphi, gradphi)
    np.add.at(vals[0], pos.reshape(-1), loc.reshape(-1))
    for m in range(1, M + 1):
        gx, gy = coeff.gradam(m, xq[:, :, 0], xq[:, :, 1])  # (nc, nq)
        conv = gx[:, :, None] * gradphi[:, :, :, 0] + gy[:, :, None] * gradphi[:, :, :, 1]  # grad a_m . grad phi_j: (nc, nq, nd)
        loc = -np.einsum("c,q,qi,cqj->cij", vol, w, phi, conv)
        np.add.at(vals[m], pos.reshape(-1), loc.reshape(-1))
    return indptr, indices, vals


def lambda_mu(coeff, multi_indices, x, y, factor=-1.0, n_truncate=None):
    """lambda_mu of expa_PCE_mop (src/coefficients/coefficients.jl:236-262) for e^{a * factor} at points (x, y):
    exp(1/2 sum_{m <= N_truncate} a_m^2) * exp(mean * factor) * prod_d a_d^{mu_d} / (sqrt(prod_d mu_d!) * factor^{|mu|}).
    Returns an array of shape (nmodes,) + x.shape."""
    from math import factorial
    n_truncate = coeff.maxm if n_truncate is None else n_truncate
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    s = np.zeros(np.broadcast(x, y).shape)
    for m in range(1, n_truncate + 1):
        s = s + coeff.am(m, x, y) ** 2
    pref = np.exp(s / 2) * np.exp(coeff.mean_value * factor)
    M = len(multi_indices[0])
    am = [coeff.am(d, x, y) for d in range(1, M + 1)]
    out = np.empty((len(multi_indices),) + pref.shape)
    for k, mu in enumerate(multi_indices):
        val = np.ones_like(pref)
        fac = 1.0
        for d in range(M):
            val = val * am[d] ** mu[d]
            fac *= factorial(mu[d])
        out[k] = val / (np.sqrt(fac) * factor ** sum(mu)) * pref
    return out


def assemble_logprimal_rhs(space: FESpace, coeff, multi_indices, f, bonus_quadorder: int = 0):
    """b[mu] = (lambda_mu f, phi_i) for every mode (logpoisson_primal.jl:108-127, LinearOperator(kernel_fexp, [id(1)];
    bonus_quadorder = bonus_quadorder_f)).  Returns (nmodes, ndofs)."""
    mesh = space.mesh
    xref, w = quadrature_rule(space.order + bonus_quadorder)
    phi, _ = space.basis(xref)
    xq = space.physical_points(xref)
    lam = lambda_mu(coeff, multi_indices, xq[:, :, 0], xq[:, :, 1])  # (nmodes, nc, nq)
    fv = f(xq[:, :, 0], xq[:, :, 1])
    loc = np.einsum("c,q,cq,kcq,qi->kci", mesh.cellvolumes, w, fv, lam, phi)
    b = np.zeros((len(multi_indices), space.ndofs))
    for k in range(len(multi_indices)):
        np.add.at(b[k], space.celldofs.reshape(-1), loc[k].reshape(-1))
    return b
