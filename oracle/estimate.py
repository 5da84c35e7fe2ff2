"""Oracle: residual-based a-posteriori error estimator for PoissonProblemPrimal.

Test infrastructure only (see oracle/__init__.py).  Restates src/estimate.jl:260-418:

* 275-283 extended set, extended basis G, neighbour tables
* 286-287 quadrature order 2(order-1)+bonus_quadorder
* 308-366 volume terms per (cell, mode, qp) incl. the |T|^3 / |T| scaling (quirk Q5: the
  grad(a_m).grad(u_h) part of div(a_m grad u_h) is NOT included, only a_m * Laplace(u_h))
* 371-415 jump terms: sum over m=0..M_ext of a_m [[grad w_{j,m}]] face-interpolated, squared L2 norm,
  zero on boundary faces, *|F| for active modes and /|F| for boundary modes, distributed to the
  3 faces of every cell; eta4modes[j] = sqrt(vol_j + sum_F J_F)

FaceInterpolator / L2NormIntegrator are third-party (ExtendableFEM); assumed semantics (SURVEY.md
B.4, parity unpinned): per face and component, L2 projection of the kernel output onto
P_{order-1}(F) computed with the 1-D rule of `quadorder`, then its squared L2 norm over F.
"""
from __future__ import annotations

import numpy as np

from . import fem
from . import multiindices as mi_mod
from . import tensorizedbasis as tb_mod


def _legendre01(deg, s):
    """Orthonormal Legendre polynomials on [0,1] (unit measure) up to `deg` at points s: (deg+1, nq)."""
    t = 2 * s - 1
    out = [np.ones_like(t)]
    if deg >= 1:
        out.append(np.sqrt(3.0) * t)
    if deg >= 2:
        out.append(np.sqrt(5.0) * 0.5 * (3 * t * t - 1))
    return np.stack(out[: deg + 1])


def estimate_poisson_primal(space, sol, multi_indices, family, coeff, f=None, bonus_quadorder=1,
                            tail_extension=(10, 2)):
    """sol: flat n*N (reference layout).  Returns (eta4modes (N_ext,), eta4cell (ncells, N_ext),
    multi_indices_extended).  `f(x, y)` vectorised rhs (default 1)."""
    mesh = space.mesh
    order = space.order
    n = space.ndofs
    N = len(multi_indices)
    ncells = mesh.ncells
    U = sol.reshape(N, n)  # U[mu] = block mu

    mi_ext = mi_mod.add_boundary_modes([list(m) for m in multi_indices], tail_extension=tail_extension)
    M_ext = len(mi_ext[0])
    N_ext = len(mi_ext)
    G = tb_mod.coupling_matrix(family, mi_ext).tocsr()
    PLUS, MINUS = mi_mod.get_neighbours(mi_ext)

    def g_of(m, j, k):  # 0-based m, j; k 1-based neighbour id
        return G[m * N_ext + j, k - 1]

    # coupling lists per extended mode j: [(m (1-based), neighbour (0-based, < N), g)]
    couplings = []
    for j in range(N_ext):
        lst = []
        for m in range(M_ext):
            for tab in (PLUS, MINUS):
                k = tab[m, j]
                if 0 < k <= N:
                    lst.append((m + 1, k - 1, g_of(m, j, k)))
        couplings.append(lst)

    quadorder = 2 * (order - 1) + bonus_quadorder
    xref, w = fem.quadrature_rule(quadorder)
    xq = space.physical_points(xref)  # (nc, nq, 2)
    vol = mesh.cellvolumes
    cd = space.celldofs
    lap = space.laplacians()  # (nc, nd)
    fq = np.ones(xq.shape[:2]) if f is None else f(xq[:, :, 0], xq[:, :, 1])
    am_q = {m: coeff.am(m, xq[:, :, 0], xq[:, :, 1]) for m in range(M_ext + 1)}

    eta4cell = np.zeros((ncells, N_ext))
    lapU = None
    if order > 1:
        lapU = np.einsum("mcd,cd->mc", U[:, cd], lap)  # Laplace(u_mu) per cell (constant for P2)
    for j in range(N_ext):
        ftemp = np.zeros(xq.shape[:2])
        if j == 0:
            ftemp += fq
        if order > 1:
            if j < N:
                ftemp += am_q[0] * lapU[j][:, None]
            for (m, k, g) in couplings[j]:
                ftemp += am_q[m] * (g * lapU[k])[:, None]
        eta4cell[:, j] = (ftemp ** 2) @ w
        eta4cell[:, j] *= vol ** 3 if j < N else vol
    eta4modes = np.sqrt(eta4cell.sum(axis=0))

    # ---- jumps ---------------------------------------------------------------------------------
    s_q, w_f = fem.quadrature_rule_1d(quadorder)
    interior = np.where(mesh.facecells[:, 1] >= 0)[0]
    fa = mesh.coords[mesh.facenodes[interior, 0]]
    fb = mesh.coords[mesh.facenodes[interior, 1]]
    xf = fa[:, None, :] + s_q[None, :, None] * (fb - fa)[:, None, :]  # (nif, nqf, 2)
    lamg = space.lambda_gradients()
    grads = []  # per side: (N, nif, nqf, 2)
    for side in (0, 1):
        cells = mesh.facecells[interior, side]
        x1 = mesh.coords[mesh.cellnodes[cells, 0]]
        gl = lamg[cells]  # (nif, 3, 2)
        lam = np.einsum("fix,fqx->fqi", gl, xf - x1[:, None, :])
        lam[:, :, 0] += 1.0
        ref = lam[:, :, 1:3]  # reference coords (x=l2, y=l3)
        nif, nqf = ref.shape[:2]
        _, dphi = space.basis(ref.reshape(-1, 2))  # (nif*nqf, nd, 3)
        dphi = dphi.reshape(nif, nqf, -1, 3)
        gphi = np.einsum("fqdl,flx->fqdx", dphi, gl)  # (nif, nqf, nd, 2)
        grads.append(np.einsum("mfd,fqdx->mfqx", U[:, cd[cells]], gphi))
    jump = grads[0] - grads[1]  # (N, nif, nqf, 2)
    am_f = {m: coeff.am(m, xf[:, :, 0], xf[:, :, 1]) for m in range(M_ext + 1)}
    L = _legendre01(order - 1, s_q)  # (order, nqf)
    flen = mesh.facevolumes
    for j in range(N_ext):
        gq = np.zeros(jump.shape[1:])
        if j < N:
            gq += am_f[0][:, :, None] * jump[j]
        for (m, k, g) in couplings[j]:
            gq += am_f[m][:, :, None] * (g * jump[k])
        c = np.einsum("lq,q,fqx->flx", L, w_f, gq)
        jf = np.zeros(mesh.nfaces)
        jf[interior] = flen[interior] * (c ** 2).sum(axis=(1, 2))
        if j < N:
            jf *= flen
        else:
            jf /= flen
        eta4cell[:, j] += jf[mesh.cellfaces].sum(axis=1)
        eta4modes[j] = np.sqrt(eta4modes[j] ** 2 + jf.sum())
    return eta4modes, eta4cell, mi_ext


def estimate_logpoisson_primal(space, sol, multi_indices, family, coeff, f, bonus_quadorder=1, tail_extension=(5, 2),
                               lambda_at_qp=None):
    """estimate(::Type{LogTransformedPoissonProblemPrimal}, ...) (src/estimate.jl:70-257).  Returns (eta4modes, eta4cell,
    multi_indices_extended, (zeta_data, zeta_data1, zeta_data2)).

    * :134-217 volume term per (cell, mode, qp): (lambda_j f + [j <= N, order > 1] Lap u_j + sum_m grad a_m . grad w_{j,m})^2 with
      w_{j,m} = G-weighted active neighbours of j in direction m; scaling |T|^2 (active) / |T| (boundary modes); eta4modes =
      sqrt of the column sums (:219-221);
    * :156-175 zeta_data1 = sum |T| w_q f^2 exp(2 sum_{m <= maxm} a_m^2), zeta_data2 = sum |T| w_q f^2 sum_j lambda_j^2;
    * :232-244 jumps of grad u_j over interior faces times |F| for the active modes, added to the three faces of a cell;
      eta4modes[j] += sqrt(eta4modes[j]^2 + sum(jumps))  ("+=" as written in the reference).
    lambda_j(x_q): the reference evaluates an H1Pk{quadorder} interpolant of lambda_j (third-party interpolate!, parity
    unpinned); `lambda_at_qp` (N_ext, ncells, nq) supplies such values, None evaluates lambda_j directly (the reference's
    commented-out line :184)."""
    mesh = space.mesh
    order = space.order
    n = space.ndofs
    N = len(multi_indices)
    ncells = mesh.ncells
    U = sol.reshape(N, n)
    mi_ext = mi_mod.add_boundary_modes([list(m) for m in multi_indices], tail_extension=tail_extension)
    M_ext = len(mi_ext[0])
    N_ext = len(mi_ext)
    G = tb_mod.coupling_matrix(family, mi_ext).tocsr()
    PLUS, MINUS = mi_mod.get_neighbours(mi_ext)
    couplings = []
    for j in range(N_ext):
        lst = []
        for m in range(M_ext):
            for tab in (PLUS, MINUS):
                k = tab[m, j]
                if 0 < k <= N:
                    lst.append((m + 1, k - 1, G[m * N_ext + j, k - 1]))
        couplings.append(lst)
    quadorder = 2 * (order - 1) + bonus_quadorder
    xref, w = fem.quadrature_rule(quadorder)
    xq = space.physical_points(xref)
    vol = mesh.cellvolumes
    cd = space.celldofs
    fq = f(xq[:, :, 0], xq[:, :, 1])
    lam = fem.lambda_mu(coeff, mi_ext, xq[:, :, 0], xq[:, :, 1]) if lambda_at_qp is None else np.asarray(lambda_at_qp)
    kmL2 = np.zeros(xq.shape[:2])
    for m in range(1, coeff.maxm + 1):
        kmL2 = kmL2 + coeff.am(m, xq[:, :, 0], xq[:, :, 1]) ** 2
    zeta1 = float(np.einsum("cq,cq,q,c->", fq ** 2, np.exp(2 * kmL2), w, vol))
    zeta2 = float(np.einsum("jcq,cq,q,c->", lam ** 2, fq ** 2, w, vol))
    # gradients of the active modes at the quadrature points
    g = space.lambda_gradients()
    _, dphi = space.basis(xref)
    gradphi = np.einsum("qdl,clx->cqdx", dphi, g)  # (nc, nq, nd, 2)
    gradU = np.einsum("mcd,cqdx->mcqx", U[:, cd], gradphi)  # (N, nc, nq, 2)
    lapU = np.einsum("mcd,cd->mc", U[:, cd], space.laplacians()) if order > 1 else None
    gradam = {m: np.stack(coeff.gradam(m, xq[:, :, 0], xq[:, :, 1]), axis=-1) for m in range(1, M_ext + 1)}  # (nc, nq, 2)
    eta4cell = np.zeros((ncells, N_ext))
    for j in range(N_ext):
        ftemp = lam[j] * fq
        if order > 1 and j < N:
            ftemp = ftemp + lapU[j][:, None]
        sig = np.zeros(xq.shape[:2])
        for (m, k, gw) in couplings[j]:
            sig = sig + gw * np.einsum("cqx,cqx->cq", gradam[m], gradU[k])
        eta4cell[:, j] = ((ftemp + sig) ** 2) @ w
        eta4cell[:, j] *= vol ** 2 if j < N else vol
    eta4modes = np.sqrt(eta4cell.sum(axis=0))
    # jumps of grad u_j (active modes), exact face integrals with the 1-D rule
    s_q, w_f = fem.quadrature_rule_1d(max(quadorder, 2 * (order - 1)))
    interior = np.where(mesh.facecells[:, 1] >= 0)[0]
    fa = mesh.coords[mesh.facenodes[interior, 0]]
    fb = mesh.coords[mesh.facenodes[interior, 1]]
    xf = fa[:, None, :] + s_q[None, :, None] * (fb - fa)[:, None, :]
    grads = []
    for side in (0, 1):
        cells = mesh.facecells[interior, side]
        x1 = mesh.coords[mesh.cellnodes[cells, 0]]
        gl = g[cells]
        lamb = np.einsum("fix,fqx->fqi", gl, xf - x1[:, None, :])
        lamb[:, :, 0] += 1.0
        ref = lamb[:, :, 1:3]
        nif, nqf = ref.shape[:2]
        _, dph = space.basis(ref.reshape(-1, 2))
        dph = dph.reshape(nif, nqf, -1, 3)
        gphi = np.einsum("fqdl,flx->fqdx", dph, gl)
        grads.append(np.einsum("mfd,fqdx->mfqx", U[:, cd[cells]], gphi))
    jump = grads[0] - grads[1]
    flen = mesh.facevolumes
    for j in range(N):
        jf = np.zeros(mesh.nfaces)
        jf[interior] = flen[interior] * np.einsum("q,fqx,fqx->f", w_f, jump[j], jump[j])  # integral over the face
        jf *= flen                                                                       # jumps4face .*= FaceVolumes
        eta4cell[:, j] += jf[mesh.cellfaces].sum(axis=1)
        eta4modes[j] += np.sqrt(eta4modes[j] ** 2 + jf.sum())
    return eta4modes, eta4cell, mi_ext, (zeta1 - zeta2, zeta1, zeta2)
