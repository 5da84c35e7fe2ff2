"""Oracle: the tensorized SGFE operator, the mean-based preconditioner and the Krylov drivers.

Test infrastructure only (see oracle/__init__.py).  Restates
src/modelproblems/solvers_poisson_primal.jl:

* 30-44  MyPreconditionerPrimal (1e60 penalty on the boundary diagonal of the SHARED A0, then LU)
* 46-78  ldiv!  (one solve with the factor per mode block)
* 86-124 mul!   (Ax[mu] = A0 x[mu] + sum_{nu,e} G[(e-1)N+mu,nu] A_e x[nu]; boundary rows zeroed)
* 130-169 solve_primal! (rhs, Krylov.gmres(...; ldiv=true, atol, rtol, M=P), residual check)
* 172-230 solve_full_primal! (assembled block matrix + direct solve - second definition of the operator)

Krylov.gmres itself is third-party (Krylov.jl 0.10.1, not vendored); `gmres` below restates its
published algorithm (left-preconditioned, MGS Arnoldi, Givens rotations, stop when
||M^-1 r_k|| <= atol + rtol ||M^-1 r_0||).  PARITY UNPINNED by the reference's tests; the three
definitions (mul!, assembled matrix, direct solve) cross-check each other in tests/.

Vectors are flat length n*N arrays in the reference layout: block mu = x[mu*n:(mu+1)*n]
(column-major n x N, sgfevector.jl:97-101).
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla


class SystemPrimal:
    """MySystemPrimal (solvers_poisson_primal.jl:14-21)."""

    def __init__(self, A0, Am, G, bdofs, nmodes):
        self.A0 = A0.tocsr()
        self.Am = [A.tocsr() for A in Am]
        self.G = G.tocsr()  # (M*N, N)
        self.bdofs = np.asarray(bdofs, dtype=np.int64)  # 0-based here
        self.nmodes = nmodes
        self.n = A0.shape[0]

    def mul(self, x):
        """mul!(Ax, S, x): same loop nest and accumulation order as :101-122 (only the nonzero G
        entries are visited; they are visited nu-major, e-minor like `for nu in 1:N, e in 1:M`)."""
        n, N = self.n, self.nmodes
        M = len(self.Am)
        Ax = np.zeros(n * N)
        G = self.G
        for mu in range(N):
            blk = slice(mu * n, (mu + 1) * n)
            Ax[blk] += self.A0 @ x[blk]
            entries = []
            for e in range(M):
                row = e * N + mu
                for p in range(G.indptr[row], G.indptr[row + 1]):
                    entries.append((G.indices[p], e, G.data[p]))
            for nu, e, g in sorted(entries):
                if g != 0:
                    Ax[blk] += g * (self.Am[e] @ x[nu * n:(nu + 1) * n])
            Ax[mu * n + self.bdofs] = 0
        return Ax

    def assembled(self, penalty=None):
        """bigS of solve_full_primal! (:183-201); with `penalty` the 1e60 boundary diagonal of :215-219,
        otherwise the operator exactly as mul! applies it (boundary rows zeroed)."""
        n, N = self.n, self.nmodes
        M = len(self.Am)
        blocks = [[None] * N for _ in range(N)]
        for j in range(N):
            blocks[j][j] = self.A0.copy()
        Gc = self.G.tocoo()
        for r, k, g in zip(Gc.row, Gc.col, Gc.data):
            e, j = divmod(r, N)
            if abs(g) > 1.0e-14:
                blocks[j][k] = g * self.Am[e] if blocks[j][k] is None else blocks[j][k] + g * self.Am[e]
        S = sp.bmat(blocks, format="lil")
        rows = (np.arange(N)[:, None] * n + self.bdofs[None, :]).reshape(-1)
        if penalty is not None:
            S[rows, rows] = penalty
            return S.tocsr()
        S = S.tocsr()
        mask = np.ones(n * N)
        mask[rows] = 0
        return sp.diags(mask) @ S


class PreconditionerPrimal:
    """MyPreconditionerPrimal: LU of A0 with 1e60 on the boundary diagonal (:30-44), applied block by
    block (:56-76).  UMFPACK's LU is replaced by SuperLU - any exact factorisation is equivalent to
    rounding."""

    def __init__(self, A0, bdofs, nmodes):
        A = A0.tolil(copy=True)
        for d in bdofs:
            A[d, d] = 1.0e60
        self.lu = spla.splu(A.tocsc())
        self.n = A0.shape[0]
        self.nmodes = nmodes

    def ldiv(self, b):
        n = self.n
        y = np.empty_like(b)
        for mu in range(self.nmodes):
            y[mu * n:(mu + 1) * n] = self.lu.solve(b[mu * n:(mu + 1) * n])
        return y


def gmres(S, b, x0, P, atol=1.0e-14, rtol=1.0e-14, itmax=0):
    """Full (non-restarted) left-preconditioned GMRES as Krylov.gmres(S, b, x0; ldiv=true, M=P)."""
    nn = b.shape[0]
    itmax = itmax or 2 * nn
    r0 = P.ldiv(b - S.mul(x0))
    beta = np.linalg.norm(r0)
    eps = atol + rtol * beta
    residuals = [beta]
    if beta <= eps:
        return x0.copy(), dict(niter=0, solved=True, residuals=residuals)
    V = [r0 / beta]
    H = []  # columns of the rotated Hessenberg (upper triangular R)
    cs, sn = [], []
    z = [beta]
    k = 0
    solved = False
    while k < itmax:
        w = P.ldiv(S.mul(V[k]))
        h = np.zeros(k + 2)
        for i in range(k + 1):  # modified Gram-Schmidt
            h[i] = np.dot(V[i], w)
            w = w - h[i] * V[i]
        h[k + 1] = np.linalg.norm(w)
        for i in range(k):  # previous rotations
            t = cs[i] * h[i] + sn[i] * h[i + 1]
            h[i + 1] = -sn[i] * h[i] + cs[i] * h[i + 1]
            h[i] = t
        denom = np.hypot(h[k], h[k + 1])
        c, s = (1.0, 0.0) if denom == 0 else (h[k] / denom, h[k + 1] / denom)
        cs.append(c)
        sn.append(s)
        h[k] = c * h[k] + s * h[k + 1]
        hk1 = h[k + 1]
        h[k + 1] = 0.0
        z.append(-s * z[k])
        z[k] = c * z[k]
        H.append(h[:k + 1].copy())
        k += 1
        residuals.append(abs(z[k]))
        if abs(z[k]) <= eps:
            solved = True
            break
        if hk1 <= np.finfo(float).tiny:  # happy breakdown
            solved = True
            break
        V.append(w / hk1)
    y = np.zeros(k)
    for i in range(k - 1, -1, -1):
        acc = z[i]
        for j in range(i + 1, k):
            acc -= H[j][i] * y[j]
        y[i] = acc / H[i][i]
    x = x0.copy()
    for i in range(k):
        x += y[i] * V[i]
    return x, dict(niter=k, solved=solved, residuals=residuals)


def pcg(S, b, x0, P, atol=1.0e-14, rtol=1.0e-14, itmax=1000):
    """Mean-preconditioned CG on the Dirichlet-reduced SPD system (north_star item 4).
    Boundary rows of x are held at the values of x0 there (zero in every script).  Stops when
    sqrt(r.z) <= atol + rtol*sqrt(r0.z0).  Same arithmetic order as the device driver (csrc/pcg.cu)."""
    x = x0.copy()
    r = b - S.mul(x)
    z = P.ldiv(r)
    n, N = S.n, S.nmodes
    rows = (np.arange(N)[:, None] * n + S.bdofs[None, :]).reshape(-1)
    z[rows] = 0
    p = z.copy()
    rz = float(np.dot(r, z))
    rz0 = rz
    hist = [np.sqrt(max(rz, 0.0))]
    k = 0
    eps = atol + rtol * np.sqrt(max(rz0, 0.0))
    while k < itmax and np.sqrt(max(rz, 0.0)) > eps:
        Ap = S.mul(p)
        alpha = rz / float(np.dot(p, Ap))
        x += alpha * p
        r -= alpha * Ap
        z = P.ldiv(r)
        z[rows] = 0
        rz_new = float(np.dot(r, z))
        beta = rz_new / rz
        p = z + beta * p
        rz = rz_new
        k += 1
        hist.append(np.sqrt(max(rz, 0.0)))
    return x, dict(niter=k, solved=np.sqrt(max(rz, 0.0)) <= eps, residuals=hist)


def make_rhs(sol0, b0, bdofs, n, N):
    """solvers_poisson_primal.jl:149-155: b = deepcopy(sol); b[1] += b0; b[m][bdofs] = 0."""
    b = sol0.copy()
    b[:n] += b0
    rows = (np.arange(N)[:, None] * n + np.asarray(bdofs)[None, :]).reshape(-1)
    b[rows] = 0
    return b


def solve_primal(sol, A0, Am, b0, G, nmodes, bdofs, atol=1.0e-14, rtol=1.0e-14, method="gmres"):
    """solve_primal! (:130-169).  `sol` (flat n*N, warm start) is overwritten; returns stats.
    NOTE quirk Q2: the preconditioner constructor puts 1e60 on the diagonal of the SAME A0 the system
    uses (:37-39 before :160); irrelevant after the row zeroing, reproduced here for fidelity."""
    n = A0.shape[0]
    A0p = A0.tolil(copy=True)
    for d in bdofs:
        A0p[d, d] = 1.0e60
    A0p = A0p.tocsr()
    S = SystemPrimal(A0p, Am, G, bdofs, nmodes)
    P = PreconditionerPrimal(A0, bdofs, nmodes)
    b = make_rhs(sol, b0, bdofs, n, nmodes)
    if method == "gmres":
        x, stats = gmres(S, b, sol, P, atol=atol, rtol=rtol)
    else:
        x, stats = pcg(S, b, sol, P, atol=atol, rtol=rtol)
    sol[:] = x
    stats["residual"] = float(np.linalg.norm(S.mul(sol) - b))
    return stats


def solve_full_primal(A0, Am, b0, G, nmodes, bdofs):
    """solve_full_primal! (:172-230): assembled block matrix with 1e60 penalty, direct solve."""
    n = A0.shape[0]
    S = SystemPrimal(A0, Am, G, bdofs, nmodes)
    big = S.assembled(penalty=1.0e60)
    bigb = make_rhs(np.zeros(n * nmodes), b0, bdofs, n, nmodes)
    return spla.spsolve(big.tocsc(), bigb)


# ---- log-transformed primal problem -------------------------------------------------
class SystemLogPrimal(SystemPrimal):
    """MySystemLogPrimal (src/modelproblems/solvers_logpoisson_primal.jl:14-22) and its mul! (:87-127):
    Ax[mu] = A x[mu] + N0 x[mu] + sum_{nu,e} G[(e-1)N+mu,nu] N_e x[nu]; boundary rows zeroed.  Same tensor structure as
    the primal operator with the (nonsymmetric) convection matrices N_e in place of the K_e and A + N0 on the diagonal."""

    def __init__(self, A, N0, Nm, G, bdofs, nmodes):
        super().__init__(A + N0, Nm, G, bdofs, nmodes)
        self.A = A.tocsr()
        self.N0 = N0.tocsr()

    def mul(self, x):
        n, N = self.n, self.nmodes
        M = len(self.Am)
        Ax = np.zeros(n * N)
        G = self.G
        for mu in range(N):
            blk = slice(mu * n, (mu + 1) * n)
            Ax[blk] += self.A @ x[blk]    # :111
            Ax[blk] += self.N0 @ x[blk]   # :112
            entries = []
            for e in range(M):
                row = e * N + mu
                for p in range(G.indptr[row], G.indptr[row + 1]):
                    entries.append((G.indices[p], e, G.data[p]))
            for nu, e, g in sorted(entries):  # for nu in 1:nmodes, e in 1:M (:115)
                if g != 0:
                    Ax[blk] += g * (self.Am[e] @ x[nu * n:(nu + 1) * n])
            Ax[mu * n + self.bdofs] = 0       # :123-125
        return Ax


def make_rhs_log(sol0, b0, bdofs, n, N):
    """solvers_logpoisson_primal.jl:149-156: b = deepcopy(sol); b[m] += b0[m] for EVERY mode; b[m][bdofs] = 0."""
    b = sol0.copy() + np.concatenate([np.asarray(v, dtype=np.float64) for v in b0])
    rows = (np.arange(N)[:, None] * n + np.asarray(bdofs)[None, :]).reshape(-1)
    b[rows] = 0
    return b


def solve_logpoisson_primal(sol, A, N0, Nm, b0, G, nmodes, bdofs, atol=1.0e-14, rtol=1.0e-14):
    """solve_logpoisson_primal! (:130-172): GMRES left-preconditioned with I (x) LU(A) (A with 1e60 on the boundary
    diagonal, :31-45).  `sol` (warm start) is overwritten; returns stats."""
    n = A.shape[0]
    S = SystemLogPrimal(A, N0, Nm, G, bdofs, nmodes)
    P = PreconditionerPrimal(A, bdofs, nmodes)
    b = make_rhs_log(sol, b0, bdofs, n, nmodes)
    x, stats = gmres(S, b, sol, P, atol=atol, rtol=rtol)
    sol[:] = x
    stats["residual"] = float(np.linalg.norm(S.mul(sol) - b))
    return stats


def solve_logpoisson_primal_full(A, N0, Nm, b0, G, nmodes, bdofs):
    """solve_logpoisson_primal_full! (:175-230): assembled block matrix with the 1e60 boundary penalty, direct solve."""
    n = A.shape[0]
    S = SystemLogPrimal(A, N0, Nm, G, bdofs, nmodes)
    big = S.assembled(penalty=1.0e60)
    bigb = make_rhs_log(np.zeros(n * nmodes), b0, bdofs, n, nmodes)
    return spla.spsolve(big.tocsc(), bigb)


def deterministic_sample_solutions(A0, Am, b, bdofs, samples):
    """Deterministic reference solutions of calculate_sampling_error (src/sampling_error.jl:112-128: one
    ExtendableFEM.solve per sample) for the affine coefficient: K(xi) = A0 + sum_m xi_m Am[m], homogeneous Dirichlet data
    on bdofs (the reference's penalty in exact arithmetic: eliminated rows / columns).  samples: (Msamples, nsamples).
    Returns u of shape (n, nsamples)."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    n = A0.shape[0]
    keep = np.ones(n, dtype=bool)
    keep[np.asarray(bdofs)] = False
    out = np.zeros((n, samples.shape[1]))
    for s in range(samples.shape[1]):
        K = sp.csr_matrix(A0, copy=True)
        for m in range(samples.shape[0]):
            K = K + samples[m, s] * Am[m]
        K = sp.csc_matrix(K)[keep][:, keep]
        out[keep, s] = spla.spsolve(K, np.asarray(b)[keep])
    return out
