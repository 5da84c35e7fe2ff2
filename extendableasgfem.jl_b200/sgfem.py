"""Host-side mirror of the reference's operator interface for the solve path, same names and argument meaning:

* TensorizedBasis        src/tensorizedbasis.jl:28-34 (multi_indices, nmodes, G)
* SGFEVector             src/sgfevector.jl:18-27 (flat `entries`, block view per mode via sol[j])
* solve_primal           solve_primal!(sol, A0, Am, b0, G, nmodes, bfac; atol, rtol)
                         src/modelproblems/solvers_poisson_primal.jl:130-169
* solve                  solve!(PoissonProblemPrimal, sol, C; rhs, ...)   src/modelproblems/poisson_primal.jl:38-81
* solve_logpoisson_primal  solve_logpoisson_primal!(sol, A, N0, Nm, b0, G, nmodes, bfac; atol, rtol)
                         src/modelproblems/solvers_logpoisson_primal.jl:130-172 (matrices and load vectors from the caller)
* set_samples            set_sample!(SGFEV, S) for a batch of samples   src/sgfevector.jl:43-69
* mul / ldiv             LinearAlgebra.mul! (:86-124) / ldiv! (:46-78) on host vectors
* estimate               estimate(PoissonProblemPrimal, sol, C; rhs, bonus_quadorder, tail_extension)
                         src/estimate.jl:260-418

All arithmetic runs in libasgfem_cuda.so; nothing here computes on the CPU except index bookkeeping.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

from . import context as _ctx
from . import grids as _grids
from . import multiindices as _mi

LegendrePolynomials = _ctx.LEGENDRE
HermitePolynomials = _ctx.HERMITE


class TensorizedBasis:
    def __init__(self, OBT, multi_indices, ctx: _ctx.Context | None = None):
        self.OBT = OBT
        self.multi_indices = [list(m) for m in multi_indices]
        _mi.prepare_multi_indices(self.multi_indices)
        self.nmodes = len(self.multi_indices)
        self.ctx = ctx or _ctx.Context()
        self.ctx.set_multiindices(OBT, np.array(self.multi_indices, dtype=np.int64))
        self._G = None

    def maxlength_multiindices(self):
        return len(self.multi_indices[0])

    @property
    def G(self):
        """(M*N) x N CSC exactly like TensorizedBasis.G.cscmatrix (built by the native library)."""
        if self._G is None:
            colptr, rowval, nzval = self.ctx.coupling_csc()
            M = self.maxlength_multiindices()
            self._G = sp.csc_matrix((nzval, rowval - 1, colptr - 1), shape=(M * self.nmodes, self.nmodes))
        return self._G

    def get_coupling_coefficient(self, m, j, k):  # 1-based like tensorizedbasis.jl:74
        return self.G[(m - 1) * self.nmodes + j - 1, k - 1]


class SGFEVector:
    def __init__(self, FES: _grids.FESpace, TB: TensorizedBasis):
        self.FES_space = FES
        self.TB = TB
        self.entries = np.zeros(FES.ndofs * TB.nmodes)

    def __getitem__(self, j):  # 1-based mode id -> block view (sgfevector.jl:118)
        n = self.FES_space.ndofs
        return self.entries[(j - 1) * n: j * n]

    def __len__(self):
        return len(self.entries)


def _install_matrices(ctx, A0, Am, extra=()):
    """A0, Am: scipy sparse matrices as a Julia caller holds FEMatrix.entries.cscmatrix; `extra`: further matrices whose
    entries must lie in the shared pattern (e.g. the preconditioner matrix of the log-transformed problem)."""
    mats = [sp.csc_matrix(A0)] + [sp.csc_matrix(A) for A in Am]
    n = mats[0].shape[0]
    union = mats[0].copy()
    union.data = np.ones_like(union.data)
    for A in mats[1:] + [sp.csc_matrix(E) for E in extra]:
        B = A.copy()
        B.data = np.ones_like(B.data)
        union = union + B
    union = sp.csc_matrix(union)
    union.sort_indices()
    ctx.set_pattern_csc(n, union.indptr.astype(np.int64) + 1, union.indices.astype(np.int64) + 1)
    ctx.set_num_stiffness(len(Am))
    for m, A in enumerate(mats):
        A.sort_indices()
        ctx.set_stiffness_csc(m, A.indptr.astype(np.int64) + 1, A.indices.astype(np.int64) + 1, A.data)


def solve_primal(sol: SGFEVector, A0, Am, b0, G=None, nmodes=None, bfac=1, atol=1.0e-14, rtol=1.0e-14,
                 itmax=0, return_stats=False):
    """Drop-in for solve_primal!: overwrites sol.entries, returns bdofs (1-based, first-occurrence order)."""
    ctx = sol.TB.ctx
    _install_matrices(ctx, A0, Am)
    bdofs = sol.FES_space.bdofs + 1
    ctx.set_bdofs(bdofs)
    stats = ctx.solve_primal_host(sol.entries, np.asarray(b0, dtype=np.float64), atol, rtol, itmax)
    return (bdofs, stats) if return_stats else bdofs


def solve_logpoisson_primal(sol: SGFEVector, A, N0, Nm, b0, G=None, nmodes=None, bfac=1, atol=1.0e-14, rtol=1.0e-14,
                            itmax=0, return_stats=False):
    """Drop-in for solve_logpoisson_primal! (log-transformed primal problem): A = Laplacian, N0 / Nm = convection
    matrices (scipy sparse, as the Julia caller holds FEMatrix.entries.cscmatrix), b0 = list of nmodes load vectors.
    Overwrites sol.entries, returns bdofs.  The operator kernels are those of the primal problem with A + N0 as matrix 0
    and N_e as matrices 1..M; the preconditioner is factorised from A alone; the Krylov method is BiCGStab on the device
    instead of the reference's GMRES (same solution of the nonsingular system)."""
    ctx = sol.TB.ctx
    A = sp.csc_matrix(A)
    _install_matrices(ctx, A + sp.csc_matrix(N0), Nm, extra=(A,))
    A.sort_indices()
    ctx.set_precond_matrix_csc(A.indptr.astype(np.int64) + 1, A.indices.astype(np.int64) + 1, A.data)
    bdofs = sol.FES_space.bdofs + 1
    ctx.set_bdofs(bdofs)
    b = np.concatenate([np.asarray(v, dtype=np.float64).reshape(-1) for v in b0])
    stats = ctx.solve_logprimal_host(sol.entries, b, atol, rtol, itmax)
    return (bdofs, stats) if return_stats else bdofs


def solve_logpoisson(sol: SGFEVector, C, rhs, bonus_quadorder_a=2, bonus_quadorder_f=0, atol=1.0e-14, rtol=1.0e-14, itmax=0,
                     return_stats=False):
    """solve!(LogTransformedPoissonProblemPrimal, sol, C; rhs, ...) (src/modelproblems/logpoisson_primal.jl:75-142) with the
    Laplacian A, the convection matrices N_m and the load vectors b[mu] = (lambda_mu f, v) assembled on the device, and the
    system of solve_logpoisson_primal! solved by the device BiCGStab with the preconditioner I (x) chol(A).  The warm start
    semantics of the reference (b = deepcopy(sol) + b0, :149-152) are kept.  Overwrites sol.entries, returns bdofs."""
    if rhs is None:
        raise ValueError("need right-hand side")  # logpoisson_primal.jl:129
    ctx, FES = sol.TB.ctx, sol.FES_space
    g = FES.grid
    ctx.set_mesh(g.coords, g.cellnodes + 1)
    ctx.set_space(FES.order, FES.ndofs, FES.celldofs + 1)
    ctx.set_coefficient_cosinus(C.mean_value, C.decay_factors, C.b1, C.b2)
    xref, w = _grids.quadrature_rule(2 * FES.order - 1 + bonus_quadorder_a)
    ctx.assemble_logprimal(sol.TB.maxlength_multiindices(), xref, w)
    bdofs = FES.bdofs + 1
    ctx.set_bdofs(bdofs)
    xf, wf = _grids.quadrature_rule(FES.order + bonus_quadorder_f)
    c = g.cellnodes
    x1, x2, x3 = g.coords[c[:, 0]], g.coords[c[:, 1]], g.coords[c[:, 2]]
    xq = x1[:, None, :] + xf[None, :, 0:1] * (x2 - x1)[:, None, :] + xf[None, :, 1:2] * (x3 - x1)[:, None, :]
    fq = rhs(xq[:, :, 0], xq[:, :, 1])  # (ncells, nq) C-order == nq x ncells column-major
    ctx.vec_ensure(2)
    ctx.vec_upload(0, sol.entries)
    ctx.assemble_logprimal_rhs(xf, wf, fq, len(C.decay_factors), 1)
    ctx.vec_axpy(1.0, 0, 1)  # b = deepcopy(sol) + b0   (:149-152)
    stats = ctx.bicgstab(1, 0, atol, rtol, itmax)
    sol.entries[:] = ctx.vec_download(0)
    return (bdofs, stats) if return_stats else bdofs


def set_samples(sol: SGFEVector, vals):
    """Batched set_sample!(SGFEV, S) (src/sgfevector.jl:43-69): vals[s, m, :] = TB.vals[m] after set_sample!(TB, xi_s)
    (tensorizedbasis.jl:226-236, the caller's polynomial tables).  Returns the (nsamples, n) array of evaluated spatial
    coefficient vectors (row s = SGFEV.FEV entries for sample s); the sum over the modes runs on the device."""
    ctx = sol.TB.ctx
    ctx.vec_ensure(1)  # grow-only: the caller's other device vectors survive (slot 0 is overwritten)
    ctx.vec_upload(0, sol.entries)
    return ctx.evaluate_samples(0, vals)


def setup_device_problem(sol: SGFEVector, C, bonus_quadorder_a=2):
    """Device-side part of solve!: mesh/space/coefficient upload and assembly of K_0..K_M on the device."""
    ctx, FES = sol.TB.ctx, sol.FES_space
    g = FES.grid
    ctx.set_mesh(g.coords, g.cellnodes + 1)
    ctx.set_space(FES.order, FES.ndofs, FES.celldofs + 1)
    ctx.set_coefficient_cosinus(C.mean_value, C.decay_factors, C.b1, C.b2)
    xref, w = _grids.quadrature_rule(2 * (FES.order - 1) + bonus_quadorder_a)
    ctx.assemble_stiffness(sol.TB.maxlength_multiindices(), xref, w)
    ctx.set_bdofs(FES.bdofs + 1)


def solve(sol: SGFEVector, C, rhs=None, bonus_quadorder_a=2, bonus_quadorder_f=0, atol=1.0e-14, rtol=1.0e-14,
          itmax=0, return_stats=False):
    """solve!(PoissonProblemPrimal, sol, C; rhs, ...) with K_m assembled on the device."""
    setup_device_problem(sol, C, bonus_quadorder_a)
    b0 = sol.FES_space.rhs(rhs, bonus_quadorder_f)
    stats = sol.TB.ctx.solve_primal_host(sol.entries, b0, atol, rtol, itmax)
    bdofs = sol.FES_space.bdofs + 1
    return (bdofs, stats) if return_stats else bdofs


def mul(ctx: _ctx.Context, x):
    """mul!(Ax, S, x) on host vectors in the reference layout."""
    return ctx.apply_host(x)


def ldiv(ctx: _ctx.Context, b):
    """ldiv!(y, P, b) on host vectors in the reference layout."""
    return ctx.precond_apply_host(b)


def estimate(sol: SGFEVector, C, rhs=None, bonus_quadorder=1, tail_extension=(10, 2), marking_columns=None):
    """Returns (eta4modes, eta4cell, multi_indices_extended) like estimate.jl:417.  Requires the device problem
    (setup_device_problem / solve) so that mesh, space and coefficient are resident.  With marking_columns (1-based column
    ids, e.g. the active modes) the second return value is the per-cell sum of eta4cell over these columns - what the
    adaptive loop hands to bulk_mark (scripts/poisson.jl:402) - and the ncells x N_ext matrix never leaves the device."""
    if rhs is None:  # the reference calls rhs(ftemp, x) unconditionally (estimate.jl:322): no silent default
        raise ValueError("estimate: the right-hand side function rhs(x, y) is required")
    ctx, FES = sol.TB.ctx, sol.FES_space
    g = FES.grid
    mi_ext = _mi.add_boundary_modes(sol.TB.multi_indices, tail_extension=tail_extension)
    quadorder = 2 * (FES.order - 1) + bonus_quadorder
    xref, w = _grids.quadrature_rule(quadorder)
    sf, wf = _grids.quadrature_rule_1d(quadorder)
    fq = None
    if rhs is not None:
        c = g.cellnodes
        x1, x2, x3 = g.coords[c[:, 0]], g.coords[c[:, 1]], g.coords[c[:, 2]]
        xq = x1[:, None, :] + xref[None, :, 0:1] * (x2 - x1)[:, None, :] + xref[None, :, 1:2] * (x3 - x1)[:, None, :]
        fq = rhs(xq[:, :, 0], xq[:, :, 1])  # (ncells, nq) C-order == nq x ncells column-major
    ctx.vec_ensure(1)  # grow-only: the caller's other device vectors survive (slot 0 is overwritten)
    ctx.vec_upload(0, sol.entries)
    if marking_columns is not None:
        eta4modes, cellsum = ctx.estimate_poisson_primal_marking(0, np.array(mi_ext, dtype=np.int64), xref, w, sf, wf,
                                                                 g.ncells, marking_columns, fq)
        return eta4modes, cellsum, mi_ext
    eta4modes, eta4cell = ctx.estimate_poisson_primal(0, np.array(mi_ext, dtype=np.int64), xref, w, sf, wf,
                                                      g.ncells, fq)
    return eta4modes, eta4cell, mi_ext


def estimate_logpoisson(sol: SGFEVector, C, rhs, bonus_quadorder=1, tail_extension=(5, 2), lambda_at_qp=None):
    """estimate(LogTransformedPoissonProblemPrimal, sol, C; rhs, bonus_quadorder, tail_extension) (src/estimate.jl:70-257):
    returns (eta4modes, eta4cell, multi_indices_extended, zeta_data).  tail_extension is the two-element form the scripts
    pass (the reference's scalar default 5 would fail at tail_extension[2] in add_boundary_modes).  Mesh, space and coefficient must be resident
    (solve_logpoisson).  lambda_at_qp (N_ext, ncells, nq): the caller's interpolated <e^-a, H_nu> at the quadrature points
    (what the reference's H1Pk{quadorder} interpolation yields); None evaluates lambda_nu directly on the device."""
    if rhs is None:
        raise ValueError("estimate: the right-hand side function rhs(x, y) is required")
    ctx, FES = sol.TB.ctx, sol.FES_space
    g = FES.grid
    mi_ext = _mi.add_boundary_modes(sol.TB.multi_indices, tail_extension=tail_extension)
    quadorder = 2 * (FES.order - 1) + bonus_quadorder
    xref, w = _grids.quadrature_rule(quadorder)
    sf, wf = _grids.quadrature_rule_1d(max(quadorder, 2 * (FES.order - 1)))
    c = g.cellnodes
    x1, x2, x3 = g.coords[c[:, 0]], g.coords[c[:, 1]], g.coords[c[:, 2]]
    xq = x1[:, None, :] + xref[None, :, 0:1] * (x2 - x1)[:, None, :] + xref[None, :, 1:2] * (x3 - x1)[:, None, :]
    fq = rhs(xq[:, :, 0], xq[:, :, 1])
    lam = None
    if lambda_at_qp is not None:  # (N_ext, ncells, nq) -> [cell][q][j]
        lam = np.ascontiguousarray(np.transpose(np.asarray(lambda_at_qp, dtype=np.float64), (1, 2, 0)))
    ctx.vec_ensure(1)
    ctx.vec_upload(0, sol.entries)
    em, ec, zeta = ctx.estimate_logpoisson_primal(0, np.array(mi_ext, dtype=np.int64), xref, w, sf, wf, g.ncells, fq,
                                                  len(C.decay_factors), lam)
    return em, ec, mi_ext, float(zeta[0])


def deterministic_sample_solutions(FES, C, samples, rhs=None, bonus_quadorder_a=2, device=0, atol=1e-14, rtol=1e-14):
    """The deterministic reference solutions of calculate_sampling_error (src/sampling_error.jl:112-128) for the affine
    coefficient: for every column xi of `samples` (Msamples x nsamples) solve  -div((a_0 + sum_m xi_m a_m) grad u) = f  on
    the space FES.  The reference runs ExtendableFEM.solve per sample on host threads; here all samples are the columns of
    one device block system (asgfem_set_samples / asgfem_solve_samples_host).  Returns (u (ndofs x nsamples), stats)."""
    samples = np.asarray(samples, dtype=np.float64)
    Ms = samples.shape[0]
    ctx = _ctx.Context(device)
    try:
        g = FES.grid
        ctx.set_multiindices(0, np.zeros((1, max(Ms, 1)), dtype=np.int64))  # placeholder set: sizes the matrices
        ctx.set_mesh(g.coords, g.cellnodes + 1)
        ctx.set_space(FES.order, FES.ndofs, FES.celldofs + 1)
        ctx.set_coefficient_cosinus(C.mean_value, C.decay_factors, C.b1, C.b2)
        xref, w = _grids.quadrature_rule(2 * (FES.order - 1) + bonus_quadorder_a)
        ctx.assemble_stiffness(Ms, xref, w)
        ctx.set_bdofs(FES.bdofs + 1)
        ctx.set_samples(samples)
        b = FES.rhs(rhs)
        b[FES.bdofs] = 0.0
        return ctx.solve_samples_host(b, atol, rtol)
    finally:
        ctx.close()
