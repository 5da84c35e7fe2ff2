"""Oracle: problem builders for the BASELINE.json configs (test infrastructure only).

* `poisson_simple`  - config 1 = scripts/poisson_simple.jl:18-71 defaults (unit square, nrefs=3, P2,
  modes [[0],[1,0],[0,1],[2,0],[0,0,1]], Legendre, tau=0.9, mean=1, decay=2)
* `synthetic`       - configs 4/5 of SURVEY.md §8(d): structured mesh, cosinus KLE with maxm=M,
  graded-lex multi-indices
"""
from __future__ import annotations

import numpy as np

from . import coefficient, fem, mesh as mesh_mod, multiindices as mi_mod, polynomials as poly, tensorizedbasis as tb_mod


class Problem:
    pass


def build(mesh, order, multi_indices, family, coeff, bonus_quadorder_a=2, f=None, assemble=True):
    P = Problem()
    P.mesh = mesh
    P.space = fem.FESpace(mesh, order)
    P.multi_indices = [list(m) for m in multi_indices]
    mi_mod.prepare_multi_indices(P.multi_indices)
    P.family = family
    P.coeff = coeff
    P.N = len(P.multi_indices)
    P.M = len(P.multi_indices[0])
    P.n = P.space.ndofs
    P.G = tb_mod.coupling_matrix(family, P.multi_indices)
    P.bdofs = P.space.bdofs
    if assemble:
        P.indptr, P.indices, P.vals = fem.assemble_stiffness(P.space, coeff, P.M, bonus_quadorder_a)
        P.A0 = fem.csr(P.indptr, P.indices, P.vals[0], P.n)
        P.Am = [fem.csr(P.indptr, P.indices, P.vals[m], P.n) for m in range(1, P.M + 1)]
        P.b0 = fem.assemble_rhs(P.space, f)
    return P


def poisson_simple(nrefs=3, order=2, domain="square", initial_modes=None, decay=2.0, mean=1.0):
    base = mesh_mod.grid_unitsquare() if domain == "square" else mesh_mod.grid_lshape()
    m = mesh_mod.uniform_refine(base, nrefs)
    modes = initial_modes or [[0], [1, 0], [0, 1], [2, 0], [0, 0, 1]]
    C = coefficient.StochasticCoefficientCosinus(tau=0.9, decay=decay, mean=mean)
    return build(m, order, modes, poly.LEGENDRE, C)


def synthetic(nx, order, N, M=20, assemble=True):
    m = mesh_mod.structured_unitsquare(nx)
    modes = mi_mod.graded_lex_multiindices(M, N)
    C = coefficient.StochasticCoefficientCosinus(tau=0.9, decay=2.0, mean=1.0, maxm=M)
    return build(m, order, modes, poly.LEGENDRE, C, assemble=assemble)


def logprimal_like(nrefs=3, order=2):
    """Test problem with the STRUCTURE of the log-transformed primal system (solvers_logpoisson_primal.jl:14-22): Hermite
    coupling G, SPD Laplacian-like A, nonsymmetric convection-like N0 and N_e on the shared pattern, one load vector per
    mode.  The matrices of the real problem come from the caller's assembly (logpoisson_primal.jl:95-128, third-party
    operators + expa_PCE_mop) and are data at the seam; here they are derived from the cosinus stiffness matrices:
    N_e = 0.5 K_e + 0.25 skew(K_e), N0 = 0.1 skew(K_0), with skew(B) = triu(B, 1) - triu(B, 1)^T."""
    import scipy.sparse as sp
    m = mesh_mod.uniform_refine(mesh_mod.grid_unitsquare(), nrefs)
    C = coefficient.StochasticCoefficientCosinus(tau=0.9, decay=2.0, mean=1.0)
    P = build(m, order, [[0], [1, 0], [0, 1], [2, 0], [1, 1], [0, 0, 1]], poly.HERMITE, C)

    def skew(B):
        U = sp.triu(B, 1)
        return (U - U.T).tocsr()

    P.A = P.A0
    P.N0 = (0.1 * skew(P.A0)).tocsr()
    P.Nm = [(0.5 * K + 0.25 * skew(K)).tocsr() for K in P.Am]
    P.b0m = [P.b0 * (0.5 ** j) * (1 + 0.1 * j) for j in range(P.N)]
    return P


def splitmix64_uniform(idx, seed=20240):
    """x = 2*u01(splitmix64(seed xor idx)) - 1, generated identically on host and device
    (SURVEY.md §8(d)); idx is the flat reference-layout index i + n*mu."""
    z = (np.asarray(idx, dtype=np.uint64) ^ np.uint64(seed)) + np.uint64(0x9E3779B97F4A7C15)
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    z = z ^ (z >> np.uint64(31))
    u = (z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
    return 2.0 * u - 1.0
