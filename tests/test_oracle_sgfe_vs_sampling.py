"""Derivable pin of the Poisson SGFE path of the oracle (rows a1-a9 of SURVEY.md section 8) that does not depend on the
restatement itself: for the affine coefficient a(x, xi) = a_0(x) + sum_m xi_m a_m(x), xi ~ U(-1, 1)^M, the stochastic
Galerkin solution sum_mu u_mu(x) H_mu(xi) in the total-degree-p Legendre space converges spectrally to the deterministic
finite element solution of (K_0 + sum_m xi_m K_m) u = b at any sample xi.  A wrong triple-product weight, polynomial
normalisation, block ordering or right-hand-side placement in oracle/tensorizedbasis.py / oracle/solver.py breaks the
convergence at the first degree.  The checker is one sparse direct solve per sample (scipy)."""
import numpy as np
import scipy.sparse.linalg as spla

from oracle import coefficient as ocoef
from oracle import mesh as omesh
from oracle import multiindices as omi
from oracle import polynomials as opoly
from oracle import problem as oproblem
from oracle import solver as osolver


def _total_degree_set(M, p):
    out = [[0] * M]
    frontier = [[0] * M]
    for _ in range(p):
        nxt = []
        for mu in frontier:
            for m in range(M):
                nu = list(mu)
                nu[m] += 1
                if nu not in out:
                    out.append(nu)
                    nxt.append(nu)
        frontier = nxt
    return out


def _sgfe_error(order, M, p, samples, family=opoly.LEGENDRE):
    mesh = omesh.uniform_refine(omesh.grid_unitsquare(), 2)
    coeff = ocoef.StochasticCoefficientCosinus(tau=0.9, decay=2.0, mean=1.0, maxm=M)
    P = oproblem.build(mesh, order, _total_degree_set(M, p), family, coeff)
    sol = np.zeros(P.n * P.N)
    st = osolver.solve_primal(sol, P.A0, P.Am, P.b0, P.G, P.N, P.bdofs, method="pcg")
    assert st["solved"]
    U = sol.reshape(P.N, P.n)
    interior = np.setdiff1d(np.arange(P.n), P.bdofs)
    worst = 0.0
    # oracle.polynomials.evaluate gives the un-normalised polynomials; their norms are taken from a Gauss rule here
    t, w = np.polynomial.legendre.leggauss(p + 2)
    nrm = np.sqrt((opoly.evaluate(family, p, t) ** 2 * (w / 2.0)[:, None]).sum(axis=0))
    for xi in samples:
        H = [opoly.evaluate(family, p, float(x)) / nrm for x in xi]  # orthonormal H_0..H_p at xi_m
        u_sg = sum(np.prod([H[m][mu[m]] for m in range(M)]) * U[k] for k, mu in enumerate(P.multi_indices))
        K = (P.A0 + sum(float(xi[m]) * P.Am[m] for m in range(M))).tocsc()
        u_det = np.zeros(P.n)
        u_det[interior] = spla.splu(K[interior][:, interior]).solve(P.b0[interior])
        worst = max(worst, np.abs(u_sg - u_det).max() / np.abs(u_det).max())
    return worst


def test_sgfe_solution_converges_spectrally_to_the_sample_solutions():
    rng = np.random.default_rng(3)
    samples = np.vstack([rng.uniform(-1, 1, size=(4, 2)), [[0.95, -0.95], [-1.0, 1.0]]])
    errs = [_sgfe_error(1, 2, p, samples) for p in (1, 3, 5, 7)]  # observed 5.7e-2, 4.9e-3, 3.9e-4, 3e-5 (worst: the corner)
    assert errs[0] < 0.1 and all(errs[k + 1] < errs[k] / 8 for k in range(3)), errs
    assert errs[3] < 1e-4, errs


def test_sgfe_mean_is_the_first_block_and_variance_is_parseval():
    """E[u] = u_0 and Var[u] = sum_{mu != 0} u_mu^2 for an orthonormal basis: checked against a tensor Gauss-Legendre
    quadrature of the sample solutions (exact for the polynomial SGFE solution, spectrally accurate for the sample solves)."""
    M, p, order = 2, 5, 2
    mesh = omesh.uniform_refine(omesh.grid_unitsquare(), 1)
    coeff = ocoef.StochasticCoefficientCosinus(tau=0.9, decay=2.0, mean=1.0, maxm=M)
    P = oproblem.build(mesh, order, _total_degree_set(M, p), opoly.LEGENDRE, coeff)
    sol = np.zeros(P.n * P.N)
    assert osolver.solve_primal(sol, P.A0, P.Am, P.b0, P.G, P.N, P.bdofs, method="pcg")["solved"]
    U = sol.reshape(P.N, P.n)
    t, w = np.polynomial.legendre.leggauss(8)
    w = w / 2.0  # uniform density on (-1, 1)
    pts = np.array([[a, b] for a in t for b in t])
    wts = np.array([wa * wb for wa in w for wb in w])
    det = osolver.deterministic_sample_solutions(P.A0, P.Am, P.b0, P.bdofs, pts.T)  # (n, nsamples)
    det = det if det.shape[0] == P.n else det.T
    mean = det @ wts
    var = (det ** 2) @ wts - mean ** 2
    scale = np.abs(mean).max()
    assert np.abs(U[0] - mean).max() < 1e-6 * scale
    assert np.abs((U[1:] ** 2).sum(axis=0) - var).max() < 1e-6 * scale ** 2
