"""CPU checks of the oracle's restatement of the log-transformed primal problem (oracle/fem.py, oracle/estimate.py) against
identities that do not depend on the restatement itself ("derivable pins": the reference has no tests for this path)."""
import numpy as np

from oracle import coefficient as ocoef
from oracle import estimate as oest
from oracle import fem as ofem
from oracle import mesh as omesh
from oracle import polynomials as opoly


def _hermite_normalised(kmax, xi):
    """He_k / sqrt(k!) at the points xi: (kmax + 1, len(xi))."""
    H = [np.ones_like(xi), xi.copy()]
    for k in range(1, kmax):
        H.append(xi * H[k] - k * H[k - 1])
    from math import factorial
    return np.stack([H[k] / np.sqrt(factorial(k)) for k in range(kmax + 1)])


def test_lambda_mu_is_the_pce_coefficient_of_exp_minus_a():
    """expa_PCE_mop (src/coefficients/coefficients.jl:236-262, factor = -1): lambda_mu(x) = E[exp(-a(x, xi)) H_mu(xi)] for
    a = mean + sum_m a_m(x) xi_m, xi ~ N(0, I), H = normalised Hermite polynomials - checked by tensor Gauss-Hermite quadrature."""
    C = ocoef.StochasticCoefficientCosinus(tau=0.8, decay=2.0, mean=0.3, maxm=3)
    modes = [[0, 0, 0], [1, 0, 0], [0, 2, 0], [1, 1, 1], [3, 0, 1]]
    x, y = np.array([0.21, 0.77]), np.array([0.35, 0.6])
    lam = ofem.lambda_mu(C, modes, x, y, factor=-1.0, n_truncate=3)
    t, w = np.polynomial.hermite_e.hermegauss(40)
    w = w / np.sqrt(2 * np.pi)
    Hn = _hermite_normalised(3, t)
    for p in range(len(x)):
        am = [C.am(m, x[p], y[p]) for m in range(1, 4)]
        for k, mu in enumerate(modes):
            val = np.exp(-C.mean_value)
            for d in range(3):
                val *= np.sum(w * np.exp(-am[d] * t) * Hn[mu[d]])
            assert abs(lam[k][p] - val) <= 1e-12 * max(1.0, abs(val))


def test_logprimal_matrices_identities():
    """A is the symmetric Laplacian with constants in its kernel; N_m annihilates constants (grad of a constant is zero) and
    its row sums give -(grad a_m . grad u, 1) = 0 only through the column sum identity sum_i phi_i = 1:
    sum_i N_m[i, j] = - int grad a_m . grad phi_j."""
    m = omesh.uniform_refine(omesh.grid_unitsquare(), 2)
    C = ocoef.StochasticCoefficientCosinus(tau=0.5, decay=2.0, mean=0.0, maxm=6)
    for order in (1, 2):
        space = ofem.FESpace(m, order)
        ip, idx, vals = ofem.assemble_logprimal_matrices(space, C, 3, bonus_quadorder=4)
        n = space.ndofs
        A = ofem.csr(ip, idx, vals[0], n)
        assert abs(A - A.T).max() < 1e-13
        assert np.abs(A @ np.ones(n)).max() < 1e-12
        # quadrature of - grad a_m . grad phi_j over the mesh
        xref, w = ofem.quadrature_rule(2 * order - 1 + 4)
        _, dphi = space.basis(xref)
        gphi = np.einsum("qdl,clx->cqdx", dphi, space.lambda_gradients())
        xq = space.physical_points(xref)
        for k in (1, 2, 3):
            N = ofem.csr(ip, idx, vals[k], n)
            assert np.abs(N @ np.ones(n)).max() < 1e-12
            gx, gy = C.gradam(k, xq[:, :, 0], xq[:, :, 1])
            loc = -np.einsum("c,q,cqj->cj", m.cellvolumes, w, gx[:, :, None] * gphi[..., 0] + gy[:, :, None] * gphi[..., 1])
            colsum = np.zeros(n)
            np.add.at(colsum, space.celldofs.reshape(-1), loc.reshape(-1))
            assert np.abs(np.asarray(N.sum(axis=0)).reshape(-1) - colsum).max() < 1e-12


def test_logprimal_estimator_data_term_is_a_parseval_gap():
    """zeta_data = zeta_data1 - zeta_data2 (src/estimate.jl:156-175, :248) is the Parseval gap of the truncated PCE of exp(-a):
    nonnegative with exact lambda_nu, and shrinking when the multi-index set grows."""
    m = omesh.uniform_refine(omesh.grid_unitsquare(), 2)
    space = ofem.FESpace(m, 1)
    C = ocoef.StochasticCoefficientCosinus(tau=0.4, decay=2.0, mean=0.0, maxm=8)
    f = lambda x, y: 1.0 + 0 * x  # noqa: E731
    gaps = []
    for modes in ([[0, 0]], [[0, 0], [1, 0], [0, 1]], [[0, 0], [1, 0], [0, 1], [2, 0], [1, 1], [0, 2]]):
        u = np.zeros(len(modes) * space.ndofs)
        _, _, _, zeta = oest.estimate_logpoisson_primal(space, u, modes, opoly.HERMITE, C, f, bonus_quadorder=2, tail_extension=(2, 1))
        assert zeta[0] >= -1e-14 and zeta[2] <= zeta[1] * (1 + 1e-14)
        gaps.append(zeta[0])
    assert gaps[0] > gaps[1] > gaps[2]
