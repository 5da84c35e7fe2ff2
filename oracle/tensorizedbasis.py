"""Oracle: TensorizedBasis, the coupling matrix G and its triple products.

Test infrastructure only (see oracle/__init__.py).  Restates

* src/tensorizedbasis.jl:165-181 (constructor), 183-191 (triple_product),
  193-218 (get_tensor_multiplication_with_ym -> G), 74 (get_coupling_coefficient)

triple_product is PINNED by runtests.jl:104-127.  G is unpinned by the reference's tests; the
derivable pin G[(m-1)N+j,k] == triple_product_y(ONB, mu_j[m], mu_k[m]) * prod_{d!=m} delta is
checked in tests/test_oracle_basis.py.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

from . import multiindices as mi_mod
from . import polynomials as poly


def coupling_matrix(family: int, multi_indices):
    """G as scipy CSC of shape (M*N, N) with sorted row indices per column, i.e. exactly the
    `cscmatrix` of the flushed ExtendableSparseMatrix{Float64,Int64} (tensorizedbasis.jl:198-216).
    Entry [(m-1)*N + j, k] (1-based) = 1/b if mu_k = mu_j + e_m, c/b if mu_k = mu_j - e_m, with
    (a,b,c) = normalise_recurrence_coefficients(OBT, mu_j[m])."""
    N = len(multi_indices)
    M = max(len(m) for m in multi_indices)
    PLUS, MINUS = mi_mod.get_neighbours(multi_indices)
    maxdeg = max(max(m) for m in multi_indices)
    gp = np.zeros(maxdeg + 2)
    gm = np.zeros(maxdeg + 2)
    for k in range(maxdeg + 1):
        gp[k], gm[k] = poly.coupling_weights(family, k)
    rows, cols, vals = [], [], []
    for j in range(N):
        for m in range(M):
            deg = multi_indices[j][m]
            if PLUS[m, j] > 0:
                rows.append(m * N + j)
                cols.append(PLUS[m, j] - 1)
                vals.append(gp[deg])
            if MINUS[m, j] > 0:
                rows.append(m * N + j)
                cols.append(MINUS[m, j] - 1)
                vals.append(gm[deg])
    G = sp.csc_matrix((vals, (rows, cols)), shape=(M * N, N), dtype=np.float64)
    G.sort_indices()
    return G


class TensorizedBasis:
    """tensorizedbasis.jl:28-34; `multi_indices="full"` generates the full tensor set (172-174)."""

    def __init__(self, family, M, order, maxorder, maxquadorder=None, multi_indices="full"):
        assert order <= maxorder
        if maxquadorder is None:
            maxquadorder = 2 * maxorder
        self.family = family
        self.ONB = poly.ONBasis(family, maxorder, maxquadorder)
        if isinstance(multi_indices, str) and multi_indices == "full":
            multi_indices = mi_mod.generate_multiindices(M, order)
        self.multi_indices = multi_indices
        self.nmodes = len(multi_indices)
        self.G = coupling_matrix(family, multi_indices)

    def maxlength_multiindices(self):
        return max(len(m) for m in self.multi_indices)

    def get_coupling_coefficient(self, m, j, k):  # 1-based like the reference (:74)
        return self.G[(m - 1) * self.nmodes + j - 1, k - 1]

    def triple_product(self, j, k, l, normalize=True):  # 1-based mode ids (:183-191)
        val = 1.0
        mi = self.multi_indices
        for d in range(self.maxlength_multiindices()):
            val *= self.ONB.triple_product(mi[j - 1][d], mi[k - 1][d], mi[l - 1][d], normalize=normalize)
        return val

    def evaluate_all(self, sample, normalize=True):
        """H_mu(sample) for all mu (set_sample!/evaluate, tensorizedbasis.jl:226-252)."""
        M = self.maxlength_multiindices()
        vals = []
        for d in range(M):
            if d < len(sample):
                vals.append(self.ONB.evaluate(sample[d], normalize=normalize))
            else:
                v = np.zeros(self.ONB.maxorder + 1)
                v[0] = 1
                vals.append(v)
        out = np.ones(self.nmodes)
        for j, mu in enumerate(self.multi_indices):
            for d in range(len(mu)):
                out[j] *= vals[d][mu[d]]
        return out
