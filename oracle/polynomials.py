"""Oracle: orthogonal polynomials, Gauss rules and the ONBasis.

Test infrastructure only (see oracle/__init__.py).  Restates

* src/orthogonal_polynomials/Legendre_uniform.jl:20-21,32
* src/orthogonal_polynomials/Hermite_normal.jl:20-21,32
* src/orthogonal_polynomials/orthogonal_polynomials.jl:26-42 (rc_array_monic),
  63-124 (evaluate), 162-192 (gauss_rule), 211-221 (normalise_recurrence_coefficients)
* src/onbasis.jl:89-207 (ONBasis and its integrals)

PINNED by the known answers of test/runtests.jl:33-101 (tests/test_oracle_polynomials.py).
"""
from __future__ import annotations

import math
from fractions import Fraction

import mpmath
import numpy as np

LEGENDRE = 0
HERMITE = 1

_FAMILY_NAMES = {LEGENDRE: "LegendrePolynomials", HERMITE: "HermitePolynomials"}


def recurrence_coefficients(family: int, k: int):
    """(a, b, c) with P_{k+1} = (a + b x) P_k - c P_{k-1}.

    Legendre_uniform.jl:20-21 returns Rationals, Hermite_normal.jl:20-21 Ints."""
    if family == LEGENDRE:
        return 0, Fraction(2 * k + 1, k + 1), Fraction(k, k + 1)
    if family == HERMITE:
        return 0, 1, k
    raise ValueError("unknown polynomial family")


def norms(family: int, k: int):
    """Hard-coded norms: Legendre_uniform.jl:32 (Float64, Lebesgue norm sqrt(2/(2k+1)),
    quirk Q10) and Hermite_normal.jl:32 (BigFloat sqrt(k!), quirk Q11)."""
    if family == LEGENDRE:
        return math.sqrt(2 / (2 * k + 1))
    if family == HERMITE:
        with mpmath.workprec(256):  # Julia BigFloat default precision
            return mpmath.sqrt(mpmath.mpf(math.factorial(k)))
    raise ValueError("unknown polynomial family")


def _mul(x, y):
    # Julia promotes Rational*Float64 to Float64 by converting the Rational first
    if isinstance(x, Fraction):
        x = x.numerator / x.denominator if not isinstance(y, mpmath.mpf) else mpmath.mpf(x.numerator) / x.denominator
    return x * y


def normalise_recurrence_coefficients(family: int, k: int):
    """orthogonal_polynomials.jl:211-221, evaluated left to right like Julia."""
    a, b, c = recurrence_coefficients(family, k)
    with mpmath.workprec(256):
        h2 = norms(family, k + 1)
        h1 = norms(family, k)
        h0 = norms(family, k - 1) if k > 0 else 0
        return (_mul(a, h1) / h2, _mul(b, h1) / h2, _mul(c, h0) / h2)


def coupling_weights(family: int, k: int):
    """(g_plus, g_minus) = (1/b, c/b) as written into G at tensorizedbasis.jl:209,211,
    rounded to Float64 on assignment into the Float64 sparse matrix."""
    with mpmath.workprec(256):
        _, b, c = normalise_recurrence_coefficients(family, k)
        return float(1 / b), float(c / b)


def evaluate(family: int, n: int, x):
    """First n+1 (un-normalised) polynomials at scalar or vector x
    (orthogonal_polynomials.jl:94-124).  Returns shape (n+1,) or (len(x), n+1)."""
    xv = np.atleast_1d(np.asarray(x, dtype=np.float64))
    y = np.ones((xv.shape[0], n + 1))
    for k in range(n):
        a, b, c = recurrence_coefficients(family, k)
        bf = float(b)
        y[:, k + 1] = (a + bf * xv) * y[:, k]
        if k > 0:
            y[:, k + 1] = y[:, k + 1] - float(c) * y[:, k - 1]
    return y[0] if np.ndim(x) == 0 else y


def rc_array_monic(family: int, n: int):
    """orthogonal_polynomials.jl:31-42."""
    rc = [recurrence_coefficients(family, k) for k in range(n + 1)]
    a = -np.array([float(r[0]) for r in rc])
    b = np.array([float(r[2]) for r in rc])[1:]
    c = np.array([float(r[1]) for r in rc])
    alpha = a / c
    beta = b / (c[:-1] * c[1:])
    return alpha, beta


def gauss_rule(family: int, n: int):
    """n-point Golub-Welsch rule, weights normalised to sum 1
    (orthogonal_polynomials.jl:162-192).  The reference's symmetrisation branch
    `if all(α == 0)` compares a Vector with a scalar and is therefore never taken;
    it is (faithfully) not applied here either."""
    alpha, beta = rc_array_monic(family, n - 1)
    J = np.diag(alpha) + np.diag(np.sqrt(beta), 1) + np.diag(np.sqrt(beta), -1)
    x, V = np.linalg.eigh(J)
    w = V[0, :] ** 2
    w = w / w.sum()
    return x, w


class ONBasis:
    """onbasis.jl:13-18, ctor 89-105."""

    def __init__(self, family: int, maxorder: int, maxquadorder: int | None = None):
        if maxquadorder is None:
            maxquadorder = 2 * maxorder
        self.family = family
        self.maxorder = maxorder
        self.qp, self.qw = gauss_rule(family, maxquadorder)
        self.vals4xref = evaluate(family, maxorder, self.qp)  # (nquad, npoly)
        nrm = np.zeros(maxorder + 1)
        for m in range(maxorder + 1):
            for k in range(len(self.qw)):
                nrm[m] += self.vals4xref[k, m] ** 2 * self.qw[k]
        self.norms = np.sqrt(nrm)

    def norm4poly(self, p):
        return self.norms[p]

    def triple_product(self, j, k, l, normalize=True):  # onbasis.jl:107-127
        v = self.vals4xref
        val = 0.0
        for q in range(len(self.qw)):
            val += v[q, j] * v[q, k] * v[q, l] * self.qw[q]
        if normalize:
            val = val / self.norms[j] / self.norms[k] / self.norms[l]
        return val

    def triple_product_y(self, j, k, normalize=True):  # onbasis.jl:129-149
        v = self.vals4xref
        val = 0.0
        for q in range(len(self.qw)):
            val += v[q, j] * v[q, k] * self.qp[q] * self.qw[q]
        if normalize:
            val = val / self.norms[j] / self.norms[k]
        return val

    def scalar_product(self, j, k, normalize=True):  # onbasis.jl:151-171
        v = self.vals4xref
        val = 0.0
        for q in range(len(self.qw)):
            val += v[q, j] * v[q, k] * self.qw[q]
        if normalize:
            val = val / self.norms[j] / self.norms[k]
        return val

    def integral(self, j, normalize=True):  # onbasis.jl:174-194
        val = 0.0
        for q in range(len(self.qw)):
            val += self.vals4xref[q, j] * self.qw[q]
        if normalize:
            val = val / self.norms[j]
        return val

    def evaluate(self, x, normalize=True):  # onbasis.jl:201-207
        val = evaluate(self.family, self.maxorder, float(x))
        if normalize:
            val = val / self.norms
        return val
