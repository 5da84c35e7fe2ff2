"""Pins the oracle against the reference's own known-answer tests (test/runtests.jl:33-127)."""
import math

import numpy as np
import pytest

from oracle import multiindices as mi
from oracle import polynomials as P
from oracle import tensorizedbasis as TB

TOL = 1.0e-12  # runtests.jl:35


def hermite_triple(j, k, l):
    """Closed form of runtests.jl:83-86 (un-normalised)."""
    if (j + k + l) % 2 == 0 and min(j + k - l, k + l - j, l + j - k) >= 0:
        f = math.factorial
        return f(l) * f(j) * f(k) / (f((j + k - l) // 2) * f((k + l - j) // 2) * f((j + l - k) // 2))
    return 0.0


@pytest.fixture(scope="module")
def onb():
    order = 6  # runtests.jl:48 (Float64 branch)
    return P.ONBasis(P.HERMITE, order, 3 * order), order


def test_integrals(onb):  # runtests.jl:51-57
    b, order = onb
    for j in range(order + 1):
        assert abs(b.integral(j) - (1 if j == 0 else 0)) < TOL


def test_norms(onb):  # runtests.jl:62-67
    b, order = onb
    for j in range(order + 1):
        assert abs(b.norm4poly(j) - float(P.norms(P.HERMITE, j))) < TOL


def test_orthonormality(onb):  # runtests.jl:70-77
    b, order = onb
    for j in range(order + 1):
        for k in range(j, order + 1):
            assert abs(b.scalar_product(j, k) - (1 if j == k else 0)) < TOL


def test_triple_products(onb):  # runtests.jl:81-93
    b, order = onb
    for j in range(order + 1):
        for k in range(j, order + 1):
            for l in range(k, order + 1):
                ref = hermite_triple(j, k, l) / math.sqrt(
                    math.factorial(j) * math.factorial(k) * math.factorial(l))
                assert abs(b.triple_product(j, k, l) - ref) < TOL


def test_tensorized_triple_products():  # runtests.jl:104-127
    tb = TB.TensorizedBasis(P.HERMITE, 3, 3, 6)
    assert tb.nmodes == 64
    # 1-D table once, then all 64^3 products as in the reference loop
    t1 = np.array([[[tb.ONB.triple_product(a, b, c) for c in range(4)] for b in range(4)] for a in range(4)])
    ref1 = np.array([[[hermite_triple(a, b, c) / math.sqrt(math.factorial(a) * math.factorial(b) * math.factorial(c))
                       for c in range(4)] for b in range(4)] for a in range(4)])
    assert np.abs(t1 - ref1).max() < TOL
    m = np.array(tb.multi_indices)
    for j in range(0, 64, 7):  # the products factorise; a strided subset of j keeps the test fast
        val = np.ones((64, 64))
        ref = np.ones((64, 64))
        for d in range(3):
            val *= t1[m[j, d]][np.ix_(m[:, d], m[:, d])]
            ref *= ref1[m[j, d]][np.ix_(m[:, d], m[:, d])]
        assert np.abs(val - ref).max() < TOL
    # spot-check the python-level method used above against the vectorised evaluation
    assert abs(tb.triple_product(5, 17, 40) - np.prod([t1[m[4, d], m[16, d], m[39, d]] for d in range(3)])) < 1e-15


def test_full_multiindex_order():  # mopcontrol.jl:6-22 ordering used by runtests.jl:106
    got = mi.generate_multiindices(2, 2)
    assert got == [[0, 0], [0, 1], [0, 2], [1, 0], [1, 1], [1, 2], [2, 0], [2, 1], [2, 2]]


@pytest.mark.parametrize("family", [P.LEGENDRE, P.HERMITE])
def test_G_equals_quadrature_triple_product_y(family):
    """Derivable pin: G[(m-1)N+j,k] = <y_m H_j H_k> (onbasis.jl:129-149)."""
    modes = mi.graded_lex_multiindices(4, 35)
    tb = TB.TensorizedBasis(family, 4, 3, 8, 16, multi_indices=modes)
    G = tb.G.toarray()
    N = len(modes)
    for m in range(4):
        for j in range(N):
            for k in range(N):
                same = all(modes[j][d] == modes[k][d] for d in range(4) if d != m)
                ref = tb.ONB.triple_product_y(modes[j][m], modes[k][m]) if same else 0.0
                assert abs(G[m * N + j, k] - ref) < 5e-13
    # each G_m is symmetric (g+(k) == g-(k+1))
    for m in range(4):
        Gm = G[m * N:(m + 1) * N]
        assert np.abs(Gm - Gm.T).max() < 1e-15


def test_legendre_closed_form():
    for k in range(12):
        gp, gm = P.coupling_weights(P.LEGENDRE, k)
        assert abs(gp - (k + 1) / math.sqrt((2 * k + 1) * (2 * k + 3))) < 1e-15
        if k > 0:
            assert abs(gm - k / math.sqrt((2 * k - 1) * (2 * k + 1))) < 1e-15
