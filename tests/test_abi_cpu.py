"""CPU-side checks of the boundary: the shared library loads, exports every symbol include/asgfem.h declares,
and the context-free integer/index entry points are bit-exact against the oracle.  No compute calls."""
import os
import re

import numpy as np
import pytest

import asgfem_b200 as A
from asgfem_b200 import _lib
from oracle import multiindices as omi
from oracle import polynomials as opoly

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "asgfem.h")).read()
    declared = set(re.findall(r"\b(asgfem_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_lib.PROTOTYPES), declared ^ set(_lib.PROTOTYPES)
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name)
    assert b"sm_100a" in lib.asgfem_version()


def test_create_without_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.AsgfemError) as e:
        A.Context()
    assert "no CPU fallback" in str(e.value)


@pytest.mark.parametrize("family", [A.LEGENDRE, A.HERMITE])
def test_coupling_weights_bit_exact(family):
    gp, gm = A.coupling_weights(family, 12)
    for k in range(13):
        ogp, ogm = opoly.coupling_weights(family, k)
        assert gp[k] == ogp and gm[k] == ogm, (k, gp[k], ogp, gm[k], ogm)


def _random_sets(seed):
    rng = np.random.default_rng(seed)
    sets = [[[0, 0, 0], [1, 0, 0], [0, 1, 0], [2, 0, 0], [0, 0, 1]], [[0]], [[0, 0], [0, 1]],
            omi.graded_lex_multiindices(4, 35), omi.generate_multiindices(3, 2)]
    for _ in range(6):  # downward-closed-ish random sets grown like the adaptive loop would
        M = int(rng.integers(1, 6))
        s = [[0] * M]
        for _ in range(int(rng.integers(1, 25))):
            base = list(s[int(rng.integers(len(s)))])
            base[int(rng.integers(M))] += 1
            if base not in s:
                s.append(base)
        sets.append(s)
    return sets


@pytest.mark.parametrize("seed", [0, 1])
def test_add_boundary_modes_and_classify_bit_exact(seed):
    for s in _random_sets(seed):
        for tail in [(10, 2), (3, 1), (1, 1)]:
            ref = omi.add_boundary_modes([list(m) for m in s], tail_extension=tail)
            got = A.add_boundary_modes(s, tail_extension=tail)
            assert got == ref
            ref_cls = omi.classify_modes(ref, ref[:len(s)])
            got_cls = A.classify_modes(got, len(s))
            assert tuple(list(c) for c in got_cls) == tuple(list(c) for c in ref_cls)


def test_host_mirrors_match_oracle():
    from oracle import coefficient as ocoef
    C = A.StochasticCoefficientCosinus(tau=0.9, decay=2, mean=1, maxm=40)
    O = ocoef.StochasticCoefficientCosinus(tau=0.9, decay=2, mean=1, maxm=40)
    assert np.array_equal(C.decay_factors, O.decay_factors) and np.array_equal(C.b1, O.b1) and np.array_equal(C.b2, O.b2)
    assert A.generate_multiindices(3, 3) == omi.generate_multiindices(3, 3)
    assert A.graded_lex_multiindices(20, 2000) == omi.graded_lex_multiindices(20, 2000)
    from oracle import mesh as omesh, fem as ofem
    g, m = A.uniform_refine(A.grid_unitsquare(), 2), omesh.uniform_refine(omesh.grid_unitsquare(), 2)
    assert np.array_equal(g.cellnodes, m.cellnodes) and np.array_equal(g.cellfaces, m.cellfaces)
    assert np.array_equal(g.coords, m.coords) and np.array_equal(g.bfacefaces, m.bfacefaces)
    for order in (1, 2):
        fs, os_ = A.FESpace(g, order), ofem.FESpace(m, order)
        assert np.array_equal(fs.celldofs, os_.celldofs) and np.array_equal(fs.bdofs, os_.bdofs)
        assert np.allclose(fs.rhs(), ofem.assemble_rhs(os_), rtol=0, atol=1e-16)
    for order in (1, 2, 4):
        assert np.allclose(A.quadrature_rule(order)[0], ofem.quadrature_rule(order)[0])
        assert np.allclose(A.quadrature_rule(order)[1], ofem.quadrature_rule(order)[1])


def test_cxx_exceptions_do_not_cross_the_boundary():
    """include/asgfem.h: "nothing throws or aborts across the boundary".  An absurd tail extension makes the index code ask
    for vectors of 2^58 and 2^62 entries: std::bad_alloc must come back as ASGFEM_ENOMEM (-4) and std::length_error as
    ASGFEM_EINTERNAL (-6) instead of terminating the host process."""
    import ctypes as C
    lib = _lib.load()
    mi = np.zeros(1, dtype=np.int64)
    n_ext, m_ext = C.c_int64(), C.c_int64()
    out = np.zeros(4, dtype=np.int64)
    codes = [lib.asgfem_add_boundary_modes(1, 1, mi.ctypes.data_as(C.c_void_p), 1, tail, 2, C.byref(n_ext), C.byref(m_ext),
                                           out.ctypes.data_as(C.c_void_p), 4) for tail in (2 ** 58, 2 ** 62)]
    assert codes == [-4, -6]
