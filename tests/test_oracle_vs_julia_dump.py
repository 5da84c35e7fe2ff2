"""Pins the CPU oracle against a dump of the reference's own run (julia/dump_fixtures.jl).
Julia is not installed in the build image, so no dump is committed: the test SKIPS unless ASGFEM_JULIA_DUMP (or
tests/golden/julia_dump/) holds one.  A maintainer with the reference checkout runs

    julia --project=<ExtendableASGFEM.jl> julia/dump_fixtures.jl /tmp/asgfem_dump
    ASGFEM_JULIA_DUMP=/tmp/asgfem_dump python -m pytest tests/test_oracle_vs_julia_dump.py

and every 'parity unpinned' row of DESIGN.md section 2 (assembly, operator, Krylov solution, estimator, extended
multi-index set) becomes a comparison with the reference itself."""
import json
import os

import numpy as np
import pytest
import scipy.sparse as sp

from oracle import estimate as oest
from oracle import problem as oproblem
from oracle import solver as osolver

DUMP = os.environ.get("ASGFEM_JULIA_DUMP", os.path.join(os.path.dirname(__file__), "golden", "julia_dump"))
pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(DUMP, "manifest.json")),
                                reason="no dump of the Julia reference (julia/dump_fixtures.jl) available")


def load(name):
    man = json.load(open(os.path.join(DUMP, "manifest.json")))
    e = man[name]
    if e.get("csc"):
        cp, rv, nz = load(name + "_colptr"), load(name + "_rowval"), load(name + "_nzval")
        return sp.csc_matrix((nz, rv - 1, cp - 1), shape=(e["m"], e["n"]))
    a = np.fromfile(os.path.join(DUMP, name + ".bin"), dtype="<" + e["dtype"])
    return a.reshape(e["shape"], order="F")


@pytest.fixture(scope="module")
def P():
    return oproblem.poisson_simple()  # the oracle's config 1


def test_mesh_space_and_index_structures(P):
    assert np.allclose(load("coords").T, P.mesh.coords, atol=1e-15)
    assert np.array_equal(load("cellnodes").T - 1, P.mesh.cellnodes)
    assert np.array_equal(load("celldofs").T - 1, P.space.celldofs)
    assert np.array_equal(load("multi_indices").T, np.array(P.multi_indices))
    G = load("G")
    Go = sp.csc_matrix(P.G)
    Go.sort_indices()
    G.sort_indices()
    assert np.array_equal(G.indptr, Go.indptr) and np.array_equal(G.indices, Go.indices)
    assert np.allclose(G.data, Go.data, rtol=1e-14, atol=0)
    assert np.array_equal(np.sort(load("bdofs").ravel() - 1), np.sort(P.bdofs))


def test_assembly(P):
    for m in range(P.M + 1):
        ref = load(f"A{m}").toarray()
        got = (P.A0 if m == 0 else P.Am[m - 1]).toarray()
        assert np.abs(got - ref).max() <= 1e-13 * np.abs(ref).max(), m
    assert np.allclose(load("b0").ravel(), P.b0, rtol=1e-13, atol=1e-16)


def test_operator(P):
    x = load("apply_x").ravel()
    ref = load("apply_Ax").ravel()
    got = osolver.SystemPrimal(P.A0, P.Am, P.G, P.bdofs, P.N).mul(x)
    assert np.linalg.norm(got - ref) <= 1e-12 * np.linalg.norm(ref)


def test_solution_and_estimator(P):
    sol = np.zeros(P.n * P.N)
    osolver.solve_primal(sol, P.A0, P.Am, P.b0, P.G, P.N, P.bdofs)
    ref = load("solution").ravel()
    assert np.linalg.norm(sol - ref) <= 1e-10 * np.linalg.norm(ref)
    em, ec, ext = oest.estimate_poisson_primal(P.space, ref, P.multi_indices, P.family, P.coeff, bonus_quadorder=1)
    assert np.array_equal(load("multi_indices_extended").T, np.array(ext))
    assert np.allclose(em, load("eta4modes").ravel(), rtol=1e-10, atol=1e-300)
    assert np.allclose(ec, load("eta4cell"), rtol=1e-10, atol=1e-14 * np.abs(ec).max())
