"""In-repo identities for the unpinned oracle rows (SURVEY.md §4 build-side plan (ii))."""
import numpy as np
import pytest

from oracle import coefficient, estimate, mesh, multiindices as mi, problem, solver


@pytest.fixture(scope="module")
def c1():
    return problem.poisson_simple()


def test_config1_sizes(c1):
    # C1 = 256 cells, 145 nodes, 400 edges, P2 -> n = 545, N = 5, M = 3, nnz(G) = 8
    assert (c1.mesh.ncells, c1.mesh.nnodes, c1.mesh.nfaces) == (256, 145, 400)
    assert (c1.n, c1.N, c1.M, c1.G.nnz) == (545, 5, 3, 8)
    assert c1.G.shape == (15, 5)


def test_coefficient_tables():
    C = coefficient.StochasticCoefficientCosinus(tau=0.9, decay=2, mean=1, maxm=9)
    assert list(zip(C.b1, C.b2)) == [(0, 1), (1, 0), (0, 2), (1, 1), (2, 0), (0, 3), (1, 2), (2, 1), (3, 0)]
    assert abs(C.decay_factors[0] - 0.34887287) < 1e-8
    C150 = coefficient.StochasticCoefficientCosinus(tau=0.9, decay=2, mean=1, maxm=150)
    assert abs(C150.decay_factors.sum() - 0.8908) < 1e-4


def test_mul_equals_assembled(c1):
    S = solver.SystemPrimal(c1.A0, c1.Am, c1.G, c1.bdofs, c1.N)
    x = np.random.default_rng(1).standard_normal(c1.n * c1.N)
    y = S.mul(x)
    assert np.abs(y - S.assembled() @ x).max() <= 1e-13 * np.abs(y).max()


def test_three_solvers_agree(c1):
    sols = []
    for method in ("gmres", "pcg"):
        s = np.zeros(c1.n * c1.N)
        st = solver.solve_primal(s, c1.A0, c1.Am, c1.b0, c1.G, c1.N, c1.bdofs, method=method)
        assert st["solved"] and st["residual"] < 1e-12
        sols.append(s)
    sols.append(solver.solve_full_primal(c1.A0, c1.Am, c1.b0, c1.G, c1.N, c1.bdofs))
    ref = np.linalg.norm(sols[0])
    assert np.linalg.norm(sols[0] - sols[1]) < 1e-12 * ref
    assert np.linalg.norm(sols[0] - sols[2]) < 1e-12 * ref


def test_mesh_counts():
    m = mesh.uniform_refine(mesh.grid_unitsquare(), 4)
    assert (m.ncells, m.nnodes) == (4 ** 5, 545)
    assert abs(m.cellvolumes.sum() - 1.0) < 1e-14
    l = mesh.uniform_refine(mesh.grid_lshape(), 2)
    assert abs(l.cellvolumes.sum() - 3.0) < 1e-13
    s = mesh.structured_unitsquare(9)
    assert (s.nnodes, s.ncells, s.nfaces) == (81, 128, 81 + 128 - 1)
    assert len(s.bfacefaces) == 32 and np.all(s.facecells[s.bfacefaces, 1] == -1)


def test_add_boundary_modes_and_classify():
    ext = mi.add_boundary_modes([[0, 0, 0], [1, 0, 0], [0, 1, 0], [2, 0, 0], [0, 0, 1]])
    assert len(ext[0]) == 13 and ext[:5] == [[0] * 13, [1] + [0] * 12, [0, 1] + [0] * 11, [2] + [0] * 12,
                                            [0, 0, 1] + [0] * 10]
    assert ext[5] == [0, 0, 0, 1] + [0] * 9  # first appended tail mode
    ie, ib, ib2, ab, ai = mi.classify_modes(ext, ext[:5])
    assert ai == [1] and ab == [2, 3, 4, 5] and ib == list(range(6, len(ext) + 1)) and ie == [] and ib2 == []
    PL, MN = mi.get_neighbours(ext)
    assert PL[0, 0] == 2 and MN[0, 1] == 1 and PL[0, 1] == 4 and MN[0, 0] == 0


def test_estimator_runs_and_is_consistent(c1):
    s = np.zeros(c1.n * c1.N)
    solver.solve_primal(s, c1.A0, c1.Am, c1.b0, c1.G, c1.N, c1.bdofs)
    em, ec, ext = estimate.estimate_poisson_primal(c1.space, s, c1.multi_indices, c1.family, c1.coeff,
                                                   bonus_quadorder=2)
    assert ec.shape == (c1.mesh.ncells, len(ext)) and np.all(ec >= 0)
    # eta4modes^2 = volume part + every interior face once; eta4cell counts interior faces twice
    assert np.all(em ** 2 <= ec.sum(axis=0) * (1 + 1e-12))
    assert np.all(ec.sum(axis=0) <= 2 * em ** 2 * (1 + 1e-12))


def test_logprimal_oracle_definitions_agree():
    """Row f1 (solvers_logpoisson_primal.jl): mul! == assembled block matrix, GMRES == direct solve of the full system."""
    P = problem.logprimal_like(nrefs=2, order=1)
    S = solver.SystemLogPrimal(P.A, P.N0, P.Nm, P.G, P.bdofs, P.N)
    x = np.random.default_rng(2).standard_normal(P.n * P.N)
    y = S.mul(x)
    assert np.abs(y - S.assembled() @ x).max() <= 1e-13 * np.abs(y).max()
    sol = np.zeros(P.n * P.N)
    st = solver.solve_logpoisson_primal(sol, P.A, P.N0, P.Nm, P.b0m, P.G, P.N, P.bdofs)
    assert st["solved"]
    ref = solver.solve_logpoisson_primal_full(P.A, P.N0, P.Nm, P.b0m, P.G, P.N, P.bdofs)
    assert np.linalg.norm(sol - ref) <= 1e-10 * np.linalg.norm(ref)
    assert np.abs((S.N0 - S.N0.T)).max() > 0  # the test system really is nonsymmetric
