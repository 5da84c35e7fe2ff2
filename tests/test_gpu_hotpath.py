"""GPU parity tests of the hot path, through the C ABI, against the oracle on the same seeded inputs.
Tolerances are BASELINE.json's: index structures bit-exact; 1e-12 relative for one operator application;
1e-10 relative for the converged solution."""
import numpy as np
import pytest
import scipy.sparse as sp

import asgfem_b200 as A
from oracle import fem as ofem
from oracle import mesh as omesh
from oracle import multiindices as omi
from oracle import polynomials as opoly
from oracle import problem as oproblem
from oracle import solver as osolver
from oracle import tensorizedbasis as otb
from oracle import coefficient as ocoef

pytestmark = pytest.mark.gpu

TOL_APPLY = 1.0e-12
TOL_SOLVE = 1.0e-10


def relerr(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def make_ctx(P):
    """Loads an oracle-built problem into a fresh context exactly as the Julia shim would (1-based CSC)."""
    ctx = A.Context()
    ctx.set_multiindices(P.family, np.array(P.multi_indices, dtype=np.int64))
    A0 = sp.csc_matrix(P.A0)
    A0.sort_indices()
    ctx.set_pattern_csc(P.n, A0.indptr.astype(np.int64) + 1, A0.indices.astype(np.int64) + 1)
    ctx.set_num_stiffness(P.M)
    ctx.set_stiffness(0, A0.data)
    for m, Am in enumerate(P.Am, start=1):
        Am = sp.csc_matrix(Am)
        Am.sort_indices()
        ctx.set_stiffness(m, Am.data)
    ctx.set_bdofs(P.bdofs + 1)
    ctx.vec_alloc(3)
    return ctx


@pytest.fixture(scope="module")
def c1():
    return oproblem.poisson_simple()  # config 1: n=545 (P2), N=5, M=3


@pytest.fixture(scope="module")
def mid():
    """P1 on a 40x40 structured mesh with 120 graded-lex modes in 6 dimensions: exercises the tiled kernel."""
    m = omesh.structured_unitsquare(40)
    modes = omi.graded_lex_multiindices(6, 120)
    C = ocoef.StochasticCoefficientCosinus(tau=0.9, decay=2.0, mean=1.0, maxm=6)
    return oproblem.build(m, 1, modes, opoly.LEGENDRE, C)


def test_index_structures_bit_exact(c1, mid):
    for P in (c1, mid):
        ctx = make_ctx(P)
        colptr, rowval, nzval = ctx.coupling_csc()
        G = P.G
        assert np.array_equal(colptr - 1, G.indptr) and np.array_equal(rowval - 1, G.indices)
        assert np.array_equal(nzval, G.data)  # bit-exact values as well
        plus, minus = ctx.neighbours()
        oplus, ominus = omi.get_neighbours(P.multi_indices)
        assert np.array_equal(plus, oplus) and np.array_equal(minus, ominus)
        cp, rv = ctx.pattern_csc()
        A0 = sp.csc_matrix(P.A0)
        A0.sort_indices()
        assert np.array_equal(cp - 1, A0.indptr) and np.array_equal(rv - 1, A0.indices)
        ctx.close()


def test_vector_layout_roundtrip_and_random_fill(c1):
    ctx = make_ctx(c1)
    x = np.random.default_rng(3).standard_normal(c1.n * c1.N)
    ctx.vec_upload(0, x)
    assert np.array_equal(ctx.vec_download(0), x)
    ctx.vec_fill_random(1, 20240)
    ref = oproblem.splitmix64_uniform(np.arange(c1.n * c1.N), 20240)
    assert np.array_equal(ctx.vec_download(1), ref)
    ctx.vec_upload(2, x)
    assert abs(ctx.vec_dot(0, 2) - float(x @ x)) <= 1e-13 * float(x @ x)
    ctx.close()


@pytest.mark.parametrize("variant", [1, 7, 9])
def test_apply_matches_oracle(c1, mid, variant):
    for P in (c1, mid):
        ctx = make_ctx(P)
        ctx.set_apply_variant(variant)
        S = osolver.SystemPrimal(P.A0, P.Am, P.G, P.bdofs, P.N)
        x = np.random.default_rng(7).standard_normal(P.n * P.N)
        ref = S.mul(x)
        ctx.vec_upload(0, x)
        ctx.apply(0, 1)
        got = ctx.vec_download(1)
        assert relerr(got, ref) < TOL_APPLY
        assert np.all(got.reshape(P.N, P.n)[:, P.bdofs] == 0)  # Dirichlet rows exactly zero
        # host seam (mul!)
        assert relerr(ctx.apply_host(x), ref) < TOL_APPLY
        ctx.close()


@pytest.mark.parametrize("nx,order,N,M", [(33, 1, 2000, 20), (17, 1, 2048, 40), (9, 2, 300, 10), (13, 2, 2000, 20)])
def test_apply_matches_oracle_on_bench_sets(nx, order, N, M):
    """The benchmark's own multi-index set (2000 graded-lex Legendre modes, M = 20) and the shapes next to it (2048 modes /
    M = 40; P2 rows with up to 19 entries) against the ORACLE (SystemPrimal.mul = mul!, solvers_poisson_primal.jl:86-124),
    not against another CUDA kernel: every kernel that accepts the shape must agree to 1e-12."""
    P = oproblem.synthetic(nx, order, N, M)
    S = osolver.SystemPrimal(P.A0, P.Am, P.G, P.bdofs, P.N)
    x = np.random.default_rng(11).standard_normal(P.n * P.N)
    ref = S.mul(x)
    ctx = make_ctx(P)
    ran = []
    from asgfem_b200 import _lib
    for variant in (0, 1, 7, 9):
        ctx.set_apply_variant(variant)
        ctx.vec_upload(0, x)
        try:
            ctx.apply(0, 1)
        except _lib.AsgfemError as e:  # a kernel may decline a shape; it must say so
            assert variant in (7, 9) and "not available" in str(e)
            continue
        ran.append(variant)
        got = ctx.vec_download(1)
        assert relerr(got, ref) < TOL_APPLY
        assert np.max(np.abs(got - ref)) <= 1e-12 * np.max(np.abs(ref))
    assert 0 in ran and 1 in ran
    ctx.close()


@pytest.mark.parametrize("variant", [0, 1, 7, 9])
def test_apply_host_pipelined_row_blocks(c1, mid, variant, monkeypatch):
    """The mul! seam overlaps upload / operator / download block by block; small blocks force the multi-block
    schedule, including rows whose columns live in much later blocks (P2 edge dofs of config 1)."""
    for P, rb in ((c1, 64), (mid, 256), (mid, 96)):
        monkeypatch.setenv("ASGFEM_HOST_BLOCK_ROWS", str(rb))
        ctx = make_ctx(P)
        ctx.set_apply_variant(variant)
        S = osolver.SystemPrimal(P.A0, P.Am, P.G, P.bdofs, P.N)
        x = np.random.default_rng(rb).standard_normal(P.n * P.N)
        got = ctx.apply_host(x)
        assert relerr(got, S.mul(x)) < TOL_APPLY
        ctx.vec_upload(0, x)
        ctx.apply(0, 1)
        assert np.array_equal(got, ctx.vec_download(1))
        ctx.close()


@pytest.mark.parametrize("family", [opoly.LEGENDRE, opoly.HERMITE])
def test_apply_random_sets_lshape(family):
    m = omesh.uniform_refine(omesh.grid_lshape(), 3)
    rng = np.random.default_rng(11)
    modes = [[0, 0, 0, 0]]
    while len(modes) < 40:
        b = list(modes[int(rng.integers(len(modes)))])
        b[int(rng.integers(4))] += 1
        if b not in modes:
            modes.append(b)
    C = ocoef.StochasticCoefficientCosinus(tau=0.9, decay=2.0, mean=1.0, maxm=4)
    for order in (1, 2):
        P = oproblem.build(m, order, modes, family, C)
        S = osolver.SystemPrimal(P.A0, P.Am, P.G, P.bdofs, P.N)
        x = rng.standard_normal(P.n * P.N)
        ref = S.mul(x)
        for variant in (1, 7, 9):
            ctx = make_ctx(P)
            ctx.set_apply_variant(variant)
            assert relerr(ctx.apply_host(x), ref) < TOL_APPLY
            ctx.close()


def test_apply_linearity_at_scale():
    """Size-independent property on a problem too large for the oracle's Python loops: A(ax+by) = aAx + bAy and
    the two kernels agree."""
    g = A.structured_unitsquare(257)
    fes = A.FESpace(g, 1)
    modes = A.graded_lex_multiindices(10, 300)
    TB = A.TensorizedBasis(A.LegendrePolynomials, modes)
    sol = A.SGFEVector(fes, TB)
    A.setup_device_problem(sol, A.StochasticCoefficientCosinus(tau=0.9, decay=2, mean=1, maxm=10))
    ctx = TB.ctx
    ctx.vec_alloc(5)
    ctx.vec_fill_random(0, 1)
    ctx.vec_fill_random(1, 2)
    ctx.set_apply_variant(9)
    ctx.apply(0, 2)
    ctx.apply(1, 3)
    ctx.vec_axpy(0.5, 1, 0)      # x0 += 0.5 x1
    ctx.apply(0, 4)              # A(x0 + 0.5 x1)
    ctx.vec_axpy(0.5, 3, 2)      # A x0 + 0.5 A x1
    ctx.vec_axpy(-1.0, 4, 2)
    nrm = np.sqrt(ctx.vec_dot(4, 4))
    assert np.sqrt(ctx.vec_dot(2, 2)) < 1e-13 * nrm
    ctx.set_apply_variant(1)
    ctx.apply(0, 3)
    ctx.vec_axpy(-1.0, 4, 3)
    assert np.sqrt(ctx.vec_dot(3, 3)) < 1e-13 * nrm
    ctx.close()


@pytest.mark.parametrize("ts_variant", [7, 9])
@pytest.mark.parametrize("nx,M,N", [(129, 20, 2000), (65, 12, 700), (97, 6, 100), (33, 20, 2048), (33, 40, 1500), (33, 3, 35)])
def test_apply_mode_stationary_matches_gather_at_bench_shape(nx, M, N, ts_variant):
    """The benchmark's mode set (2000 graded-lex modes in 20 dimensions) and other shapes (wide: M > 31, tiny, one and two
    passes) on meshes too large for the oracle: the MMA kernel must agree with the reference-order gather kernel
    (variant 1) to rounding level, and be bit-reproducible."""
    g = A.structured_unitsquare(nx)
    fes = A.FESpace(g, 1)
    modes = A.graded_lex_multiindices(M, N)
    TB = A.TensorizedBasis(A.LegendrePolynomials, modes)
    sol = A.SGFEVector(fes, TB)
    A.setup_device_problem(sol, A.StochasticCoefficientCosinus(tau=0.9, decay=2, mean=1, maxm=M))
    ctx = TB.ctx
    ctx.vec_alloc(4)
    ctx.vec_fill_random(0, 3)
    ctx.set_apply_variant(1)
    ctx.apply(0, 1)
    ctx.set_apply_variant(ts_variant)
    ctx.apply(0, 2)
    ctx.apply(0, 3)
    a, b, c = ctx.vec_download(1), ctx.vec_download(2), ctx.vec_download(3)
    assert np.array_equal(b, c)
    assert relerr(b, a) < TOL_APPLY
    assert np.max(np.abs(b - a)) <= 1e-12 * np.max(np.abs(a))
    ctx.close()


@pytest.mark.parametrize("order,nx,M,N", [(2, 17, 20, 5000), (2, 33, 8, 300), (1, 17, 70, 600)])
def test_apply_block_kernel_long_rows_and_many_modes(order, nx, M, N):
    """Shapes only the block kernel (variant 9) covers: P2 rows (up to 24 entries -> 6 k-steps), the mode count of config 5
    (N = 5000 -> several passes per dof row, lists in global memory), more than 63 directions.  Reference: the gather
    kernel (variant 1)."""
    g = A.structured_unitsquare(nx)
    fes = A.FESpace(g, order)
    modes = A.graded_lex_multiindices(M, N)
    TB = A.TensorizedBasis(A.LegendrePolynomials, modes)
    sol = A.SGFEVector(fes, TB)
    A.setup_device_problem(sol, A.StochasticCoefficientCosinus(tau=0.9, decay=2, mean=1, maxm=M))
    ctx = TB.ctx
    ctx.vec_alloc(4)
    ctx.vec_fill_random(0, 11)
    ctx.set_apply_variant(1)
    ctx.apply(0, 1)
    ctx.set_apply_variant(9)
    ctx.apply(0, 2)
    ctx.apply(0, 3)
    a, b, c = ctx.vec_download(1), ctx.vec_download(2), ctx.vec_download(3)
    assert np.array_equal(b, c)
    assert relerr(b, a) < TOL_APPLY
    ctx.close()


@pytest.mark.parametrize("order", [1, 2])
def test_device_assembly_matches_oracle(order):
    m = omesh.uniform_refine(omesh.grid_unitsquare(), 3)
    C = ocoef.StochasticCoefficientCosinus(tau=0.9, decay=2.0, mean=1.0, maxm=7)
    space = ofem.FESpace(m, order)
    indptr, indices, vals = ofem.assemble_stiffness(space, C, 7, 2)
    ctx = A.Context()
    ctx.set_mesh(m.coords, m.cellnodes + 1)
    ctx.set_space(order, space.ndofs, space.celldofs + 1)
    ctx.set_coefficient_cosinus(C.mean_value, C.decay_factors, C.b1, C.b2)
    xref, w = ofem.quadrature_rule(2 * (order - 1) + 2)
    ctx.assemble_stiffness(7, xref, w)
    cp, rv = ctx.pattern_csc()
    assert np.array_equal(cp - 1, indptr) and np.array_equal(rv - 1, indices)  # symmetric pattern: CSC == CSR
    for mm in range(8):
        K = sp.csr_matrix((vals[mm], indices, indptr), shape=(space.ndofs,) * 2)
        Kt = sp.csc_matrix(K)
        Kt.sort_indices()
        got = ctx.get_stiffness(mm)
        assert np.abs(got - Kt.data).max() <= 1e-13 * np.abs(Kt.data).max()
    ctx.close()


def test_preconditioner_matches_oracle(c1, mid):
    for P in (c1, mid):
        ctx = make_ctx(P)
        Pre = osolver.PreconditionerPrimal(P.A0, P.bdofs, P.N)
        b = np.random.default_rng(5).standard_normal(P.n * P.N)
        b.reshape(P.N, P.n)[:, P.bdofs] = 0
        ref = Pre.ldiv(b)
        ref.reshape(P.N, P.n)[:, P.bdofs] = 0  # the reference leaves O(1e-60) there
        got = ctx.precond_apply_host(b)
        assert relerr(got, ref) < 1e-11
        ctx.close()


@pytest.mark.parametrize("nx,order,N", [(257, 1, 40), (129, 2, 24), (513, 1, 17)])
def test_preconditioner_direct_solve_at_scale(nx, order, N):
    """Meshes whose dissection tree has long separators (sub-blocks split into solve-only and push-only tasks, chunked
    separators at 513^2): every mode block of ldiv! must equal the sparse direct solve with K_0 restricted to the
    interior dofs (the oracle's LU of the penalised matrix is the same thing up to 1e-60)."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    g = A.structured_unitsquare(nx)
    fes = A.FESpace(g, order)
    modes = A.graded_lex_multiindices(3, N)
    TB = A.TensorizedBasis(A.LegendrePolynomials, modes)
    sol = A.SGFEVector(fes, TB)
    A.setup_device_problem(sol, A.StochasticCoefficientCosinus(tau=0.9, decay=2, mean=1, maxm=3))
    ctx = TB.ctx
    n = fes.ndofs
    colptr, rowval = ctx.pattern_csc()
    K0 = sp.csc_matrix((ctx.get_stiffness(0), rowval - 1, colptr - 1), shape=(n, n))
    interior = np.setdiff1d(np.arange(n), fes.bdofs)
    lu = spla.splu(sp.csc_matrix(K0[interior][:, interior]))
    b = np.random.default_rng(nx).standard_normal(n * N)
    b.reshape(N, n)[:, fes.bdofs] = 0
    ref = np.zeros((N, n))
    ref[:, interior] = lu.solve(b.reshape(N, n)[:, interior].T).T
    got = ctx.precond_apply_host(b).reshape(N, n)
    assert np.all(got[:, fes.bdofs] == 0)
    assert relerr(got.ravel(), ref.ravel()) < 1e-10
    ctx.close()


def test_pcg_matches_reference_gmres_solution(c1, mid):
    for P in (c1, mid):
        ref = np.zeros(P.n * P.N)
        st = osolver.solve_primal(ref, P.A0, P.Am, P.b0, P.G, P.N, P.bdofs, method="gmres")
        assert st["solved"]
        ctx = make_ctx(P)
        sol = np.zeros(P.n * P.N)
        stats = ctx.solve_primal_host(sol, P.b0)
        assert stats["solved"] == 1 and stats["niter"] < 200
        assert relerr(sol, ref) < TOL_SOLVE
        ctx.close()


def test_reference_interface_solve_primal(c1):
    """Same call shape as solve_primal!(sol, A0, Am, b0, G, nmodes, bfac) with scipy CSC matrices."""
    g = A.uniform_refine(A.grid_unitsquare(), 3)
    fes = A.FESpace(g, 2)
    TB = A.TensorizedBasis(A.LegendrePolynomials, [[0], [1, 0], [0, 1], [2, 0], [0, 0, 1]])
    assert (TB.G != c1.G).nnz == 0
    sol = A.SGFEVector(fes, TB)
    bdofs = A.solve_primal(sol, c1.A0, c1.Am, c1.b0, TB.G, TB.nmodes, 1)
    assert np.array_equal(bdofs, c1.bdofs + 1)
    ref = np.zeros(c1.n * c1.N)
    osolver.solve_primal(ref, c1.A0, c1.Am, c1.b0, c1.G, c1.N, c1.bdofs)
    assert relerr(sol.entries, ref) < TOL_SOLVE
    # and the all-device route (assembly on the GPU) gives the same solution
    sol2 = A.SGFEVector(fes, A.TensorizedBasis(A.LegendrePolynomials, TB.multi_indices))
    A.solve(sol2, A.StochasticCoefficientCosinus(tau=0.9, decay=2.0, mean=1.0))
    assert relerr(sol2.entries, ref) < TOL_SOLVE
    TB.ctx.close()
    sol2.TB.ctx.close()


@pytest.mark.parametrize("order,domain", [(1, "square"), (2, "square"), (1, "lshape")])
def test_estimator_matches_oracle(order, domain):
    from oracle import estimate as oest
    P = oproblem.poisson_simple(nrefs=3, order=order, domain=domain)
    ref_sol = np.zeros(P.n * P.N)
    osolver.solve_primal(ref_sol, P.A0, P.Am, P.b0, P.G, P.N, P.bdofs)
    f = lambda x, y: 1.0 + x * y  # noqa: E731  non-constant rhs exercises f_at_qp
    em, ec, ext = oest.estimate_poisson_primal(P.space, ref_sol, P.multi_indices, P.family, P.coeff, f=f,
                                               bonus_quadorder=2, tail_extension=(10, 2))
    g = A.uniform_refine(A.grid_unitsquare() if domain == "square" else A.grid_lshape(), 3)
    fes = A.FESpace(g, order)
    TB = A.TensorizedBasis(A.LegendrePolynomials, P.multi_indices)
    sol = A.SGFEVector(fes, TB)
    A.setup_device_problem(sol, A.StochasticCoefficientCosinus(tau=0.9, decay=2.0, mean=1.0))
    sol.entries[:] = ref_sol
    gm, gc, gext = A.estimate(sol, None, rhs=f, bonus_quadorder=2, tail_extension=(10, 2))
    assert gext == ext                                   # extended multi-index set bit-exact
    assert gc.shape == ec.shape
    assert np.abs(gm - em).max() <= TOL_SOLVE * em.max()
    assert np.abs(gc - ec).max() <= TOL_SOLVE * ec.max()
    # marked cells of the script's spatial refinement (poisson.jl:402): Doerfler marking on the active-mode sum
    act = list(range(P.N))

    def marked(ind, theta=0.5):
        order_ = np.argsort(-ind, kind="stable")
        csum = np.cumsum(ind[order_])
        return set(order_[: int(np.searchsorted(csum, theta * csum[-1]) + 1)].tolist())

    if domain == "lshape":  # non-degenerate indicators (no exact ties) -> the marked set must be identical
        assert marked(gc[:, act].sum(axis=1)) == marked(ec[:, act].sum(axis=1))
    # the marking-oriented call: eta4modes and the active-mode row sums only (no ncells x N_ext transfer)
    gm2, cs, _ = A.estimate(sol, None, rhs=f, bonus_quadorder=2, tail_extension=(10, 2), marking_columns=np.arange(1, P.N + 1))
    assert np.array_equal(gm2, gm)
    assert np.abs(cs - gc[:, act].sum(axis=1)).max() <= 1e-14 * np.abs(cs).max()
    assert np.abs(cs - ec[:, act].sum(axis=1)).max() <= TOL_SOLVE * ec.max()
    TB.ctx.close()


def test_edge_cases_and_error_codes():
    """Ragged / degenerate inputs and the error behaviour of the C ABI (negative codes + message, never a crash)."""
    from asgfem_b200 import _lib
    m = omesh.uniform_refine(omesh.grid_unitsquare(), 2)
    C = ocoef.StochasticCoefficientCosinus(tau=0.9, decay=2.0, mean=1.0, maxm=3)
    # --- a single mode (no couplings at all): the operator is K_0, the solve is one Poisson problem ---------------
    P1 = oproblem.build(m, 1, [[0]], opoly.LEGENDRE, C)
    ctx = make_ctx(P1)
    x = np.random.default_rng(0).standard_normal(P1.n)
    S = osolver.SystemPrimal(P1.A0, P1.Am, P1.G, P1.bdofs, 1)
    for variant in (1, 7, 9):
        ctx.set_apply_variant(variant)
        assert relerr(ctx.apply_host(x), S.mul(x)) < TOL_APPLY
    sol = np.zeros(P1.n)
    st = ctx.solve_primal_host(sol, P1.b0)
    assert st["solved"] == 1 and st["niter"] <= 2  # the mean preconditioner is the exact inverse here
    ref = np.zeros(P1.n)
    osolver.solve_primal(ref, P1.A0, P1.Am, P1.b0, P1.G, 1, P1.bdofs)
    assert relerr(sol, ref) < TOL_SOLVE
    # --- non-zero incoming vector: the reference adds it to the right-hand side (b = deepcopy(sol); b[1] += b0,
    #     solvers_poisson_primal.jl:149-150) and uses it as warm start; the oracle does the same ---------------------
    warm = 0.1 * np.random.default_rng(1).standard_normal(P1.n)
    warm[P1.bdofs] = 0
    sol_w, ref_w = warm.copy(), warm.copy()
    ctx.solve_primal_host(sol_w, P1.b0)
    osolver.solve_primal(ref_w, P1.A0, P1.Am, P1.b0, P1.G, 1, P1.bdofs)
    assert relerr(sol_w, ref_w) < TOL_SOLVE
    # --- call-order and argument errors --------------------------------------------------------------------------
    with pytest.raises(_lib.AsgfemError) as e:
        ctx.apply(0, 0)
    assert e.value.code == -1
    with pytest.raises(_lib.AsgfemError) as e:
        ctx.set_stiffness(9, np.zeros(3))
    assert e.value.code == -1
    with pytest.raises(_lib.AsgfemError) as e:
        ctx.set_bdofs(np.array([0]))  # 0 is not a valid 1-based dof
    assert e.value.code == -1
    ctx.close()
    fresh = A.Context()
    with pytest.raises(_lib.AsgfemError) as e:
        fresh.vec_alloc(1)
    assert e.value.code == -2
    fresh.set_multiindices(A.LEGENDRE, np.array([[0, 0], [1, 0], [0, 1]], dtype=np.int64))
    with pytest.raises(_lib.AsgfemError) as e:
        fresh.set_multiindices(A.LEGENDRE, np.array([[0, -1]], dtype=np.int64))
    assert e.value.code == -1
    fresh.close()
    # --- multi-indices that couple in a direction without a stiffness matrix (maxlength > length(Am)) -----------------
    P2 = oproblem.build(m, 1, [[0, 0], [1, 0], [0, 1]], opoly.LEGENDRE, C)
    ctx = A.Context()
    ctx.set_multiindices(P2.family, np.array(P2.multi_indices, dtype=np.int64))
    A0 = sp.csc_matrix(P2.A0)
    A0.sort_indices()
    ctx.set_pattern_csc(P2.n, A0.indptr.astype(np.int64) + 1, A0.indices.astype(np.int64) + 1)
    ctx.set_num_stiffness(1)  # only K_0, K_1 although direction 2 is coupled
    ctx.set_stiffness(0, A0.data)
    ctx.vec_alloc(2)
    with pytest.raises(_lib.AsgfemError) as e:
        ctx.apply(0, 1)
    assert e.value.code == -1 and "direction" in str(e.value)
    # --- no Dirichlet dofs: K_0 is singular -> the factorisation must report it, not crash ----------------------------
    ctx.set_num_stiffness(2)
    for mm, Am in enumerate([P2.A0] + P2.Am):
        Am = sp.csc_matrix(Am)
        Am.sort_indices()
        ctx.set_stiffness(mm, Am.data)
    ctx.set_bdofs(np.zeros(0, dtype=np.int64))
    with pytest.raises(_lib.AsgfemError) as e:
        ctx.precond_setup()
    assert e.value.code == -5
    ctx.close()


def test_operator_properties_p2_hermite_at_scale():
    """P2 space (rows up to 19+ entries), Hermite couplings, N > one shared-memory tile: symmetry <x, A y> = <A x, y>
    of the Dirichlet-reduced operator and agreement of all kernels."""
    g = A.uniform_refine(A.grid_lshape(), 5)
    fes = A.FESpace(g, 2)
    modes = A.graded_lex_multiindices(6, 400)
    TB = A.TensorizedBasis(A.HermitePolynomials, modes)
    sol = A.SGFEVector(fes, TB)
    A.setup_device_problem(sol, A.StochasticCoefficientCosinus(tau=0.3, decay=2, mean=1, maxm=6))
    ctx = TB.ctx
    ctx.vec_alloc(5)
    n, N = fes.ndofs, len(modes)
    rng = np.random.default_rng(2)
    x = rng.standard_normal(n * N)
    y = rng.standard_normal(n * N)
    x.reshape(N, n)[:, fes.bdofs] = 0
    y.reshape(N, n)[:, fes.bdofs] = 0
    ctx.vec_upload(0, x)
    ctx.vec_upload(1, y)
    ref = None
    for variant in (1, 7, 9):
        ctx.set_apply_variant(variant)
        try:
            ctx.apply(0, 2)
        except Exception as e:  # a kernel may decline a pattern it cannot hold; it must say so
            assert "plan not available" in str(e)
            continue
        ctx.apply(1, 3)
        xAy, yAx = ctx.vec_dot(0, 3), ctx.vec_dot(1, 2)
        assert abs(xAy - yAx) <= 1e-12 * abs(xAy)
        out = ctx.vec_download(2)
        if ref is None:
            ref = out
        else:
            assert relerr(out, ref) < TOL_APPLY
    assert ref is not None
    ctx.close()


def test_apply_rows_ranges_compose(c1):
    """asgfem_apply_rows on a partition of the rows reproduces asgfem_apply bit for bit and leaves the other rows alone."""
    ctx = make_ctx(c1)
    ctx.vec_alloc(3)
    ctx.vec_fill_random(0, 11)
    ctx.apply(0, 1)
    full = ctx.vec_download(1)
    ctx.vec_fill_random(2, 5)
    before = ctx.vec_download(2).copy()
    n = c1.n
    ctx.apply_rows(0, 2, n // 3, 2 * n // 3)
    part = ctx.vec_download(2).reshape(c1.N, n)  # reference layout: mode blocks of n dofs
    assert np.array_equal(part[:, n // 3:2 * n // 3], full.reshape(c1.N, n)[:, n // 3:2 * n // 3])
    assert np.array_equal(part[:, :n // 3], before.reshape(c1.N, n)[:, :n // 3])
    ctx.apply_rows(0, 2, 0, n // 3)
    ctx.apply_rows(0, 2, 2 * n // 3, n + 100)  # clipped to the owned rows
    assert np.array_equal(ctx.vec_download(2), full)
    from asgfem_b200 import _lib
    with pytest.raises(_lib.AsgfemError):
        ctx.apply_rows(0, 2, 5, 3)
    ctx.close()


@pytest.mark.parametrize("order", [1, 2])
def test_logprimal_seam_matches_oracle(order):
    """Row f1: solve_logpoisson_primal!(sol, A, N0, Nm, b0, G, nmodes, bfac) with Hermite coupling, nonsymmetric N_e and a
    load vector per mode.  Operator vs the oracle's mul! (1e-12), device BiCGStab vs the direct solve of the assembled
    block system (1e-10), preconditioner built from A alone."""
    P = oproblem.logprimal_like(nrefs=3, order=order)
    g = A.uniform_refine(A.grid_unitsquare(), 3)
    fes = A.FESpace(g, order)
    TB = A.TensorizedBasis(A.HermitePolynomials, P.multi_indices)
    assert (TB.G != P.G).nnz == 0
    sol = A.SGFEVector(fes, TB)
    bdofs, stats = A.solve_logpoisson_primal(sol, P.A, P.N0, P.Nm, P.b0m, TB.G, TB.nmodes, 1, return_stats=True)
    assert np.array_equal(bdofs, P.bdofs + 1)
    assert stats["solved"] == 1
    ref = osolver.solve_logpoisson_primal_full(P.A, P.N0, P.Nm, P.b0m, P.G, P.N, P.bdofs)
    assert relerr(sol.entries, ref) < TOL_SOLVE
    # operator parity on the installed matrices (A + N0 on the diagonal, N_e off it), every kernel variant
    S = osolver.SystemLogPrimal(P.A, P.N0, P.Nm, P.G, P.bdofs, P.N)
    x = np.random.default_rng(4).standard_normal(P.n * P.N)
    y = S.mul(x)
    ctx = TB.ctx
    for variant in (1, 7, 9):
        ctx.set_apply_variant(variant)
        assert relerr(ctx.apply_host(x), y) < TOL_APPLY
    # the preconditioner is A^-1 per mode, not (A + N0)^-1
    import scipy.sparse.linalg as spla
    keep = np.setdiff1d(np.arange(P.n), P.bdofs)
    r = np.zeros(P.n * P.N)
    r[:P.n][keep] = np.random.default_rng(5).standard_normal(len(keep))
    z = A.ldiv(ctx, r)
    zr = np.zeros(P.n)
    zr[keep] = spla.spsolve(P.A.tocsc()[keep][:, keep], r[:P.n][keep])
    assert relerr(z[:P.n], zr) < 1e-11
    ctx.close()


def test_apply_auto_variant_beyond_register_capacity():
    """More modes than one pass of the MMA kernel holds in registers (N > 2048): more passes, or the automatic choice falls
    back to the gather kernel; either way the result agrees with the reference-order gather kernel."""
    g = A.structured_unitsquare(17)
    fes = A.FESpace(g, 1)
    modes = A.graded_lex_multiindices(20, 2600)
    TB = A.TensorizedBasis(A.LegendrePolynomials, modes)
    sol = A.SGFEVector(fes, TB)
    A.setup_device_problem(sol, A.StochasticCoefficientCosinus(tau=0.9, decay=2, mean=1, maxm=20))
    ctx = TB.ctx
    ctx.vec_alloc(3)
    ctx.vec_fill_random(0, 7)
    ctx.set_apply_variant(1)
    ctx.apply(0, 1)
    ctx.set_apply_variant(0)
    ctx.apply(0, 2)
    a, b = ctx.vec_download(1), ctx.vec_download(2)
    assert relerr(b, a) < TOL_APPLY
    from asgfem_b200 import _lib
    with pytest.raises(_lib.AsgfemError):
        ctx.set_apply_variant(5)
    ctx.close()


@pytest.mark.parametrize("order", [1, 2])
def test_deterministic_sample_solutions_match_oracle(order):
    """Row f4 (src/sampling_error.jl:112-128): the deterministic reference solutions at samples, all samples as columns
    of one device block system, against one sparse direct solve per sample (1e-10 relative)."""
    m = omesh.uniform_refine(omesh.grid_unitsquare(), 3)
    Cc = ocoef.StochasticCoefficientCosinus(tau=0.9, decay=2.0, mean=1.0, maxm=6)
    Ms, ns = 6, 37
    P = oproblem.build(m, order, [[0] * Ms], opoly.LEGENDRE, Cc, bonus_quadorder_a=2)
    rng = np.random.default_rng(5)
    samples = rng.uniform(-1, 1, size=(Ms, ns))
    ref = osolver.deterministic_sample_solutions(P.A0, P.Am, P.b0, P.bdofs, samples)
    g = A.Grid(m.coords, m.cellnodes, m.bfacenodes)
    fes = A.FESpace(g, order)
    from asgfem_b200 import sgfem
    u, st = sgfem.deterministic_sample_solutions(fes, A.StochasticCoefficientCosinus(tau=0.9, decay=2.0, mean=1.0, maxm=6), samples)
    assert st["solved"] and st["niter"] < 60
    assert u.shape == ref.shape
    assert np.max(np.abs(u - ref)) <= 1e-10 * np.max(np.abs(ref))
    assert np.all(u[P.bdofs] == 0.0)


@pytest.mark.parametrize("family", [opoly.LEGENDRE, opoly.HERMITE])
def test_evaluate_samples_matches_oracle(family):
    """Row f4, set_sample! half (sgfevector.jl:43-69): u(x, xi_s) = sum_k H_k(xi_s) u_k for a batch of samples, with the
    univariate values TB.vals of set_sample!(TB, xi) (tensorizedbasis.jl:226-236) as data, incl. samples shorter than M."""
    modes = A.graded_lex_multiindices(6, 150)
    g = A.structured_unitsquare(21)
    fes = A.FESpace(g, 1)
    TB = A.TensorizedBasis(family, modes)
    sol = A.SGFEVector(fes, TB)
    rng = np.random.default_rng(12)
    sol.entries[:] = rng.standard_normal(sol.entries.shape)
    n, N, M = fes.ndofs, TB.nmodes, 6
    maxdeg = max(max(m) for m in modes)
    oTB = otb.TensorizedBasis(family, M, maxdeg, maxdeg + 2, multi_indices=[list(m) for m in TB.multi_indices])
    S = 13
    samples = [rng.uniform(-1, 1, size=(M if s % 3 else 4)) for s in range(S)]  # every third sample is shorter than M
    vals = np.zeros((S, M, oTB.ONB.maxorder + 1))
    ref = np.zeros((S, n))
    for s, xi in enumerate(samples):
        for d in range(M):
            if d < len(xi):
                vals[s, d] = oTB.ONB.evaluate(xi[d], normalize=True)
            else:
                vals[s, d, 0] = 1.0
        H = oTB.evaluate_all(xi, normalize=True)
        ref[s] = H @ sol.entries.reshape(N, n)
    ctx = TB.ctx
    # minimal device problem: the evaluation needs the vector layout only (pattern of any matrix on the space)
    A.setup_device_problem(sol, A.StochasticCoefficientCosinus(tau=0.9, decay=2, mean=1, maxm=M))
    ctx.vec_alloc(1)
    ctx.vec_upload(0, sol.entries)
    got = ctx.evaluate_samples(0, vals)
    assert got.shape == ref.shape
    assert np.abs(got - ref).max() <= 1e-12 * np.abs(ref).max()
    ctx.close()


def test_new_space_on_a_context_drops_the_derived_pattern():
    """a second asgfem_set_mesh / asgfem_set_space on a context must not assemble on the pattern derived
    from the first space (same ndofs, other connectivity)."""
    g1 = A.structured_unitsquare(9)
    perm = np.random.default_rng(3).permutation(g1.nnodes)
    inv = np.empty_like(perm)
    inv[perm] = np.arange(len(perm))
    coords2 = np.empty_like(g1.coords)
    coords2[inv] = g1.coords
    g2 = A.Grid(coords2, inv[g1.cellnodes], inv[g1.bfacenodes])
    Cf = A.StochasticCoefficientCosinus(tau=0.9, decay=2.0, mean=1.0, maxm=3)
    xref, w = A.quadrature_rule(2)

    def assemble(ctx, g):
        fes = A.FESpace(g, 1)
        ctx.set_mesh(g.coords, g.cellnodes + 1)
        ctx.set_space(1, fes.ndofs, fes.celldofs + 1)
        ctx.set_coefficient_cosinus(Cf.mean_value, Cf.decay_factors, Cf.b1, Cf.b2)
        ctx.assemble_stiffness(3, xref, w)
        cp, rv = ctx.pattern_csc()
        return cp.copy(), rv.copy(), [ctx.get_stiffness(m).copy() for m in range(4)]

    ctx = A.Context()
    ctx.set_multiindices(A.LEGENDRE, np.array([[0, 0, 0], [1, 0, 0]], dtype=np.int64))
    assemble(ctx, g1)
    cp2, rv2, v2 = assemble(ctx, g2)   # reuse of the context
    fresh = A.Context()
    fresh.set_multiindices(A.LEGENDRE, np.array([[0, 0, 0], [1, 0, 0]], dtype=np.int64))
    cpf, rvf, vf = assemble(fresh, g2)
    assert np.array_equal(cp2, cpf) and np.array_equal(rv2, rvf)
    for a, b in zip(v2, vf):
        assert np.array_equal(a, b)
    ctx.close()
    fresh.close()


@pytest.mark.parametrize("order", [1, 2])
def test_logprimal_device_assembly_and_solve_match_oracle(order):
    """Row f1 completed: the Laplacian, the convection matrices N_m = -(grad a_m . grad u, v) and the load vectors
    b[mu] = (lambda_mu f, v) (logpoisson_primal.jl:95-128, expa_PCE_mop) assembled on the device against the oracle's
    restatement (1e-12), then the full solve against the oracle's solve_logpoisson_primal (GMRES) on the oracle's matrices."""
    m = omesh.uniform_refine(omesh.grid_unitsquare(), 3)
    Cc = ocoef.StochasticCoefficientCosinus(tau=0.5, decay=2.0, mean=0.0, maxm=10)
    modes = [[0, 0, 0], [1, 0, 0], [0, 1, 0], [2, 0, 0], [0, 0, 1], [1, 1, 0], [0, 2, 1]]
    f = lambda x, y: 1.0 + x * y  # noqa: E731
    space = ofem.FESpace(m, order)
    ip, idx, vals = ofem.assemble_logprimal_matrices(space, Cc, 3, bonus_quadorder=2)
    bref = ofem.assemble_logprimal_rhs(space, Cc, modes, f, bonus_quadorder=1)
    g = A.Grid(m.coords, m.cellnodes, m.bfacenodes)
    fes = A.FESpace(g, order)
    TB = A.TensorizedBasis(A.HermitePolynomials, modes)
    sol = A.SGFEVector(fes, TB)
    ctx = TB.ctx
    Cd = A.StochasticCoefficientCosinus(tau=0.5, decay=2.0, mean=0.0, maxm=10)
    ctx.set_mesh(g.coords, g.cellnodes + 1)
    ctx.set_space(order, fes.ndofs, fes.celldofs + 1)
    ctx.set_coefficient_cosinus(Cd.mean_value, Cd.decay_factors, Cd.b1, Cd.b2)
    xref, w = A.quadrature_rule(2 * order - 1 + 2)
    ctx.assemble_logprimal(3, xref, w)
    cp, rv = ctx.pattern_csc()
    n = fes.ndofs
    for k in range(4):
        got = sp.csc_matrix((ctx.get_stiffness(k), rv - 1, cp - 1), shape=(n, n))
        ref = ofem.csr(ip, idx, vals[k], n)
        assert abs(got - ref).max() <= 1e-12 * abs(ref).max(), k
    # load vectors
    xf, wf = A.quadrature_rule(order + 1)
    c = g.cellnodes
    x1, x2, x3 = g.coords[c[:, 0]], g.coords[c[:, 1]], g.coords[c[:, 2]]
    xq = x1[:, None, :] + xf[None, :, 0:1] * (x2 - x1)[:, None, :] + xf[None, :, 1:2] * (x3 - x1)[:, None, :]
    ctx.vec_alloc(2)
    ctx.assemble_logprimal_rhs(xf, wf, f(xq[:, :, 0], xq[:, :, 1]), 10, 1)
    bgot = ctx.vec_download(1).reshape(len(modes), n)
    assert np.abs(bgot - bref).max() <= 1e-12 * np.abs(bref).max()
    # full solve through the mirror of solve!(LogTransformedPoissonProblemPrimal, ...)
    sol.entries[:] = 0.0
    bd, st = A.solve_logpoisson(sol, Cd, f, bonus_quadorder_a=2, bonus_quadorder_f=1, return_stats=True)
    assert st["solved"]
    Aor = ofem.csr(ip, idx, vals[0], n)
    Nm = [ofem.csr(ip, idx, vals[k], n) for k in range(1, 4)]
    G = otb.coupling_matrix(opoly.HERMITE, modes)
    ref_sol = np.zeros(n * len(modes))
    osolver.solve_logpoisson_primal(ref_sol, Aor, 0 * Aor, Nm, list(bref), G, len(modes), space.bdofs)
    assert np.abs(sol.entries - ref_sol).max() <= 1e-10 * np.abs(ref_sol).max()
    assert np.array_equal(np.sort(np.asarray(bd) - 1), np.sort(space.bdofs))
    ctx.close()


@pytest.mark.parametrize("order,domain", [(1, "square"), (2, "square"), (2, "lshape")])
def test_logprimal_estimator_matches_oracle(order, domain):
    """Row f2: estimate(LogTransformedPoissonProblemPrimal, ...) (src/estimate.jl:70-257) on the device against the oracle's
    restatement: eta4cell, eta4modes (with the reference's '+=' for the active modes), zeta_data; 1e-10 relative."""
    from oracle import estimate as oest
    base = omesh.grid_unitsquare() if domain == "square" else omesh.grid_lshape()
    m = omesh.uniform_refine(base, 3)
    Cc = ocoef.StochasticCoefficientCosinus(tau=0.5, decay=2.0, mean=0.0, maxm=12)
    modes = [[0, 0, 0], [1, 0, 0], [0, 1, 0], [2, 0, 0], [0, 0, 1], [1, 1, 0]]
    f = lambda x, y: 1.0 + x * y  # noqa: E731
    space = ofem.FESpace(m, order)
    rng = np.random.default_rng(11)
    u = rng.standard_normal(len(modes) * space.ndofs) * 0.1
    em, ec, ext, zeta = oest.estimate_logpoisson_primal(space, u, modes, opoly.HERMITE, Cc, f, bonus_quadorder=2, tail_extension=(5, 2))
    g = A.Grid(m.coords, m.cellnodes, m.bfacenodes)
    fes = A.FESpace(g, order)
    TB = A.TensorizedBasis(A.HermitePolynomials, modes)
    sol = A.SGFEVector(fes, TB)
    Cd = A.StochasticCoefficientCosinus(tau=0.5, decay=2.0, mean=0.0, maxm=12)
    ctx = TB.ctx
    ctx.set_mesh(g.coords, g.cellnodes + 1)
    ctx.set_space(order, fes.ndofs, fes.celldofs + 1)
    ctx.set_coefficient_cosinus(Cd.mean_value, Cd.decay_factors, Cd.b1, Cd.b2)
    sol.entries[:] = u
    gm, gc, gext, gz = A.estimate_logpoisson(sol, Cd, f, bonus_quadorder=2, tail_extension=(5, 2))
    assert gext == ext
    assert gc.shape == ec.shape
    assert np.abs(gm - em).max() <= TOL_SOLVE * em.max()
    assert np.abs(gc - ec).max() <= TOL_SOLVE * ec.max()
    assert abs(gz - zeta[0]) <= 1e-9 * max(abs(zeta[1]), abs(zeta[2]))
    # caller-supplied lambda values at the quadrature points (the reference's interpolation hand-off): same numbers when
    # the exact values are passed in
    xref, _ = A.quadrature_rule(2 * (order - 1) + 2)
    xq = space.physical_points(xref)
    lam = ofem.lambda_mu(Cc, ext, xq[:, :, 0], xq[:, :, 1])
    gm2, gc2, _, gz2 = A.estimate_logpoisson(sol, Cd, f, bonus_quadorder=2, tail_extension=(5, 2), lambda_at_qp=lam)
    assert np.abs(gm2 - gm).max() <= 1e-12 * gm.max() and np.abs(gc2 - gc).max() <= 1e-12 * gc.max()
    ctx.close()
