"""CPU checks of the host side of the mean preconditioner (csrc/chol.cpp: nested dissection + multifrontal Cholesky with
the dense kernels of csrc/dense_chol.cpp), through the context-free C-ABI entry asgfem_host_factor_solve.  The checker is
scipy's sparse LU on the Dirichlet-reduced K_0 assembled by the oracle; tolerance 1e-9 relative to |x|_inf (fp64
factorisation of matrices with condition numbers up to ~1e5)."""
import os
import subprocess
import sys

import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

import asgfem_b200 as A
from asgfem_b200 import _lib
from oracle import coefficient as ocoef
from oracle import fem as ofem
from oracle import mesh as omesh

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _mean_stiffness(mesh, order):
    space = ofem.FESpace(mesh, order)
    coeff = ocoef.StochasticCoefficientCosinus(tau=0.9, decay=2.0, mean=1.0, maxm=2)
    indptr, indices, vals = ofem.assemble_stiffness(space, coeff, 0)
    return space, indptr, indices, vals[0]


def _dof_coords(space):
    m = space.mesh
    xy = m.coords.T  # (2, nnodes)
    if space.order == 1:
        return xy
    fn = m.facenodes
    mid = 0.5 * (xy[:, fn[:, 0]] + xy[:, fn[:, 1]])
    return np.hstack([xy, mid])


def _reference_solve(indptr, indices, k0, bdofs1, b):
    n = len(indptr) - 1
    K = sp.csr_matrix((k0, indices, indptr), shape=(n, n)).tocsc()
    keep = np.setdiff1d(np.arange(n), np.asarray(bdofs1) - 1)
    x = np.zeros(n)
    x[keep] = spla.splu(K[keep][:, keep]).solve(b[keep])
    return x


@pytest.mark.parametrize("case", ["p1_square", "p2_square", "p1_lshape", "p2_lshape_nocoords", "p1_threads"])
def test_host_factor_solve_matches_sparse_lu(case, monkeypatch):
    order = 2 if case.startswith("p2") else 1
    base = omesh.grid_lshape() if "lshape" in case else omesh.grid_unitsquare()
    nref = {"p1_square": 5, "p2_square": 4, "p1_lshape": 5, "p2_lshape_nocoords": 4, "p1_threads": 8}[case]
    if case == "p1_threads":
        monkeypatch.setenv("ASGFEM_CHOL_THREADS", "4")  # n > 20000: subtrees in parallel + threaded fronts on top
    space, indptr, indices, k0 = _mean_stiffness(omesh.uniform_refine(base, nref), order)
    rng = np.random.default_rng(7)
    b = rng.standard_normal(space.ndofs)
    coords = None if "nocoords" in case else _dof_coords(space)
    bd0 = space.bdofs  # 0-based in the oracle
    x, lnz = A.host_factor_solve(indptr, indices, k0, bd0 + 1, b, coords)
    ref = _reference_solve(indptr, indices, k0, bd0 + 1, b)
    assert lnz > 0
    assert np.all(x[bd0] == 0.0)
    assert np.max(np.abs(x - ref)) <= 1e-9 * np.max(np.abs(ref))


def test_host_factor_rejects_indefinite_matrix():
    space, indptr, indices, k0 = _mean_stiffness(omesh.uniform_refine(omesh.grid_unitsquare(), 3), 1)
    bd0 = space.bdofs
    with pytest.raises(_lib.AsgfemError) as e:
        A.host_factor_solve(indptr, indices, -k0, bd0 + 1, np.ones(space.ndofs))
    assert "not positive definite" in str(e.value)
    # no Dirichlet dofs: K_0 is singular (constants), the pivot test must catch it
    with pytest.raises(_lib.AsgfemError):
        A.host_factor_solve(indptr, indices, k0, np.zeros(0, dtype=np.int64), np.ones(space.ndofs))


@pytest.mark.parametrize("isa", ["c", "avx2"])
def test_dense_kernel_variants_agree(isa):
    """The portable and the AVX2 register tiles (ASGFEM_CHOL_ISA, read once per process) give the same solve."""
    code = (
        "import numpy as np, sys; sys.path[:0] = [%r, %r]\n"
        "import asgfem_b200 as A\n"
        "from test_host_cholesky_cpu import _mean_stiffness, _dof_coords, _reference_solve\n"
        "from oracle import mesh as omesh\n"
        "space, ip, ix, k0 = _mean_stiffness(omesh.uniform_refine(omesh.grid_unitsquare(), 5), 2)\n"
        "b = np.linspace(-1, 1, space.ndofs) ** 3\n"
        "bd0 = space.bdofs\n"
        "x, _ = A.host_factor_solve(ip, ix, k0, bd0 + 1, b, _dof_coords(space))\n"
        "ref = _reference_solve(ip, ix, k0, bd0 + 1, b)\n"
        "assert np.max(np.abs(x - ref)) <= 1e-9 * np.max(np.abs(ref)), np.max(np.abs(x - ref))\n" % (ROOT, os.path.join(ROOT, "tests")))
    env = dict(os.environ, ASGFEM_CHOL_ISA=isa)
    subprocess.run([sys.executable, "-c", code], check=True, cwd=ROOT, env=env, timeout=300)


def test_factorisation_is_bit_reproducible_across_thread_counts(monkeypatch):
    """Every entry of a front is updated by one register tile per panel, in panel order, and the extend-add visits the
    children in tree order: the factor (hence the solve) does not depend on how the work was spread over the threads."""
    space, indptr, indices, k0 = _mean_stiffness(omesh.uniform_refine(omesh.grid_unitsquare(), 8), 1)
    b = np.cos(np.arange(space.ndofs) * 0.01)
    xs = []
    for threads in ("1", "3", "8"):
        monkeypatch.setenv("ASGFEM_CHOL_THREADS", threads)
        xs.append(A.host_factor_solve(indptr, indices, k0, space.bdofs + 1, b, _dof_coords(space))[0])
    assert np.array_equal(xs[0], xs[1]) and np.array_equal(xs[0], xs[2])


def test_multifrontal_and_task_builder_against_their_serial_references():
    """tools/chol_bench.cpp (host functions of libasgfem_cuda.so, no GPU): the multifrontal factor has the patterns of
    the scalar up-looking factorisation and its values to 1e-9 (observed 1e-14), and the parallel sweep-task builder
    produces byte for byte the records, descriptors and launch list of the serial builder it replaced."""
    import shutil
    if shutil.which("g++") is None:
        pytest.skip("no host compiler")
    env = dict(os.environ, ASGFEM_CHOL_VERBOSE="", TMPDIR=os.environ.get("TMPDIR", "/tmp"))
    env.pop("ASGFEM_CHOL_UPLOOKING", None)
    r = subprocess.run([os.path.join(ROOT, "tools", "chol_bench.sh"), "180", "check"], cwd=ROOT, env=env, timeout=600,
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "patterns identical: yes" in r.stdout
    assert r.stdout.count("identical: yes") == 3  # factor patterns + the task records of both factors
