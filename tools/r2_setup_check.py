"""Preconditioner setup on a GPU box after the host-side rewrite (multifrontal Cholesky, parallel task builder): setup
time and the residual |K_0 z - b| / |b| of ldiv! on the interior dofs (scipy matvec only - independent of the factor).
No torch import: starts fast on a fresh box."""
import json
import os
import sys
import time

import numpy as np
import scipy.sparse as sp

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import asgfem_b200 as A


def run(nx, order, N):
    t0 = time.time()
    g = A.structured_unitsquare(nx)
    fes = A.FESpace(g, order)
    modes = A.graded_lex_multiindices(3, N)
    TB = A.TensorizedBasis(A.LegendrePolynomials, modes)
    sol = A.SGFEVector(fes, TB)
    A.setup_device_problem(sol, A.StochasticCoefficientCosinus(tau=0.9, decay=2, mean=1, maxm=3))
    ctx = TB.ctx
    t_problem = time.time() - t0
    t0 = time.time()
    ctx.precond_setup()
    t_setup = time.time() - t0
    n = fes.ndofs
    colptr, rowval = ctx.pattern_csc()
    K0 = sp.csc_matrix((ctx.get_stiffness(0), rowval - 1, colptr - 1), shape=(n, n)).tocsr()
    interior = np.setdiff1d(np.arange(n), fes.bdofs)
    b = np.random.default_rng(nx).standard_normal(n * N)
    b.reshape(N, n)[:, fes.bdofs] = 0
    t0 = time.time()
    z = ctx.precond_apply_host(b).reshape(N, n)
    t_apply = time.time() - t0
    r = (K0 @ z.T).T - b.reshape(N, n)
    res = float(np.abs(r[:, interior]).max() / np.abs(b).max())
    bzero = bool(np.all(z[:, fes.bdofs] == 0))
    print(json.dumps(dict(nx=nx, order=order, n=n, N=N, problem_s=round(t_problem, 2), precond_setup_s=round(t_setup, 3),
                          apply_host_s=round(t_apply, 3), residual=res, boundary_rows_zero=bzero)), flush=True)
    ctx.close()
    return res < 1e-9 and bzero


if __name__ == "__main__":
    ok = True
    for nx, order, N in [(129, 2, 24), (513, 1, 17), (1024, 1, 16)]:
        ok = run(nx, order, N) and ok
    print("SETUP CHECK", "OK" if ok else "FAILED", flush=True)
    sys.exit(0 if ok else 1)
