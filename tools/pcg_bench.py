"""PCG solve timing on synthetic problems (device-assembled K_m, f = 1): prints one JSON line per size."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import asgfem_b200 as A


def run(nx, N, M=20, variant=0):
    g = A.structured_unitsquare(nx)
    fes = A.FESpace(g, 1)
    modes = A.graded_lex_multiindices(M, N)
    TB = A.TensorizedBasis(A.LegendrePolynomials, modes)
    sol = A.SGFEVector(fes, TB)
    t0 = time.time()
    A.setup_device_problem(sol, A.StochasticCoefficientCosinus(tau=0.9, decay=2.0, mean=1.0, maxm=M))
    ctx = TB.ctx
    ctx.set_apply_variant(variant)
    t_asm = time.time() - t0
    t0 = time.time()
    ctx.precond_setup()
    t_fac = time.time() - t0
    ctx.vec_alloc(1)
    ctx.vec_zero(0)
    b0 = fes.rhs()
    t0 = time.time()
    st = ctx.pcg(b0, 0, 1e-14, 1e-14, 500)
    t_solve = time.time() - t0
    out = dict(n=fes.ndofs, N=N, M=M, assemble_s=round(t_asm, 2), factor_s=round(t_fac, 2), solve_s=round(t_solve, 2), **st)
    print(json.dumps(out), flush=True)
    ctx.close()


if __name__ == "__main__":
    sizes = [(129, 200), (257, 500), (513, 1000)] + ([(1024, 2000)] if len(sys.argv) > 1 else [])
    if len(sys.argv) > 1 and sys.argv[1] == "c4":  # config 4 only
        sizes = [(1024, 2000)]
    for nx, N in sizes:
        run(nx, N)
