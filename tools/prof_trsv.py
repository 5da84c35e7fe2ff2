"""One config-4 preconditioner setup + applications (for ncu launch lists / captures of the triangular sweeps)."""
import os
import sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import asgfem_b200 as A
nx = int(os.environ.get("NX", 1025))
nm = int(os.environ.get("NMODES", 2000))
g = A.structured_unitsquare(nx)
fes = A.FESpace(g, 1)
TB = A.TensorizedBasis(A.LegendrePolynomials, A.graded_lex_multiindices(20, nm))
sol = A.SGFEVector(fes, TB)
A.setup_device_problem(sol, A.StochasticCoefficientCosinus(tau=0.9, decay=2.0, mean=1.0, maxm=20))
ctx = TB.ctx
ctx.precond_setup()
ctx.vec_alloc(2)
ctx.vec_fill_random(0, 1)
import time, torch
reps = int(os.environ.get("REPS", 2))
ctx.precond_apply(0, 1)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(reps):
    ctx.precond_apply(0, 1)
torch.cuda.synchronize()
print("precond_apply ms:", (time.perf_counter() - t0) / reps * 1e3, "leaf", os.environ.get("ASGFEM_CHOL_LEAF", "24"), flush=True)
