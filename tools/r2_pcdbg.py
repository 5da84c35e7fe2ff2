import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import asgfem_b200 as A
g = A.structured_unitsquare(33)
fes = A.FESpace(g, 1)
TB = A.TensorizedBasis(A.LegendrePolynomials, A.graded_lex_multiindices(3, 20))
sol = A.SGFEVector(fes, TB)
A.setup_device_problem(sol, A.StochasticCoefficientCosinus(tau=0.9, decay=2.0, mean=1.0, maxm=3))
ctx = TB.ctx
ctx.precond_setup()
ctx.vec_alloc(4)
ctx.vec_fill_random(0, 1)
ctx.vec_fill_random(2, 2)
ctx.precond_apply(0, 1)
a1 = ctx.vec_download(1).copy()
ctx.precond_apply(2, 3)
ctx.precond_apply(0, 1)
a2 = ctx.vec_download(1).copy()
print("repeat diff", np.abs(a1 - a2).max(), np.abs(a1).max())
for split in ("1", "3"):
    pass
