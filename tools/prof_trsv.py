import os
import sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import asgfem_b200 as A
g = A.structured_unitsquare(513)
fes = A.FESpace(g, 1)
TB = A.TensorizedBasis(A.LegendrePolynomials, A.graded_lex_multiindices(20, 1000))
sol = A.SGFEVector(fes, TB)
A.setup_device_problem(sol, A.StochasticCoefficientCosinus(tau=0.9, decay=2.0, mean=1.0, maxm=20))
ctx = TB.ctx
ctx.precond_setup()
ctx.vec_alloc(2)
ctx.vec_fill_random(0, 1)
for _ in range(3):
    ctx.precond_apply(0, 1)
