import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import asgfem_b200 as A
g = A.structured_unitsquare(9)
fes = A.FESpace(g, 1)
TB = A.TensorizedBasis(A.LegendrePolynomials, A.graded_lex_multiindices(3, 5))
sol = A.SGFEVector(fes, TB)
A.setup_device_problem(sol, A.StochasticCoefficientCosinus(tau=0.9, decay=2.0, mean=1.0, maxm=3))
ctx = TB.ctx
ctx.vec_alloc(2)
ctx.vec_fill_random(0, 1)
ctx.set_apply_variant(7)
ctx.apply(0, 1)
print("done")
