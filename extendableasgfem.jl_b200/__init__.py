"""B200-native SGFE solve hot path behind the ExtendableASGFEM.jl interface.

The compute lives in libasgfem_cuda.so (hand-written CUDA for sm_100a, C ABI in include/asgfem.h); this
package is the host-side mirror of the reference's operator interface plus the ctypes binding.
"""
from . import _lib
from .context import Context, LEGENDRE, HERMITE, coupling_weights, host_factor_solve  # noqa: F401
from .coefficients import StochasticCoefficientCosinus  # noqa: F401
from .multiindices import (generate_multiindices, prepare_multi_indices, add_boundary_modes,  # noqa: F401
                           classify_modes, graded_lex_multiindices)
from .grids import (Grid, FESpace, grid_unitsquare, grid_lshape, uniform_refine, structured_unitsquare,  # noqa: F401
                    quadrature_rule, quadrature_rule_1d)
from .sgfem import (TensorizedBasis, SGFEVector, solve_primal, solve, solve_logpoisson_primal, solve_logpoisson, estimate_logpoisson, set_samples,  # noqa: F401
                    setup_device_problem, mul, ldiv,
                    estimate, LegendrePolynomials, HermitePolynomials)

__version__ = "0.1.0"
