"""ctypes binding of libasgfem_cuda.so (include/asgfem.h).  Fails loudly: there is no Python/CPU fallback."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libasgfem_cuda.so")

c_i32, c_i64, c_f64, c_u64 = C.c_int32, C.c_int64, C.c_double, C.c_uint64
P = C.POINTER
vp = C.c_void_p


class Stats(C.Structure):
    _fields_ = [("niter", c_i64), ("solved", c_i32), ("_pad", c_i32), ("rz0", c_f64), ("rzk", c_f64),
                ("residual", c_f64), ("ms_setup", c_f64), ("ms_iterations", c_f64), ("ms_apply", c_f64),
                ("ms_precond", c_f64)]


# every symbol include/asgfem.h declares: name -> (restype, argtypes)
PROTOTYPES = {
    "asgfem_create": (c_i32, [P(vp), c_i32]),
    "asgfem_destroy": (c_i32, [vp]),
    "asgfem_last_error": (C.c_char_p, [vp]),
    "asgfem_version": (C.c_char_p, []),
    "asgfem_coupling_weights": (c_i32, [c_i32, c_i64, vp, vp]),
    "asgfem_set_multiindices": (c_i32, [vp, c_i32, c_i64, c_i64, vp]),
    "asgfem_get_coupling_nnz": (c_i32, [vp, P(c_i64)]),
    "asgfem_get_coupling_csc": (c_i32, [vp, vp, vp, vp]),
    "asgfem_get_neighbours": (c_i32, [vp, vp, vp]),
    "asgfem_add_boundary_modes": (c_i32, [c_i64, c_i64, vp, c_i64, c_i64, c_i64, P(c_i64), P(c_i64), vp, c_i64]),
    "asgfem_classify_modes": (c_i32, [c_i64, c_i64, vp, c_i64, vp]),
    "asgfem_set_pattern_csc": (c_i32, [vp, c_i64, vp, vp]),
    "asgfem_set_num_stiffness": (c_i32, [vp, c_i32]),
    "asgfem_set_stiffness": (c_i32, [vp, c_i32, vp]),
    "asgfem_set_stiffness_csc": (c_i32, [vp, c_i32, vp, vp, vp]),
    "asgfem_get_stiffness": (c_i32, [vp, c_i32, vp]),
    "asgfem_get_pattern_nnz": (c_i32, [vp, P(c_i64)]),
    "asgfem_get_pattern_csc": (c_i32, [vp, vp, vp]),
    "asgfem_set_bdofs": (c_i32, [vp, c_i64, vp]),
    "asgfem_set_mesh": (c_i32, [vp, c_i64, c_i64, vp, vp]),
    "asgfem_set_space": (c_i32, [vp, c_i32, c_i64, c_i32, vp]),
    "asgfem_set_coefficient_cosinus": (c_i32, [vp, c_i64, c_f64, vp, vp, vp]),
    "asgfem_assemble_stiffness": (c_i32, [vp, c_i32, c_i32, vp, vp]),
    "asgfem_vec_alloc": (c_i32, [vp, c_i32]),
    "asgfem_vec_upload": (c_i32, [vp, c_i32, vp]),
    "asgfem_vec_download": (c_i32, [vp, c_i32, vp]),
    "asgfem_vec_zero": (c_i32, [vp, c_i32]),
    "asgfem_vec_fill_random": (c_i32, [vp, c_i32, c_u64]),
    "asgfem_vec_dot": (c_i32, [vp, c_i32, c_i32, P(c_f64)]),
    "asgfem_vec_axpy": (c_i32, [vp, c_f64, c_i32, c_i32]),
    "asgfem_vec_xpay": (c_i32, [vp, c_i32, c_f64, c_i32]),
    "asgfem_vec_copy": (c_i32, [vp, c_i32, c_i32]),
    "asgfem_apply": (c_i32, [vp, c_i32, c_i32]),
    "asgfem_apply_host": (c_i32, [vp, vp, vp]),
    "asgfem_set_apply_variant": (c_i32, [vp, c_i32]),
    "asgfem_apply_rows": (c_i32, [vp, c_i32, c_i32, c_i64, c_i64]),
    "asgfem_last_apply_ms": (c_i32, [vp, P(c_f64)]),
    "asgfem_last_estimate_ms": (c_i32, [vp, P(c_f64)]),
    "asgfem_precond_setup": (c_i32, [vp]),
    "asgfem_host_factor_solve": (c_i32, [c_i64, vp, vp, vp, vp, vp, vp, vp, vp, vp, c_i32]),
    "asgfem_precond_apply": (c_i32, [vp, c_i32, c_i32]),
    "asgfem_precond_apply_host": (c_i32, [vp, vp, vp]),
    "asgfem_pcg": (c_i32, [vp, vp, c_i32, c_f64, c_f64, c_i64, P(Stats)]),
    "asgfem_solve_primal_host": (c_i32, [vp, vp, vp, c_f64, c_f64, c_i64, P(Stats)]),
    "asgfem_evaluate_samples": (c_i32, [vp, c_i32, c_i64, c_i64, c_i32, vp, vp]),
    "asgfem_set_precond_matrix_csc": (c_i32, [vp, vp, vp, vp]),
    "asgfem_bicgstab": (c_i32, [vp, c_i32, c_i32, c_f64, c_f64, c_i64, P(Stats)]),
    "asgfem_solve_logprimal_host": (c_i32, [vp, vp, vp, c_f64, c_f64, c_i64, P(Stats)]),
    "asgfem_estimate_poisson_primal": (c_i32, [vp, c_i32, c_i64, c_i64, vp, c_i32, vp, vp, vp, c_i32, vp, vp, vp, vp]),
    "asgfem_estimate_logpoisson_primal": (c_i32, [vp, c_i32, c_i64, c_i64, vp, c_i32, vp, vp, vp, vp, c_i32, c_i32, vp, vp, vp, vp, vp]),
    "asgfem_estimate_poisson_primal_marking": (c_i32, [vp, c_i32, c_i64, c_i64, vp, c_i32, vp, vp, vp, c_i32, vp, vp, c_i64, vp, vp, vp]),
    "asgfem_set_owned_rows": (c_i32, [vp, c_i64]),
    "asgfem_set_owned_cells": (c_i32, [vp, c_i64, vp]),
    "asgfem_set_samples": (c_i32, [vp, c_i64, c_i64, vp]),
    "asgfem_assemble_logprimal": (c_i32, [vp, c_i32, c_i32, vp, vp]),
    "asgfem_assemble_logprimal_rhs": (c_i32, [vp, c_i32, vp, vp, vp, c_i32, c_i32]),
    "asgfem_solve_samples_host": (c_i32, [vp, vp, vp, c_f64, c_f64, c_i64, vp]),
    "asgfem_halo_exchange": (c_i32, [vp, c_i32]),
    "asgfem_vec_device_ptr": (c_i32, [vp, c_i32, P(vp), P(c_i64)]),
    "asgfem_pack_rows": (c_i32, [vp, c_i32, c_i64, vp, vp]),
    "asgfem_unpack_rows": (c_i32, [vp, c_i32, c_i64, vp, vp]),
    "asgfem_vec_dot_owned": (c_i32, [vp, c_i32, c_i32, P(c_f64)]),
    "asgfem_comm_unique_id": (c_i32, [vp]),
    "asgfem_comm_init": (c_i32, [vp, c_i32, c_i32, vp]),
    "asgfem_comm_destroy": (c_i32, [vp]),
    "asgfem_set_halo": (c_i32, [vp, c_i32, vp, vp, vp, vp, vp, c_i64, c_i64]),
    "asgfem_vec_dot_global": (c_i32, [vp, c_i32, c_i32, P(c_f64)]),
    "asgfem_precond_setup_global": (c_i32, [vp, c_i64, vp, vp, vp, c_i64, vp, vp, vp]),
}

_lib = None


class AsgfemError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libasgfem_cuda error {code}: {msg}")
        self.code = code


def load():
    """Loads the shared library and installs the prototypes.  Raises if it is missing - the CUDA extension
    is the product, nothing silently replaces it."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C extendableasgfem.jl_b200/csrc`); there is no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
