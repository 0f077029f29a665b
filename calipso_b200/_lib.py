"""ctypes binding of libcalipso_b200.so (include/calipso_b200.h).

The product loads ONLY the CUDA library built by calipso_b200/build.py; if it is missing or no CUDA device is
present the constructors raise -- there is no CPU fallback.  (``Binding(path)`` takes an explicit path so that the
test-suite can point the same binding at its host-emulation build of the device code; nothing in this package does.)
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcalipso_b200.so")

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)
c_llp = C.POINTER(C.c_longlong)

# enums of include/calipso_b200.h
ARRAYS = ["POINT", "CANDIDATE", "STEP", "RESIDUAL", "GRADIENT", "EQ_DUAL_GRAD", "CONE_DUAL_GRAD", "EQUALITY", "CONE",
          "W_VALUES", "G_VALUES", "C_VALUES", "CONE_PRODUCT", "BARRIER_GRADIENT", "DUAL", "LQ_Q", "LQ_G0", "LQ_H0",
          "SCALARS", "MERIT_GRADIENT", "RESIDUAL_SYMMETRIC", "STEP_SYMMETRIC", "PIVOTS", "MATRIX_VALUES", "RHS",
          "PANELS"]
A = {name: i for i, name in enumerate(ARRAYS)}
SCALARS = ["kappa", "tau", "rho", "eps_p", "eps_d", "eps_p_last", "objective", "barrier", "residual_violation",
           "optimality_violation", "slack_violation", "theta", "merit", "step_size", "step_size_t",
           "equality_violation", "cone_product_violation", "refine_norm", "refine_norm_initial", "merit_candidate",
           "theta_candidate", "merit_slope", "step_size_cone"]
S = {name: i for i, name in enumerate(SCALARS)}
S_COUNT = 24
STATS = ["inertia_pos", "inertia_neg", "inertia_zero", "n_trials", "n_refine", "refine_ok", "k_s", "k_t", "status",
         "used_fallback", "fallbacks", "total_iterations", "outer", "line_search", "converged", "gmres_iters",
         "filter_index", "inner", "factorizations", "solves", "unrefined_steps"]
I = {name: i for i, name in enumerate(STATS)}
I_COUNT = 24
STATUS_TEXT = {0: "ok", 1: "inertia correction failure", 2: "iterative refinement failure", 3: "cone search failure",
               4: "zero pivot"}

PROFILE = ["assemble", "factor_leaves", "factor_small", "factor_big_stage", "factor_big_gemm", "factor_big_panel",
           "factor_big_generic", "solve_fwd", "solve_bwd", "rhs_recover", "jtimes", "eval_linesearch", "cone_residual",
           "inertia", "total", "sf_bulk", "sf_pull", "sf_sweep", "sf_push", "sf_other", "sb_gather", "sb_sweep", "sb_other",
           "sb_leaves", "chain_fwd_tma", "chain_fwd_sweep", "chain_fwd_wait", "chain_bwd_tma", "chain_bwd_wait",
           "chain_bwd_sweep", "spare0", "spare1"]
PROF_COUNT = 32

EV_OBJECTIVE, EV_GRADIENT, EV_EQUALITY, EV_CONE, EV_EQUALITY_DUAL_GRAD, EV_CONE_DUAL_GRAD = 1, 2, 4, 8, 16, 32
EV_HESSIAN, EV_EQUALITY_JAC, EV_CONE_JAC = 64, 128, 256
CONE_BARRIER, CONE_BARRIER_GRADIENT, CONE_PRODUCT = 1, 2, 4


class COptions(C.Structure):
    """cb200_options: src/solver/options.jl:6-59 hot-path subset + GMRES fallback knobs."""
    _fields_ = ([(k, C.c_int) for k in ("max_outer_iterations", "max_residual_iterations", "max_residual_line_search",
                                        "max_cone_line_search", "iterative_refinement", "max_iterative_refinement",
                                        "min_iterative_refinement")] +
                [(k, C.c_double) for k in (
                    "scaling_line_search", "iterative_refinement_tolerance", "central_path_initial",
                    "central_path_update_tolerance", "central_path_scaling", "central_path_exponent",
                    "penalty_initial", "penalty_scaling", "dual_initial", "residual_tolerance", "optimality_tolerance",
                    "slack_tolerance", "equality_tolerance", "complementarity_tolerance", "min_regularization",
                    "primal_regularization_initial", "dual_regularization_initial", "max_regularization",
                    "dual_regularization", "dual_regularization_exponent", "scaling_regularization_initial",
                    "scaling_regularization", "scaling_regularization_last", "max_penalty", "violation_tolerance",
                    "violation_exponent", "merit_tolerance", "merit_exponent", "armijo_tolerance",
                    "machine_tolerance")] +
                [(k, C.c_int) for k in ("max_filter", "gmres_restart", "gmres_max_cycles")])


# every symbol include/calipso_b200.h declares: name -> (restype, argtypes)
vp = C.c_void_p
SYMBOLS = {
    "cb200_options_default": (None, [C.POINTER(COptions)]),
    "cb200_last_error": (C.c_char_p, []),
    "cb200_device_count": (C.c_int, []),
    "cb200_create": (vp, [C.c_int] * 6 + [c_ip] * 8 + [C.POINTER(COptions), C.c_int]),
    "cb200_ldl_create": (vp, [C.c_int, C.c_int, c_ip, c_ip, c_ip, C.c_int]),
    "cb200_destroy": (None, [vp]),
    "cb200_info": (C.c_int, [vp, c_llp]),
    "cb200_amd_order": (C.c_int, [C.c_int, c_ip, c_ip, c_ip]),
    "cb200_path_info": (C.c_int, [vp, c_llp]),
    "cb200_get_symbolic": (C.c_int, [vp, c_ip, c_ip, c_ip]),
    "cb200_get_factor": (C.c_int, [vp, C.c_int, c_ip, c_ip, c_dp, c_dp]),
    "cb200_set_array": (C.c_int, [vp, C.c_int, c_dp, C.c_int, C.c_int]),
    "cb200_get_array": (C.c_int, [vp, C.c_int, c_dp, C.c_int, C.c_int]),
    "cb200_initialize": (C.c_int, [vp, c_dp, C.c_int, C.c_int]),
    "cb200_get_stats": (C.c_int, [vp, c_ip, C.c_int, C.c_int]),
    "cb200_get_array_async": (C.c_int, [vp, C.c_int, c_dp, C.c_int, C.c_int]),
    "cb200_get_stats_async": (C.c_int, [vp, c_ip, C.c_int, C.c_int]),
    "cb200_array_length": (C.c_int, [vp, C.c_int]),
    "cb200_get_profile": (C.c_int, [vp, c_llp, C.c_int]),
    "cb200_device_ptr": (vp, [vp, C.c_int]),
    "cb200_values_changed": (C.c_int, [vp]),
    "cb200_stream": (vp, [vp]),
    "cb200_synchronize": (C.c_int, [vp]),
    "cb200_set_options": (C.c_int, [vp, C.POINTER(COptions)]),
    "cb200_cone": (C.c_int, [vp, C.c_int, C.c_int]),
    "cb200_residual": (C.c_int, [vp]),
    "cb200_search_direction": (C.c_int, [vp]),
    "cb200_cone_search": (C.c_int, [vp]),
    "cb200_apply_step": (C.c_int, [vp]),
    "cb200_kkt_factor_solve": (C.c_int, [vp, C.c_int]),
    "cb200_jacobian_times": (C.c_int, [vp, c_dp, c_dp]),
    "cb200_differentiate": (C.c_int, [vp, C.c_int, c_dp, c_dp]),
    "cb200_scatter_plan": (C.c_int, [vp, C.c_int, C.c_int, c_ip, c_ip, c_ip]),
    "cb200_scatter": (C.c_int, [vp, C.c_int, c_dp, C.c_int, C.c_int]),
    "cb200_scatter_buffer": (vp, [vp, C.c_int]),
    "cb200_stage_plan": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, c_ip]),
    "cb200_stage_scatter": (C.c_int, [vp, C.c_int, c_dp, C.c_int, C.c_int]),
    "cb200_stage_buffer": (vp, [vp, C.c_int]),
    "cb200_lq_evaluate": (C.c_int, [vp, C.c_int, C.c_int]),
    "cb200_lq_begin": (C.c_int, [vp, C.c_int]),
    "cb200_lq_step": (C.c_int, [vp, C.c_int]),
    "cb200_lq_solve": (C.c_int, [vp, C.c_int, C.c_int, c_llp, c_ip]),
    "cb200_lq_set_order": (C.c_int, [vp, c_ip]),
    "cb200_filter_reset": (C.c_int, [vp]),
    "cb200_filter_search": (C.c_int, [vp, C.c_int, C.c_int, c_dp, c_dp, c_dp, c_ip]),
    "cb200_ldl_factorize": (C.c_int, [vp]),
    "cb200_ldl_inertia": (C.c_int, [vp, c_ip]),
    "cb200_ldl_solve": (C.c_int, [vp]),
    "cb200_ldl_linear_solve": (C.c_int, [vp, c_dp, c_dp, c_dp, C.c_int]),
    "cb200_nccl_unique_id": (C.c_int, [C.c_char_p]),
    "cb200_comm_init": (C.c_int, [vp, C.c_int, C.c_int, C.c_char_p]),
    "cb200_allreduce_counts": (C.c_int, [vp, c_llp]),
}


class CalipsoB200Error(RuntimeError):
    pass


class Binding:
    def __init__(self, path: str = LIB_PATH):
        if not os.path.exists(path):
            raise CalipsoB200Error(
                f"{path} not found: build the CUDA extension first (python -m calipso_b200.build); there is no CPU fallback")
        self.path = path
        self.lib = C.CDLL(path)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(self.lib, name)    # AttributeError if the library does not export a declared symbol
            fn.restype = res
            fn.argtypes = args

    def check(self, rc):
        if rc != 0:
            raise CalipsoB200Error(self.lib.cb200_last_error().decode())

    def default_options(self) -> COptions:
        o = COptions()
        self.lib.cb200_options_default(C.byref(o))
        return o


_DEFAULT = None


def default_binding() -> Binding:
    global _DEFAULT
    if _DEFAULT is None:
        _DEFAULT = Binding(LIB_PATH)
    return _DEFAULT


def i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def ip(a):
    return a.ctypes.data_as(c_ip)


def dp(a):
    return a.ctypes.data_as(c_dp)
