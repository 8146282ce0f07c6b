"""ctypes binding of libnonlin_b200.so — the C ABI declared in include/nonlin_batch.h.

There is no Python or CPU implementation behind this module: if the shared library is missing
or does not load, importing the package fails.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NLB_LIB") or os.path.join(HERE, "libnonlin_b200.so")   # NLB_LIB: tuning builds only

# API-level errors (include/nonlin_batch.h)
NLB_OK = 0
NLB_ERR_INVALID_ARGUMENT = 1
NLB_ERR_UNKNOWN_FCN = 2
NLB_ERR_SIZE = 3
NLB_ERR_UNSUPPORTED = 4
NLB_ERR_CUDA = 5
NLB_ERR_NO_DEVICE = 6

# per-system status codes (reference src/nonlin_error_handling.f90:10-38)
NL_NO_ERROR = 0
NL_INVALID_INPUT_ERROR = 201
NL_ARRAY_SIZE_ERROR = 202
NL_CONVERGENCE_ERROR = 106
NL_DIVERGENT_BEHAVIOR_ERROR = 206
NL_SPURIOUS_CONVERGENCE_ERROR = 207
NL_TOLERANCE_TOO_SMALL_ERROR = 208
NL_UNDEFINED_FUNCTION_ERROR = 211
NL_UNDERDEFINED_PROBLEM_ERROR = 212
LA_INVALID_OPERATION_ERROR = 107   # linalg: rank-deficient solve_least_squares (polynomial fit)
NL_LA_INVALID_OPERATION_ERROR = LA_INVALID_OPERATION_ERROR   # header name: NLB_LA_INVALID_OPERATION_ERROR

NLB_STAT_NAMES = ["systems", "converged", "converged_fcn", "converged_chng", "converged_zero_diff", "failed",
                  "sum_iter", "sum_fcn", "sum_jac", "max_iter"]
NLB_STAT_COUNT = 16

EXPORTS = [
    "nlb_create", "nlb_destroy", "nlb_last_error", "nlb_kernel_launch_count", "nlb_params_default",
    "nlb_vecfcn_count", "nlb_vecfcn_lookup", "nlb_vecfcn_name", "nlb_vecfcn_info",
    "nlb_least_squares_solve_batch", "nlb_newton_solve_batch", "nlb_quasi_newton_solve_batch",
    "nlb_vecfcn_eval_batch", "nlb_jacobian_batch", "nlb_reduce_stats", "nlb_measure_fp64_peak",
    "nlb_measure_fp64_latency", "nlb_constrained_options_default", "nlb_constrained_least_squares_solve_batch",
    "nlb_polynomial_fit_batch", "nlb_polynomial_evaluate_batch",
    "nlb_params_1var_default", "nlb_fcn1var_count", "nlb_fcn1var_lookup", "nlb_fcn1var_name", "nlb_fcn1var_info",
    "nlb_brent_solve_batch", "nlb_newton_1var_solve_batch", "nlb_solve_sharded", "nlb_register_vecfcn", "nlb_load_plugin",
]

NLB_SOLVER_LEAST_SQUARES, NLB_SOLVER_NEWTON, NLB_SOLVER_QUASI_NEWTON = 0, 1, 2


class nlb_params(C.Structure):
    _fields_ = [
        ("max_fcn_evals", C.c_int32),
        ("fcn_tol", C.c_double),
        ("var_tol", C.c_double),
        ("grad_tol", C.c_double),
        ("lm_factor", C.c_double),
        ("jacobian_interval", C.c_int32),
        ("use_line_search", C.c_int32),
        ("ls_max_fcn_evals", C.c_int32),
        ("ls_alpha", C.c_double),
        ("ls_factor", C.c_double),
        ("use_analytic_jacobian", C.c_int32),
        ("max_iter_guard", C.c_int32),
    ]


class nlb_params_1var(C.Structure):
    _fields_ = [
        ("max_fcn_evals", C.c_int32),
        ("fcn_tol", C.c_double),
        ("var_tol", C.c_double),
        ("diff_tol", C.c_double),
        ("use_analytic_diff", C.c_int32),
    ]


class nlb_constrained_options(C.Structure):
    _fields_ = [
        ("trust_region_radius", C.c_double),
        ("step_scaling_factor", C.c_double),
        ("lower", C.c_void_p),
        ("upper", C.c_void_p),
    ]


IB_DTYPE = np.dtype(
    [
        ("iter_count", "<i4"),
        ("fcn_count", "<i4"),
        ("jacobian_count", "<i4"),
        ("gradient_count", "<i4"),
        ("converge_on_fcn", "<i4"),
        ("converge_on_chng", "<i4"),
        ("converge_on_zero_diff", "<i4"),
    ]
)
assert IB_DTYPE.itemsize == 28


def load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "nonlin_b200: %s is missing. Build it with `python nonlin_b200/build.py` "
            "(nvcc, sm_100a). The engine has no CPU fallback." % LIB_PATH
        )
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, dbl = C.c_void_p, C.c_int, C.c_int64, C.c_double
    lib.nlb_create.argtypes = [C.POINTER(vp), i32]
    lib.nlb_destroy.argtypes = [vp]
    lib.nlb_last_error.argtypes = [vp]
    lib.nlb_last_error.restype = C.c_char_p
    lib.nlb_kernel_launch_count.argtypes = [vp]
    lib.nlb_kernel_launch_count.restype = i64
    lib.nlb_params_default.argtypes = [C.POINTER(nlb_params)]
    lib.nlb_params_default.restype = None
    lib.nlb_vecfcn_lookup.argtypes = [C.c_char_p]
    lib.nlb_vecfcn_name.argtypes = [i32]
    lib.nlb_vecfcn_name.restype = C.c_char_p
    lib.nlb_vecfcn_info.argtypes = [i32] + [C.POINTER(i32)] * 5
    solve_args = [vp, C.POINTER(nlb_params), i32, i64, i32, i32, vp, vp, vp, vp, vp, vp, vp]
    for name in ("nlb_least_squares_solve_batch", "nlb_newton_solve_batch", "nlb_quasi_newton_solve_batch"):
        getattr(lib, name).argtypes = solve_args
    lib.nlb_constrained_options_default.argtypes = [C.POINTER(nlb_constrained_options)]
    lib.nlb_constrained_options_default.restype = None
    lib.nlb_constrained_least_squares_solve_batch.argtypes = (
        [vp, C.POINTER(nlb_params), C.POINTER(nlb_constrained_options)] + solve_args[2:])
    lib.nlb_polynomial_fit_batch.argtypes = [vp, i64, i32, i32, i32, i32, vp, vp, vp, vp, vp]
    lib.nlb_polynomial_evaluate_batch.argtypes = [vp, i64, i32, i32, i32, vp, vp, vp, vp]
    lib.nlb_params_1var_default.argtypes = [C.POINTER(nlb_params_1var)]
    lib.nlb_params_1var_default.restype = None
    lib.nlb_fcn1var_lookup.argtypes = [C.c_char_p]
    lib.nlb_fcn1var_name.argtypes = [i32]
    lib.nlb_fcn1var_name.restype = C.c_char_p
    lib.nlb_fcn1var_info.argtypes = [i32, C.POINTER(i32), C.POINTER(i32)]
    for name in ("nlb_brent_solve_batch", "nlb_newton_1var_solve_batch"):
        getattr(lib, name).argtypes = [vp, C.POINTER(nlb_params_1var), i32, i64, vp, vp, vp, vp, vp, vp, vp, vp]
    lib.nlb_vecfcn_eval_batch.argtypes = [vp, i32, i64, i32, i32, vp, vp, vp, vp, vp]
    lib.nlb_jacobian_batch.argtypes = [vp, C.POINTER(nlb_params), i32, i64, i32, i32, vp, vp, vp, vp, vp]
    lib.nlb_reduce_stats.argtypes = [vp, i64, vp, vp, vp, vp]
    lib.nlb_load_plugin.argtypes = [C.c_char_p]
    lib.nlb_solve_sharded.argtypes = [C.POINTER(vp), i32, i32, C.POINTER(nlb_params), i32, i64, i32, i32, vp, vp, vp, vp, vp,
                                      vp, vp]
    lib.nlb_measure_fp64_peak.argtypes = [vp, C.POINTER(dbl), C.POINTER(dbl)]
    lib.nlb_measure_fp64_latency.argtypes = [vp, C.POINTER(dbl)]
    return lib
