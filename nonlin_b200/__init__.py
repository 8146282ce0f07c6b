"""nonlin_b200 — B200-native batched nonlinear-solver engine.

Re-exports, like the reference's umbrella module (src/nonlin.f90:13-63), the names of the
"M equations / N unknowns" solver family, here as the batch extension backed by
libnonlin_b200.so (hand-written sm_100a CUDA behind the C ABI of include/nonlin_batch.h).
Importing this package loads the shared library and fails loudly if it is missing.
"""
from ._lib import (  # noqa: F401
    IB_DTYPE,
    LA_INVALID_OPERATION_ERROR,
    NL_CONVERGENCE_ERROR,
    NL_DIVERGENT_BEHAVIOR_ERROR,
    NL_INVALID_INPUT_ERROR,
    NL_NO_ERROR,
    NL_SPURIOUS_CONVERGENCE_ERROR,
    NL_TOLERANCE_TOO_SMALL_ERROR,
    NL_UNDEFINED_FUNCTION_ERROR,
    NL_UNDERDEFINED_PROBLEM_ERROR,
    NLB_ERR_CUDA,
    NLB_ERR_INVALID_ARGUMENT,
    NLB_ERR_NO_DEVICE,
    NLB_ERR_SIZE,
    NLB_ERR_UNKNOWN_FCN,
    NLB_ERR_UNSUPPORTED,
)
from .api import (  # noqa: F401
    Engine,
    NonlinError,
    constrained_equation_solver,
    constrained_least_squares_solver,
    brent_solver,
    default_engine,
    equation_solver_1var,
    fcn1var_helper,
    fcn1var_names,
    equation_solver,
    ib_view,
    iteration_behavior,
    least_squares_solver,
    line_search,
    line_search_solver,
    load_plugin,
    newton_1var_solver,
    newton_solver,
    polynomial,
    quasi_newton_solver,
    value_pair,
    vecfcn_helper,
    vecfcn_names,
)
