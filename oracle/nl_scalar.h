// oracle/nl_scalar.h — TEST INFRASTRUCTURE ONLY (see oracle/README.md).
//
// CPU restatement of the reference's one-variable solvers:
//   brent_solve     <- brent_solve      src/nonlin_solve.f90:643-835
//   newton1_solve   <- newt1var_solve   src/nonlin_solve.f90:840-1032
//   fd_diff         <- f1h_diff_fcn     src/nonlin_single_var.f90:154-200
// and the one-variable test functions (tests/nonlin_test_solve.f90:165-183, examples/example_problems.f90).
// Pinned by the roots the reference's tests assert (pi on [1.5, 5] for sin(x)/x, tests/nonlin_test_solve.f90:729-
// 790, 898-970); the reference publishes no counts or digits for these solvers.
#ifndef NL_SCALAR_H
#define NL_SCALAR_H

#include <cstdint>

#include "nl_solvers.h"

namespace nlo {

typedef real (*Fcn1)(real x, const real* args);

struct Problem1 {
    int id;
    const char* name;
    int args_len;
    Fcn1 fcn;
    Fcn1 diff;      // analytic derivative or null
};

enum {
    NL_FCN1_SINX_DIV_X = 0,      // sin(x)/x                         tests/nonlin_test_solve.f90:165-170 (libm sin)
    NL_FCN1_SINX_DIV_X_A = 1,    // a sin(x)/x, args = a             tests/nonlin_test_solve.f90:172-183
    NL_FCN1_CUBIC_WALLIS = 2,    // x**3 - 2x - 5                    (+ - * only: bitwise set)
    NL_FCN1_EXP_MINUS_X = 3,     // exp(-x) - x                      (shared software exp: bitwise set)
    NL_FCN1_CUBIC_ARGS = 4,      // a0 + a1 x + a2 x**2 + a3 x**3, args = a0..a3 (Horner; bitwise set)
    NL_FCN1_COUNT = 5
};

const Problem1* nl_problem1(int id);
const Problem1* nl_problem1_by_name(const char* name);

// equation_solver_1var members (src/nonlin_single_var.f90:44-54) + "set_diff was called"
struct Params1 {
    int32_t max_fcn_evals;     // m_maxEval = 100
    double fcn_tol;            // m_fcnTol  = 1e-8
    double var_tol;            // m_xtol    = 1e-12
    double diff_tol;           // m_difftol = 1e-12
    int32_t use_analytic_diff;
};
void params1_default(Params1* p);

real fd_diff(const Problem1* p, const real* args, const Params1* prm, real x, real f);
// f_present mirrors the optional `f` argument (newt1var_solve counts one more evaluation when it is given).
int brent_solve(const Problem1* p, const real* args, const Params1* prm, real lim1, real lim2, real* x, real* f,
                IterBehavior* ib);
int newton1_solve(const Problem1* p, const real* args, const Params1* prm, real lim1, real lim2, real* x, real* f,
                  bool f_present, IterBehavior* ib);

}  // namespace nlo
#endif
