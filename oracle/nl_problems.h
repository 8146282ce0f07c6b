// oracle/nl_problems.h — TEST INFRASTRUCTURE ONLY (see oracle/README.md).
//
// CPU restatement of the residual functions ("vecfcn", src/nonlin_multi_eqn_mult_var.f90:14-25)
// and optional Jacobians ("jacobianfcn", :27-38) that the reference's tests and examples use
// on the hot path, plus the synthetic models of BASELINE.json configs 4 and 5.  Each keeps
// the Fortran expression's evaluation order (left to right, x**2 = x*x, x**3 = (x*x)*x).
//
// The ids and names are the same ones the CUDA engine's registry uses
// (nonlin_b200/csrc/vecfcn_registry.cuh) — two independent implementations of one table.
#ifndef NL_PROBLEMS_H
#define NL_PROBLEMS_H

#include "nl_numerics.h"

namespace nlo {

struct FcnCtx {
    int m, n;
    const real* sys;      // per-system data block, contiguous (args analogue), may be null
    const real* shared;   // data shared by every system of the batch (e.g. abscissae), may be null
};

typedef void (*vecfcn_t)(const real* x, real* f, const FcnCtx* c);
typedef void (*jacfcn_t)(const real* x, real* jac, const FcnCtx* c);   // jac column-major m x n

struct Problem {
    int id;
    const char* name;
    int m, n;          // 0 = taken from the call (runtime-sized family)
    int sys_len;       // doubles of per-system data; -1 = m (one observation per equation)
    int shared_len;    // doubles of shared data; -1 = m
    vecfcn_t fcn;
    jacfcn_t jac;      // null = no analytic Jacobian registered
};

enum {
    NL_FCN_MISC_2FCN = 0,        // tests/nonlin_test_solve.f90:41-47 (fcn1), examples/example_problems.f90 misc_2fcn
    NL_FCN_MISC_2FCN_A = 1,      // tests/nonlin_test_solve.f90:49-60 (fcn1a), args = a
    NL_FCN_POORLY_SCALED = 2,    // tests/nonlin_test_solve.f90:109-115 (fcn2)
    NL_FCN_POWELL_BADLY_SCALED = 3,  // tests/powell_badly_scaled.f90:9-15
    NL_FCN_LSQ_POLY_FIT = 4,     // tests/nonlin_test_solve.f90:133-159 (lsfcn1) with per-system yp
    NL_FCN_POLAR = 5,            // tests/nonlin_test_jacobian.f90 fcn1
    NL_FCN_POLAR_SCALED = 6,     // tests/nonlin_test_jacobian.f90 fcn2, args = y
    NL_FCN_MISC_2FCN_01 = 7,     // examples/example_problems.f90 misc_2fcn_01
    NL_FCN_RATIONAL_7_8 = 8,     // BASELINE config 4 parity model
    NL_FCN_EXP_SUM_8 = 9,        // BASELINE config 4 throughput model
    NL_FCN_EXT_ROSENBROCK = 10,  // BASELINE config 5
    NL_FCN_EXP_DECAY_4 = 11,     // 4-parameter double-exponential curve fit (SURVEY §6 probe)
    NL_FCN_COUNT = 12
};

// Residuals supplied by the TEST as C callbacks (ids NL_FCN_COUNT ... NL_FCN_COUNT + NL_MAX_CALLBACKS - 1): the checker for
// residuals that reach the engine as plug-ins (nonlin_b200/csrc/nlb_plugin.cuh).  The callback works on plain doubles.
typedef void (*callback_t)(const double* x, double* f, const double* sys, const double* shared, int m, int n);
enum { NL_MAX_CALLBACKS = 4 };
int nl_register_callback(const char* name, int m, int n, int sys_len, int shared_len, callback_t fcn);   // id or -1

const Problem* nl_problem(int id);
const Problem* nl_problem_by_name(const char* name);

// abscissae and ordinates of README Example 2 (examples/example_problems.f90 lsq_poly_fit_fcn)
extern const double NL_POLYFIT_XP[21];
extern const double NL_POLYFIT_YP[21];

}  // namespace nlo
#endif
