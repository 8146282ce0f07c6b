// oracle/nl_solvers.h — TEST INFRASTRUCTURE ONLY (see oracle/README.md).
//
// CPU restatement, one system at a time, of the reference's hot path:
//   fd_jacobian     <- vfh_jac_fcn            src/nonlin_multi_eqn_mult_var.f90:198-277
//   lm_solve        <- lss_solve              src/nonlin_least_squares.f90:118-391
//   lm_par          <- lmpar                  src/nonlin_least_squares.f90:394-566
//   lm_factor       <- lmfactor               src/nonlin_least_squares.f90:569-667
//   lm_qrsolve      <- lmsolve                src/nonlin_least_squares.f90:670-791
//   newton_solve    <- ns_solve               src/nonlin_solve.f90:452-638
//   broyden_solve   <- qns_solve              src/nonlin_solve.f90:156-425
//   line_search     <- ls_search_mimo         src/nonlin_linesearch.f90:152-326
//   backtrack_min   <- min_backtrack_search   src/nonlin_linesearch.f90:495-551
//   limit_vector    <- limit_search_vector    src/nonlin_linesearch.f90:554-572
//   test_convergence<- test_convergence       src/nonlin_helper.f90:36-124
//   cls_solve       <- cls_solve              src/nonlin_least_squares.f90:938-1176
//   dogleg          <- dogleg                 src/nonlin_least_squares.f90:1301-1403
//   alpha_box, coleman_li_scaling             src/nonlin_least_squares.f90:1181-1260
// Where the reference executes `error stop <code>` the restatement returns <code> as the
// per-system status and leaves x / fvec / ib at the state they had at that point.
#ifndef NL_SOLVERS_H
#define NL_SOLVERS_H

#include <cstdint>
#include <vector>

#include "nl_problems.h"

namespace nlo {

// NL_* codes, src/nonlin_error_handling.f90:10-38.  NL_CONVERGENCE_ERROR aliases linalg's
// LA_CONVERGENCE_ERROR, whose value (106 in linalg's published C header) is not in the tree.
enum {
    NL_NO_ERROR = 0,
    NL_INVALID_INPUT_ERROR = 201,
    NL_ARRAY_SIZE_ERROR = 202,
    NL_CONVERGENCE_ERROR = 106,
    NL_DIVERGENT_BEHAVIOR_ERROR = 206,
    NL_SPURIOUS_CONVERGENCE_ERROR = 207,
    NL_TOLERANCE_TOO_SMALL_ERROR = 208,
    NL_UNDEFINED_FUNCTION_ERROR = 211,
    NL_UNDERDEFINED_PROBLEM_ERROR = 212
};

// Solver settings: the private members behind the reference's getters/setters.
struct Params {
    int32_t max_fcn_evals;     // equation_solver%m_maxEval   = 100    multi_eqn:69
    double fcn_tol;            // m_fcnTol                    = 1e-8   multi_eqn:71
    double var_tol;            // m_xtol                      = 1e-12  multi_eqn:73
    double grad_tol;           // m_gtol                      = 1e-12  multi_eqn:75
    double lm_factor;          // least_squares_solver%m_factor = 100  least_squares:25
    int32_t jacobian_interval; // quasi_newton_solver%m_jDelta = 5     solve:51
    int32_t use_line_search;   // line_search_solver%m_useLineSearch = true  solve:30
    int32_t ls_max_fcn_evals;  // line_search%m_maxEval = 100          linesearch:35
    double ls_alpha;           // line_search%m_alpha   = 1e-4         linesearch:38
    double ls_factor;          // line_search%m_factor  = 0.1          linesearch:46
    int32_t use_analytic_jacobian;  // vecfcn_helper%set_jacobian called   multi_eqn:143
    int32_t max_iter_guard;    // NOT in the reference: bound on Broyden's uncounted uphill restarts
};
void params_default(Params* p);

// iteration_behavior, src/nonlin_types.f90:8-29 (default LOGICAL = 4 bytes under gfortran)
struct IterBehavior {
    int32_t iter_count, fcn_count, jacobian_count, gradient_count;
    int32_t converge_on_fcn, converge_on_chng, converge_on_zero_diff;
};

struct Workspace {
    std::vector<real> buf;
    std::vector<int> ibuf;
    real* get(size_t n) { if (buf.size() < n) buf.resize(n); return buf.data(); }
    int* geti(size_t n) { if (ibuf.size() < n) ibuf.resize(n); return ibuf.data(); }
};

void fd_jacobian(const Problem* p, const FcnCtx* c, const Params* prm, real* x, real* jac, const real* fv, real* wrk);

void test_convergence(int nvar, int neqn, const real* x, const real* xo, const real* f, const real* g, bool lg,
                      real xtol, real ftol, real gtol, bool* c, bool* cx, bool* cf, bool* cg, real* xnorm,
                      real* fnorm);

void limit_vector(int n, real* x, real lim);
real backtrack_min(int mode, real f0, real f, real f1, real alam, real alam1, real slope);
int line_search(const Problem* p, const FcnCtx* c, const Params* prm, const real* xold, const real* grad,
                const real* dir, real* x, real* fvec, real fold, real* fx, IterBehavior* ib);

void lm_factor(int m, int n, real* a, bool pivot, int* ipvt, real* rdiag, real* acnorm, real* wa);
void lm_qrsolve(int n, real* r, int ldr, const int* ipvt, const real* diag, const real* qtb, real* x, real* sdiag,
                real* wa);
void lm_par(int m, int n, real* r, int ldr, const int* ipvt, const real* diag, const real* qtb, real delta,
            real* par, real* x, real* sdiag, real* wa1, real* wa2);

int lm_solve(const Problem* p, const FcnCtx* c, const Params* prm, real* x, real* fvec, IterBehavior* ib,
             Workspace* ws);
// constrained_least_squares_solver's own members (m_delta, m_scaling, least_squares:64-66) and the limits of
// constrained_equation_solver (multi_eqn: m_lower / m_upper); null limits = the +-huge arrays the reference
// fills in when none were set (least_squares:1014-1024).
struct ClsOptions {
    double trust_region_radius;
    double step_scaling_factor;
    const double* lower;
    const double* upper;
};
void cls_options_default(ClsOptions* o);
real alpha_box(int n, const real* x, const real* p, const real* xl, const real* xu);
void coleman_li_scaling(int n, const real* x, const real* xl, const real* xu, real* s);
void dogleg(int m, int n, real delta, const real* x, const real* f, const real* jac, real* qr, const real* tau,
            const real* s, const real* xl, const real* xu, real* p, real* g, real* Jp, real* prered, real* wrk);
int cls_solve(const Problem* p, const FcnCtx* c, const Params* prm, const ClsOptions* opt, real* x, real* fvec,
              IterBehavior* ib, Workspace* ws);
int newton_solve(const Problem* p, const FcnCtx* c, const Params* prm, real* x, real* fvec, IterBehavior* ib,
                 Workspace* ws);
int broyden_solve(const Problem* p, const FcnCtx* c, const Params* prm, real* x, real* fvec, IterBehavior* ib,
                  Workspace* ws);

}  // namespace nlo
#endif
