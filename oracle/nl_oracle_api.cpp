// oracle/nl_oracle_api.cpp — TEST INFRASTRUCTURE ONLY.
//
// extern "C" surface of the CPU oracle, loaded with ctypes by tests/, by
// __graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs.  The
// product library (nonlin_b200/libnonlin_b200.so) never links or loads this file.
//
// Batch layout is the same SoA the engine's C ABI uses: x[j*B + b], fvec[i*B + b],
// sys[k*B + b] (system index fastest — a Fortran array declared x(B, n)).
#include <atomic>
#include <cstdint>
#include <cstring>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

#include "nl_lapack.h"
#include "nl_polynomial.h"
#include "nl_scalar.h"
#include "nl_solvers.h"

namespace nlo {
static int g_libm_exp = 0;
void nl_set_libm_exp(int on) { g_libm_exp = on; }
int nl_get_libm_exp() { return g_libm_exp; }
static int g_sum_mode = 0;
void nl_set_sum_mode(int mode) { g_sum_mode = mode; }
int nl_get_sum_mode() { return g_sum_mode; }
#ifdef NL_COUNT_FLOPS
thread_local unsigned long long g_flops = 0;
#endif
}  // namespace nlo

using namespace nlo;

static std::atomic<unsigned long long> g_flops_total{0};

static inline real* R(double* p) { return reinterpret_cast<real*>(p); }
static inline const real* R(const double* p) { return reinterpret_cast<const real*>(p); }

static int resolve_sizes(const Problem* p, int* m, int* n, int* sys_len) {
    if (!p) return NL_UNDEFINED_FUNCTION_ERROR;
    if (p->m != 0) { if (*m != 0 && *m != p->m) return NL_ARRAY_SIZE_ERROR; *m = p->m; }
    if (p->n != 0) { if (*n != 0 && *n != p->n) return NL_ARRAY_SIZE_ERROR; *n = p->n; }
    if (p->id == NL_FCN_EXT_ROSENBROCK) { if (*m == 0) *m = *n; if (*m != *n || (*n & 1)) return NL_INVALID_INPUT_ERROR; }
    if (*m <= 0 || *n <= 0) return NL_INVALID_INPUT_ERROR;
    *sys_len = p->sys_len < 0 ? *m : p->sys_len;
    return 0;
}

static int solve_one(int solver, const Problem* p, const FcnCtx* c, const Params* prm, real* x, real* fvec,
                     IterBehavior* ib, Workspace* ws) {
    switch (solver) {
        case 0: return lm_solve(p, c, prm, x, fvec, ib, ws);
        case 1: return newton_solve(p, c, prm, x, fvec, ib, ws);
        case 2: return broyden_solve(p, c, prm, x, fvec, ib, ws);
    }
    return NL_INVALID_INPUT_ERROR;
}

extern "C" {

int nlo_is_counting_build(void) {
#ifdef NL_COUNT_FLOPS
    return 1;
#else
    return 0;
#endif
}

unsigned long long nlo_flops_total(void) { return g_flops_total.load(); }
void nlo_flops_reset(void) { g_flops_total.store(0); }

void nlo_set_libm_exp(int on) { nl_set_libm_exp(on); }
void nlo_set_sum_mode(int mode) { nl_set_sum_mode(mode); }

int nlo_fcn_lookup(const char* name) {
    const Problem* p = nl_problem_by_name(name);
    return p ? p->id : -1;
}

int nlo_fcn_info(int id, int* m, int* n, int* sys_len, int* shared_len, int* has_jac) {
    const Problem* p = nl_problem(id);
    if (!p) return -1;
    *m = p->m; *n = p->n; *sys_len = p->sys_len; *shared_len = p->shared_len; *has_jac = p->jac != nullptr;
    return 0;
}

void nlo_params_default(Params* p) { params_default(p); }

// a residual supplied by the test as a C callback (checker for plug-in residuals); returns its id or -1
int nlo_register_callback(const char* name, int m, int n, int sys_len, int shared_len, callback_t fcn) {
    return nl_register_callback(name, m, n, sys_len, shared_len, fcn);
}

double nlo_soft_exp(double x) { return soft_exp_d(x); }
double nlo_norm2(const double* v, int n) { return dval(f_norm2(R(v), n)); }
double nlo_dnrm2(const double* v, int n) { return dval(la_dnrm2(n, R(v), 1)); }

int nlo_eval_fcn(int fcn_id, int m, int n, const double* x, const double* sys, const double* shared, double* f) {
    const Problem* p = nl_problem(fcn_id);
    int sl;
    int rc = resolve_sizes(p, &m, &n, &sl);
    if (rc) return rc;
    FcnCtx c = {m, n, R(sys), R(shared)};
    p->fcn(R(x), R(f), &c);
    return 0;
}

// vecfcn_helper%jacobian without fv (reference computes f(x) itself, multi_eqn:257-259)
int nlo_jacobian(int fcn_id, int m, int n, const Params* prm, double* x, const double* sys, const double* shared,
                 double* jac) {
    const Problem* p = nl_problem(fcn_id);
    int sl;
    int rc = resolve_sizes(p, &m, &n, &sl);
    if (rc) return rc;
    FcnCtx c = {m, n, R(sys), R(shared)};
    std::vector<real> f(m), w(m);
    p->fcn(R(x), f.data(), &c);
    fd_jacobian(p, &c, prm, R(x), R(jac), f.data(), w.data());
    return 0;
}

// pieces exposed for unit tests against scipy's LAPACK
void nlo_dgeqr2_dorg2r(int n, const double* a, double* q, double* r) {
    std::vector<real> tau(n), work(n);
    for (long e = 0; e < (long)n * n; ++e) R(q)[e] = R(a)[e];
    la_dgeqr2(n, n, R(q), n, tau.data(), work.data());
    for (int j = 0; j < n; ++j)
        for (int i = 0; i < n; ++i) R(r)[i + (long)j * n] = (i <= j) ? R(q)[i + (long)j * n] : real(0.0);
    la_dorg2r(n, n, n, R(q), n, tau.data(), work.data());
}
void nlo_dqr1up(int n, double* q, double* r, const double* u, const double* v) {
    std::vector<real> w(2 * n);
    la_dqr1up(n, n, R(q), n, R(r), n, R(u), R(v), w.data());
}
int nlo_dgesv(int n, double* a, int* ipiv, double* b) {
    int info = la_dgetrf(n, n, R(a), n, ipiv);
    la_dgetrs(n, R(a), n, ipiv, R(b));
    return info;
}
void nlo_lmfactor(int m, int n, double* a, int* ipvt, double* rdiag, double* acnorm) {
    std::vector<real> wa(n);
    lm_factor(m, n, R(a), true, ipvt, R(rdiag), R(acnorm), wa.data());
}

// DGELS restatement on a caller's m x n matrix (column-major) and right-hand side; returns LAPACK info
int nlo_dgels(int m, int n, double* a, double* b) {
    std::vector<real> tau(n), work(n);
    return la_dgels(m, n, R(a), m, R(b), tau.data(), work.data());
}
// qr_factor(a, tau=, qr=) + solve_qr(qr, tau, b) as the constrained solver uses them: x(1:n) in b
void nlo_qr_solve(int m, int n, double* a, double* b) {
    std::vector<real> tau(n), work(n);
    la_dgeqr2(m, n, R(a), m, tau.data(), work.data());
    la_dorm2r_lt_vec(m, n, R(a), m, tau.data(), R(b));
    la_dtrsv_unn(n, R(a), m, R(b));
}
void nlo_dgemv(int trans, int m, int n, const double* a, const double* x, double* y) {
    if (trans) la_dgemv_t(m, n, real(1.0), R(a), m, R(x), R(y));
    else la_dgemv_n(m, n, real(1.0), R(a), m, R(x), R(y));
}

// one system, contiguous x(n), fvec(m), sys(sys_len)
int nlo_solve(int solver, int fcn_id, int m, int n, const Params* prm, double* x, double* fvec, const double* sys,
              const double* shared, IterBehavior* ib) {
    const Problem* p = nl_problem(fcn_id);
    int sl;
    int rc = resolve_sizes(p, &m, &n, &sl);
    if (rc) return rc;
    FcnCtx c = {m, n, R(sys), R(shared)};
    Workspace ws;
#ifdef NL_COUNT_FLOPS
    g_flops = 0;
#endif
    int st = solve_one(solver, p, &c, prm, R(x), R(fvec), ib, &ws);
#ifdef NL_COUNT_FLOPS
    g_flops_total += g_flops;
#endif
    return st;
}

// constrained_least_squares_solver, one system.  lower / upper: n entries each or null.
void nlo_cls_options_default(ClsOptions* o) { cls_options_default(o); }

int nlo_cls_solve(int fcn_id, int m, int n, const Params* prm, const ClsOptions* opt, double* x, double* fvec,
                  const double* sys, const double* shared, IterBehavior* ib) {
    const Problem* p = nl_problem(fcn_id);
    int sl;
    int rc = resolve_sizes(p, &m, &n, &sl);
    if (rc) return rc;
    FcnCtx c = {m, n, R(sys), R(shared)};
    Workspace ws;
#ifdef NL_COUNT_FLOPS
    g_flops = 0;
#endif
    int st = cls_solve(p, &c, prm, opt, R(x), R(fvec), ib, &ws);
#ifdef NL_COUNT_FLOPS
    g_flops_total += g_flops;
#endif
    return st;
}

int nlo_cls_solve_batch(int fcn_id, long B, int m, int n, const Params* prm, const ClsOptions* opt, double* x,
                        double* fvec, const double* sys, const double* shared, IterBehavior* ib, int32_t* status,
                        int nthreads) {
    const Problem* p = nl_problem(fcn_id);
    int sl;
    int rc = resolve_sizes(p, &m, &n, &sl);
    if (rc) return rc;
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
    nthreads = 1;
#endif
#pragma omp parallel num_threads(nthreads)
    {
        Workspace ws;
        std::vector<real> xl(n), fl(m), sl_buf(sl > 0 ? sl : 1);
#ifdef NL_COUNT_FLOPS
        g_flops = 0;
#endif
#pragma omp for schedule(dynamic, 64)
        for (long b = 0; b < B; ++b) {
            for (int j = 0; j < n; ++j) xl[j] = R(x)[(long)j * B + b];
            for (int k = 0; k < sl; ++k) sl_buf[k] = R(sys)[(long)k * B + b];
            FcnCtx c = {m, n, sl > 0 ? sl_buf.data() : nullptr, R(shared)};
            IterBehavior lib;
            int st = cls_solve(p, &c, prm, opt, xl.data(), fl.data(), &lib, &ws);
            for (int j = 0; j < n; ++j) R(x)[(long)j * B + b] = xl[j];
            for (int i = 0; i < m; ++i) R(fvec)[(long)i * B + b] = fl[i];
            if (ib) ib[b] = lib;
            if (status) status[b] = st;
        }
#ifdef NL_COUNT_FLOPS
        g_flops_total += g_flops;
#endif
    }
    return 0;
}

// polynomial%fit / fit_thru_zero over B data sets.  x: npts (x_is_shared) or npts x B; y: npts x B (not modified
// here: each fit works on a copy, the reference overwrites its y); coeffs: (order + 1) x B, c0 first.
int nlo_polyfit_batch(long B, int npts, int order, int thru_zero, int x_is_shared, const double* x, const double* y,
                      double* coeffs, int32_t* status, int nthreads) {
    if (order < 1 || order >= npts || order + 1 > 64) return NL_INVALID_INPUT_ERROR;   // error stop 4 (:160-163)
    const int ncols = thru_zero ? order : order + 1;
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
    nthreads = 1;
#endif
#pragma omp parallel num_threads(nthreads)
    {
        std::vector<real> xl(npts), yl(npts), a((size_t)npts * ncols), c(order + 1);
#ifdef NL_COUNT_FLOPS
        g_flops = 0;
#endif
#pragma omp for schedule(dynamic, 64)
        for (long b = 0; b < B; ++b) {
            for (int i = 0; i < npts; ++i) {
                xl[i] = x_is_shared ? R(x)[i] : R(x)[(long)i * B + b];
                yl[i] = R(y)[(long)i * B + b];
            }
            int st = thru_zero ? poly_fit_thru_zero(npts, order, xl.data(), yl.data(), c.data(), a.data())
                               : poly_fit(npts, order, xl.data(), yl.data(), c.data(), a.data());
            for (int j = 0; j <= order; ++j) R(coeffs)[(long)j * B + b] = c[j];
            if (status) status[b] = st;
        }
#ifdef NL_COUNT_FLOPS
        g_flops_total += g_flops;
#endif
    }
    return 0;
}

// polynomial%evaluate: y[i*B + b] = p_b(x_i)
int nlo_polyval_batch(long B, int order, int npts, int x_is_shared, const double* coeffs, const double* x, double* y) {
    std::vector<real> c(order + 1);
    for (long b = 0; b < B; ++b) {
        for (int j = 0; j <= order; ++j) c[j] = R(coeffs)[(long)j * B + b];
        for (int i = 0; i < npts; ++i) {
            real xv = x_is_shared ? R(x)[i] : R(x)[(long)i * B + b];
            R(y)[(long)i * B + b] = poly_eval(order, c.data(), xv);
        }
    }
    return 0;
}

// brent_solver (solver = 0) / newton_1var_solver (solver = 1) over B equations.  lim1 / lim2: the value_pair of
// each equation; x in/out; f out, or null (= the optional argument absent); args[k*B + b].
int nlo_fcn1_lookup(const char* name) {
    const Problem1* p = nl_problem1_by_name(name);
    return p ? p->id : -1;
}
int nlo_fcn1_info(int id, int* args_len, int* has_diff) {
    const Problem1* p = nl_problem1(id);
    if (!p) return -1;
    *args_len = p->args_len; *has_diff = p->diff != nullptr;
    return 0;
}
void nlo_params1_default(Params1* p) { params1_default(p); }
double nlo_fcn1_eval(int id, double x, const double* args) { return dval(nl_problem1(id)->fcn(real(x), R(args))); }

int nlo_solve_1var_batch(int solver, int fcn_id, long B, const Params1* prm, const double* lim1, const double* lim2,
                         double* x, double* f, const double* args, IterBehavior* ib, int32_t* status, int nthreads) {
    const Problem1* p = nl_problem1(fcn_id);
    if (!p) return NL_UNDEFINED_FUNCTION_ERROR;
    if (solver < 0 || solver > 1) return NL_INVALID_INPUT_ERROR;
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
    nthreads = 1;
#endif
#pragma omp parallel num_threads(nthreads)
    {
        std::vector<real> al(p->args_len > 0 ? p->args_len : 1);
#ifdef NL_COUNT_FLOPS
        g_flops = 0;
#endif
#pragma omp for schedule(dynamic, 256)
        for (long b = 0; b < B; ++b) {
            for (int k = 0; k < p->args_len; ++k) al[k] = R(args)[(long)k * B + b];
            real xv = R(x)[b], fv = 0.0;
            IterBehavior lib;
            int st = solver == 0 ? brent_solve(p, al.data(), prm, R(lim1)[b], R(lim2)[b], &xv, &fv, &lib)
                                 : newton1_solve(p, al.data(), prm, R(lim1)[b], R(lim2)[b], &xv, &fv, f != nullptr, &lib);
            R(x)[b] = xv;
            if (f) R(f)[b] = fv;
            if (ib) ib[b] = lib;
            if (status) status[b] = st;
        }
#ifdef NL_COUNT_FLOPS
        g_flops_total += g_flops;
#endif
    }
    return 0;
}

// B systems, SoA; one system per OpenMP thread at a time (schedule(dynamic)), workspaces
// allocated once per thread.  Returns 0 or an API-level error; per-system codes in status[].
int nlo_solve_batch(int solver, int fcn_id, long B, int m, int n, const Params* prm, double* x, double* fvec,
                    const double* sys, const double* shared, IterBehavior* ib, int32_t* status, int nthreads) {
    const Problem* p = nl_problem(fcn_id);
    int sl;
    int rc = resolve_sizes(p, &m, &n, &sl);
    if (rc) return rc;
    if (solver < 0 || solver > 2) return NL_INVALID_INPUT_ERROR;
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
    nthreads = 1;
#endif
#pragma omp parallel num_threads(nthreads)
    {
        Workspace ws;
        std::vector<real> xl(n), fl(m), sl_buf(sl > 0 ? sl : 1);
#ifdef NL_COUNT_FLOPS
        g_flops = 0;
#endif
#pragma omp for schedule(dynamic, 64)
        for (long b = 0; b < B; ++b) {
            for (int j = 0; j < n; ++j) xl[j] = R(x)[(long)j * B + b];
            for (int k = 0; k < sl; ++k) sl_buf[k] = R(sys)[(long)k * B + b];
            FcnCtx c = {m, n, sl > 0 ? sl_buf.data() : nullptr, R(shared)};
            IterBehavior lib;
            int st = solve_one(solver, p, &c, prm, xl.data(), fl.data(), &lib, &ws);
            for (int j = 0; j < n; ++j) R(x)[(long)j * B + b] = xl[j];
            for (int i = 0; i < m; ++i) R(fvec)[(long)i * B + b] = fl[i];
            if (ib) ib[b] = lib;
            if (status) status[b] = st;
        }
#ifdef NL_COUNT_FLOPS
        g_flops_total += g_flops;
#endif
    }
    return 0;
}

}  // extern "C"
