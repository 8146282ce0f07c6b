// scalar_solvers.cuh — the reference's one-variable solvers, one CUDA thread per equation.
//
// Reference behaviour reproduced here:
//   brent_solve_dev    brent_solve     src/nonlin_solve.f90:643-835   (Brent's method on a bracketing value_pair)
//   newton1_solve_dev  newt1var_solve  src/nonlin_solve.f90:840-1032  (Newton safeguarded by bisection)
//   fd_diff            f1h_diff_fcn    src/nonlin_single_var.f90:154-200 (forward difference, h = sqrt(eps)|x|)
// Quirks kept: Brent returns x = 0 when the evaluation budget runs out (x is only assigned on convergence, :691);
// Newton's early returns for a root at an end point set only converge_on_fcn and fcn_count = 2 (:906-923); when the
// optional `f` is requested Newton counts one more evaluation and still reports the older residual (:1011-1017).
// Brent's c, d, e are undefined on entry in the reference; they start at 0 here.
//
// One-variable functions are registered functors like the vector residuals (the `fcn1var` / `fcn1var_helper`
// analogue): `eval(x, a)` and, when HAS_DIFF, `diff(x, a)`; a.v(k) = args[k*B + b].
#pragma once
#include "nlb_math.cuh"
#include "nlb_types.h"

namespace nlb {

struct Args1 {
    const double* p;      // args + b
    long long B;
    NLB_DEV double v(int k) const { return p[(long long)k * B]; }
};

enum Fcn1Id { FCN1_SINX_DIV_X = 0, FCN1_SINX_DIV_X_A, FCN1_CUBIC_WALLIS, FCN1_EXP_MINUS_X, FCN1_CUBIC_ARGS, FCN1_COUNT };

struct SinxDivX {            // tests/nonlin_test_solve.f90:165-170; libm sin: outside the bitwise-parity set
    static constexpr int ID = FCN1_SINX_DIV_X, ARGS_LEN = 0;
    static constexpr bool HAS_DIFF = false;
    NLB_DEV static double eval(double x, const Args1&) { return sin(x) / x; }
    NLB_DEV static double diff(double, const Args1&) { return 0.0; }
};
struct SinxDivXA {           // tests/nonlin_test_solve.f90:172-183, args = a
    static constexpr int ID = FCN1_SINX_DIV_X_A, ARGS_LEN = 1;
    static constexpr bool HAS_DIFF = false;
    NLB_DEV static double eval(double x, const Args1& a) { return a.v(0) * sin(x) / x; }
    NLB_DEV static double diff(double, const Args1&) { return 0.0; }
};
struct CubicWallis {         // x**3 - 2x - 5
    static constexpr int ID = FCN1_CUBIC_WALLIS, ARGS_LEN = 0;
    static constexpr bool HAS_DIFF = true;
    NLB_DEV static double eval(double x, const Args1&) { return (x * x) * x - 2.0 * x - 5.0; }
    NLB_DEV static double diff(double x, const Args1&) { return 3.0 * (x * x) - 2.0; }
};
struct ExpMinusX {           // exp(-x) - x, shared software exp
    static constexpr int ID = FCN1_EXP_MINUS_X, ARGS_LEN = 0;
    static constexpr bool HAS_DIFF = true;
    NLB_DEV static double eval(double x, const Args1&) { return nl_exp(-x) - x; }
    NLB_DEV static double diff(double x, const Args1&) { return -nl_exp(-x) - 1.0; }
};
struct CubicArgs {           // a0 + a1 x + a2 x**2 + a3 x**3 (Horner), args = a0..a3
    static constexpr int ID = FCN1_CUBIC_ARGS, ARGS_LEN = 4;
    static constexpr bool HAS_DIFF = true;
    NLB_DEV static double eval(double x, const Args1& a) { return ((a.v(3) * x + a.v(2)) * x + a.v(1)) * x + a.v(0); }
    NLB_DEV static double diff(double x, const Args1& a) { return (3.0 * a.v(3) * x + 2.0 * a.v(2)) * x + a.v(1); }
};

struct Fcn1Info {
    const char* name;
    int args_len;
    int has_diff;
};
inline const Fcn1Info* fcn1_table() {
    static const Fcn1Info t[FCN1_COUNT] = {
        {"sinx_div_x", 0, 0}, {"sinx_div_x_a", 1, 0}, {"cubic_wallis", 0, 1}, {"exp_minus_x", 0, 1}, {"cubic_args", 4, 1},
    };
    return t;
}

// equation_solver_1var settings in the form the kernels read them
struct DevParams1 {
    int max_fcn_evals;
    double fcn_tol, var_tol, diff_tol;
    int use_analytic_diff;
};

struct Solve1Result {
    double x, f;
    int iter = 0, nfev = 0, ndiff = 0, cf = 0, cx = 0, cd = 0, status = 0;
};

template <class F>
NLB_DEV double fd_diff(const DevParams1& p, const Args1& a, double x, double f) {
    if (F::HAS_DIFF && p.use_analytic_diff) return F::diff(x, a);
    const double epsmch = 0x1p-52, eps = 0x1p-26;
    double h = eps * fabs(x);
    if (h < epsmch) h = eps;
    const double temp = x + h;
    const double f1 = F::eval(temp, a);
    return (f1 - f) / h;
}

template <class F>
NLB_DEV void brent_solve_dev(const DevParams1& p, const Args1& args, double lim1, double lim2, Solve1Result& r) {
    const double eps = 0x1p-52, ftol = p.fcn_tol, xtol = p.var_tol;
    r.x = 0.0;
    r.f = 0.0;
    double a = nl_min(lim1, lim2), b = nl_max(lim1, lim2);
    if (fabs(a - b) < eps) { r.status = NLB_INVALID_INPUT_ERROR; return; }
    double fa = F::eval(a, args), fb = F::eval(b, args);
    int neval = 2, iter = 0;
    double fc = fb, c = 0.0, d = 0.0, e = 0.0;
    bool flag = false;
    for (;;) {
        ++iter;
        if ((fb > 0.0 && fc >= 0.0) || (fb < 0.0 && fc < 0.0)) { c = a; fc = fa; d = b - a; e = d; }
        if (fabs(fc) < fabs(fb)) {
            a = b; b = c; c = a;
            fa = fb; fb = fc; fc = fa;
        }
        const double tol1 = 2.0 * eps * fabs(b) + 0.5 * xtol;
        const double xm = 0.5 * (c - b);
        if (fabs(fb) < ftol) { r.x = b; r.cf = 1; break; }
        if (fabs(xm) <= tol1) { r.x = b; r.cx = 1; break; }
        if (fabs(e) >= tol1 && fabs(fa) > fabs(fb)) {
            const double s = fb / fa;
            double pp, q;
            if (fabs(a - c) < eps) {
                pp = 2.0 * xm * s;
                q = 1.0 - s;
            } else {
                q = fa / fc;
                const double rr = fb / fc;
                pp = s * (2.0 * xm * q * (q - rr) - (b - a) * (rr - 1.0));
                q = (q - 1.0) * (rr - 1.0) * (s - 1.0);
            }
            if (pp > 0.0) q = -q;
            pp = fabs(pp);
            const double mn1 = 3.0 * xm * q - fabs(tol1 * q);
            const double mn2 = fabs(e * q);
            const double temp = (mn1 < mn2) ? mn1 : mn2;
            if (2.0 * pp < temp) { e = d; d = pp / q; }
            else { d = xm; e = d; }
        } else {
            d = xm; e = d;
        }
        a = b;
        fa = fb;
        if (fabs(d) > tol1) b = b + d;
        else b = b + nl_sign(tol1, xm);
        fb = F::eval(b, args);
        ++neval;
        if (neval >= p.max_fcn_evals) { flag = true; break; }
    }
    r.f = fb;
    r.iter = iter;
    r.nfev = neval;
    r.status = flag ? NLB_CONVERGENCE_ERROR : 0;
}

template <class F>
NLB_DEV void newton1_solve_dev(const DevParams1& p, const Args1& args, double lim1, double lim2, bool f_present,
                               Solve1Result& r) {
    const double eps = 0x1p-52, ftol = p.fcn_tol, xtol = p.var_tol, dtol = p.diff_tol;
    r.f = 0.0;
    const double x1 = nl_min(lim1, lim2), x2 = nl_max(lim1, lim2);
    if (fabs(x1 - x2) < eps) { r.status = NLB_INVALID_INPUT_ERROR; return; }
    const double fl = F::eval(x1, args), fh = F::eval(x2, args);
    int neval = 2, ndiff = 0, iter = 0;
    if (fabs(fl) < ftol) { r.x = x1; r.f = fl; r.cf = 1; r.nfev = 2; return; }
    if (fabs(fh) < ftol) { r.x = x2; r.f = fh; r.cf = 1; r.nfev = 2; return; }
    double xl, xh;
    if (fl < 0.0) { xl = x1; xh = x2; }
    else { xl = x2; xh = x1; }
    double x = 0.5 * (x1 + x2);
    double dxold = fabs(x2 - x1), dx = dxold;
    double ff = F::eval(x, args);
    double df = fd_diff<F>(p, args, x, ff);
    ++neval; ++ndiff;
    bool flag = false;
    for (;;) {
        ++iter;
        if ((((x - xh) * df - ff) * ((x - xl) * df - ff) > 0.0) || (fabs(2.0 * ff) > fabs(dxold * df))) {
            dxold = dx;
            dx = 0.5 * (xh - xl);
            x = xl + dx;
            if (fabs(xl - x) < xtol) { r.cx = 1; break; }
        } else {
            dxold = dx;
            dx = ff / df;
            const double temp = x;
            x = x - dx;
            if (fabs(temp - x) < xtol) { r.cx = 1; break; }
        }
        ff = F::eval(x, args);
        df = fd_diff<F>(p, args, x, ff);
        ++neval; ++ndiff;
        if (fabs(ff) < ftol) { r.cf = 1; break; }
        if (fabs(dx) < xtol) { r.cx = 1; break; }
        if (fabs(df) < dtol) { r.cd = 1; break; }
        if (ff < 0.0) xl = x;
        else xh = x;
        if (neval >= p.max_fcn_evals) { flag = true; break; }
    }
    if (f_present) ++neval;
    r.x = x;
    r.f = ff;
    r.iter = iter;
    r.nfev = neval;
    r.ndiff = ndiff;
    r.status = flag ? NLB_CONVERGENCE_ERROR : 0;
}

// SOLVER: 0 = brent_solver, 1 = newton_1var_solver
template <class F, int SOLVER>
__global__ void __launch_bounds__(128)
solve_1var_kernel(DevParams1 p, long long B, const double* __restrict__ lim1, const double* __restrict__ lim2,
                  double* __restrict__ x, double* __restrict__ f, const double* __restrict__ args,
                  nlb_iteration_behavior* __restrict__ ib, int32_t* __restrict__ status) {
    const long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (b >= B) return;
    Args1 a{args ? args + b : nullptr, B};
    Solve1Result r;
    r.x = x[b];                          // Newton leaves x untouched when the limits are rejected
    if (SOLVER == 0) brent_solve_dev<F>(p, a, lim1[b], lim2[b], r);
    else newton1_solve_dev<F>(p, a, lim1[b], lim2[b], f != nullptr, r);
    x[b] = r.x;
    if (f) f[b] = r.f;
    if (ib) {
        nlb_iteration_behavior o;
        o.iter_count = r.iter;
        o.fcn_count = r.nfev;
        o.jacobian_count = r.ndiff;
        o.gradient_count = 0;
        o.converge_on_fcn = r.cf;
        o.converge_on_chng = r.cx;
        o.converge_on_zero_diff = r.cd;
        ib[b] = o;
    }
    if (status) status[b] = r.status;
}

}  // namespace nlb
