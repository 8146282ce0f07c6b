// oracle/nl_scalar.cpp — TEST INFRASTRUCTURE ONLY.  See nl_scalar.h.
#include "nl_scalar.h"

#include <cfloat>
#include <cmath>
#include <cstring>

namespace nlo {

static const real ZERO = 0.0, HALF = 0.5, ONE = 1.0, TWO = 2.0, THREE = 3.0;

// ---- one-variable test functions --------------------------------------------------------
static real sinx_div_x(real x, const real*) { return real(std::sin(dval(x))) / x; }
static real sinx_div_x_a(real x, const real* a) { return a[0] * real(std::sin(dval(x))) / x; }
static real cubic_wallis(real x, const real*) { return (x * x) * x - TWO * x - real(5.0); }
static real cubic_wallis_d(real x, const real*) { return THREE * (x * x) - TWO; }
static real exp_minus_x(real x, const real*) { return f_exp(-x) - x; }
static real exp_minus_x_d(real x, const real*) { return -f_exp(-x) - ONE; }
static real cubic_args(real x, const real* a) { return ((a[3] * x + a[2]) * x + a[1]) * x + a[0]; }
static real cubic_args_d(real x, const real* a) { return (THREE * a[3] * x + TWO * a[2]) * x + a[1]; }

static const Problem1 kProblems1[NL_FCN1_COUNT] = {
    {NL_FCN1_SINX_DIV_X, "sinx_div_x", 0, sinx_div_x, nullptr},
    {NL_FCN1_SINX_DIV_X_A, "sinx_div_x_a", 1, sinx_div_x_a, nullptr},
    {NL_FCN1_CUBIC_WALLIS, "cubic_wallis", 0, cubic_wallis, cubic_wallis_d},
    {NL_FCN1_EXP_MINUS_X, "exp_minus_x", 0, exp_minus_x, exp_minus_x_d},
    {NL_FCN1_CUBIC_ARGS, "cubic_args", 4, cubic_args, cubic_args_d},
};

const Problem1* nl_problem1(int id) { return (id >= 0 && id < NL_FCN1_COUNT) ? &kProblems1[id] : nullptr; }
const Problem1* nl_problem1_by_name(const char* name) {
    for (int i = 0; i < NL_FCN1_COUNT; ++i)
        if (std::strcmp(kProblems1[i].name, name) == 0) return &kProblems1[i];
    return nullptr;
}

void params1_default(Params1* p) {
    p->max_fcn_evals = 100;
    p->fcn_tol = 1.0e-8;
    p->var_tol = 1.0e-12;
    p->diff_tol = 1.0e-12;
    p->use_analytic_diff = 0;
}

static void ib_zero(IterBehavior* ib) {
    ib->iter_count = 0; ib->fcn_count = 0; ib->jacobian_count = 0; ib->gradient_count = 0;
    ib->converge_on_fcn = 0; ib->converge_on_chng = 0; ib->converge_on_zero_diff = 0;
}

// f1h_diff_fcn, src/nonlin_single_var.f90:154-200 (f always supplied by newt1var_solve)
real fd_diff(const Problem1* p, const real* args, const Params1* prm, real x, real f) {
    if (prm->use_analytic_diff && p->diff) return p->diff(x, args);      // :184-186
    const real epsmch = DBL_EPSILON;
    const real eps = f_sqrt(epsmch);
    real h = eps * f_abs(x);                                              // :189-190
    if (h < epsmch) h = eps;
    real temp = x + h;
    real f1 = p->fcn(temp, args);
    return (f1 - f) / h;                                                  // :198
}

// brent_solve, src/nonlin_solve.f90:643-835.  c, d, e are undefined on entry in the reference (they are set by the
// first pass of the loop unless f(b) is exactly 0 or NaN); they start at 0 here.
int brent_solve(const Problem1* p, const real* args, const Params1* prm, real lim1, real lim2, real* x, real* f,
                IterBehavior* ib) {
    bool fcnvrg = false, xcnvrg = false;
    *x = ZERO;                                                            // :691
    real a = f_min(lim1, lim2), b = f_max(lim1, lim2);
    int neval = 0, iter = 0, flag = 0;
    const real eps = DBL_EPSILON, ftol = prm->fcn_tol, xtol = prm->var_tol;
    const int maxeval = prm->max_fcn_evals;
    *f = ZERO;
    ib_zero(ib);
    if (f_abs(a - b) < eps) return NL_INVALID_INPUT_ERROR;                // :713
    real fa = p->fcn(a, args), fb = p->fcn(b, args);                      // :717-720
    neval = 2;
    real fc = fb, c = ZERO, d = ZERO, e = ZERO;
    for (;;) {
        ++iter;
        if ((fb > ZERO && fc >= ZERO) || (fb < ZERO && fc < ZERO)) {      // :726-732
            c = a; fc = fa; d = b - a; e = d;
        }
        if (f_abs(fc) < f_abs(fb)) {                                      // :733-740
            a = b; b = c; c = a;
            fa = fb; fb = fc; fc = fa;
        }
        real tol1 = TWO * eps * f_abs(b) + HALF * xtol;                   // :743-754
        real xm = HALF * (c - b);
        if (f_abs(fb) < ftol) { *x = b; fcnvrg = true; break; }
        if (f_abs(xm) <= tol1) { *x = b; xcnvrg = true; break; }
        if (f_abs(e) >= tol1 && f_abs(fa) > f_abs(fb)) {                  // :757-794
            real s = fb / fa, pp, q;
            if (f_abs(a - c) < eps) {
                pp = TWO * xm * s;
                q = ONE - s;
            } else {
                q = fa / fc;
                real r = fb / fc;
                pp = s * (TWO * xm * q * (q - r) - (b - a) * (r - ONE));
                q = (q - ONE) * (r - ONE) * (s - ONE);
            }
            if (pp > ZERO) q = -q;
            pp = f_abs(pp);
            real mn1 = THREE * xm * q - f_abs(tol1 * q);
            real mn2 = f_abs(e * q);
            real temp = (mn1 < mn2) ? mn1 : mn2;
            if (TWO * pp < temp) { e = d; d = pp / q; }
            else { d = xm; e = d; }
        } else {
            d = xm; e = d;
        }
        a = b;                                                            // :797-804
        fa = fb;
        if (f_abs(d) > tol1) b = b + d;
        else b = b + f_sign(tol1, xm);
        fb = p->fcn(b, args);
        ++neval;
        if (neval >= maxeval) { flag = 1; break; }                        // :813-816
    }
    *f = fb;
    ib->iter_count = iter; ib->fcn_count = neval;
    ib->converge_on_fcn = fcnvrg; ib->converge_on_chng = xcnvrg;
    return flag ? NL_CONVERGENCE_ERROR : NL_NO_ERROR;
}

// newt1var_solve, src/nonlin_solve.f90:840-1032
int newton1_solve(const Problem1* p, const real* args, const Params1* prm, real lim1, real lim2, real* x, real* f,
                  bool f_present, IterBehavior* ib) {
    bool fcnvrg = false, xcnvrg = false, dcnvrg = false;
    int neval = 0, ndiff = 0, iter = 0, flag = 0;
    const real ftol = prm->fcn_tol, xtol = prm->var_tol, dtol = prm->diff_tol, eps = DBL_EPSILON;
    const int maxeval = prm->max_fcn_evals;
    *f = ZERO;
    ib_zero(ib);
    real x1 = f_min(lim1, lim2), x2 = f_max(lim1, lim2);
    if (f_abs(x1 - x2) < eps) return NL_INVALID_INPUT_ERROR;              // :899
    real fl = p->fcn(x1, args), fh = p->fcn(x2, args);                    // :903-905
    neval = 2;
    if (f_abs(fl) < ftol) {                                               // :906-914: only these two ib fields are set
        *x = x1; *f = fl; ib->converge_on_fcn = 1; ib->fcn_count = 2;
        return NL_NO_ERROR;
    }
    if (f_abs(fh) < ftol) {                                               // :915-923
        *x = x2; *f = fh; ib->converge_on_fcn = 1; ib->fcn_count = 2;
        return NL_NO_ERROR;
    }
    real xl, xh;
    if (fl < ZERO) { xl = x1; xh = x2; }                                  // :926-932
    else { xl = x2; xh = x1; }
    *x = HALF * (x1 + x2);
    real dxold = f_abs(x2 - x1), dx = dxold;
    real ff = p->fcn(*x, args);
    real df = fd_diff(p, args, prm, *x, ff);
    ++neval; ++ndiff;
    for (;;) {
        ++iter;
        if ((((*x - xh) * df - ff) * ((*x - xl) * df - ff) > ZERO) || (f_abs(TWO * ff) > f_abs(dxold * df))) {
            dxold = dx;                                                   // bisection :949-958
            dx = HALF * (xh - xl);
            *x = xl + dx;
            if (f_abs(xl - *x) < xtol) { xcnvrg = true; break; }
        } else {
            dxold = dx;                                                   // Newton :959-970
            dx = ff / df;
            real temp = *x;
            *x = *x - dx;
            if (f_abs(temp - *x) < xtol) { xcnvrg = true; break; }
        }
        ff = p->fcn(*x, args);                                            // :972-975
        df = fd_diff(p, args, prm, *x, ff);
        ++neval; ++ndiff;
        if (f_abs(ff) < ftol) { fcnvrg = true; break; }                   // :978-989
        if (f_abs(dx) < xtol) { xcnvrg = true; break; }
        if (f_abs(df) < dtol) { dcnvrg = true; break; }
        if (ff < ZERO) xl = *x;                                           // :992-996
        else xh = *x;
        if (neval >= maxeval) { flag = 1; break; }                        // :1004-1007
    }
    if (f_present) ++neval;        // :1011-1014 evaluates f(x) once more; :1017 then stores the older ff anyway
    *f = ff;
    ib->iter_count = iter; ib->fcn_count = neval; ib->jacobian_count = ndiff;
    ib->converge_on_fcn = fcnvrg; ib->converge_on_chng = xcnvrg; ib->converge_on_zero_diff = dcnvrg;
    return flag ? NL_CONVERGENCE_ERROR : NL_NO_ERROR;
}

}  // namespace nlo
