// oracle/nl_solvers.cpp — TEST INFRASTRUCTURE ONLY.  See nl_solvers.h for the map from each
// function to the reference lines it follows.  Indices in the code are 1-based through the
// macros below so that loops read like the Fortran they restate.
#include "nl_solvers.h"

#include <cfloat>

#include "nl_lapack.h"

namespace nlo {

#define V(a, i) (a)[(i) - 1]
#define M2(a, ld, i, j) (a)[((long)(i) - 1) + ((long)(j) - 1) * (long)(ld)]

static const real ZERO = 0.0, ONE = 1.0, HALF = 0.5;
static const double EPSMCH = DBL_EPSILON;            // epsilon(1d0) = 2^-52
static const double DWARF = DBL_MIN;                 // tiny(1d0)    = 2^-1022

void params_default(Params* p) {
    p->max_fcn_evals = 100;
    p->fcn_tol = 1.0e-8;
    p->var_tol = 1.0e-12;
    p->grad_tol = 1.0e-12;
    p->lm_factor = 100.0;
    p->jacobian_interval = 5;
    p->use_line_search = 1;
    p->ls_max_fcn_evals = 100;
    p->ls_alpha = 1.0e-4;
    p->ls_factor = 0.1;
    p->use_analytic_jacobian = 0;
    p->max_iter_guard = 100000;
}

// ---------------------------------------------------------------------------------------
// vfh_jac_fcn, src/nonlin_multi_eqn_mult_var.f90:198-277.  fv is always supplied by the
// three solvers (:221 LM, solve.f90:282 Broyden, :561 Newton).
// ---------------------------------------------------------------------------------------
void fd_jacobian(const Problem* p, const FcnCtx* c, const Params* prm, real* x, real* jac, const real* fv,
                 real* wrk) {
    const int m = c->m, n = c->n;
    if (prm->use_analytic_jacobian && p->jac) {       // :241-243
        p->jac(x, jac, c);
        return;
    }
    const real eps = f_sqrt(real(EPSMCH));            // :263-264
    for (int j = 1; j <= n; ++j) {                    // :267-275
        real temp = V(x, j);
        real h = eps * f_abs(temp);
        if (h == ZERO) h = eps;
        V(x, j) = temp + h;
        p->fcn(x, wrk, c);
        V(x, j) = temp;
        for (int i = 1; i <= m; ++i) M2(jac, m, i, j) = (V(wrk, i) - V(fv, i)) / h;
    }
}

// ---------------------------------------------------------------------------------------
// test_convergence, src/nonlin_helper.f90:36-124
// ---------------------------------------------------------------------------------------
void test_convergence(int nvar, int neqn, const real* x, const real* xo, const real* f, const real* g, bool lg,
                      real xtol, real ftol, real gtol, bool* c, bool* cx, bool* cf, bool* cg, real* xnorm,
                      real* fnorm) {
    *cx = false; *cf = false; *cg = false; *c = false;
    real fc = HALF * f_dot(f, f, neqn);               // :87
    *fnorm = ZERO;
    *xnorm = ZERO;
    for (int i = 1; i <= neqn; ++i) *fnorm = f_max(f_abs(V(f, i)), *fnorm);   // :92-94
    if (*fnorm < ftol) { *cf = true; *c = true; return; }
    for (int i = 1; i <= nvar; ++i) {                 // :102-105
        real test = f_abs(V(x, i) - V(xo, i)) / f_max(f_abs(V(x, i)), ONE);
        *xnorm = f_max(test, *xnorm);
    }
    if (*xnorm < xtol) { *cx = true; *c = true; return; }
    if (lg) {                                         // :113-123
        real test = ZERO;
        real den = f_max(fc, HALF * real(nvar));
        for (int i = 1; i <= nvar; ++i) {
            real dxmax = f_abs(V(g, i)) * f_max(f_abs(V(x, i)), ONE) / den;
            test = f_max(test, dxmax);
        }
        if (test < gtol) *cg = true;
    }
}

// limit_search_vector, src/nonlin_linesearch.f90:554-572
void limit_vector(int n, real* x, real lim) {
    real mag = f_norm2(x, n);
    if (mag == ZERO) return;
    if (mag > lim) {
        real s = lim / mag;
        for (int i = 0; i < n; ++i) x[i] = s * x[i];
    }
}

// min_backtrack_search, src/nonlin_linesearch.f90:495-551
real backtrack_min(int mode, real f0, real f, real f1, real alam, real alam1, real slope) {
    const real p5 = 0.5, two = 2.0, three = 3.0;
    real lam;
    if (mode == 1) {
        lam = -slope / (two * (f - f0 - slope));                                   // :529
    } else {
        real rhs1 = f - f0 - alam * slope;                                         // :532
        real rhs2 = f1 - f0 - alam1 * slope;                                       // :533
        real a = (rhs1 / (alam * alam) - rhs2 / (alam1 * alam1)) / (alam - alam1); // :534
        real b = (-(alam1 * rhs1 / (alam * alam)) + alam * rhs2 / (alam1 * alam1)) / (alam - alam1);  // :535
        if (a == ZERO) {
            lam = -slope / (two * b);
        } else {
            real disc = b * b - three * a * slope;                                 // :540
            if (disc < ZERO) lam = p5 * alam;
            else if (b <= ZERO) lam = (-b + f_sqrt(disc)) / (three * a);
            else lam = -slope / (b + f_sqrt(disc));
        }
        if (lam > p5 * alam) lam = p5 * alam;                                      // :549
    }
    return lam;
}

// ---------------------------------------------------------------------------------------
// ls_search_mimo, src/nonlin_linesearch.f90:152-326 (fold and fx always present at the two
// call sites, solve.f90:347 and :587).  Returns 0 or the code of the `error stop` reached.
// ---------------------------------------------------------------------------------------
int line_search(const Problem* p, const FcnCtx* c, const Params* prm, const real* xold, const real* grad,
                const real* dir, real* x, real* fvec, real fold, real* fx, IterBehavior* ib) {
    const int m = c->m, n = c->n;
    bool xcnvrg = false, fcnvrg = false;
    int neval = 0, niter = 0;
    const real tolx = real(2.0) * real(EPSMCH);       // :209
    const real alpha = prm->ls_alpha, lambdamin = prm->ls_factor;
    const int maxeval = prm->ls_max_fcn_evals;
    *fx = ZERO;
    ib->iter_count = 0; ib->fcn_count = 0;
    ib->converge_on_fcn = 0; ib->converge_on_chng = 0; ib->converge_on_zero_diff = 0;

    real fo = fold;                                   // :239-240
    real slope = f_dot(grad, dir, n);                 // :249
    if (slope >= ZERO) return NL_DIVERGENT_BEHAVIOR_ERROR;   // :250-253

    real test = ZERO;                                 // :256-262
    for (int i = 1; i <= n; ++i) {
        real temp = f_abs(V(dir, i)) / f_max(f_abs(V(xold, i)), ONE);
        if (temp > test) test = temp;
    }
    real alamin = tolx / test;
    real alam = ONE;
    real alam1 = ZERO, f1 = ZERO, f = ZERO, tmplam;

    int flag = 0;
    for (;;) {                                        // :266-310
        for (int i = 0; i < n; ++i) x[i] = xold[i] + alam * dir[i];
        p->fcn(x, fvec, c);
        f = real(0.5) * f_dot(fvec, fvec, m);
        ++neval;
        ++niter;
        if (alam < alamin) {                          // :275-287
            bool same = true;                         // norm2(x - xold) == 0  <=>  x == xold elementwise
            for (int i = 0; i < n; ++i) if ((x[i] - xold[i]) != ZERO) same = false;
            if (same) {
                ib->iter_count = niter; ib->fcn_count = neval; *fx = f;
                return NL_CONVERGENCE_ERROR;          // :281-284
            }
            for (int i = 0; i < n; ++i) x[i] = xold[i];
            xcnvrg = true;
            break;
        } else if (f <= fo + alpha * alam * slope) {  // :288-291
            fcnvrg = true;
            break;
        } else {
            tmplam = backtrack_min(niter, fo, f, f1, alam, alam1, slope);   // :294
        }
        alam1 = alam;                                 // :300-302
        f1 = f;
        alam = f_max(tmplam, lambdamin * alam);
        if (neval >= maxeval) { flag = 1; break; }    // :305-309
    }
    *fx = f;                                          // :311
    ib->iter_count = niter; ib->fcn_count = neval;    // :314-320
    ib->converge_on_fcn = fcnvrg; ib->converge_on_chng = xcnvrg; ib->converge_on_zero_diff = 0;
    if (flag != 0) return NL_CONVERGENCE_ERROR;       // :323-325
    return NL_NO_ERROR;
}

// ---------------------------------------------------------------------------------------
// lmfactor, src/nonlin_least_squares.f90:569-667 (MINPACK QRFAC lineage)
// ---------------------------------------------------------------------------------------
void lm_factor(int m, int n, real* a, bool pivot, int* ipvt, real* rdiag, real* acnorm, real* wa) {
    const real p05 = 5.0e-2;
    const real epsmch = EPSMCH;
    const int minmn = m < n ? m : n;
    for (int j = 1; j <= n; ++j) {                    // :611-616
        V(acnorm, j) = f_norm2(&M2(a, m, 1, j), m);
        V(rdiag, j) = V(acnorm, j);
        V(wa, j) = V(rdiag, j);
        if (pivot) V(ipvt, j) = j;
    }
    for (int j = 1; j <= minmn; ++j) {                // :619-666
        if (pivot) {
            int kmax = j;                             // :622-625
            for (int k = j; k <= n; ++k) if (V(rdiag, k) > V(rdiag, kmax)) kmax = k;
            if (kmax != j) {                          // :626-637
                for (int i = 1; i <= m; ++i) {
                    real temp = M2(a, m, i, j);
                    M2(a, m, i, j) = M2(a, m, i, kmax);
                    M2(a, m, i, kmax) = temp;
                }
                V(rdiag, kmax) = V(rdiag, j);
                V(wa, kmax) = V(wa, j);
                int k = V(ipvt, j);
                V(ipvt, j) = V(ipvt, kmax);
                V(ipvt, kmax) = k;
            }
        }
        real ajnorm = f_norm2(&M2(a, m, j, j), m - j + 1);   // :642
        if (ajnorm != ZERO) {
            if (M2(a, m, j, j) < ZERO) ajnorm = -ajnorm;
            for (int i = j; i <= m; ++i) M2(a, m, i, j) = M2(a, m, i, j) / ajnorm;   // :645
            M2(a, m, j, j) = M2(a, m, j, j) + ONE;
            const int jp1 = j + 1;
            if (n >= jp1) {
                for (int k = jp1; k <= n; ++k) {      // :652-662
                    real sm = f_dot(&M2(a, m, j, j), &M2(a, m, j, k), m - j + 1);   // sequential, i = j..m
                    real temp = sm / M2(a, m, j, j);
                    for (int i = j; i <= m; ++i) M2(a, m, i, k) = M2(a, m, i, k) - temp * M2(a, m, i, j);
                    if (!pivot || V(rdiag, k) == ZERO) continue;
                    temp = M2(a, m, j, k) / V(rdiag, k);
                    V(rdiag, k) = V(rdiag, k) * f_sqrt(f_max(ZERO, ONE - temp * temp));
                    real q = V(rdiag, k) / V(wa, k);
                    if (p05 * (q * q) > epsmch) continue;
                    V(rdiag, k) = f_norm2(&M2(a, m, jp1, k), m - jp1 + 1);
                    V(wa, k) = V(rdiag, k);
                }
            }
        }
        V(rdiag, j) = -ajnorm;                        // :665
    }
}

// ---------------------------------------------------------------------------------------
// lmsolve, src/nonlin_least_squares.f90:670-791 (MINPACK QRSOLV lineage).  r is the leading
// n-by-n block of the m-by-n factored Jacobian (ldr = m).
// ---------------------------------------------------------------------------------------
void lm_qrsolve(int n, real* r, int ldr, const int* ipvt, const real* diag, const real* qtb, real* x, real* sdiag,
                real* wa) {
    const real qtr = 0.25, half = 0.5;
    for (int j = 1; j <= n; ++j) {                    // :710-714
        for (int i = j; i <= n; ++i) M2(r, ldr, i, j) = M2(r, ldr, j, i);
        V(x, j) = M2(r, ldr, j, j);
        V(wa, j) = V(qtb, j);
    }
    for (int j = 1; j <= n; ++j) {                    // :717-765
        int l = V(ipvt, j);
        if (V(diag, l) != ZERO) {
            for (int k = j; k <= n; ++k) V(sdiag, k) = ZERO;
            V(sdiag, j) = V(diag, l);
            real qtbpj = ZERO;
            for (int k = j; k <= n; ++k) {
                if (V(sdiag, k) == ZERO) continue;
                real cs, sn;
                if (f_abs(M2(r, ldr, k, k)) < f_abs(V(sdiag, k))) {      // :733-741
                    real ctan = M2(r, ldr, k, k) / V(sdiag, k);
                    sn = half / f_sqrt(qtr + qtr * (ctan * ctan));
                    cs = sn * ctan;
                } else {
                    real tn = V(sdiag, k) / M2(r, ldr, k, k);
                    cs = half / f_sqrt(qtr + qtr * (tn * tn));
                    sn = cs * tn;
                }
                M2(r, ldr, k, k) = cs * M2(r, ldr, k, k) + sn * V(sdiag, k);   // :745-748
                real temp = cs * V(wa, k) + sn * qtbpj;
                qtbpj = -sn * V(wa, k) + cs * qtbpj;
                V(wa, k) = temp;
                for (int i = k + 1; i <= n; ++i) {    // :753-757
                    temp = cs * M2(r, ldr, i, k) + sn * V(sdiag, i);
                    V(sdiag, i) = -sn * M2(r, ldr, i, k) + cs * V(sdiag, i);
                    M2(r, ldr, i, k) = temp;
                }
            }
        }
        V(sdiag, j) = M2(r, ldr, j, j);               // :763-764
        M2(r, ldr, j, j) = V(x, j);
    }
    int nsing = n;                                    // :769-784
    for (int j = 1; j <= n; ++j) {
        if (V(sdiag, j) == ZERO && nsing == n) nsing = j - 1;
        if (nsing < n) V(wa, j) = ZERO;
    }
    for (int k = 1; k <= nsing; ++k) {
        int j = nsing - k + 1;
        real sm = ZERO;
        for (int i = j + 1; i <= nsing; ++i) sm += M2(r, ldr, i, j) * V(wa, i);
        V(wa, j) = (V(wa, j) - sm) / V(sdiag, j);
    }
    for (int j = 1; j <= n; ++j) V(x, V(ipvt, j)) = V(wa, j);   // :787-790
}

// ---------------------------------------------------------------------------------------
// lmpar, src/nonlin_least_squares.f90:394-566 (MINPACK LMPAR lineage, with the reference's
// two departures kept: the Newton correction updates the whole vector (:552), and inside
// the iteration dxnorm is NORM2 over *all m* entries of the work array wa2 (:531) — wa2 is
// the caller's m-element wa4, whose entries n+1..m still hold the tail of Q^T f (first
// inner pass) or of the last trial residual (later passes).
// ---------------------------------------------------------------------------------------
void lm_par(int m, int n, real* r, int ldr, const int* ipvt, const real* diag, const real* qtb, real delta,
            real* par, real* x, real* sdiag, real* wa1, real* wa2) {
    const real p001 = 1.0e-3, p1 = 0.1;
    const real dwarf = DWARF;
    int nsing = n;
    for (int j = 1; j <= n; ++j) {                    // :447-451
        V(wa1, j) = V(qtb, j);
        if (M2(r, ldr, j, j) == ZERO && nsing == n) nsing = j - 1;
        if (nsing < n) V(wa1, j) = ZERO;
    }
    for (int k = 1; k <= nsing; ++k) {                // :453-463
        int j = nsing - k + 1;
        V(wa1, j) = V(wa1, j) / M2(r, ldr, j, j);
        real temp = V(wa1, j);
        for (int i = 1; i <= j - 1; ++i) V(wa1, i) = V(wa1, i) - M2(r, ldr, i, j) * temp;
    }
    for (int j = 1; j <= n; ++j) V(x, V(ipvt, j)) = V(wa1, j);   // :466-469

    int iter = 0;                                     // :473-481
    for (int j = 1; j <= n; ++j) V(wa2, j) = V(diag, j) * V(x, j);
    real dxnorm = f_norm2(wa2, n);
    real fp = dxnorm - delta;
    if (fp <= p1 * delta) {
        *par = ZERO;
        return;
    }

    real parl = ZERO;                                 // :486-503
    if (nsing == n) {
        for (int j = 1; j <= n; ++j) {
            int l = V(ipvt, j);
            V(wa1, j) = V(diag, l) * (V(wa2, l) / dxnorm);
        }
        for (int j = 1; j <= n; ++j) {
            real sm = ZERO;
            for (int i = 1; i <= j - 1; ++i) sm += M2(r, ldr, i, j) * V(wa1, i);
            V(wa1, j) = (V(wa1, j) - sm) / M2(r, ldr, j, j);
        }
        real temp = f_norm2(wa1, n);
        parl = ((fp / delta) / temp) / temp;
    }

    for (int j = 1; j <= n; ++j) {                    // :506-513
        real sm = ZERO;
        for (int i = 1; i <= j; ++i) sm += M2(r, ldr, i, j) * V(qtb, i);
        int l = V(ipvt, j);
        V(wa1, j) = sm / V(diag, l);
    }
    real gnorm = f_norm2(wa1, n);
    real paru = gnorm / delta;
    if (paru == ZERO) paru = dwarf / f_min(delta, p1);

    *par = f_max(*par, parl);                         // :517-519
    *par = f_min(*par, paru);
    if (*par == ZERO) *par = gnorm / dxnorm;

    for (;;) {                                        // :522-563
        ++iter;
        if (*par == ZERO) *par = f_max(dwarf, p001 * paru);
        real temp = f_sqrt(*par);
        for (int j = 1; j <= n; ++j) V(wa1, j) = temp * V(diag, j);
        lm_qrsolve(n, r, ldr, ipvt, wa1, qtb, x, sdiag, wa2);            // :529
        for (int j = 1; j <= n; ++j) V(wa2, j) = V(diag, j) * V(x, j);   // :530
        dxnorm = f_norm2(wa2, m);                     // :531  (whole m-element work array)
        temp = fp;
        fp = dxnorm - delta;

        if (f_abs(fp) <= p1 * delta || (parl == ZERO && fp <= temp && temp < ZERO) || iter == 10) break;   // :538-540

        for (int j = 1; j <= n; ++j) {                // :543-546
            int l = V(ipvt, j);
            V(wa1, j) = V(diag, l) * (V(wa2, l) / dxnorm);
        }
        for (int j = 1; j <= n; ++j) {                // :547-553
            V(wa1, j) = V(wa1, j) / V(sdiag, j);
            temp = V(wa1, j);
            if (n < j + 1) continue;
            for (int i = 1; i <= n; ++i) V(wa1, i) = V(wa1, i) - M2(r, ldr, i, j) * temp;
        }
        temp = f_norm2(wa1, n);
        real parc = ((fp / delta) / temp) / temp;     // :555

        if (fp > ZERO) parl = f_max(parl, *par);      // :558-559
        if (fp < ZERO) paru = f_min(paru, *par);
        *par = f_max(parl, *par + parc);              // :562
    }
}

// ---------------------------------------------------------------------------------------
// lss_solve, src/nonlin_least_squares.f90:118-391
// ---------------------------------------------------------------------------------------
int lm_solve(const Problem* p, const FcnCtx* c, const Params* prm, real* x, real* fvec, IterBehavior* ib,
             Workspace* ws) {
    const real p0001 = 1.0e-4, p1 = 0.1, qtr = 0.25, half = 0.5, p75 = 0.75, one = 1.0, zero = 0.0;
    const int neqn = c->m, nvar = c->n;
    bool xcnvrg = false, fcnvrg = false, gcnvrg = false;
    int neval = 0, iter = 0, njac = 0, flag = 0;
    const real fac = prm->lm_factor, ftol = prm->fcn_tol, xtol = prm->var_tol, gtol = prm->grad_tol;
    const int maxeval = prm->max_fcn_evals;
    const real eps = EPSMCH;
    ib->iter_count = 0; ib->fcn_count = 0; ib->jacobian_count = 0; ib->gradient_count = 0;
    ib->converge_on_fcn = 0; ib->converge_on_chng = 0; ib->converge_on_zero_diff = 0;
    if (nvar > neqn) return NL_UNDERDEFINED_PROBLEM_ERROR;              // :189

    real* base = ws->get((size_t)neqn * nvar + 5 * (size_t)nvar + 2 * (size_t)neqn);   // :199-208
    real* jac = base;
    real* diag = jac + (size_t)neqn * nvar;
    real* qtf = diag + nvar;
    real* wa1 = qtf + nvar;
    real* wa2 = wa1 + nvar;
    real* wa3 = wa2 + nvar;
    real* wa4 = wa3 + nvar;
    real* fdw = wa4 + neqn;                            // vfh_jac_fcn's own work array (:253)
    int* jpvt = ws->geti(nvar);

    p->fcn(x, fvec, c);                                // :211-213
    neval = 1;
    real fnorm = f_norm2(fvec, neqn);

    real par = zero, xnorm = zero, delta = zero, gnorm = zero, temp = zero;
    iter = 1;
    for (;;) {                                         // :219-375
        fd_jacobian(p, c, prm, x, jac, fvec, fdw);     // :221-222
        ++njac;
        lm_factor(neqn, nvar, jac, true, jpvt, wa1, wa2, wa3);   // :225

        if (iter == 1) {                               // :229-238
            for (int j = 1; j <= nvar; ++j) {
                V(diag, j) = V(wa2, j);
                if (V(wa2, j) == zero) V(diag, j) = one;
            }
            for (int j = 1; j <= nvar; ++j) V(wa3, j) = V(diag, j) * V(x, j);
            xnorm = f_norm2(wa3, nvar);
            delta = fac * xnorm;
            if (delta == zero) delta = fac;
        }

        for (int i = 1; i <= neqn; ++i) V(wa4, i) = V(fvec, i);   // :241-253
        for (int j = 1; j <= nvar; ++j) {
            if (M2(jac, neqn, j, j) != zero) {
                real sm = f_dot(&M2(jac, neqn, j, j), &V(wa4, j), neqn - j + 1);   // sequential, i = j..neqn
                temp = -sm / M2(jac, neqn, j, j);
                for (int i = j; i <= neqn; ++i) V(wa4, i) = V(wa4, i) + M2(jac, neqn, i, j) * temp;
            }
            M2(jac, neqn, j, j) = V(wa1, j);
            V(qtf, j) = V(wa4, j);
        }

        gnorm = zero;                                  // :256-267
        if (fnorm != zero) {
            for (int j = 1; j <= nvar; ++j) {
                int l = V(jpvt, j);
                if (V(wa2, l) == zero) continue;
                real sm = zero;
                for (int i = 1; i <= j; ++i) sm += M2(jac, neqn, i, j) * (V(qtf, i) / fnorm);
                gnorm = f_max(gnorm, f_abs(sm / V(wa2, l)));
            }
        }

        if (gnorm <= gtol) { gcnvrg = true; break; }   // :270-273

        for (int j = 1; j <= nvar; ++j) V(diag, j) = f_max(V(diag, j), V(wa2, j));   // :276-278

        for (;;) {                                     // :281-366
            lm_par(neqn, nvar, jac, neqn, jpvt, diag, qtf, delta, &par, wa1, wa2, wa3, wa4);   // :283

            for (int j = 1; j <= nvar; ++j) {          // :286-291
                V(wa1, j) = -V(wa1, j);
                V(wa2, j) = V(x, j) + V(wa1, j);
                V(wa3, j) = V(diag, j) * V(wa1, j);
            }
            real pnorm = f_norm2(wa3, nvar);
            if (iter == 1) delta = f_min(delta, pnorm);   // :294

            p->fcn(wa2, wa4, c);                       // :297-299
            ++neval;
            real fnorm1 = f_norm2(wa4, neqn);

            real actred = -one;                        // :302-303
            if (p1 * fnorm1 < fnorm) { real q = fnorm1 / fnorm; actred = one - q * q; }

            for (int j = 1; j <= nvar; ++j) {          // :307-312
                V(wa3, j) = zero;
                int l = V(jpvt, j);
                temp = V(wa1, l);
                for (int i = 1; i <= j; ++i) V(wa3, i) = V(wa3, i) + M2(jac, neqn, i, j) * temp;
            }
            real temp1 = f_norm2(wa3, nvar) / fnorm;   // :313-316
            real temp2 = (f_sqrt(par) * pnorm) / fnorm;
            real prered = temp1 * temp1 + temp2 * temp2 / half;
            real dirder = -(temp1 * temp1 + temp2 * temp2);

            real ratio = zero;                         // :319-320
            if (prered != zero) ratio = actred / prered;

            if (ratio <= qtr) {                        // :323-337
                if (actred >= zero) temp = half;
                if (actred < zero) temp = half * dirder / (dirder + half * actred);
                if (p1 * fnorm1 >= fnorm || temp < p1) temp = p1;
                delta = temp * f_min(delta, pnorm / p1);
                par = par / temp;
            } else {
                if (par != zero && ratio < p75) {
                } else {
                    delta = pnorm / half;
                    par = half * par;
                }
            }

            if (ratio >= p0001) {                      // :340-349
                for (int j = 1; j <= nvar; ++j) {
                    V(x, j) = V(wa2, j);
                    V(wa2, j) = V(diag, j) * V(x, j);
                }
                for (int i = 1; i <= neqn; ++i) V(fvec, i) = V(wa4, i);
                xnorm = f_norm2(wa2, nvar);
                fnorm = fnorm1;
                ++iter;
            }

            if (f_abs(actred) <= ftol && prered <= ftol && half * ratio <= one) fcnvrg = true;   // :352-355
            if (delta <= xtol * xnorm) xcnvrg = true;
            if (fcnvrg || xcnvrg) break;

            if (neval >= maxeval) flag = NL_CONVERGENCE_ERROR;            // :358-363
            if (f_abs(actred) <= eps && prered <= eps && half * ratio <= one) flag = NL_TOLERANCE_TOO_SMALL_ERROR;
            if (delta <= eps * xnorm) flag = NL_TOLERANCE_TOO_SMALL_ERROR;
            if (gnorm <= eps) flag = NL_TOLERANCE_TOO_SMALL_ERROR;
            if (flag != 0) break;

            if (ratio >= p0001) break;                 // :365
        }
        if (fcnvrg || xcnvrg || gcnvrg || flag != 0) break;   // :369
    }

    ib->iter_count = iter; ib->fcn_count = neval; ib->jacobian_count = njac;   // :378-385
    ib->converge_on_fcn = fcnvrg; ib->converge_on_chng = xcnvrg; ib->converge_on_zero_diff = gcnvrg;
    if (flag != 0) return NL_CONVERGENCE_ERROR;        // :388-390 (always this code)
    return NL_NO_ERROR;
}

// ---------------------------------------------------------------------------------------
// ns_solve, src/nonlin_solve.f90:452-638.  The uncounted Jacobian evaluated on an undefined
// fvec before the loop (:535) has no observable effect and is not restated.
// ---------------------------------------------------------------------------------------
int newton_solve(const Problem* p, const FcnCtx* c, const Params* prm, real* x, real* fvec, IterBehavior* ib,
                 Workspace* ws) {
    const real zero = 0.0, half = 0.5, factor = 1.0e2;
    const int neqn = c->m, nvar = c->n;
    bool check = false, xcnvrg = false, fcnvrg = false, gcnvrg = false;
    int neval = 0, iter = 0, njac = 0, flag = 0, status = NL_NO_ERROR;
    const real ftol = prm->fcn_tol, xtol = prm->var_tol, gtol = prm->grad_tol;
    const int maxeval = prm->max_fcn_evals;
    ib->iter_count = 0; ib->fcn_count = 0; ib->jacobian_count = 0; ib->gradient_count = 0;
    ib->converge_on_fcn = 0; ib->converge_on_chng = 0; ib->converge_on_zero_diff = 0;
    if (nvar != neqn) return NL_INVALID_INPUT_ERROR;   // :519

    real* base = ws->get((size_t)nvar * nvar + 5 * (size_t)nvar);
    real* jac = base;
    real* dir = jac + (size_t)nvar * nvar;
    real* grad = dir + nvar;
    real* xold = grad + nvar;
    real* fdw = xold + nvar;
    int* ipvt = ws->geti(nvar);
    IterBehavior lib = {0, 0, 0, 0, 0, 0, 0};

    p->fcn(x, fvec, c);                                // :538-547
    real f = half * f_dot(fvec, fvec, neqn);
    ++neval;
    real test = zero;
    for (int i = 1; i <= neqn; ++i) test = f_max(f_abs(V(fvec, i)), test);
    if (test < ftol) fcnvrg = true;

    if (!fcnvrg) {
        const real stpmax = factor * f_max(f_norm2(x, nvar), real(nvar));   // :553
        for (;;) {                                     // :556-620
            ++iter;
            fd_jacobian(p, c, prm, x, jac, fvec, fdw); // :561-562
            ++njac;
            for (int i = 1; i <= nvar; ++i) V(grad, i) = f_dot(&M2(jac, neqn, 1, i), fvec, neqn);   // :565-567
            la_dgetrf(nvar, nvar, jac, nvar, ipvt);    // :570   lu_factor
            for (int i = 0; i < nvar; ++i) xold[i] = x[i];   // :573-574
            real fold = f;
            for (int i = 0; i < nvar; ++i) dir[i] = -fvec[i];   // :577   solve_lu(lu, ipvt, -fvec)
            la_dgetrs(nvar, jac, nvar, ipvt, dir);

            if (prm->use_line_search) {                // :580-589
                real temp = f_dot(dir, dir, nvar);
                if (temp > stpmax) { real s = stpmax / temp; for (int i = 0; i < nvar; ++i) dir[i] = dir[i] * s; }
                limit_vector(nvar, dir, stpmax);
                int ls = line_search(p, c, prm, xold, grad, dir, x, fvec, fold, &f, &lib);
                neval += lib.fcn_count;
                if (ls != NL_NO_ERROR) { status = ls; break; }
            } else {                                   // :590-596
                for (int i = 0; i < nvar; ++i) x[i] = x[i] + dir[i];
                p->fcn(x, fvec, c);
                f = half * f_dot(fvec, fvec, neqn);
                ++neval;
            }

            real xnorm, fnorm;                         // :599-608
            test_convergence(nvar, neqn, x, xold, fvec, grad, true, xtol, ftol, gtol, &check, &xcnvrg, &fcnvrg,
                             &gcnvrg, &xnorm, &fnorm);
            if (check) break;
            else if (gcnvrg) { status = NL_SPURIOUS_CONVERGENCE_ERROR; break; }

            if (neval >= maxeval) { flag = 1; break; } // :616-619
        }
    }
    ib->iter_count = iter; ib->fcn_count = neval; ib->jacobian_count = njac; ib->gradient_count = 0;   // :624-632
    ib->converge_on_fcn = fcnvrg; ib->converge_on_chng = xcnvrg; ib->converge_on_zero_diff = gcnvrg;
    if (status != NL_NO_ERROR) return status;
    if (flag != 0) return NL_CONVERGENCE_ERROR;        // :635-637
    return NL_NO_ERROR;
}

// ---------------------------------------------------------------------------------------
// qns_solve, src/nonlin_solve.f90:156-425
// ---------------------------------------------------------------------------------------
int broyden_solve(const Problem* p, const FcnCtx* c, const Params* prm, real* x, real* fvec, IterBehavior* ib,
                  Workspace* ws) {
    const real zero = 0.0, half = 0.5, one = 1.0, factor = 1.0e2;
    const int neqn = c->m, nvar = c->n;
    bool restart = true, xcnvrg = false, fcnvrg = false, gcnvrg = false, check = false;
    int neval = 0, iter = 0, njac = 0, flag = 0, jcount = 0, status = NL_NO_ERROR;
    const real ftol = prm->fcn_tol, xtol = prm->var_tol, gtol = prm->grad_tol;
    const int maxeval = prm->max_fcn_evals;
    ib->iter_count = 0; ib->fcn_count = 0; ib->jacobian_count = 0; ib->gradient_count = 0;
    ib->converge_on_fcn = 0; ib->converge_on_chng = 0; ib->converge_on_zero_diff = 0;
    if (nvar != neqn) return NL_INVALID_INPUT_ERROR;   // :237

    const size_t nn = (size_t)nvar * nvar;
    real* base = ws->get(3 * nn + 10 * (size_t)nvar);
    real* b = base;
    real* q = b + nn;
    real* r = q + nn;
    real* df = r + nn;
    real* fvold = df + nvar;
    real* xold = fvold + nvar;
    real* dx = xold + nvar;
    real* s = dx + nvar;
    real* tau = s + nvar;
    real* work = tau + nvar;       // n (LAPACK work)
    real* w2 = work + nvar;        // 2n (DQR1UP work)
    real* fdw = w2 + 2 * nvar;     // n (FD work)
    IterBehavior lib = {0, 0, 0, 0, 0, 0, 0};   // uninitialised in the reference when the line search is off

    p->fcn(x, fvec, c);                                // :257-266
    real f = half * f_dot(fvec, fvec, neqn);
    ++neval;
    real test = zero;
    for (int i = 1; i <= neqn; ++i) test = f_max(f_abs(V(fvec, i)), test);
    if (test < ftol) fcnvrg = true;

    if (!fcnvrg) {
        const real stpmax = factor * f_max(f_norm2(x, nvar), real(nvar));   // :272
        real fold = f;
        for (;;) {                                     // :275-407
            ++iter;
            if (iter > prm->max_iter_guard) { flag = 1; break; }   // guard, not in the reference
            if (restart) {                             // :280-289
                fd_jacobian(p, c, prm, x, b, fvec, fdw);
                ++njac;
                // qr_factor(b, q = q, r = r): DGEQR2 on a copy, R = upper triangle, Q by DORG2R
                for (size_t e = 0; e < nn; ++e) q[e] = b[e];
                la_dgeqr2(nvar, nvar, q, nvar, tau, work);
                for (int j = 1; j <= nvar; ++j)
                    for (int i = 1; i <= nvar; ++i) M2(r, nvar, i, j) = (i <= j) ? M2(q, nvar, i, j) : zero;
                la_dorg2r(nvar, nvar, nvar, q, nvar, tau, work);
                jcount = 0;
            } else {                                   // :292-307
                for (int i = 0; i < nvar; ++i) df[i] = fvec[i] - fvold[i];
                for (int i = 0; i < nvar; ++i) dx[i] = x[i] - xold[i];
                real x2 = f_dot(dx, dx, nvar);
                // s = df - matmul(b, dx): c(i) accumulates b(i,j)*dx(j) over j from zero
                for (int i = 1; i <= nvar; ++i) V(s, i) = zero;
                for (int j = 1; j <= nvar; ++j)
                    for (int i = 1; i <= nvar; ++i) V(s, i) = V(s, i) + M2(b, nvar, i, j) * V(dx, j);
                for (int i = 1; i <= nvar; ++i) V(s, i) = V(df, i) - V(s, i);
                la_drscl(nvar, x2, s);                 // :298   recip_mult_array
                la_dger(nvar, nvar, one, s, dx, b, nvar);           // :302   rank1_update
                la_dqr1up(nvar, nvar, q, nvar, r, nvar, s, dx, w2); // :303   qr_rank1_update
                ++jcount;
            }

            la_dgemv_t(nvar, nvar, one, b, nvar, fvec, dx);   // :311   grad = B^T f -> dx

            for (int i = 0; i < nvar; ++i) { xold[i] = x[i]; fvold[i] = fvec[i]; }   // :314-316
            fold = f;

            la_dgemv_t(nvar, nvar, -one, q, nvar, fvec, df);  // :320   -Q^T f -> df
            la_dtrsv_unn(nvar, r, nvar, df);                  // :325

            real temp = f_dot(dx, df, nvar);           // :330-337
            if (temp >= zero) { restart = true; continue; }

            if (prm->use_line_search) {                // :340-349
                temp = f_dot(df, df, nvar);
                if (temp > stpmax) { real sc = stpmax / temp; for (int i = 0; i < nvar; ++i) df[i] = df[i] * sc; }
                limit_vector(nvar, df, stpmax);
                int ls = line_search(p, c, prm, xold, dx, df, x, fvec, fold, &f, &lib);
                neval += lib.fcn_count;
                if (ls != NL_NO_ERROR) { status = ls; break; }
            } else {                                   // :350-356
                for (int i = 0; i < nvar; ++i) x[i] = x[i] + df[i];
                p->fcn(x, fvec, c);
                f = half * f_dot(fvec, fvec, neqn);
                ++neval;
            }

            real xnorm, fnorm;                         // :359-366
            bool lg = lib.converge_on_zero_diff && prm->use_line_search;
            test_convergence(nvar, neqn, x, xold, fvec, dx, lg, xtol, ftol, gtol, &check, &xcnvrg, &fcnvrg, &gcnvrg,
                             &xnorm, &fnorm);
            if (!check) {                              // :367-395
                if (gcnvrg) {
                    if (restart) { status = NL_SPURIOUS_CONVERGENCE_ERROR; break; }
                    else restart = true;
                } else {
                    restart = (jcount >= prm->jacobian_interval);
                }
            } else {
                break;
            }

            if (neval >= maxeval) { flag = 1; break; } // :403-406
        }
    }
    ib->iter_count = iter; ib->fcn_count = neval; ib->jacobian_count = njac; ib->gradient_count = 0;   // :411-419
    ib->converge_on_fcn = fcnvrg; ib->converge_on_chng = xcnvrg; ib->converge_on_zero_diff = gcnvrg;
    if (status != NL_NO_ERROR) return status;
    if (flag != 0) return NL_CONVERGENCE_ERROR;        // :422-424
    return NL_NO_ERROR;
}

// ---------------------------------------------------------------------------------------
// constrained_least_squares_solver: bounded trust-region dogleg.
//   cls_solve            src/nonlin_least_squares.f90:938-1176
//   ces_apply_limits     :858-883
//   alpha_box            :1181-1219
//   coleman_li_scaling   :1222-1260
//   scaled_norm          :1263-1273
//   is_finite_array      :1276-1298   (tests NaN and |x| == huge only: +-Inf passes as finite)
//   dogleg               :1301-1403
// linalg boundary (qr_factor(a, tau=, qr=), solve_qr(qr, tau, b)): see nl_lapack.h.
// ---------------------------------------------------------------------------------------
void cls_options_default(ClsOptions* o) {
    o->trust_region_radius = 1.0;      // m_delta   least_squares:64
    o->step_scaling_factor = 1.0;      // m_scaling least_squares:66
    o->lower = nullptr;
    o->upper = nullptr;
}

static const double HUGE_D = DBL_MAX;  // huge(0d0)

static void apply_limits(int n, real* x, const real* xl, const real* xu) {   // :858-883
    for (int i = 1; i <= n; ++i)
        if (V(x, i) < V(xl, i)) V(x, i) = V(xl, i);
    for (int i = 1; i <= n; ++i)
        if (V(x, i) > V(xu, i)) V(x, i) = V(xu, i);
}

static bool is_finite_array(int n, const real* x) {                          // :1276-1298
    for (int i = 1; i <= n; ++i) {
        if (!(V(x, i) == V(x, i))) return false;
        if (f_abs(V(x, i)) == real(HUGE_D)) return false;
    }
    return true;
}

real alpha_box(int n, const real* x, const real* p, const real* xl, const real* xu) {   // :1181-1219
    real rst = HUGE_D;
    for (int i = 1; i <= n; ++i) {
        if (V(p, i) > ZERO) {
            if (V(xu, i) < V(x, i)) return ZERO;
            real a = (V(xu, i) - V(x, i)) / V(p, i);
            if (a < rst) rst = a;
        } else if (V(p, i) < ZERO) {
            if (V(xl, i) > V(x, i)) return ZERO;
            real a = (V(xl, i) - V(x, i)) / V(p, i);
            if (a < rst) rst = a;
        }
    }
    if (rst < ZERO) rst = ZERO;
    return rst;
}

void coleman_li_scaling(int n, const real* x, const real* xl, const real* xu, real* s) {   // :1222-1260
    const real min_scale = 1.0e-8, max_scale = 1.0e8, big = HUGE_D;
    for (int i = 1; i <= n; ++i) {
        real di;
        if (V(xl, i) > -big && V(xu, i) < big) di = f_min(V(x, i) - V(xl, i), V(xu, i) - V(x, i));
        else if (V(xl, i) > -big) di = V(x, i) - V(xl, i);
        else if (V(xu, i) < big) di = V(xu, i) - V(x, i);
        else di = ONE;
        di = f_max(di, min_scale);
        V(s, i) = ONE / di;
        if (V(s, i) > max_scale) V(s, i) = max_scale;
    }
}

static real scaled_norm(int n, const real* x, const real* s, real* tmp) {    // :1263-1273
    for (int i = 1; i <= n; ++i) V(tmp, i) = V(x, i) * V(s, i);
    return f_norm2(tmp, n);
}

// wrk: 5n + 2m entries (pgn, psd, u, v, tmp, Jg, rhs)
void dogleg(int m, int n, real delta, const real* x, const real* f, const real* jac, real* qr, const real* tau,
            const real* s, const real* xl, const real* xu, real* p, real* g, real* Jp, real* prered, real* wrk) {
    real* pgn = wrk;
    real* psd = pgn + n;
    real* u = psd + n;
    real* v = u + n;
    real* tmp = v + n;
    real* Jg = tmp + n;
    real* rhs = Jg + m;
    real alpha;
    la_dgemv_t(m, n, ONE, jac, m, f, g);                                     // :1341
    for (int i = 1; i <= m; ++i) V(rhs, i) = V(f, i);                        // :1344  u = solve_qr(qr, tau, f)
    la_dorm2r_lt_vec(m, n, qr, m, tau, rhs);
    la_dtrsv_unn(n, qr, m, rhs);
    for (int i = 1; i <= n; ++i) V(u, i) = V(rhs, i);
    for (int i = 1; i <= n; ++i) V(pgn, i) = -V(u, i);
    real pgnnorm = scaled_norm(n, pgn, s, tmp);
    if (pgnnorm > delta) {                                                   // :1349
        la_dgemv_n(m, n, ONE, jac, m, g, Jg);
        real c1 = f_dot(g, g, n);
        real c2 = f_dot(Jg, Jg, m);
        if (c2 > ZERO && c1 > ZERO) alpha = c1 / c2;
        else alpha = ZERO;
        for (int i = 1; i <= n; ++i) V(psd, i) = -alpha * V(g, i);
        real psdnorm = scaled_norm(n, psd, s, tmp);
        if (psdnorm >= delta && psdnorm > ZERO) {
            real sc = delta / psdnorm;
            for (int i = 1; i <= n; ++i) V(p, i) = sc * V(psd, i);
        } else {
            for (int i = 1; i <= n; ++i) V(u, i) = V(pgn, i) - V(psd, i);
            for (int i = 1; i <= n; ++i) V(u, i) = V(s, i) * V(u, i);
            for (int i = 1; i <= n; ++i) V(v, i) = V(s, i) * V(psd, i);
            real a = f_dot(u, u, n);
            real b = real(2.0) * f_dot(u, v, n);
            real cq = f_dot(v, v, n) - delta * delta;
            if (a <= ZERO) {
                for (int i = 1; i <= n; ++i) V(p, i) = V(psd, i);
            } else {
                real t;
                real arg = f_max(ZERO, b * b - real(4.0) * a * cq);
                if (arg == ZERO) {
                    t = -b / (real(2.0) * a);
                } else {
                    t = (-b + f_sqrt(arg)) / (real(2.0) * a);
                    if (t < ZERO || t > ONE) t = (-b - f_sqrt(arg)) / (real(2.0) * a);
                }
                t = f_max(ZERO, f_min(ONE, t));
                for (int i = 1; i <= n; ++i) V(p, i) = V(psd, i) + t * V(u, i);   // u is the *scaled* difference (:1384)
            }
        }
    } else {
        for (int i = 1; i <= n; ++i) V(p, i) = V(pgn, i);
    }
    alpha = alpha_box(n, x, p, xl, xu);                                      // :1393-1396
    if (alpha < ONE)
        for (int i = 1; i <= n; ++i) V(p, i) = alpha * V(p, i);
    la_dgemv_n(m, n, ONE, jac, m, p, Jp);                                    // :1399-1402
    real c1 = f_dot(g, p, n);
    real c2 = HALF * f_dot(Jp, Jp, m);
    *prered = -c1 - c2;
}

int cls_solve(const Problem* p, const FcnCtx* c, const Params* prm, const ClsOptions* opt, real* x, real* fvec,
              IterBehavior* ib, Workspace* ws) {
    const real delta_max = 1.0e3, eta = 1.0e-1, ls_cl = 1.0e-4, ls_beta = 0.5;
    const int ls_max_iter = 10;
    const int neqn = c->m, nvar = c->n;
    bool converged = false, xcnvrg = false, fcnvrg = false, gcnvrg = false;
    int neval = 0, iter = 0, njac = 0;
    const real ftol = prm->fcn_tol, xtol = prm->var_tol, gtol = prm->grad_tol;
    const int maxeval = prm->max_fcn_evals;
    ib->iter_count = 0; ib->fcn_count = 0; ib->jacobian_count = 0; ib->gradient_count = 0;   // :993-1001
    ib->converge_on_fcn = 0; ib->converge_on_chng = 0; ib->converge_on_zero_diff = 0;
    if (nvar > neqn) return NL_UNDERDEFINED_PROBLEM_ERROR;                   // :1005

    const size_t mn = (size_t)neqn * nvar;
    real* base = ws->get(2 * mn + 12 * (size_t)nvar + 5 * (size_t)neqn);
    real* jac = base;
    real* qr = jac + mn;
    real* tau = qr + mn;
    real* s = tau + nvar;
    real* g = s + nvar;
    real* pv = g + nvar;
    real* xnew = pv + nvar;
    real* xl = xnew + nvar;
    real* xu = xl + nvar;
    real* Jp = xu + nvar;
    real* fnew = Jp + neqn;
    real* fdw = fnew + neqn;
    real* dwrk = fdw + neqn;                            // 5n + 2m
    real qrwork[1];
    (void)qrwork;

    for (int i = 1; i <= nvar; ++i) {                                        // :1014-1024
        V(xl, i) = opt->lower ? real(opt->lower[i - 1]) : real(-HUGE_D);
        V(xu, i) = opt->upper ? real(opt->upper[i - 1]) : real(HUGE_D);
    }

    apply_limits(nvar, x, xl, xu);                                           // :1038-1045
    p->fcn(x, fvec, c);
    neval = 1;
    real fnorm = f_norm2(fvec, neqn);
    real xnorm = f_norm2(x, nvar);
    if (!is_finite_array(nvar, x) || !is_finite_array(neqn, fvec)) return NL_NO_ERROR;   // early return, ib stays zero

    real delta = opt->trust_region_radius;                                   // :1048
    iter = 1;
    for (;;) {
        fd_jacobian(p, c, prm, x, jac, fvec, fdw);                           // :1052-1053
        ++njac;
        for (size_t e = 0; e < mn; ++e) qr[e] = jac[e];                      // :1061
        la_dgeqr2(neqn, nvar, qr, neqn, tau, dwrk);
        coleman_li_scaling(nvar, x, xl, xu, s);                              // :1064
        real prered;
        dogleg(neqn, nvar, delta, x, fvec, jac, qr, tau, s, xl, xu, pv, g, Jp, &prered, dwrk);   // :1067-1068
        xnorm = scaled_norm(nvar, pv, s, dwrk);
        real gnorm = f_norm2(g, nvar);
        for (int i = 1; i <= nvar; ++i) V(xnew, i) = V(x, i) + V(pv, i);

        p->fcn(xnew, fnew, c);                                               // :1074-1076
        real fnewnorm = f_norm2(fnew, neqn);
        ++neval;

        real actred = HALF * (fnorm * fnorm - fnewnorm * fnewnorm);          // :1079-1084
        real rho;
        if (prered > ZERO && actred >= ZERO) rho = actred / prered;
        else rho = ZERO;

        if (rho < real(0.25)) {                                              // :1087-1091
            delta = f_max(real(0.25), real(1.0e-12));
        } else if (rho > real(0.75) && f_abs(xnorm - delta) < real(1.0e-12) * delta) {
            delta = f_min(real(2.0) * delta, delta_max);
        }

        if (rho > eta && fnewnorm <= fnorm) {                                // :1094-1100
            for (int i = 1; i <= nvar; ++i) V(x, i) = V(xnew, i);
            apply_limits(nvar, x, xl, xu);
            for (int i = 1; i <= neqn; ++i) V(fvec, i) = V(fnew, i);
            fnorm = fnewnorm;
            ++iter;
        } else {                                                             // :1101-1134
            real dderiv = f_dot(g, pv, nvar);
            if (dderiv >= ZERO) {
                delta = f_max(HALF * delta, real(1.0e-12));
            } else {
                real stepscale = opt->step_scaling_factor;
                int k;
                for (k = 1; k <= ls_max_iter; ++k) {
                    for (int i = 1; i <= nvar; ++i) V(xnew, i) = V(x, i) + stepscale * V(pv, i);
                    apply_limits(nvar, xnew, xl, xu);
                    p->fcn(xnew, fnew, c);
                    ++neval;
                    fnewnorm = f_norm2(fnew, neqn);
                    if (fnewnorm <= fnorm + ls_cl * stepscale * dderiv) {
                        for (int i = 1; i <= nvar; ++i) V(x, i) = V(xnew, i);
                        for (int i = 1; i <= neqn; ++i) V(fvec, i) = V(fnew, i);
                        fnorm = fnewnorm;
                        ++iter;
                        delta = f_max(stepscale * xnorm, real(1.0e-12));
                        break;
                    }
                    stepscale = stepscale * ls_beta;
                }
                if (k > ls_max_iter) delta = f_max(HALF * delta, real(1.0e-12));
            }
        }

        if (!is_finite_array(nvar, x) || !is_finite_array(neqn, fvec)) break;   // :1137-1139

        if (xnorm <= xtol) { converged = true; xcnvrg = true; break; }       // :1142-1159
        if (f_abs(actred) <= ftol && f_abs(prered) <= ftol && HALF * rho <= ONE) {
            converged = true; fcnvrg = true; break;
        }
        if (gnorm <= gtol) { converged = true; gcnvrg = true; break; }
        if (neval >= maxeval) break;
    }
    ib->iter_count = iter; ib->fcn_count = neval; ib->jacobian_count = njac;    // :1163-1170 (gradient_count stays 0)
    ib->converge_on_fcn = fcnvrg; ib->converge_on_chng = xcnvrg; ib->converge_on_zero_diff = gcnvrg;
    if (!converged) return NL_CONVERGENCE_ERROR;                             // :1173-1175
    return NL_NO_ERROR;
}


}  // namespace nlo
