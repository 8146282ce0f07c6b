// tps_common.cuh — building blocks of the thread-per-system kernels: one CUDA thread owns
// one system of the batch, all of its state lives in registers / local memory, and every
// sum runs in the index order of the Fortran loop it replaces.
//
// Reference behaviour reproduced here:
//   fd_jacobian       vfh_jac_fcn            src/nonlin_multi_eqn_mult_var.f90:198-277
//   test_convergence  test_convergence       src/nonlin_helper.f90:36-124
//   limit_vector      limit_search_vector    src/nonlin_linesearch.f90:554-572
//   backtrack_min     min_backtrack_search   src/nonlin_linesearch.f90:495-551
//   line_search       ls_search_mimo         src/nonlin_linesearch.f90:152-326
#pragma once
#include "nlb_types.h"
#include "vecfcn_registry.cuh"

namespace nlb {

// Settings in the form the kernels read them (copied from nlb_params by the launcher).
struct DevParams {
    int max_fcn_evals;
    double fcn_tol, var_tol, grad_tol;
    double lm_factor;
    int jacobian_interval;
    int use_line_search;
    int ls_max_fcn_evals;
    double ls_alpha, ls_factor;
    int use_analytic_jacobian;
    int max_iter_guard;
};

// What one solve reports besides x and fvec (iteration_behavior, src/nonlin_types.f90:8-29).
struct SolveStats {
    int iter = 0, nfev = 0, njac = 0;
    int cf = 0, cx = 0, cg = 0;
    int status = 0;
};

// Forward-difference Jacobian.  h = sqrt(eps)*|x_j| (sqrt(eps) if that is zero); divide by the
// nominal h; columns in order j = 1..n; the caller's f(x) is reused.
// Matrix stored one element every STRIDE doubles: a thread's slice of a [element][thread] shared-memory tile
// (conflict-free: consecutive lanes hold consecutive doubles).  Indexable like a per-thread array.
template <int STRIDE>
struct StridedMat {
    double* p;
    NLB_DEV double& operator[](int e) const { return p[e * STRIDE]; }
};

template <class F, int K>
NLB_DEV void analytic_jacobian(const double (&x)[F::N], double (&jac)[K], const SysCtx& c) {
    F::jac(x, JacView<F::M>{jac}, c);
}
template <class F, int STRIDE>
NLB_DEV void analytic_jacobian(const double (&)[F::N], const StridedMat<STRIDE>&, const SysCtx&) {
    static_assert(!F::HAS_JAC, "registered Jacobians write per-thread arrays");
}

template <class F, class A>
NLB_DEV void fd_jacobian(double (&x)[F::N], A& jac, const double (&fv)[F::M],
                         double (&wrk)[F::M], const SysCtx& c, bool analytic) {
    constexpr int M = F::M, N = F::N;
    if constexpr (F::HAS_JAC) {
        if (analytic) {
            analytic_jacobian<F>(x, jac, c);
            return;
        }
    }
    const double eps = 0x1p-26;   // sqrt(epsilon(1d0))
#pragma unroll
    for (int j = 0; j < N; ++j) {
        const double temp = x[j];
        double h = eps * fabs(temp);
        if (h == 0.0) h = eps;
        x[j] = temp + h;
        F::eval(x, wrk, c);
        x[j] = temp;
#pragma unroll(M <= 8 ? M : 1)
        for (int i = 0; i < M; ++i) jac[i + j * M] = (wrk[i] - fv[i]) / h;
    }
}

struct ConvFlags {
    bool c, cx, cf, cg;
};

template <int NV, int NE>
NLB_DEV ConvFlags test_convergence(const double (&x)[NV], const double (&xo)[NV], const double (&f)[NE],
                                   const double (&g)[NV], bool lg, double xtol, double ftol, double gtol) {
    ConvFlags r{false, false, false, false};
    double fc = 0.0;
#pragma unroll
    for (int i = 0; i < NE; ++i) fc += f[i] * f[i];
    fc = 0.5 * fc;
    double fnorm = 0.0;
#pragma unroll
    for (int i = 0; i < NE; ++i) fnorm = nl_max(fabs(f[i]), fnorm);
    if (fnorm < ftol) { r.cf = true; r.c = true; return r; }
    double xnorm = 0.0;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const double test = fabs(x[i] - xo[i]) / nl_max(fabs(x[i]), 1.0);
        xnorm = nl_max(test, xnorm);
    }
    if (xnorm < xtol) { r.cx = true; r.c = true; return r; }
    if (lg) {
        double test = 0.0;
        const double den = nl_max(fc, 0.5 * (double)NV);
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const double dxmax = fabs(g[i]) * nl_max(fabs(x[i]), 1.0) / den;
            test = nl_max(test, dxmax);
        }
        if (test < gtol) r.cg = true;
    }
    return r;
}

template <int N>
NLB_DEV void limit_vector(double (&x)[N], double lim) {
    const double mag = norm2_vec(x);
    if (mag == 0.0) return;
    if (mag > lim) {
        const double s = lim / mag;
#pragma unroll
        for (int i = 0; i < N; ++i) x[i] = s * x[i];
    }
}

NLB_DEV double backtrack_min(int mode, double f0, double f, double f1, double alam, double alam1, double slope) {
    double lam;
    if (mode == 1) {
        lam = -slope / (2.0 * (f - f0 - slope));
    } else {
        const double rhs1 = f - f0 - alam * slope;
        const double rhs2 = f1 - f0 - alam1 * slope;
        const double a = (rhs1 / (alam * alam) - rhs2 / (alam1 * alam1)) / (alam - alam1);
        const double b = (-(alam1 * rhs1 / (alam * alam)) + alam * rhs2 / (alam1 * alam1)) / (alam - alam1);
        // the three quotient forms of the cubic model share one division
        double num = -slope, den = 2.0 * b;
        bool divide = true;
        if (a != 0.0) {
            const double disc = b * b - 3.0 * a * slope;
            if (disc < 0.0) {
                divide = false;
            } else {
                const double sq = sqrt(disc);
                if (b <= 0.0) { num = -b + sq; den = 3.0 * a; }
                else { den = b + sq; }
            }
        }
        lam = divide ? num / den : 0.5 * alam;
        if (lam > 0.5 * alam) lam = 0.5 * alam;
    }
    return lam;
}

// Backtracking line search on 0.5*|F|^2 along dir from xold.  Returns 0 or the NL_* code of
// the `error stop` the reference would reach; nfev_out = residual evaluations spent.
template <class F>
NLB_DEV int line_search(const DevParams& p, const SysCtx& c, const double (&xold)[F::N], const double (&grad)[F::N],
                        const double (&dir)[F::N], double (&x)[F::N], double (&fvec)[F::M], double fold, double& fx,
                        int& nfev_out) {
    constexpr int M = F::M, N = F::N;
    int neval = 0, niter = 0;
    const double tolx = 0x1p-51;   // 2 * epsilon(1d0)
    fx = 0.0;
    nfev_out = 0;
    const double slope = dot_vec(grad, dir);
    if (slope >= 0.0) return NLB_DIVERGENT_BEHAVIOR_ERROR;
    double test = 0.0;
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const double temp = fabs(dir[i]) / nl_max(fabs(xold[i]), 1.0);
        if (temp > test) test = temp;
    }
    const double alamin = tolx / test;
    double alam = 1.0, alam1 = 0.0, f1 = 0.0, f = 0.0, tmplam;
    int status = 0;
    for (;;) {
#pragma unroll
        for (int i = 0; i < N; ++i) x[i] = xold[i] + alam * dir[i];
        F::eval(x, fvec, c);
        f = 0.5 * dot_vec(fvec, fvec);
        ++neval;
        ++niter;
        if (alam < alamin) {
            bool same = true;
#pragma unroll
            for (int i = 0; i < N; ++i) same = same && ((x[i] - xold[i]) == 0.0);
            if (same) { status = NLB_CONVERGENCE_ERROR; break; }
#pragma unroll
            for (int i = 0; i < N; ++i) x[i] = xold[i];
            break;
        } else if (f <= fold + p.ls_alpha * alam * slope) {
            break;
        } else {
            tmplam = backtrack_min(niter, fold, f, f1, alam, alam1, slope);
        }
        alam1 = alam;
        f1 = f;
        alam = nl_max(tmplam, p.ls_factor * alam);
        if (neval >= p.ls_max_fcn_evals) { status = NLB_CONVERGENCE_ERROR; break; }
    }
    fx = f;
    nfev_out = neval;
    return status;
}

}  // namespace nlb
