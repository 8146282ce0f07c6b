// tps_cls.cuh — constrained_least_squares_solver, one CUDA thread per system: a bounded
// trust-region dogleg on the Gauss-Newton model with Coleman-Li scaling and an Armijo
// fall-back along the step.
//
// Reference behaviour reproduced here (src/nonlin_least_squares.f90):
//   tps_cls_solve       cls_solve            :938-1176
//   apply_limits        ces_apply_limits     :858-883
//   alpha_box           alpha_box            :1181-1219
//   coleman_li_scaling  coleman_li_scaling   :1222-1260
//   scaled_norm         scaled_norm          :1263-1273
//   is_finite_array     is_finite_array      :1276-1298  (NaN and |x| == huge only; +-Inf counts as finite)
//   dogleg              dogleg               :1301-1403
// Quirks kept on purpose: `delta = max(0.25, 1e-12)` resets the radius to the constant 0.25
// on a poor step (:1088); the dogleg interpolation adds the *scaled* difference s*(pgn-psd)
// to the unscaled Cauchy step (:1384); a non-finite start returns without an error and with
// an all-zero iteration_behavior (:1043-1045).
//
// The linalg boundary (qr_factor(a, tau=, qr=) :1061, solve_qr(qr, tau, f) :1344) and the two
// BLAS calls (dgemv 'T' :1341, 'N' :1351,1399) follow the unblocked Reference-LAPACK forms
// those calls reach for min(m, n) < 32: DGEQR2 (DLARFG + DLARF), DORM2R('L','T') + DTRSV, and
// the reference DGEMV loops (same forms as tps_dense.cuh, here for an M x N matrix).
#pragma once
#include "tps_common.cuh"

namespace nlb {

constexpr int CLS_MAX_N = 16;

// constrained_least_squares_solver's own members and the limit arrays of constrained_equation_solver.
struct DevCls {
    double radius, scaling;
    double xl[CLS_MAX_N], xu[CLS_MAX_N];
};

#define NLB_CLS_UNROLL_M _Pragma("unroll(M <= 8 ? M : 1)")

template <int N>
NLB_DEV void apply_limits(double (&x)[N], const DevCls& o) {
#pragma unroll
    for (int i = 0; i < N; ++i)
        if (x[i] < o.xl[i]) x[i] = o.xl[i];
#pragma unroll
    for (int i = 0; i < N; ++i)
        if (x[i] > o.xu[i]) x[i] = o.xu[i];
}

template <int N>
NLB_DEV bool is_finite_array(const double (&x)[N]) {
    const double huge = 1.7976931348623157e+308;
    bool ok = true;
#pragma unroll(N <= 8 ? N : 1)
    for (int i = 0; i < N; ++i) ok = ok && (x[i] == x[i]) && (fabs(x[i]) != huge);
    return ok;
}

template <int N>
NLB_DEV double alpha_box(const double (&x)[N], const double (&p)[N], const DevCls& o) {
    double rst = 1.7976931348623157e+308;
    bool zero = false;                    // the reference returns 0 at the first violated bound
#pragma unroll
    for (int i = 0; i < N; ++i) {
        if (!zero) {
            if (p[i] > 0.0) {
                if (o.xu[i] < x[i]) zero = true;
                else {
                    const double a = (o.xu[i] - x[i]) / p[i];
                    if (a < rst) rst = a;
                }
            } else if (p[i] < 0.0) {
                if (o.xl[i] > x[i]) zero = true;
                else {
                    const double a = (o.xl[i] - x[i]) / p[i];
                    if (a < rst) rst = a;
                }
            }
        }
    }
    if (zero) return 0.0;
    if (rst < 0.0) rst = 0.0;
    return rst;
}

template <int N>
NLB_DEV void coleman_li_scaling(const double (&x)[N], const DevCls& o, double (&s)[N]) {
    const double min_scale = 1.0e-8, max_scale = 1.0e8, big = 1.7976931348623157e+308;
#pragma unroll
    for (int i = 0; i < N; ++i) {
        double di;
        if (o.xl[i] > -big && o.xu[i] < big) di = nl_min(x[i] - o.xl[i], o.xu[i] - x[i]);
        else if (o.xl[i] > -big) di = x[i] - o.xl[i];
        else if (o.xu[i] < big) di = o.xu[i] - x[i];
        else di = 1.0;
        di = nl_max(di, min_scale);
        s[i] = 1.0 / di;
        if (s[i] > max_scale) s[i] = max_scale;
    }
}

template <int N>
NLB_DEV double scaled_norm(const double (&x)[N], const double (&s)[N]) {
    double t[N];
#pragma unroll
    for (int i = 0; i < N; ++i) t[i] = x[i] * s[i];
    return norm2_vec(t);
}

// DLARFG on column i of the M x N matrix a: alpha = a(i,i), x = a(i+1:M-1, i).  Returns tau.
template <int M, int N>
NLB_DEV double cls_make_reflector(double (&a)[M * N], int i) {
    if (M - i <= 1) return 0.0;
    Dnrm2 acc;
    NLB_CLS_UNROLL_M
    for (int r = 0; r < M; ++r)
        if (r > i) acc.add(a[r + i * M]);
    double xnorm = acc.value();
    if (xnorm == 0.0) return 0.0;
    double alpha = a[i + i * M];
    double beta = -nl_sign(dlapy2(alpha, xnorm), alpha);
    const double safmin = 0x1p-969;                 // dlamch('S') / dlamch('E')
    int knt = 0;
    if (fabs(beta) < safmin) {
        const double rsafmn = 1.0 / safmin;
        do {
            ++knt;
            NLB_CLS_UNROLL_M
            for (int r = 0; r < M; ++r)
                if (r > i) a[r + i * M] = rsafmn * a[r + i * M];
            beta = beta * rsafmn;
            alpha = alpha * rsafmn;
        } while (fabs(beta) < safmin && knt < 20);
        Dnrm2 acc2;
        NLB_CLS_UNROLL_M
        for (int r = 0; r < M; ++r)
            if (r > i) acc2.add(a[r + i * M]);
        xnorm = acc2.value();
        beta = -nl_sign(dlapy2(alpha, xnorm), alpha);
    }
    const double tau = (beta - alpha) / beta;
    const double sc = 1.0 / (alpha - beta);
    NLB_CLS_UNROLL_M
    for (int r = 0; r < M; ++r)
        if (r > i) a[r + i * M] = sc * a[r + i * M];
    for (int j = 0; j < knt; ++j) beta = beta * safmin;
    a[i + i * M] = beta;
    return tau;
}

// Number of leading rows of v = a(i:M-1, i) up to its last non-zero entry (DLARF's ILADLR scan).
template <int M, int N>
NLB_DEV int cls_lastv(const double (&a)[M * N], int i) {
    int lastv = M - i;
    bool scanning = true;
    NLB_CLS_UNROLL_M
    for (int r = M - 1; r >= 0; --r) {
        if (r >= i && scanning) {
            if (a[r + i * M] == 0.0) lastv = r - i;
            else scanning = false;
        }
    }
    return lastv;
}

// DLARF('L'): H(i) applied to the trailing columns a(i:M-1, i+1:N-1); a(i,i) holds 1 for the call.
template <int M, int N>
NLB_DEV void cls_reflect_trailing(double (&a)[M * N], int i, double tau) {
    if (tau == 0.0) return;
    const int lastv = cls_lastv<M, N>(a, i);
    int lastc = 0;
#pragma unroll
    for (int c = 0; c < N; ++c) {
        if (c > i) {
            NLB_CLS_UNROLL_M
            for (int r = 0; r < M; ++r)
                if (r >= i && (r - i) < lastv && a[r + c * M] != 0.0) lastc = c - i;
        }
    }
    if (lastv <= 0 || lastc <= 0) return;
#pragma unroll
    for (int c = 0; c < N; ++c) {
        if (c > i && (c - i) <= lastc) {
            double temp = 0.0;
            NLB_CLS_UNROLL_M
            for (int r = 0; r < M; ++r)
                if (r >= i && (r - i) < lastv) temp += a[r + c * M] * a[r + i * M];
            const double w = 0.0 + 1.0 * temp;      // DGEMV: y = beta*y (0), then y += alpha*temp
            if (w != 0.0) {                          // DGER skips zero entries of y
                const double t = (-tau) * w;
                NLB_CLS_UNROLL_M
                for (int r = 0; r < M; ++r)
                    if (r >= i && (r - i) < lastv) a[r + c * M] = a[r + c * M] + a[r + i * M] * t;
            }
        }
    }
}

// DLARF('L') with one column: H(i) applied to the vector rows c(i:M-1); a(i,i) holds 1 for the call.
template <int M, int N>
NLB_DEV void cls_reflect_vec(const double (&a)[M * N], int i, double tau, double (&c)[M]) {
    if (tau == 0.0) return;
    const int lastv = cls_lastv<M, N>(a, i);
    bool any = false;
    NLB_CLS_UNROLL_M
    for (int r = 0; r < M; ++r)
        if (r >= i && (r - i) < lastv && c[r] != 0.0) any = true;
    if (lastv <= 0 || !any) return;
    double temp = 0.0;
    NLB_CLS_UNROLL_M
    for (int r = 0; r < M; ++r)
        if (r >= i && (r - i) < lastv) temp += c[r] * a[r + i * M];
    const double w = 0.0 + 1.0 * temp;
    if (w != 0.0) {
        const double t = (-tau) * w;
        NLB_CLS_UNROLL_M
        for (int r = 0; r < M; ++r)
            if (r >= i && (r - i) < lastv) c[r] = c[r] + a[r + i * M] * t;
    }
}

// DGEQR2 in place: R on and above the diagonal, reflector vectors below it.
template <int M, int N>
NLB_DEV void cls_qr_factor(double (&a)[M * N], double (&tau)[N]) {
#pragma unroll
    for (int i = 0; i < N; ++i) {
        tau[i] = cls_make_reflector<M, N>(a, i);
        if (i < N - 1) {
            const double aii = a[i + i * M];
            a[i + i * M] = 1.0;
            cls_reflect_trailing<M, N>(a, i, tau[i]);
            a[i + i * M] = aii;
        }
    }
}

// solve_qr(qr, tau, f): rhs := Q^T f by DORM2R('L','T'), then DTRSV('U','N','N') on its first N entries.
template <int M, int N>
NLB_DEV void cls_solve_qr(double (&qr)[M * N], const double (&tau)[N], const double (&f)[M], double (&u)[N]) {
    double rhs[M];
    NLB_CLS_UNROLL_M
    for (int r = 0; r < M; ++r) rhs[r] = f[r];
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const double aii = qr[i + i * M];
        qr[i + i * M] = 1.0;
        cls_reflect_vec<M, N>(qr, i, tau[i], rhs);
        qr[i + i * M] = aii;
    }
#pragma unroll
    for (int j = 0; j < N; ++j) u[j] = rhs[j];
#pragma unroll
    for (int j = N - 1; j >= 0; --j) {
        if (u[j] != 0.0) {
            u[j] = u[j] / qr[j + j * M];
            const double temp = u[j];
#pragma unroll
            for (int i = j - 1; i >= 0; --i) u[i] = u[i] - temp * qr[i + j * M];
        }
    }
}

// DGEMV('T'), alpha = 1, beta = 0: y = A^T x.
template <int M, int N>
NLB_DEV void cls_gemv_t(const double (&a)[M * N], const double (&x)[M], double (&y)[N]) {
#pragma unroll
    for (int j = 0; j < N; ++j) {
        double temp = 0.0;
        NLB_CLS_UNROLL_M
        for (int i = 0; i < M; ++i) temp += a[i + j * M] * x[i];
        y[j] = 0.0 + 1.0 * temp;
    }
}

// DGEMV('N'), alpha = 1, beta = 0: y = A x, columns outermost.
template <int M, int N>
NLB_DEV void cls_gemv_n(const double (&a)[M * N], const double (&x)[N], double (&y)[M]) {
    NLB_CLS_UNROLL_M
    for (int i = 0; i < M; ++i) y[i] = 0.0;
#pragma unroll
    for (int j = 0; j < N; ++j) {
        const double temp = 1.0 * x[j];
        NLB_CLS_UNROLL_M
        for (int i = 0; i < M; ++i) y[i] = y[i] + temp * a[i + j * M];
    }
}

template <int M>
NLB_DEV double cls_dot_m(const double (&a)[M], const double (&b)[M]) {
    double s = 0.0;
    NLB_CLS_UNROLL_M
    for (int i = 0; i < M; ++i) s += a[i] * b[i];
    return s;
}

template <int M>
NLB_DEV double cls_norm2_m(const double (&v)[M]) {
    Norm2 acc;
    NLB_CLS_UNROLL_M
    for (int i = 0; i < M; ++i) acc.add(v[i]);
    return acc.value();
}

// dogleg: step p inside the scaled trust region and the box, gradient g = J^T f, predicted reduction.
template <int M, int N>
NLB_DEV void dogleg(double delta, const double (&x)[N], const double (&f)[M], const double (&jac)[M * N],
                    double (&qr)[M * N], const double (&tau)[N], const double (&s)[N], const DevCls& o,
                    double (&p)[N], double (&g)[N], double (&Jp)[M], double& prered) {
    double pgn[N], u[N];
    cls_gemv_t<M, N>(jac, f, g);
    cls_solve_qr<M, N>(qr, tau, f, u);
#pragma unroll
    for (int i = 0; i < N; ++i) pgn[i] = -u[i];
    const double pgnnorm = scaled_norm(pgn, s);
    if (pgnnorm > delta) {
        double psd[N];
        cls_gemv_n<M, N>(jac, g, Jp);                 // Jg; Jp is free until the end
        const double c1 = dot_vec(g, g);
        const double c2 = cls_dot_m<M>(Jp, Jp);
        double alpha;
        if (c2 > 0.0 && c1 > 0.0) alpha = c1 / c2;
        else alpha = 0.0;
#pragma unroll
        for (int i = 0; i < N; ++i) psd[i] = -alpha * g[i];
        const double psdnorm = scaled_norm(psd, s);
        if (psdnorm >= delta && psdnorm > 0.0) {
            const double sc = delta / psdnorm;
#pragma unroll
            for (int i = 0; i < N; ++i) p[i] = sc * psd[i];
        } else {
            double v[N];
#pragma unroll
            for (int i = 0; i < N; ++i) u[i] = pgn[i] - psd[i];
#pragma unroll
            for (int i = 0; i < N; ++i) u[i] = s[i] * u[i];
#pragma unroll
            for (int i = 0; i < N; ++i) v[i] = s[i] * psd[i];
            const double a = dot_vec(u, u);
            const double b = 2.0 * dot_vec(u, v);
            const double cq = dot_vec(v, v) - delta * delta;
            if (a <= 0.0) {
#pragma unroll
                for (int i = 0; i < N; ++i) p[i] = psd[i];
            } else {
                double t;
                const double arg = nl_max(0.0, b * b - 4.0 * a * cq);
                if (arg == 0.0) {
                    t = -b / (2.0 * a);
                } else {
                    t = (-b + sqrt(arg)) / (2.0 * a);
                    if (t < 0.0 || t > 1.0) t = (-b - sqrt(arg)) / (2.0 * a);
                }
                t = nl_max(0.0, nl_min(1.0, t));
#pragma unroll
                for (int i = 0; i < N; ++i) p[i] = psd[i] + t * u[i];
            }
        }
    } else {
#pragma unroll
        for (int i = 0; i < N; ++i) p[i] = pgn[i];
    }
    const double ab = alpha_box(x, p, o);
    if (ab < 1.0) {
#pragma unroll
        for (int i = 0; i < N; ++i) p[i] = ab * p[i];
    }
    cls_gemv_n<M, N>(jac, p, Jp);
    const double c1 = dot_vec(g, p);
    const double c2 = 0.5 * cls_dot_m<M>(Jp, Jp);
    prered = -c1 - c2;
}

// State of one constrained solve between outer iterations.
template <class F>
struct ClsState {
    double x[F::N], fvec[F::M];
    double fnorm, delta;
    int iter, neval, njac;
    bool xcnvrg, fcnvrg, gcnvrg, converged;
};

// Everything before the outer loop (:1038-1049).  Returns false when the solve ends here: the start, after clamping
// to the limits, or the first residual is not finite - the reference returns without an error and ib stays zero.
template <class F>
NLB_DEV bool cls_begin(const DevCls& o, const SysCtx& c, ClsState<F>& s) {
    constexpr int M = F::M, N = F::N;
    static_assert(N <= CLS_MAX_N, "limit arrays are passed by value");
    s.xcnvrg = s.fcnvrg = s.gcnvrg = s.converged = false;
    s.iter = 0; s.neval = 0; s.njac = 0;
    apply_limits(s.x, o);
    F::eval(s.x, s.fvec, c);
    s.fnorm = cls_norm2_m<M>(s.fvec);
    if (!is_finite_array(s.x) || !is_finite_array(s.fvec)) { s.converged = true; return false; }
    s.neval = 1;
    s.delta = o.radius;
    s.iter = 1;
    return true;
}

// One trip of the outer loop (:1051-1160).  Returns true when the solve is over (converged or failed).
template <class F>
NLB_DEV bool cls_iterate(const DevParams& prm, const DevCls& o, const SysCtx& c, ClsState<F>& s) {
    constexpr int M = F::M, N = F::N;
    const double delta_max = 1.0e3, eta = 1.0e-1, ls_cl = 1.0e-4, ls_beta = 0.5;
    const int ls_max_iter = 10;
    double (&x)[N] = s.x;
    double (&fvec)[M] = s.fvec;
    double jac[M * N], qr[M * N], tau[N], sc[N], g[N], p[N], xnew[N], Jp[M], fnew[M];

    fd_jacobian<F>(x, jac, fvec, fnew, c, prm.use_analytic_jacobian != 0);
    ++s.njac;
#pragma unroll(M * N <= 16 ? M * N : 1)
    for (int e = 0; e < M * N; ++e) qr[e] = jac[e];
    cls_qr_factor<M, N>(qr, tau);
    coleman_li_scaling(x, o, sc);
    double prered;
    dogleg<M, N>(s.delta, x, fvec, jac, qr, tau, sc, o, p, g, Jp, prered);
    const double xnorm = scaled_norm(p, sc);
    const double gnorm = norm2_vec(g);
#pragma unroll
    for (int i = 0; i < N; ++i) xnew[i] = x[i] + p[i];

    F::eval(xnew, fnew, c);
    double fnewnorm = cls_norm2_m<M>(fnew);
    ++s.neval;

    const double actred = 0.5 * (s.fnorm * s.fnorm - fnewnorm * fnewnorm);
    double rho;
    if (prered > 0.0 && actred >= 0.0) rho = actred / prered;
    else rho = 0.0;

    if (rho < 0.25) {
        s.delta = nl_max(0.25, 1.0e-12);
    } else if (rho > 0.75 && fabs(xnorm - s.delta) < 1.0e-12 * s.delta) {
        s.delta = nl_min(2.0 * s.delta, delta_max);
    }

    if (rho > eta && fnewnorm <= s.fnorm) {
#pragma unroll
        for (int i = 0; i < N; ++i) x[i] = xnew[i];
        apply_limits(x, o);
        NLB_CLS_UNROLL_M
        for (int i = 0; i < M; ++i) fvec[i] = fnew[i];
        s.fnorm = fnewnorm;
        ++s.iter;
    } else {
        const double dderiv = dot_vec(g, p);
        if (dderiv >= 0.0) {
            s.delta = nl_max(0.5 * s.delta, 1.0e-12);
        } else {
            double stepscale = o.scaling;
            bool accepted = false;
            for (int k = 1; k <= ls_max_iter; ++k) {
#pragma unroll
                for (int i = 0; i < N; ++i) xnew[i] = x[i] + stepscale * p[i];
                apply_limits(xnew, o);
                F::eval(xnew, fnew, c);
                ++s.neval;
                fnewnorm = cls_norm2_m<M>(fnew);
                if (fnewnorm <= s.fnorm + ls_cl * stepscale * dderiv) {
#pragma unroll
                    for (int i = 0; i < N; ++i) x[i] = xnew[i];
                    NLB_CLS_UNROLL_M
                    for (int i = 0; i < M; ++i) fvec[i] = fnew[i];
                    s.fnorm = fnewnorm;
                    ++s.iter;
                    s.delta = nl_max(stepscale * xnorm, 1.0e-12);
                    accepted = true;
                    break;
                }
                stepscale = stepscale * ls_beta;
            }
            if (!accepted) s.delta = nl_max(0.5 * s.delta, 1.0e-12);
        }
    }

    if (!is_finite_array(x) || !is_finite_array(fvec)) return true;

    if (xnorm <= prm.var_tol) { s.converged = true; s.xcnvrg = true; return true; }
    if (fabs(actred) <= prm.fcn_tol && fabs(prered) <= prm.fcn_tol && 0.5 * rho <= 1.0) {
        s.converged = true; s.fcnvrg = true; return true;
    }
    if (gnorm <= prm.grad_tol) { s.converged = true; s.gcnvrg = true; return true; }
    return s.neval >= prm.max_fcn_evals;
}

template <class F>
NLB_DEV void cls_store(const ClsState<F>& s, long long B, long long b, double* __restrict__ xg, double* __restrict__ fg,
                       nlb_iteration_behavior* __restrict__ ibg, int32_t* __restrict__ statusg) {
    constexpr int M = F::M, N = F::N;
#pragma unroll
    for (int j = 0; j < N; ++j) xg[j * B + b] = s.x[j];
    NLB_CLS_UNROLL_M
    for (int i = 0; i < M; ++i) fg[i * B + b] = s.fvec[i];
    if (ibg) {
        nlb_iteration_behavior o;
        o.iter_count = s.iter;
        o.fcn_count = s.neval;
        o.jacobian_count = s.njac;
        o.gradient_count = 0;
        o.converge_on_fcn = s.fcnvrg;
        o.converge_on_chng = s.xcnvrg;
        o.converge_on_zero_diff = s.gcnvrg;
        ibg[b] = o;
    }
    if (statusg) statusg[b] = s.converged ? 0 : NLB_CONVERGENCE_ERROR;
}

// Persistent form: the grid holds only resident CTAs; one trip of the loop is ONE outer iteration of whatever
// system the lane holds, and a lane whose system ends stores it and pulls the next index from a global cursor
// (tps_fetch: one atomic per warp).  The spread of iteration counts between systems (2...17 on the 2x2 box
// problem) then no longer idles lanes; the warp re-converges at the top of every trip.  Same arithmetic per system.
NLB_DEV long long cls_fetch(unsigned long long* cursor) {
    const unsigned mask = __activemask();
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(mask) - 1;
    unsigned long long base = 0;
    if (lane == leader) base = atomicAdd(cursor, (unsigned long long)__popc(mask));
    base = __shfl_sync(mask, base, leader);
    return (long long)(base + __popc(mask & ((1u << lane) - 1u)));
}

template <class F>
NLB_DEV void tps_cls_refill(const DevParams& prm, const DevCls& o, long long nsys, long long B,
                            unsigned long long* cursor, double* __restrict__ xg, double* __restrict__ fg,
                            const double* __restrict__ sys, const double* __restrict__ shared,
                            nlb_iteration_behavior* __restrict__ ibg, int32_t* __restrict__ statusg) {
    constexpr int M = F::M, N = F::N;
    ClsState<F> s;
    long long b = -1;
    bool active = false, more = true;
    for (;;) {
        bool finished = false;
        if (!active && more) {
            b = cls_fetch(cursor);
            if (b >= nsys) more = false;
        }
        if (!active && more) {
#pragma unroll
            for (int j = 0; j < N; ++j) s.x[j] = xg[j * B + b];
            const SysCtx c{sys ? sys + b : nullptr, shared, B, M, N};
            if (cls_begin<F>(o, c, s)) active = true;
            else finished = true;
        }
        __syncwarp();
        if (!__any_sync(0xffffffffu, active || finished || more)) break;
        if (active) {
            const SysCtx c{sys ? sys + b : nullptr, shared, B, M, N};
            if (cls_iterate<F>(prm, o, c, s)) { finished = true; active = false; }
        }
        if (finished) cls_store<F>(s, B, b, xg, fg, ibg, statusg);
    }
}

}  // namespace nlb
