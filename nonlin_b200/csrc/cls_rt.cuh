// cls_rt.cuh — constrained_least_squares_solver for the curve-fit residual families (run-time number of
// observations m): one CUDA thread per system, the m-sized state (Jacobian, its QR copy, four m-vectors) in an HBM
// workspace laid out [element][thread] (a warp touches 256 contiguous bytes per element), the n-sized state in
// registers.  Same algorithm, same operation order as tps_cls.cuh (which keeps a fixed-size system in a thread's own
// arrays); the n-only pieces - limits, Coleman-Li scaling, alpha_box, scaled norm - are shared with it.
//
// Reference behaviour reproduced (src/nonlin_least_squares.f90): cls_solve :938-1176, dogleg :1301-1403, with
// qr_factor / solve_qr / dgemv as the unblocked Reference-LAPACK forms listed in tps_cls.cuh.
#pragma once
#include "tps_cls.cuh"

namespace nlb {

// element i of a per-thread array of the workspace
struct WsArr {
    double* p;
    long long s;
    NLB_DEV double& operator[](long long i) const { return p[i * s]; }
};

// DLARFG on column i of the m x N matrix a: alpha = a(i,i), x = a(i+1:m-1, i).  Returns tau.
NLB_DEV double rt_make_reflector(const WsArr& a, int m, int i) {
    if (m - i <= 1) return 0.0;
    const long long col = (long long)i * m;
    Dnrm2 acc;
    for (int r = i + 1; r < m; ++r) acc.add(a[r + col]);
    double xnorm = acc.value();
    if (xnorm == 0.0) return 0.0;
    double alpha = a[i + col];
    double beta = -nl_sign(dlapy2(alpha, xnorm), alpha);
    const double safmin = 0x1p-969;
    int knt = 0;
    if (fabs(beta) < safmin) {
        const double rsafmn = 1.0 / safmin;
        do {
            ++knt;
            for (int r = i + 1; r < m; ++r) a[r + col] = rsafmn * a[r + col];
            beta = beta * rsafmn;
            alpha = alpha * rsafmn;
        } while (fabs(beta) < safmin && knt < 20);
        Dnrm2 acc2;
        for (int r = i + 1; r < m; ++r) acc2.add(a[r + col]);
        xnorm = acc2.value();
        beta = -nl_sign(dlapy2(alpha, xnorm), alpha);
    }
    const double tau = (beta - alpha) / beta;
    const double sc = 1.0 / (alpha - beta);
    for (int r = i + 1; r < m; ++r) a[r + col] = sc * a[r + col];
    for (int j = 0; j < knt; ++j) beta = beta * safmin;
    a[i + col] = beta;
    return tau;
}

// rows of v = a(i:m-1, i) up to its last non-zero entry (DLARF's ILADLR scan)
NLB_DEV int rt_lastv(const WsArr& a, int m, int i) {
    const long long col = (long long)i * m;
    int r = m - 1;
    while (r >= i && a[r + col] == 0.0) --r;
    return r - i + 1;
}

// DLARF('L'): H(i) applied to the trailing columns; a(i,i) holds 1 for the call.
template <int N>
NLB_DEV void rt_reflect_trailing(const WsArr& a, int m, int i, double tau) {
    if (tau == 0.0) return;
    const int lastv = rt_lastv(a, m, i);
    if (lastv <= 0) return;
    const long long vi = (long long)i * m;
    int lastc = 0;                                    // ILADLC: the last trailing column with a non-zero in those rows
    for (int c = N - 1; c > i && lastc == 0; --c) {
        const long long cc = (long long)c * m;
        for (int r = i; r < i + lastv; ++r)
            if (a[r + cc] != 0.0) { lastc = c - i; break; }
    }
    if (lastc <= 0) return;
    for (int c = i + 1; c <= i + lastc; ++c) {
        const long long cc = (long long)c * m;
        double temp = 0.0;
        for (int r = i; r < i + lastv; ++r) temp += a[r + cc] * a[r + vi];
        const double w = 0.0 + 1.0 * temp;
        if (w != 0.0) {
            const double t = (-tau) * w;
            for (int r = i; r < i + lastv; ++r) a[r + cc] = a[r + cc] + a[r + vi] * t;
        }
    }
}

// DLARF('L') with one column: H(i) applied to c(i:m-1); a(i,i) holds 1 for the call.
NLB_DEV void rt_reflect_vec(const WsArr& a, int m, int i, double tau, const WsArr& c) {
    if (tau == 0.0) return;
    const int lastv = rt_lastv(a, m, i);
    if (lastv <= 0) return;
    const long long vi = (long long)i * m;
    bool any = false;
    for (int r = i; r < i + lastv && !any; ++r) any = c[r] != 0.0;
    if (!any) return;
    double temp = 0.0;
    for (int r = i; r < i + lastv; ++r) temp += c[r] * a[r + vi];
    const double w = 0.0 + 1.0 * temp;
    if (w != 0.0) {
        const double t = (-tau) * w;
        for (int r = i; r < i + lastv; ++r) c[r] = c[r] + a[r + vi] * t;
    }
}

// DGEQR2 in place
template <int N>
NLB_DEV void rt_qr_factor(const WsArr& a, int m, double (&tau)[N]) {
#pragma unroll 1
    for (int i = 0; i < N; ++i) {
        const double t = rt_make_reflector(a, m, i);
        vset(tau, i, t);
        if (i < N - 1) {
            const long long d = i + (long long)i * m;
            const double aii = a[d];
            a[d] = 1.0;
            rt_reflect_trailing<N>(a, m, i, t);
            a[d] = aii;
        }
    }
}

// solve_qr(qr, tau, f): rhs := Q^T f by DORM2R('L','T'), then DTRSV('U','N','N') on its first N entries.
template <int N>
NLB_DEV void rt_solve_qr(const WsArr& qr, int m, const double (&tau)[N], const WsArr& f, const WsArr& rhs, double (&u)[N]) {
    for (int r = 0; r < m; ++r) rhs[r] = f[r];
#pragma unroll 1
    for (int i = 0; i < N; ++i) {
        const long long d = i + (long long)i * m;
        const double aii = qr[d];
        qr[d] = 1.0;
        rt_reflect_vec(qr, m, i, vget(tau, i), rhs);
        qr[d] = aii;
    }
#pragma unroll
    for (int j = 0; j < N; ++j) u[j] = rhs[j];
#pragma unroll
    for (int j = N - 1; j >= 0; --j) {
        if (u[j] != 0.0) {
            u[j] = u[j] / qr[j + (long long)j * m];
            const double temp = u[j];
#pragma unroll
            for (int i = j - 1; i >= 0; --i) u[i] = u[i] - temp * qr[i + (long long)j * m];
        }
    }
}

// DGEMV('T'), alpha = 1, beta = 0
template <int N>
NLB_DEV void rt_gemv_t(const WsArr& a, int m, const WsArr& x, double (&y)[N]) {
#pragma unroll 1
    for (int j = 0; j < N; ++j) {
        const long long cj = (long long)j * m;
        double temp = 0.0;
        for (int i = 0; i < m; ++i) temp += a[i + cj] * x[i];
        vset(y, j, 0.0 + 1.0 * temp);
    }
}

// DGEMV('N'), alpha = 1, beta = 0, columns outermost
template <int N>
NLB_DEV void rt_gemv_n(const WsArr& a, int m, const double (&x)[N], const WsArr& y) {
    for (int i = 0; i < m; ++i) y[i] = 0.0;
#pragma unroll 1
    for (int j = 0; j < N; ++j) {
        const long long cj = (long long)j * m;
        const double temp = 1.0 * vget(x, j);
        for (int i = 0; i < m; ++i) y[i] = y[i] + temp * a[i + cj];
    }
}

NLB_DEV double rt_dot_m(const WsArr& a, const WsArr& b, int m) {
    double s = 0.0;
    for (int i = 0; i < m; ++i) s += a[i] * b[i];
    return s;
}

NLB_DEV double rt_norm2_m(const WsArr& v, int m) {
    Norm2 acc;
    for (int i = 0; i < m; ++i) acc.add(v[i]);
    return acc.value();
}

NLB_DEV bool rt_is_finite_m(const WsArr& v, int m) {
    const double huge = 1.7976931348623157e+308;
    bool ok = true;
    for (int i = 0; i < m; ++i) ok = ok && (v[i] == v[i]) && (fabs(v[i]) != huge);
    return ok;
}

template <class F>
NLB_DEV void rt_eval(const double (&x)[F::N], const WsArr& f, int m, const double* __restrict__ y, long long B,
                     const double* __restrict__ t) {
    for (int i = 0; i < m; ++i) f[i] = F::residual(x, __ldg(t + i), __ldg(y + (long long)i * B));
}

// dogleg (:1301-1403): same statements as tps_cls.cuh's, m-sized operands in the workspace
template <int N>
NLB_DEV void rt_dogleg(double delta, const double (&x)[N], const WsArr& f, const WsArr& jac, const WsArr& qr, int m,
                       const double (&tau)[N], const double (&s)[N], const DevCls& o, double (&p)[N], double (&g)[N],
                       const WsArr& Jp, const WsArr& rhs, double& prered) {
    double pgn[N], u[N];
    rt_gemv_t<N>(jac, m, f, g);
    rt_solve_qr<N>(qr, m, tau, f, rhs, u);
#pragma unroll
    for (int i = 0; i < N; ++i) pgn[i] = -u[i];
    const double pgnnorm = scaled_norm(pgn, s);
    if (pgnnorm > delta) {
        double psd[N];
        rt_gemv_n<N>(jac, m, g, Jp);
        const double c1 = dot_vec(g, g);
        const double c2 = rt_dot_m(Jp, Jp, m);
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
    rt_gemv_n<N>(jac, m, p, Jp);
    const double c1 = dot_vec(g, p);
    const double c2 = 0.5 * rt_dot_m(Jp, Jp, m);
    prered = -c1 - c2;
}

// Doubles of workspace per thread: jac, qr (m x N each), fvec, fnew, Jp, rhs (m each).
template <int N>
constexpr long long cls_rt_ws_doubles(long long m) { return 2 * m * N + 4 * m; }

// One thread per system, grid-stride over the batch.  ws: cls_rt_ws_doubles(m) x (threads of the grid) doubles.
template <class F>
__global__ void __launch_bounds__(128)
cls_rt_kernel(DevParams prm, DevCls o, long long nsys, long long B, int m, double* __restrict__ xg,
              double* __restrict__ fg, const double* __restrict__ sys, const double* __restrict__ shared,
              nlb_iteration_behavior* __restrict__ ibg, int32_t* __restrict__ statusg, double* __restrict__ ws) {
    constexpr int N = F::N;
    static_assert(N <= CLS_MAX_N, "limit arrays are passed by value");
    const long long T = (long long)gridDim.x * blockDim.x;
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long mn = (long long)m * N;
    const WsArr jac{ws + tid, T}, qr{ws + mn * T + tid, T};
    const WsArr fvec{ws + 2 * mn * T + tid, T}, fnew{ws + (2 * mn + m) * T + tid, T};
    const WsArr Jp{ws + (2 * mn + 2 * m) * T + tid, T}, rhs{ws + (2 * mn + 3 * m) * T + tid, T};
    const double delta_max = 1.0e3, eta = 1.0e-1, ls_cl = 1.0e-4, ls_beta = 0.5;
    const int ls_max_iter = 10;
    const double eps = 0x1p-26;

    for (long long b = tid; b < nsys; b += T) {
        const double* y = sys + b;
        double x[N];
#pragma unroll
        for (int j = 0; j < N; ++j) x[j] = xg[j * B + b];
        bool xcnvrg = false, fcnvrg = false, gcnvrg = false, converged = false;
        int iter = 0, neval = 0, njac = 0;
        // before the outer loop (:1038-1049)
        apply_limits(x, o);
        rt_eval<F>(x, fvec, m, y, B, shared);
        double fnorm = rt_norm2_m(fvec, m);
        bool running = true;
        double delta = o.radius;
        if (!is_finite_array(x) || !rt_is_finite_m(fvec, m)) { converged = true; running = false; }
        else { neval = 1; iter = 1; }

        while (running) {
            double tau[N], sc[N], g[N], p[N], xnew[N];
            // forward-difference Jacobian (vfh_jac_fcn :262-275), f(x) reused
#pragma unroll 1
            for (int j = 0; j < N; ++j) {
                const double temp = vget(x, j);
                double h = eps * fabs(temp);
                if (h == 0.0) h = eps;
                vset(x, j, temp + h);
                const long long cj = (long long)j * m;
                for (int i = 0; i < m; ++i)
                    jac[i + cj] = (F::residual(x, __ldg(shared + i), __ldg(y + (long long)i * B)) - fvec[i]) / h;
                vset(x, j, temp);
            }
            ++njac;
            for (long long e = 0; e < mn; ++e) qr[e] = jac[e];
            rt_qr_factor<N>(qr, m, tau);
            coleman_li_scaling(x, o, sc);
            double prered;
            rt_dogleg<N>(delta, x, fvec, jac, qr, m, tau, sc, o, p, g, Jp, rhs, prered);
            const double xnorm = scaled_norm(p, sc);
            const double gnorm = norm2_vec(g);
#pragma unroll
            for (int i = 0; i < N; ++i) xnew[i] = x[i] + p[i];

            rt_eval<F>(xnew, fnew, m, y, B, shared);
            double fnewnorm = rt_norm2_m(fnew, m);
            ++neval;

            const double actred = 0.5 * (fnorm * fnorm - fnewnorm * fnewnorm);
            double rho;
            if (prered > 0.0 && actred >= 0.0) rho = actred / prered;
            else rho = 0.0;

            if (rho < 0.25) {
                delta = nl_max(0.25, 1.0e-12);
            } else if (rho > 0.75 && fabs(xnorm - delta) < 1.0e-12 * delta) {
                delta = nl_min(2.0 * delta, delta_max);
            }

            if (rho > eta && fnewnorm <= fnorm) {
#pragma unroll
                for (int i = 0; i < N; ++i) x[i] = xnew[i];
                apply_limits(x, o);
                for (int i = 0; i < m; ++i) fvec[i] = fnew[i];
                fnorm = fnewnorm;
                ++iter;
            } else {
                const double dderiv = dot_vec(g, p);
                if (dderiv >= 0.0) {
                    delta = nl_max(0.5 * delta, 1.0e-12);
                } else {
                    double stepscale = o.scaling;
                    bool accepted = false;
                    for (int k = 1; k <= ls_max_iter; ++k) {
#pragma unroll
                        for (int i = 0; i < N; ++i) xnew[i] = x[i] + stepscale * p[i];
                        apply_limits(xnew, o);
                        rt_eval<F>(xnew, fnew, m, y, B, shared);
                        ++neval;
                        fnewnorm = rt_norm2_m(fnew, m);
                        if (fnewnorm <= fnorm + ls_cl * stepscale * dderiv) {
#pragma unroll
                            for (int i = 0; i < N; ++i) x[i] = xnew[i];
                            for (int i = 0; i < m; ++i) fvec[i] = fnew[i];
                            fnorm = fnewnorm;
                            ++iter;
                            delta = nl_max(stepscale * xnorm, 1.0e-12);
                            accepted = true;
                            break;
                        }
                        stepscale = stepscale * ls_beta;
                    }
                    if (!accepted) delta = nl_max(0.5 * delta, 1.0e-12);
                }
            }

            if (!is_finite_array(x) || !rt_is_finite_m(fvec, m)) break;
            if (xnorm <= prm.var_tol) { converged = true; xcnvrg = true; break; }
            if (fabs(actred) <= prm.fcn_tol && fabs(prered) <= prm.fcn_tol && 0.5 * rho <= 1.0) {
                converged = true; fcnvrg = true; break;
            }
            if (gnorm <= prm.grad_tol) { converged = true; gcnvrg = true; break; }
            if (neval >= prm.max_fcn_evals) break;
        }

#pragma unroll
        for (int j = 0; j < N; ++j) xg[j * B + b] = x[j];
        for (int i = 0; i < m; ++i) fg[(long long)i * B + b] = fvec[i];
        if (ibg) {
            nlb_iteration_behavior r;
            r.iter_count = iter; r.fcn_count = neval; r.jacobian_count = njac; r.gradient_count = 0;
            r.converge_on_fcn = fcnvrg; r.converge_on_chng = xcnvrg; r.converge_on_zero_diff = gcnvrg;
            ibg[b] = r;
        }
        if (statusg) statusg[b] = converged ? 0 : NLB_CONVERGENCE_ERROR;
    }
}

}  // namespace nlb
