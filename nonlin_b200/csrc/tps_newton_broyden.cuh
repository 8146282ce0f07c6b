// tps_newton_broyden.cuh — thread-per-system Newton (LU) and Broyden quasi-Newton (QR with
// rank-1 updates), both with the backtracking line search.
//
// Behaviour reproduced from the reference:
//   tps_newton_solve   ns_solve   src/nonlin_solve.f90:452-638
//   tps_broyden_solve  qns_solve  src/nonlin_solve.f90:156-425
// Where the reference executes `error stop <code>`, the thread records <code> in its status
// and returns with x / fvec / counters as they were at that point.
#pragma once
#include "tps_common.cuh"
#include "tps_dense.cuh"

namespace nlb {

template <class F>
NLB_DEV void tps_newton_solve(const DevParams& p, const SysCtx& c, double (&x)[F::N], double (&fvec)[F::M],
                              SolveStats& st) {
    constexpr int N = F::N;
    static_assert(F::M == F::N, "Newton needs a square system (src/nonlin_solve.f90:519)");
    const double ftol = p.fcn_tol, xtol = p.var_tol, gtol = p.grad_tol;
    const bool analytic = p.use_analytic_jacobian != 0;
    bool xcnvrg = false, fcnvrg = false, gcnvrg = false;
    int neval = 0, iter = 0, njac = 0, flag = 0, status = NLB_NO_ERROR;

    double jac[N * N], dir[N], grad[N], xold[N], wrk[N];
    int ipvt[N];

    // is the starting point already a root?
    F::eval(x, fvec, c);
    double f = 0.5 * dot_vec(fvec, fvec);
    ++neval;
    double test = 0.0;
#pragma unroll
    for (int i = 0; i < N; ++i) test = nl_max(fabs(fvec[i]), test);
    if (test < ftol) fcnvrg = true;

    if (!fcnvrg) {
        const double stpmax = 100.0 * nl_max(norm2_vec(x), (double)N);
        for (;;) {
            ++iter;
            fd_jacobian<F>(x, jac, fvec, wrk, c, analytic);
            ++njac;
            // gradient of 0.5 |F|^2 = J^T F (used by the line search and the convergence test)
#pragma unroll
            for (int i = 0; i < N; ++i) {
                double s = 0.0;
#pragma unroll
                for (int l = 0; l < N; ++l) s += jac[l + i * N] * fvec[l];
                grad[i] = s;
            }
            dgetrf<N>(jac, ipvt);
#pragma unroll
            for (int i = 0; i < N; ++i) xold[i] = x[i];
            const double fold = f;
#pragma unroll
            for (int i = 0; i < N; ++i) dir[i] = -fvec[i];
            dgetrs<N>(jac, ipvt, dir);

            if (p.use_line_search) {
                // pre-scale compares the SQUARED length with stpmax (:582-583), then the true clamp
                const double temp = dot_vec(dir, dir);
                if (temp > stpmax) {
                    const double s = stpmax / temp;
#pragma unroll
                    for (int i = 0; i < N; ++i) dir[i] = dir[i] * s;
                }
                limit_vector(dir, stpmax);
                int ls_nfev;
                const int ls = line_search<F>(p, c, xold, grad, dir, x, fvec, fold, f, ls_nfev);
                neval += ls_nfev;
                if (ls != NLB_NO_ERROR) { status = ls; break; }
            } else {
#pragma unroll
                for (int i = 0; i < N; ++i) x[i] = x[i] + dir[i];
                F::eval(x, fvec, c);
                f = 0.5 * dot_vec(fvec, fvec);
                ++neval;
            }

            const ConvFlags cv = test_convergence<N, N>(x, xold, fvec, grad, true, xtol, ftol, gtol);
            xcnvrg = cv.cx; fcnvrg = cv.cf; gcnvrg = cv.cg;
            if (cv.c) break;
            if (cv.cg) { status = NLB_SPURIOUS_CONVERGENCE_ERROR; break; }

            if (neval >= p.max_fcn_evals) { flag = 1; break; }
        }
    }
    st.iter = iter; st.nfev = neval; st.njac = njac;
    st.cf = fcnvrg; st.cx = xcnvrg; st.cg = gcnvrg;
    st.status = status != NLB_NO_ERROR ? status : (flag != 0 ? NLB_CONVERGENCE_ERROR : NLB_NO_ERROR);
}

// Work distribution for the persistent ("refill") kernels: lanes that need a system take consecutive
// indices from a global cursor with one atomic per warp.
NLB_DEV long long tps_fetch(unsigned long long* cursor) {
    const unsigned mask = __activemask();
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(mask) - 1;
    unsigned long long base = 0;
    if (lane == leader) base = atomicAdd(cursor, (unsigned long long)__popc(mask));
    base = __shfl_sync(mask, base, leader);
    return (long long)(base + __popc(mask & ((1u << lane) - 1u)));
}

// Persistent Newton: the same arithmetic per system as tps_newton_solve, organised so that one trip
// of the outer loop is ONE Newton iteration of whatever system the lane currently holds.  A lane
// whose system finishes stores it and takes the next one, so the iteration-count spread between
// systems (16...85 on Powell's function) no longer idles lanes.
template <class F>
NLB_DEV void tps_newton_refill(const DevParams& p, long long nsys, long long B, unsigned long long* cursor,
                               double* __restrict__ xg, double* __restrict__ fg, const double* __restrict__ sys,
                               const double* __restrict__ shared, nlb_iteration_behavior* __restrict__ ibg,
                               int32_t* __restrict__ statusg) {
    constexpr int N = F::N;
    static_assert(F::M == F::N, "Newton needs a square system (src/nonlin_solve.f90:519)");
    const double ftol = p.fcn_tol, xtol = p.var_tol, gtol = p.grad_tol;
    const bool analytic = p.use_analytic_jacobian != 0;

    double x[N], fvec[N], jac[N * N], dir[N], grad[N], xold[N], wrk[N];
    int ipvt[N];
    double f = 0.0, stpmax = 0.0;
    int neval = 0, iter = 0, njac = 0, status = NLB_NO_ERROR;
    bool xcnvrg = false, fcnvrg = false, gcnvrg = false;
    long long b = -1;
    bool active = false, more = true;      // more: the queue may still hold systems for this lane

    for (;;) {
        bool finished = false;
        if (!active && more) {
            b = tps_fetch(cursor);
            if (b >= nsys) more = false;
        }
        if (!active && more) {
#pragma unroll
            for (int j = 0; j < N; ++j) x[j] = xg[j * B + b];
            const SysCtx c{sys ? sys + b : nullptr, shared, B, N, N};
            neval = 0; iter = 0; njac = 0; status = NLB_NO_ERROR;
            xcnvrg = false; fcnvrg = false; gcnvrg = false;
            F::eval(x, fvec, c);
            f = 0.5 * dot_vec(fvec, fvec);
            ++neval;
            double test = 0.0;
#pragma unroll
            for (int i = 0; i < N; ++i) test = nl_max(fabs(fvec[i]), test);
            if (test < ftol) { fcnvrg = true; finished = true; }
            else { stpmax = 100.0 * nl_max(norm2_vec(x), (double)N); active = true; }
        }
        // No lane leaves the loop on its own: the warp re-converges here every trip (lanes that refilled and lanes
        // that did not would otherwise run the iteration body as two half-empty groups) and exits together.
        __syncwarp();
        if (!__any_sync(0xffffffffu, active || finished || more)) break;
        if (active) {
            const SysCtx c{sys ? sys + b : nullptr, shared, B, N, N};
            ++iter;
            fd_jacobian<F>(x, jac, fvec, wrk, c, analytic);
            ++njac;
#pragma unroll
            for (int i = 0; i < N; ++i) {
                double s = 0.0;
#pragma unroll
                for (int l = 0; l < N; ++l) s += jac[l + i * N] * fvec[l];
                grad[i] = s;
            }
            dgetrf<N>(jac, ipvt);
#pragma unroll
            for (int i = 0; i < N; ++i) xold[i] = x[i];
            const double fold = f;
#pragma unroll
            for (int i = 0; i < N; ++i) dir[i] = -fvec[i];
            dgetrs<N>(jac, ipvt, dir);
            bool stop = false;
            if (p.use_line_search) {
                const double temp = dot_vec(dir, dir);
                if (temp > stpmax) {
                    const double s = stpmax / temp;
#pragma unroll
                    for (int i = 0; i < N; ++i) dir[i] = dir[i] * s;
                }
                limit_vector(dir, stpmax);
                int ls_nfev;
                const int ls = line_search<F>(p, c, xold, grad, dir, x, fvec, fold, f, ls_nfev);
                neval += ls_nfev;
                if (ls != NLB_NO_ERROR) { status = ls; stop = true; }
            } else {
#pragma unroll
                for (int i = 0; i < N; ++i) x[i] = x[i] + dir[i];
                F::eval(x, fvec, c);
                f = 0.5 * dot_vec(fvec, fvec);
                ++neval;
            }
            if (!stop) {
                const ConvFlags cv = test_convergence<N, N>(x, xold, fvec, grad, true, xtol, ftol, gtol);
                xcnvrg = cv.cx; fcnvrg = cv.cf; gcnvrg = cv.cg;
                if (cv.c) stop = true;
                else if (cv.cg) { status = NLB_SPURIOUS_CONVERGENCE_ERROR; stop = true; }
                else if (neval >= p.max_fcn_evals) { status = NLB_CONVERGENCE_ERROR; stop = true; }
            }
            if (stop) { finished = true; active = false; }
        }
        if (finished) {
#pragma unroll
            for (int j = 0; j < N; ++j) xg[j * B + b] = x[j];
#pragma unroll
            for (int i = 0; i < N; ++i) fg[i * B + b] = fvec[i];
            if (ibg) {
                nlb_iteration_behavior o;
                o.iter_count = iter; o.fcn_count = neval; o.jacobian_count = njac; o.gradient_count = 0;
                o.converge_on_fcn = fcnvrg; o.converge_on_chng = xcnvrg; o.converge_on_zero_diff = gcnvrg;
                ibg[b] = o;
            }
            if (statusg) statusg[b] = status;
        }
    }
}

template <class F>
NLB_DEV void tps_broyden_solve(const DevParams& p, const SysCtx& c, double (&x)[F::N], double (&fvec)[F::M],
                               SolveStats& st) {
    constexpr int N = F::N;
    static_assert(F::M == F::N, "quasi-Newton needs a square system (src/nonlin_solve.f90:237)");
    const double ftol = p.fcn_tol, xtol = p.var_tol, gtol = p.grad_tol;
    const bool analytic = p.use_analytic_jacobian != 0;
    bool restart = true, xcnvrg = false, fcnvrg = false, gcnvrg = false;
    int neval = 0, iter = 0, njac = 0, flag = 0, jcount = 0, status = NLB_NO_ERROR;

    double b[N * N], q[N * N], r[N * N];
    double df[N], fvold[N], xold[N], dx[N], s[N];

    F::eval(x, fvec, c);
    double f = 0.5 * dot_vec(fvec, fvec);
    ++neval;
    double test = 0.0;
#pragma unroll
    for (int i = 0; i < N; ++i) test = nl_max(fabs(fvec[i]), test);
    if (test < ftol) fcnvrg = true;

    if (!fcnvrg) {
        const double stpmax = 100.0 * nl_max(norm2_vec(x), (double)N);
        double fold = f;
        for (;;) {
            ++iter;
            if (iter > p.max_iter_guard) { flag = 1; break; }
            if (restart) {
                // fresh Jacobian (df is free here and serves as the difference work vector)
                fd_jacobian<F>(x, b, fvec, df, c, analytic);
                ++njac;
                qr_full<N>(b, q, r);
                jcount = 0;
            } else {
                // Broyden: B += s dx^T with s = (df - B dx) / (dx.dx); Q, R updated, not refactored
#pragma unroll
                for (int i = 0; i < N; ++i) df[i] = fvec[i] - fvold[i];
#pragma unroll
                for (int i = 0; i < N; ++i) dx[i] = x[i] - xold[i];
                const double x2 = dot_vec(dx, dx);
#pragma unroll
                for (int i = 0; i < N; ++i) s[i] = 0.0;
#pragma unroll
                for (int j = 0; j < N; ++j) {
#pragma unroll
                    for (int i = 0; i < N; ++i) s[i] = s[i] + b[i + j * N] * dx[j];
                }
#pragma unroll
                for (int i = 0; i < N; ++i) s[i] = df[i] - s[i];
                drscl<N>(x2, s);
                // DGER, alpha = 1
#pragma unroll
                for (int j = 0; j < N; ++j) {
                    if (dx[j] != 0.0) {
                        const double temp = 1.0 * dx[j];
#pragma unroll
                        for (int i = 0; i < N; ++i) b[i + j * N] = b[i + j * N] + s[i] * temp;
                    }
                }
                qr_rank1_update<N>(q, r, s, dx);
                ++jcount;
            }

            gemv_t<N>(1.0, b, fvec, dx);            // gradient B^T F -> dx
#pragma unroll
            for (int i = 0; i < N; ++i) { xold[i] = x[i]; fvold[i] = fvec[i]; }
            fold = f;
            gemv_t<N>(-1.0, q, fvec, df);           // -Q^T F
            trsv_upper<N>(r, df);                   // R step = -Q^T F

            double temp = dot_vec(dx, df);
            if (temp >= 0.0) { restart = true; continue; }   // uphill: recompute the Jacobian (:330-337)

            if (p.use_line_search) {
                temp = dot_vec(df, df);
                if (temp > stpmax) {
                    const double sc = stpmax / temp;
#pragma unroll
                    for (int i = 0; i < N; ++i) df[i] = df[i] * sc;
                }
                limit_vector(df, stpmax);
                int ls_nfev;
                const int ls = line_search<F>(p, c, xold, dx, df, x, fvec, fold, f, ls_nfev);
                neval += ls_nfev;
                if (ls != NLB_NO_ERROR) { status = ls; break; }
            } else {
#pragma unroll
                for (int i = 0; i < N; ++i) x[i] = x[i] + df[i];
                F::eval(x, fvec, c);
                f = 0.5 * dot_vec(fvec, fvec);
                ++neval;
            }

            // the gradient test is never armed here: the line search always reports
            // converge_on_zero_diff = .false. (src/nonlin_linesearch.f90:219,319)
            const ConvFlags cv = test_convergence<N, N>(x, xold, fvec, dx, false, xtol, ftol, gtol);
            xcnvrg = cv.cx; fcnvrg = cv.cf; gcnvrg = cv.cg;
            if (cv.c) break;
            if (cv.cg) {
                if (restart) { status = NLB_SPURIOUS_CONVERGENCE_ERROR; break; }
                restart = true;
            } else {
                restart = (jcount >= p.jacobian_interval);
            }

            if (neval >= p.max_fcn_evals) { flag = 1; break; }
        }
    }
    st.iter = iter; st.nfev = neval; st.njac = njac;
    st.cf = fcnvrg; st.cx = xcnvrg; st.cg = gcnvrg;
    st.status = status != NLB_NO_ERROR ? status : (flag != 0 ? NLB_CONVERGENCE_ERROR : NLB_NO_ERROR);
}

}  // namespace nlb
