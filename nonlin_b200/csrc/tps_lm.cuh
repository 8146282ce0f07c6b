// tps_lm.cuh — thread-per-system Levenberg-Marquardt.
//
// Behaviour of the reference's least_squares_solver%solve, reproduced operation for
// operation so that iteration / function / Jacobian counts match it:
//   lm_solve    lss_solve   src/nonlin_least_squares.f90:118-391
//   lm_par      lmpar       src/nonlin_least_squares.f90:394-566
//   lm_factor   lmfactor    src/nonlin_least_squares.f90:569-667
//   lm_qrsolve  lmsolve     src/nonlin_least_squares.f90:670-791
// Layout: the M x N Jacobian is a per-thread column-major array (local memory for tall
// systems, registers when M*N <= 8); the N-vectors stay in registers (loops over N are
// unrolled, run-time indices go through compare/select chains); M-vectors are per-thread
// arrays walked by rolled loops.
#pragma once
#include "tps_common.cuh"

namespace nlb {

// Loops over the m equations: fully unrolled for small systems (registers), rolled for tall ones (measured on
// B200: unrolling the 21-row loops 7-fold raises the kernel to 206 registers and costs 30 %).
#define NLB_UNROLL_M _Pragma("unroll(M <= 8 ? M : 1)")

// Pivoted Householder QR in place (MINPACK QRFAC lineage).  ipvt is 0-based.
template <int M, int N, class A>
NLB_DEV void lm_factor(A& a, int (&ipvt)[N], double (&rdiag)[N], double (&acnorm)[N], double (&wa)[N]) {
    constexpr int MINMN = M < N ? M : N;
    const double epsmch = 0x1p-52;
#pragma unroll
    for (int j = 0; j < N; ++j) {
        Norm2 acc;
        NLB_UNROLL_M
        for (int i = 0; i < M; ++i) acc.add(a[i + j * M]);
        acnorm[j] = acc.value();
        rdiag[j] = acnorm[j];
        wa[j] = rdiag[j];
        ipvt[j] = j;
    }
#pragma unroll
    for (int j = 0; j < MINMN; ++j) {
        // bring the column of largest (down-dated) norm into the pivot position; first maximum wins
        int kmax = j;
        double rmax = rdiag[j];
#pragma unroll
        for (int k = j + 1; k < N; ++k) {
            if (rdiag[k] > rmax) { rmax = rdiag[k]; kmax = k; }
        }
        if (kmax != j) {
            if constexpr (M * N <= 8) {
#pragma unroll
                for (int i = 0; i < M; ++i) {
                    const double t = a[i + j * M];
                    const double u = vget(a, i + kmax * M);
                    a[i + j * M] = u;
                    vset(a, i + kmax * M, t);
                }
            } else {
                for (int i = 0; i < M; ++i) {
                    const double t = a[i + j * M];
                    a[i + j * M] = a[i + kmax * M];
                    a[i + kmax * M] = t;
                }
            }
            vset(rdiag, kmax, rdiag[j]);
            vset(wa, kmax, wa[j]);
            const int k = ipvt[j];
            ipvt[j] = iget(ipvt, kmax);
            iset(ipvt, kmax, k);
        }
        // Householder vector that reduces column j to a multiple of e_j
        Norm2 acc;
        NLB_UNROLL_M
        for (int i = j; i < M; ++i) acc.add(a[i + j * M]);
        double ajnorm = acc.value();
        if (ajnorm != 0.0) {
            if (a[j + j * M] < 0.0) ajnorm = -ajnorm;
            NLB_UNROLL_M
            for (int i = j; i < M; ++i) a[i + j * M] = a[i + j * M] / ajnorm;
            a[j + j * M] = a[j + j * M] + 1.0;
            // apply it to the remaining columns and down-date their norms
#pragma unroll
            for (int k = j + 1; k < N; ++k) {
                double sm = 0.0;
                NLB_UNROLL_M
                for (int i = j; i < M; ++i) sm += a[i + j * M] * a[i + k * M];
                double temp = sm / a[j + j * M];
                NLB_UNROLL_M
                for (int i = j; i < M; ++i) a[i + k * M] = a[i + k * M] - temp * a[i + j * M];
                if (rdiag[k] == 0.0) continue;
                temp = a[j + k * M] / rdiag[k];
                rdiag[k] = rdiag[k] * sqrt(nl_max(0.0, 1.0 - temp * temp));
                const double q = rdiag[k] / wa[k];
                if (0.05 * (q * q) > epsmch) continue;
                Norm2 acc2;
                NLB_UNROLL_M
                for (int i = j + 1; i < M; ++i) acc2.add(a[i + k * M]);
                rdiag[k] = acc2.value();
                wa[k] = rdiag[k];
            }
        }
        rdiag[j] = -ajnorm;
    }
}

// Solve R z = Q^T b with the rows sqrt(par) D appended, by Givens elimination (MINPACK QRSOLV
// lineage).  r = leading N x N block of the factored Jacobian `a` (leading dimension M); the
// strict lower triangle is overwritten with S^T, the upper triangle and diagonal are kept.
// `wa` is the caller's work vector: only its first N entries are touched.
template <int M, int N, int WN, class A>
NLB_DEV void lm_qrsolve(A& a, const int (&ipvt)[N], const double (&diag)[N], const double (&qtb)[N],
                        double (&x)[N], double (&sdiag)[N], double (&wa)[WN]) {
#pragma unroll
    for (int j = 0; j < N; ++j) {
#pragma unroll
        for (int i = j; i < N; ++i) a[i + j * M] = a[j + i * M];
        x[j] = a[j + j * M];
        wa[j] = qtb[j];
    }
#pragma unroll
    for (int j = 0; j < N; ++j) {
        const double dl = vget(diag, ipvt[j]);
        if (dl != 0.0) {
#pragma unroll
            for (int k = j; k < N; ++k) sdiag[k] = 0.0;
            sdiag[j] = dl;
            double qtbpj = 0.0;
#pragma unroll
            for (int k = j; k < N; ++k) {
                if (sdiag[k] == 0.0) continue;
                double cs, sn;
                const double rkk = a[k + k * M];
                {   // both cases of the reference's rotation share one quotient, one sqrt and one division
                    const bool small_r = fabs(rkk) < fabs(sdiag[k]);
                    const double q = (small_r ? rkk : sdiag[k]) / (small_r ? sdiag[k] : rkk);   // ctan or tan
                    const double v = 0.5 / sqrt(0.25 + 0.25 * (q * q));
                    const double vq = v * q;
                    sn = small_r ? v : vq;
                    cs = small_r ? vq : v;
                }
                a[k + k * M] = cs * rkk + sn * sdiag[k];
                double temp = cs * wa[k] + sn * qtbpj;
                qtbpj = -sn * wa[k] + cs * qtbpj;
                wa[k] = temp;
#pragma unroll
                for (int i = k + 1; i < N; ++i) {
                    temp = cs * a[i + k * M] + sn * sdiag[i];
                    sdiag[i] = -sn * a[i + k * M] + cs * sdiag[i];
                    a[i + k * M] = temp;
                }
            }
        }
        sdiag[j] = a[j + j * M];
        a[j + j * M] = x[j];
    }
    // back substitution; singular tail -> zero (least-squares solution)
    int nsing = N;
#pragma unroll
    for (int j = 0; j < N; ++j) {
        if (sdiag[j] == 0.0 && nsing == N) nsing = j;
        if (nsing < N) wa[j] = 0.0;
    }
#pragma unroll
    for (int j = N - 1; j >= 0; --j) {
        if (j < nsing) {
            double sm = 0.0;
#pragma unroll
            for (int i = j + 1; i < N; ++i)
                if (i < nsing) sm += a[i + j * M] * wa[i];
            wa[j] = (wa[j] - sm) / sdiag[j];
        }
    }
#pragma unroll
    for (int j = 0; j < N; ++j) vset(x, ipvt[j], wa[j]);
}

// Levenberg-Marquardt parameter (MINPACK LMPAR lineage) with the reference's two departures:
// the Newton correction subtracts r(1:n,j)*temp from the WHOLE vector (:552), and inside the
// iteration dxnorm is the norm of all M entries of the work array (:531), whose tail
// n+1..m still holds Q^T f (first pass of an outer iteration) or the last trial residual.
template <int M, int N, class A>
NLB_DEV void lm_par(A& a, const int (&ipvt)[N], const double (&diag)[N], const double (&qtb)[N],
                    double delta, double& par, double (&x)[N], double (&sdiag)[N], double (&wa1)[N],
                    double (&wa2)[M]) {
    const double dwarf = 0x1p-1022;
    int nsing = N;
#pragma unroll
    for (int j = 0; j < N; ++j) {
        wa1[j] = qtb[j];
        if (a[j + j * M] == 0.0 && nsing == N) nsing = j;
        if (nsing < N) wa1[j] = 0.0;
    }
#pragma unroll
    for (int j = N - 1; j >= 0; --j) {
        if (j < nsing) {
            wa1[j] = wa1[j] / a[j + j * M];
            const double temp = wa1[j];
#pragma unroll
            for (int i = 0; i < j; ++i) wa1[i] = wa1[i] - a[i + j * M] * temp;
        }
    }
#pragma unroll
    for (int j = 0; j < N; ++j) vset(x, ipvt[j], wa1[j]);

    int iter = 0;
#pragma unroll
    for (int j = 0; j < N; ++j) wa2[j] = diag[j] * x[j];
    double dxnorm;
    {
        Norm2 acc;
#pragma unroll
        for (int j = 0; j < N; ++j) acc.add(wa2[j]);
        dxnorm = acc.value();
    }
    double fp = dxnorm - delta;
    if (fp <= 0.1 * delta) {
        par = 0.0;
        return;
    }

    double parl = 0.0;
    if (nsing == N) {
#pragma unroll
        for (int j = 0; j < N; ++j) {
            const int l = ipvt[j];
            wa1[j] = vget(diag, l) * (vget(wa2, l) / dxnorm);
        }
#pragma unroll
        for (int j = 0; j < N; ++j) {
            double sm = 0.0;
#pragma unroll
            for (int i = 0; i < j; ++i) sm += a[i + j * M] * wa1[i];
            wa1[j] = (wa1[j] - sm) / a[j + j * M];
        }
        const double temp = norm2_vec(wa1);
        parl = ((fp / delta) / temp) / temp;
    }

#pragma unroll
    for (int j = 0; j < N; ++j) {
        double sm = 0.0;
#pragma unroll
        for (int i = 0; i <= j; ++i) sm += a[i + j * M] * qtb[i];
        wa1[j] = sm / vget(diag, ipvt[j]);
    }
    const double gnorm = norm2_vec(wa1);
    double paru = gnorm / delta;
    if (paru == 0.0) paru = dwarf / nl_min(delta, 0.1);

    par = nl_max(par, parl);
    par = nl_min(par, paru);
    if (par == 0.0) par = gnorm / dxnorm;

    for (;;) {
        ++iter;
        if (par == 0.0) par = nl_max(dwarf, 1.0e-3 * paru);
        double temp = sqrt(par);
#pragma unroll
        for (int j = 0; j < N; ++j) wa1[j] = temp * diag[j];
        lm_qrsolve<M, N, M>(a, ipvt, wa1, qtb, x, sdiag, wa2);
#pragma unroll
        for (int j = 0; j < N; ++j) wa2[j] = diag[j] * x[j];
        {
            Norm2 acc;
            NLB_UNROLL_M
            for (int i = 0; i < M; ++i) acc.add(wa2[i]);
            dxnorm = acc.value();
        }
        temp = fp;
        fp = dxnorm - delta;

        if (fabs(fp) <= 0.1 * delta || (parl == 0.0 && fp <= temp && temp < 0.0) || iter == 10) break;

#pragma unroll
        for (int j = 0; j < N; ++j) {
            const int l = ipvt[j];
            wa1[j] = vget(diag, l) * (vget(wa2, l) / dxnorm);
        }
#pragma unroll
        for (int j = 0; j < N; ++j) {
            wa1[j] = wa1[j] / sdiag[j];
            temp = wa1[j];
            if (j + 1 < N) {
#pragma unroll
                for (int i = 0; i < N; ++i) wa1[i] = wa1[i] - a[i + j * M] * temp;
            }
        }
        temp = norm2_vec(wa1);
        const double parc = ((fp / delta) / temp) / temp;

        if (fp > 0.0) parl = nl_max(parl, par);
        if (fp < 0.0) paru = nl_min(paru, par);
        par = nl_max(parl, par + parc);
    }
}

// `jac` is the caller's storage for the m x n Jacobian / its QR factors: a per-thread array (registers or local
// memory) or a StridedMat view of shared memory - anything indexable as jac[i + j * M].
template <class F, class A>
NLB_DEV void tps_lm_solve(const DevParams& p, const SysCtx& c, double (&x)[F::N], double (&fvec)[F::M],
                          SolveStats& st, A& jac) {
    constexpr int M = F::M, N = F::N;
    static_assert(M >= N, "least squares needs m >= n (src/nonlin_least_squares.f90:189)");
    const double eps = 0x1p-52;
    const double ftol = p.fcn_tol, xtol = p.var_tol, gtol = p.grad_tol, fac = p.lm_factor;
    const int maxeval = p.max_fcn_evals;
    const bool analytic = p.use_analytic_jacobian != 0;
    bool xcnvrg = false, fcnvrg = false, gcnvrg = false;
    int flag = 0;

    double wa4[M];
    double diag[N], qtf[N], wa1[N], wa2[N], wa3[N];
    int jpvt[N];

    F::eval(x, fvec, c);
    int neval = 1, njac = 0, iter = 1;
    double fnorm;
    {
        Norm2 acc;
        NLB_UNROLL_M
        for (int i = 0; i < M; ++i) acc.add(fvec[i]);
        fnorm = acc.value();
    }
    double par = 0.0, xnorm = 0.0, delta = 0.0, gnorm = 0.0, temp = 0.0;

    for (;;) {
        fd_jacobian<F>(x, jac, fvec, wa4, c, analytic);
        ++njac;
        lm_factor<M, N>(jac, jpvt, wa1, wa2, wa3);

        if (iter == 1) {
#pragma unroll
            for (int j = 0; j < N; ++j) {
                diag[j] = wa2[j];
                if (wa2[j] == 0.0) diag[j] = 1.0;
            }
#pragma unroll
            for (int j = 0; j < N; ++j) wa3[j] = diag[j] * x[j];
            xnorm = norm2_vec(wa3);
            delta = fac * xnorm;
            if (delta == 0.0) delta = fac;
        }

        // qtf = first n components of Q^T fvec; the tail stays in wa4
        NLB_UNROLL_M
        for (int i = 0; i < M; ++i) wa4[i] = fvec[i];
#pragma unroll
        for (int j = 0; j < N; ++j) {
            if (jac[j + j * M] != 0.0) {
                double sm = 0.0;
                NLB_UNROLL_M
                for (int i = j; i < M; ++i) sm += jac[i + j * M] * wa4[i];
                temp = -sm / jac[j + j * M];
                NLB_UNROLL_M
                for (int i = j; i < M; ++i) wa4[i] = wa4[i] + jac[i + j * M] * temp;
            }
            jac[j + j * M] = wa1[j];
            qtf[j] = wa4[j];
        }

        // scaled gradient norm
        gnorm = 0.0;
        if (fnorm != 0.0) {
#pragma unroll
            for (int j = 0; j < N; ++j) {
                const double acn = vget(wa2, jpvt[j]);
                if (acn == 0.0) continue;
                double sm = 0.0;
#pragma unroll
                for (int i = 0; i <= j; ++i) sm += jac[i + j * M] * (qtf[i] / fnorm);
                gnorm = nl_max(gnorm, fabs(sm / acn));
            }
        }
        if (gnorm <= gtol) { gcnvrg = true; break; }

#pragma unroll
        for (int j = 0; j < N; ++j) diag[j] = nl_max(diag[j], wa2[j]);

        for (;;) {
            lm_par<M, N>(jac, jpvt, diag, qtf, delta, par, wa1, wa2, wa3, wa4);

#pragma unroll
            for (int j = 0; j < N; ++j) {
                wa1[j] = -wa1[j];
                wa2[j] = x[j] + wa1[j];
                wa3[j] = diag[j] * wa1[j];
            }
            const double pnorm = norm2_vec(wa3);
            if (iter == 1) delta = nl_min(delta, pnorm);

            F::eval(wa2, wa4, c);
            ++neval;
            double fnorm1;
            {
                Norm2 acc;
                NLB_UNROLL_M
                for (int i = 0; i < M; ++i) acc.add(wa4[i]);
                fnorm1 = acc.value();
            }

            double actred = -1.0;
            if (0.1 * fnorm1 < fnorm) { const double q = fnorm1 / fnorm; actred = 1.0 - q * q; }

#pragma unroll
            for (int j = 0; j < N; ++j) {
                wa3[j] = 0.0;
                temp = vget(wa1, jpvt[j]);
#pragma unroll
                for (int i = 0; i <= j; ++i) wa3[i] = wa3[i] + jac[i + j * M] * temp;
            }
            const double temp1 = norm2_vec(wa3) / fnorm;
            const double temp2 = (sqrt(par) * pnorm) / fnorm;
            const double prered = temp1 * temp1 + temp2 * temp2 / 0.5;
            const double dirder = -(temp1 * temp1 + temp2 * temp2);

            double ratio = 0.0;
            if (prered != 0.0) ratio = actred / prered;

            if (ratio <= 0.25) {
                if (actred >= 0.0) temp = 0.5;
                if (actred < 0.0) temp = 0.5 * dirder / (dirder + 0.5 * actred);
                if (0.1 * fnorm1 >= fnorm || temp < 0.1) temp = 0.1;
                delta = temp * nl_min(delta, pnorm / 0.1);
                par = par / temp;
            } else if (!(par != 0.0 && ratio < 0.75)) {
                delta = pnorm / 0.5;
                par = 0.5 * par;
            }

            if (ratio >= 1.0e-4) {
#pragma unroll
                for (int j = 0; j < N; ++j) {
                    x[j] = wa2[j];
                    wa2[j] = diag[j] * x[j];
                }
                NLB_UNROLL_M
                for (int i = 0; i < M; ++i) fvec[i] = wa4[i];
                xnorm = norm2_vec(wa2);
                fnorm = fnorm1;
                ++iter;
            }

            if (fabs(actred) <= ftol && prered <= ftol && 0.5 * ratio <= 1.0) fcnvrg = true;
            if (delta <= xtol * xnorm) xcnvrg = true;
            if (fcnvrg || xcnvrg) break;

            if (neval >= maxeval) flag = NLB_CONVERGENCE_ERROR;
            if (fabs(actred) <= eps && prered <= eps && 0.5 * ratio <= 1.0) flag = NLB_TOLERANCE_TOO_SMALL_ERROR;
            if (delta <= eps * xnorm) flag = NLB_TOLERANCE_TOO_SMALL_ERROR;
            if (gnorm <= eps) flag = NLB_TOLERANCE_TOO_SMALL_ERROR;
            if (flag != 0) break;

            if (ratio >= 1.0e-4) break;
        }
        if (fcnvrg || xcnvrg || gcnvrg || flag != 0) break;
    }
    st.iter = iter;
    st.nfev = neval;
    st.njac = njac;
    st.cf = fcnvrg;
    st.cx = xcnvrg;
    st.cg = gcnvrg;
    // every non-zero flag ends in `error stop NL_CONVERGENCE_ERROR` (:388-390)
    st.status = flag != 0 ? NLB_CONVERGENCE_ERROR : NLB_NO_ERROR;
}

template <class F>
NLB_DEV void tps_lm_solve(const DevParams& p, const SysCtx& c, double (&x)[F::N], double (&fvec)[F::M],
                          SolveStats& st) {
    double jac[F::M * F::N];
    tps_lm_solve<F>(p, c, x, fvec, st, jac);
}

}  // namespace nlb
