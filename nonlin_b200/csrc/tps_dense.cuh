// tps_dense.cuh — per-thread N x N dense kernels behind the Newton and quasi-Newton solvers.
//
// The reference does this algebra in the external `linalg` package (not vendored; call sites
// src/nonlin_solve.f90:286,298,302,303,311,320,325,570,577), which reaches LAPACK's unblocked
// routines for n <= 128 and QRUPDATE's DQR1UP.  These device functions implement the same
// published algorithms with the same operation order (Reference LAPACK 3.12.0 forms):
//   lu_factor / solve_lu      DGETRF (right-looking, reciprocal pivot scaling) / DGETRS
//   qr_factor(q=, r=)         DGEQR2 (DLARFG, DLARF) then DORG2R
//   qr_rank1_update           DQR1UP (DQRTV1, DQRQH, DQROT, DAXPY, DQHQR with DLARTG)
//   rank1_update, recip_mult_array, mtx_mult('T'), solve_triangular_system
//                             DGER, DRSCL, DGEMV('T'), DTRSV('U','N','N')
// Matrices are column-major N x N per-thread arrays; every loop bound is a compile-time
// constant and data-dependent ranges (pivot rows, DLARF's trailing-zero scan) are predicates,
// so for N = 2 everything stays in registers.
#pragma once
#include "nlb_math.cuh"

namespace nlb {

// DGETRF on a[N*N]; ipiv 0-based.  Returns LAPACK info (0, or 1-based index of a zero pivot).
template <int N>
NLB_DEV int dgetrf(double (&a)[N * N], int (&ipiv)[N]) {
    const double sfmin = 0x1p-1022;
    int info = 0;
#pragma unroll
    for (int j = 0; j < N; ++j) {
        int jp = j;
        double dmax = fabs(a[j + j * N]);
#pragma unroll
        for (int i = j + 1; i < N; ++i) {
            const double v = fabs(a[i + j * N]);
            if (v > dmax) { dmax = v; jp = i; }
        }
        ipiv[j] = jp;
        double piv = a[j + j * N];
#pragma unroll
        for (int i = j + 1; i < N; ++i) piv = (i == jp) ? a[i + j * N] : piv;
        if (piv != 0.0) {
            if (jp != j) {
#pragma unroll
                for (int c = 0; c < N; ++c) {
#pragma unroll
                    for (int i = j + 1; i < N; ++i) {
                        if (i == jp) {
                            const double t = a[j + c * N];
                            a[j + c * N] = a[i + c * N];
                            a[i + c * N] = t;
                        }
                    }
                }
            }
            if (j < N - 1) {
                if (fabs(a[j + j * N]) >= sfmin) {
                    const double rp = 1.0 / a[j + j * N];
#pragma unroll
                    for (int i = j + 1; i < N; ++i) a[i + j * N] = rp * a[i + j * N];
                } else {
#pragma unroll
                    for (int i = j + 1; i < N; ++i) a[i + j * N] = a[i + j * N] / a[j + j * N];
                }
            }
        } else if (info == 0) {
            info = j + 1;
        }
        if (j < N - 1) {
#pragma unroll
            for (int c = j + 1; c < N; ++c) {
                if (a[j + c * N] != 0.0) {
                    const double temp = -a[j + c * N];
#pragma unroll
                    for (int i = j + 1; i < N; ++i) a[i + c * N] = a[i + c * N] + a[i + j * N] * temp;
                }
            }
        }
    }
    return info;
}

// DGETRS('N'), one right-hand side, in place.
template <int N>
NLB_DEV void dgetrs(const double (&a)[N * N], const int (&ipiv)[N], double (&b)[N]) {
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const int ip = ipiv[i];
#pragma unroll
        for (int r = i + 1; r < N; ++r) {
            if (r == ip) { const double t = b[i]; b[i] = b[r]; b[r] = t; }
        }
    }
#pragma unroll
    for (int k = 0; k < N; ++k) {
        if (b[k] != 0.0) {
#pragma unroll
            for (int i = k + 1; i < N; ++i) b[i] = b[i] - b[k] * a[i + k * N];
        }
    }
#pragma unroll
    for (int k = N - 1; k >= 0; --k) {
        if (b[k] != 0.0) {
            b[k] = b[k] / a[k + k * N];
#pragma unroll
            for (int i = 0; i < k; ++i) b[i] = b[i] - b[k] * a[i + k * N];
        }
    }
}

// DLARF('L') with v = a(i:N-1, i) (a(i,i) already set to 1) applied to C = a(i:N-1, i+1:N-1).
template <int N>
NLB_DEV void reflect_trailing(double (&a)[N * N], int i, double tau) {
    if (tau == 0.0) return;
    // last non-zero row of v, last non-zero column of C(1:lastv, :)
    int lastv = N - i;
    bool scanning = true;
#pragma unroll
    for (int r = N - 1; r >= 0; --r) {
        if (r >= i && scanning) {
            if (a[r + i * N] == 0.0) lastv = r - i;
            else scanning = false;
        }
    }
    int lastc = 0;
#pragma unroll
    for (int c = 0; c < N; ++c) {
        if (c > i) {
#pragma unroll
            for (int r = 0; r < N; ++r)
                if (r >= i && (r - i) < lastv && a[r + c * N] != 0.0) lastc = c - i;
        }
    }
    if (lastv <= 0 || lastc <= 0) return;
#pragma unroll
    for (int c = 0; c < N; ++c) {
        if (c > i && (c - i) <= lastc) {
            double temp = 0.0;
#pragma unroll
            for (int r = 0; r < N; ++r)
                if (r >= i && (r - i) < lastv) temp += a[r + c * N] * a[r + i * N];
            const double w = 0.0 + 1.0 * temp;      // DGEMV: y = beta*y (0) then y += alpha*temp
            if (w != 0.0) {                          // DGER skips zero entries of y
                const double t = (-tau) * w;
#pragma unroll
                for (int r = 0; r < N; ++r)
                    if (r >= i && (r - i) < lastv) a[r + c * N] = a[r + c * N] + a[r + i * N] * t;
            }
        }
    }
}

// DLARFG on column i of a: alpha = a(i,i), x = a(i+1:N-1, i).
template <int N>
NLB_DEV double make_reflector(double (&a)[N * N], int i) {
    if (N - i <= 1) return 0.0;
    Dnrm2 acc;
#pragma unroll
    for (int r = 0; r < N; ++r)
        if (r > i) acc.add(a[r + i * N]);
    double xnorm = acc.value();
    if (xnorm == 0.0) return 0.0;
    double alpha = a[i + i * N];
    double beta = -nl_sign(dlapy2(alpha, xnorm), alpha);
    const double safmin = 0x1p-969;                 // dlamch('S') / dlamch('E')
    int knt = 0;
    if (fabs(beta) < safmin) {
        const double rsafmn = 1.0 / safmin;
        do {
            ++knt;
#pragma unroll
            for (int r = 0; r < N; ++r)
                if (r > i) a[r + i * N] = rsafmn * a[r + i * N];
            beta = beta * rsafmn;
            alpha = alpha * rsafmn;
        } while (fabs(beta) < safmin && knt < 20);
        Dnrm2 acc2;
#pragma unroll
        for (int r = 0; r < N; ++r)
            if (r > i) acc2.add(a[r + i * N]);
        xnorm = acc2.value();
        beta = -nl_sign(dlapy2(alpha, xnorm), alpha);
    }
    const double tau = (beta - alpha) / beta;
    const double sc = 1.0 / (alpha - beta);
#pragma unroll
    for (int r = 0; r < N; ++r)
        if (r > i) a[r + i * N] = sc * a[r + i * N];
    for (int j = 0; j < knt; ++j) beta = beta * safmin;
    a[i + i * N] = beta;
    return tau;
}

// qr_factor(b, q = q, r = r): DGEQR2 on a copy, R = its upper triangle, Q = DORG2R.
template <int N>
NLB_DEV void qr_full(const double (&b)[N * N], double (&q)[N * N], double (&r)[N * N]) {
    double tau[N];
#pragma unroll
    for (int e = 0; e < N * N; ++e) q[e] = b[e];
#pragma unroll
    for (int i = 0; i < N; ++i) {
        tau[i] = make_reflector<N>(q, i);
        if (i < N - 1) {
            const double aii = q[i + i * N];
            q[i + i * N] = 1.0;
            reflect_trailing<N>(q, i, tau[i]);
            q[i + i * N] = aii;
        }
    }
#pragma unroll
    for (int j = 0; j < N; ++j) {
#pragma unroll
        for (int i = 0; i < N; ++i) r[i + j * N] = (i <= j) ? q[i + j * N] : 0.0;
    }
#pragma unroll
    for (int i = N - 1; i >= 0; --i) {
        if (i < N - 1) {
            q[i + i * N] = 1.0;
            reflect_trailing<N>(q, i, tau[i]);
            const double sc = -tau[i];
#pragma unroll
            for (int l = i + 1; l < N; ++l) q[l + i * N] = sc * q[l + i * N];
        }
        q[i + i * N] = 1.0 - tau[i];
#pragma unroll
        for (int l = 0; l < i; ++l) q[l + i * N] = 0.0;
    }
}

// DQR1UP, full Q: Q R + u v^T -> Q1 R1 in place.
template <int N>
NLB_DEV void qr_rank1_update(double (&q)[N * N], double (&r)[N * N], const double (&u)[N], const double (&v)[N]) {
    double w[N], cs[N], sn[N];
    // w = Q^T u, each entry a sequential dot product
#pragma unroll
    for (int i = 0; i < N; ++i) {
        double s = 0.0;
#pragma unroll
        for (int l = 0; l < N; ++l) s += q[l + i * N] * u[l];
        w[i] = s;
    }
    // DQRTV1: rotations that fold w into its first entry, bottom-up
    {
        double rr = w[N - 1];
#pragma unroll
        for (int i = N - 2; i >= 0; --i) {
            double t;
            dlartg(w[i], rr, cs[i], sn[i], t);
            rr = t;
        }
        w[0] = rr;
    }
    // DQRQH: R -> upper Hessenberg
    if (N > 1) {
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const int ii = (N - 2 < i) ? N - 2 : i;      // 0-based index of the last rotation touching column i
            double t = r[(ii + 1) + i * N];
#pragma unroll
            for (int j = N - 2; j >= 0; --j) {
                if (j <= ii) {
                    r[(j + 1) + i * N] = cs[j] * t - sn[j] * r[j + i * N];
                    t = cs[j] * r[j + i * N] + sn[j] * t;
                }
            }
            r[0 + i * N] = t;
        }
        // DQROT('B'): Q <- Q G^T, last rotation first
#pragma unroll
        for (int i = N - 2; i >= 0; --i) {
#pragma unroll
            for (int l = 0; l < N; ++l) {
                const double t = cs[i] * q[l + i * N] + sn[i] * q[l + (i + 1) * N];
                q[l + (i + 1) * N] = cs[i] * q[l + (i + 1) * N] - sn[i] * q[l + i * N];
                q[l + i * N] = t;
            }
        }
    }
    // first row of R += w(1) v^T
#pragma unroll
    for (int j = 0; j < N; ++j) r[0 + j * N] = r[0 + j * N] + w[0] * v[j];
    if (N > 1) {
        // DQHQR: back to upper triangular, rotation i generated from column i
#pragma unroll
        for (int i = 0; i < N; ++i) {
            double t = r[0 + i * N];
#pragma unroll
            for (int j = 0; j < N - 1; ++j) {
                if (j < i) {
                    r[j + i * N] = cs[j] * t + sn[j] * r[(j + 1) + i * N];
                    t = cs[j] * r[(j + 1) + i * N] - sn[j] * t;
                }
            }
            if (i < N - 1) {
                dlartg(t, r[(i + 1) + i * N], cs[i], sn[i], r[i + i * N]);
                r[(i + 1) + i * N] = 0.0;
            } else {
                r[i + i * N] = t;
            }
        }
        // DQROT('F')
#pragma unroll
        for (int i = 0; i < N - 1; ++i) {
#pragma unroll
            for (int l = 0; l < N; ++l) {
                const double t = cs[i] * q[l + i * N] + sn[i] * q[l + (i + 1) * N];
                q[l + (i + 1) * N] = cs[i] * q[l + (i + 1) * N] - sn[i] * q[l + i * N];
                q[l + i * N] = t;
            }
        }
    }
}

// DRSCL: x := (1/sa) x with LAPACK's overflow-safe splitting of the reciprocal.
template <int N>
NLB_DEV void drscl(double sa, double (&x)[N]) {
    const double smlnum = 0x1p-1022;
    const double bignum = 1.0 / smlnum;
    double cden = sa, cnum = 1.0;
    for (;;) {
        const double cden1 = cden * smlnum;
        const double cnum1 = cnum / bignum;
        double mul;
        bool done;
        if (fabs(cden1) > fabs(cnum) && cnum != 0.0) { mul = smlnum; done = false; cden = cden1; }
        else if (fabs(cnum1) > fabs(cden)) { mul = bignum; done = false; cnum = cnum1; }
        else { mul = cnum / cden; done = true; }
#pragma unroll
        for (int i = 0; i < N; ++i) x[i] = mul * x[i];
        if (done) break;
    }
}

// DGEMV('T') with beta = 0: y = alpha * A^T x, column sums in row order.
template <int N>
NLB_DEV void gemv_t(double alpha, const double (&a)[N * N], const double (&x)[N], double (&y)[N]) {
#pragma unroll
    for (int j = 0; j < N; ++j) {
        double temp = 0.0;
#pragma unroll
        for (int i = 0; i < N; ++i) temp += a[i + j * N] * x[i];
        y[j] = 0.0 + alpha * temp;
    }
}

// DTRSV('U','N','N') in place.
template <int N>
NLB_DEV void trsv_upper(const double (&a)[N * N], double (&x)[N]) {
#pragma unroll
    for (int j = N - 1; j >= 0; --j) {
        if (x[j] != 0.0) {
            x[j] = x[j] / a[j + j * N];
            const double temp = x[j];
#pragma unroll
            for (int i = j - 1; i >= 0; --i) x[i] = x[i] - temp * a[i + j * N];
        }
    }
}

}  // namespace nlb
