// oracle/nl_lapack.cpp — TEST INFRASTRUCTURE ONLY.  See nl_lapack.h for what each routine
// restates and why this boundary is "parity unpinned" at bit level.
#include "nl_lapack.h"

#include <cfloat>

namespace nlo {

#define AT(a, ld, i, j) (a)[((long)(i) - 1) + ((long)(j) - 1) * (long)(ld)]

static inline real pow2(int e) { return real(std::ldexp(1.0, e)); }

real la_dnrm2(int n, const real* x, int incx) {
    if (n <= 0) return real(0.0);
    const real tsml = pow2(-511), tbig = pow2(486), ssml = pow2(537), sbig = pow2(-538);
    const real maxn = real(DBL_MAX);
    real asml = 0.0, amed = 0.0, abig = 0.0;
    bool notbig = true;
    long ix = 0;
    if (incx < 0) ix = -(long)(n - 1) * incx;
    for (int i = 0; i < n; ++i, ix += incx) {
        real ax = f_abs(x[ix]);
        if (ax > tbig) {
            real t = ax * sbig;
            abig += t * t;
            notbig = false;
        } else if (ax < tsml) {
            if (notbig) {
                real t = ax * ssml;
                asml += t * t;
            }
        } else {
            amed += ax * ax;
        }
    }
    real scl, sumsq;
    if (abig > real(0.0)) {
        if (amed > real(0.0) || amed > maxn || f_isnan(amed)) abig += (amed * sbig) * sbig;
        scl = real(1.0) / sbig;
        sumsq = abig;
    } else if (asml > real(0.0)) {
        if (amed > real(0.0) || amed > maxn || f_isnan(amed)) {
            amed = f_sqrt(amed);
            asml = f_sqrt(asml) / ssml;
            real ymin, ymax;
            if (asml > amed) { ymin = amed; ymax = asml; }
            else { ymin = asml; ymax = amed; }
            scl = 1.0;
            real q = ymin / ymax;
            sumsq = (ymax * ymax) * (real(1.0) + q * q);
        } else {
            scl = real(1.0) / ssml;
            sumsq = asml;
        }
    } else {
        scl = 1.0;
        sumsq = amed;
    }
    return scl * f_sqrt(sumsq);
}

real la_dlapy2(real x, real y) {
    bool xn = f_isnan(x), yn = f_isnan(y);
    if (xn) return x;
    if (yn) return y;
    real xa = f_abs(x), ya = f_abs(y);
    real w = f_max(xa, ya), z = f_min(xa, ya);
    if (z == real(0.0) || w > real(DBL_MAX)) return w;
    real q = z / w;
    return w * f_sqrt(real(1.0) + q * q);
}

void la_dlartg(real f, real g, real* c, real* s, real* r) {
    const real safmin = pow2(-1022), safmax = pow2(1022);
    const real rtmin = f_sqrt(safmin), rtmax = f_sqrt(safmax / real(2.0));
    real f1 = f_abs(f), g1 = f_abs(g);
    if (g == real(0.0)) {
        *c = 1.0; *s = 0.0; *r = f;
    } else if (f == real(0.0)) {
        *c = 0.0; *s = f_sign(real(1.0), g); *r = g1;
    } else if (f1 > rtmin && f1 < rtmax && g1 > rtmin && g1 < rtmax) {
        real d = f_sqrt(f * f + g * g);
        *c = f1 / d;
        *r = f_sign(d, f);
        *s = g / *r;
    } else {
        real u = f_min(safmax, f_max(safmin, f_max(f1, g1)));
        real fs = f / u, gs = g / u;
        real d = f_sqrt(fs * fs + gs * gs);
        *c = f_abs(fs) / d;
        *r = f_sign(d, f);
        *s = gs / *r;
        *r = *r * u;
    }
}

void la_dlarfg(int n, real* alpha, real* x, int incx, real* tau) {
    if (n <= 1) { *tau = 0.0; return; }
    real xnorm = la_dnrm2(n - 1, x, incx);
    if (xnorm == real(0.0)) { *tau = 0.0; return; }
    real beta = -f_sign(la_dlapy2(*alpha, xnorm), *alpha);
    const real safmin = pow2(-1022) / pow2(-53);
    int knt = 0;
    if (f_abs(beta) < safmin) {
        const real rsafmn = real(1.0) / safmin;
        do {
            ++knt;
            for (int i = 0; i < n - 1; ++i) x[(long)i * incx] = rsafmn * x[(long)i * incx];
            beta = beta * rsafmn;
            *alpha = *alpha * rsafmn;
        } while (f_abs(beta) < safmin && knt < 20);
        xnorm = la_dnrm2(n - 1, x, incx);
        beta = -f_sign(la_dlapy2(*alpha, xnorm), *alpha);
    }
    *tau = (beta - *alpha) / beta;
    real sc = real(1.0) / (*alpha - beta);
    for (int i = 0; i < n - 1; ++i) x[(long)i * incx] = sc * x[(long)i * incx];
    for (int j = 0; j < knt; ++j) beta = beta * safmin;
    *alpha = beta;
}

// ILADLC: index of the last column of the m-by-n matrix with a non-zero entry (0 if none).
static int iladlc(int m, int n, const real* a, int lda) {
    if (n == 0) return 0;
    if (AT(a, lda, 1, n) != real(0.0) || AT(a, lda, m, n) != real(0.0)) return n;
    for (int col = n; col >= 1; --col)
        for (int i = 1; i <= m; ++i)
            if (AT(a, lda, i, col) != real(0.0)) return col;
    return 0;
}

void la_dgemv_t(int m, int n, real alpha, const real* a, int lda, const real* x, real* y) {
    if (m == 0 || n == 0) return;
    for (int j = 1; j <= n; ++j) y[j - 1] = 0.0;   // beta = 0
    if (alpha == real(0.0)) return;
    for (int j = 1; j <= n; ++j) {
        real temp = 0.0;
        for (int i = 1; i <= m; ++i) temp += AT(a, lda, i, j) * x[i - 1];
        y[j - 1] = y[j - 1] + alpha * temp;
    }
}

// DGEMV('N') with beta = 0: y := alpha A x.  Column sweep, y(i) += (alpha x(j)) a(i,j), j outermost.
void la_dgemv_n(int m, int n, real alpha, const real* a, int lda, const real* x, real* y) {
    if (m == 0 || n == 0) return;
    for (int i = 1; i <= m; ++i) y[i - 1] = 0.0;   // beta = 0
    if (alpha == real(0.0)) return;
    for (int j = 1; j <= n; ++j) {
        real temp = alpha * x[j - 1];
        for (int i = 1; i <= m; ++i) y[i - 1] = y[i - 1] + temp * AT(a, lda, i, j);
    }
}

void la_dger(int m, int n, real alpha, const real* x, const real* y, real* a, int lda) {
    if (m == 0 || n == 0 || alpha == real(0.0)) return;
    for (int j = 1; j <= n; ++j) {
        if (y[j - 1] != real(0.0)) {
            real temp = alpha * y[j - 1];
            for (int i = 1; i <= m; ++i) AT(a, lda, i, j) = AT(a, lda, i, j) + x[i - 1] * temp;
        }
    }
}

void la_dlarf_left(int m, int n, const real* v, int incv, real tau, real* c, int ldc, real* work) {
    int lastv = 0, lastc = 0;
    if (tau != real(0.0)) {
        lastv = m;
        long i = (incv > 0) ? 1 + (long)(lastv - 1) * incv : 1;
        while (lastv > 0 && v[i - 1] == real(0.0)) { --lastv; i -= incv; }
        lastc = iladlc(lastv, n, c, ldc);
    }
    if (lastv > 0 && lastc > 0) {
        // work(1:lastc) = C(1:lastv,1:lastc)^T v ; C -= tau v work^T      (incv == 1 here)
        la_dgemv_t(lastv, lastc, real(1.0), c, ldc, v, work);
        la_dger(lastv, lastc, -tau, v, work, c, ldc);
    }
}

// DORM2R('L','T') on one right-hand side: c := Q^T c = H(k) ... H(1) c, reflectors applied in the order
// i = 1..k, each through DLARF with the unit diagonal entry put in place for the call.
void la_dorm2r_lt_vec(int m, int k, real* a, int lda, const real* tau, real* c) {
    real work[1];
    for (int i = 1; i <= k; ++i) {
        real aii = AT(a, lda, i, i);
        AT(a, lda, i, i) = 1.0;
        la_dlarf_left(m - i + 1, 1, &AT(a, lda, i, i), 1, tau[i - 1], &c[i - 1], m, work);
        AT(a, lda, i, i) = aii;
    }
}

void la_dgeqr2(int m, int n, real* a, int lda, real* tau, real* work) {
    int k = m < n ? m : n;
    for (int i = 1; i <= k; ++i) {
        int ip1 = (i + 1 < m) ? i + 1 : m;
        la_dlarfg(m - i + 1, &AT(a, lda, i, i), &AT(a, lda, ip1, i), 1, &tau[i - 1]);
        if (i < n) {
            real aii = AT(a, lda, i, i);
            AT(a, lda, i, i) = 1.0;
            la_dlarf_left(m - i + 1, n - i, &AT(a, lda, i, i), 1, tau[i - 1], &AT(a, lda, i, i + 1), lda, work);
            AT(a, lda, i, i) = aii;
        }
    }
}

void la_dorg2r(int m, int n, int k, real* a, int lda, const real* tau, real* work) {
    if (n <= 0) return;
    for (int j = k + 1; j <= n; ++j) {
        for (int l = 1; l <= m; ++l) AT(a, lda, l, j) = 0.0;
        AT(a, lda, j, j) = 1.0;
    }
    for (int i = k; i >= 1; --i) {
        if (i < n) {
            AT(a, lda, i, i) = 1.0;
            la_dlarf_left(m - i + 1, n - i, &AT(a, lda, i, i), 1, tau[i - 1], &AT(a, lda, i, i + 1), lda, work);
        }
        if (i < m) {
            real sc = -tau[i - 1];
            for (int l = i + 1; l <= m; ++l) AT(a, lda, l, i) = sc * AT(a, lda, l, i);
        }
        AT(a, lda, i, i) = real(1.0) - tau[i - 1];
        for (int l = 1; l <= i - 1; ++l) AT(a, lda, l, i) = 0.0;
    }
}

int la_dgetrf(int m, int n, real* a, int lda, int* ipiv) {
    const real sfmin = pow2(-1022);
    int info = 0;
    int mn = m < n ? m : n;
    for (int j = 1; j <= mn; ++j) {
        // IDAMAX over a(j:m, j): first entry of largest magnitude
        int jp = j;
        real dmax = f_abs(AT(a, lda, j, j));
        for (int i = j + 1; i <= m; ++i) {
            real v = f_abs(AT(a, lda, i, j));
            if (v > dmax) { dmax = v; jp = i; }
        }
        ipiv[j - 1] = jp;
        if (AT(a, lda, jp, j) != real(0.0)) {
            if (jp != j) {
                for (int c = 1; c <= n; ++c) {
                    real t = AT(a, lda, j, c);
                    AT(a, lda, j, c) = AT(a, lda, jp, c);
                    AT(a, lda, jp, c) = t;
                }
            }
            if (j < m) {
                if (f_abs(AT(a, lda, j, j)) >= sfmin) {
                    real rp = real(1.0) / AT(a, lda, j, j);
                    for (int i = j + 1; i <= m; ++i) AT(a, lda, i, j) = rp * AT(a, lda, i, j);
                } else {
                    for (int i = j + 1; i <= m; ++i) AT(a, lda, i, j) = AT(a, lda, i, j) / AT(a, lda, j, j);
                }
            }
        } else if (info == 0) {
            info = j;
        }
        if (j < mn) {
            // trailing update a(j+1:m, j+1:n) -= a(j+1:m, j) a(j, j+1:n)     (DGER, alpha = -1)
            for (int c = j + 1; c <= n; ++c) {
                if (AT(a, lda, j, c) != real(0.0)) {
                    real temp = -AT(a, lda, j, c);
                    for (int i = j + 1; i <= m; ++i)
                        AT(a, lda, i, c) = AT(a, lda, i, c) + AT(a, lda, i, j) * temp;
                }
            }
        }
    }
    return info;
}

void la_dgetrs(int n, const real* a, int lda, const int* ipiv, real* b) {
    for (int i = 1; i <= n; ++i) {
        int ip = ipiv[i - 1];
        if (ip != i) { real t = b[i - 1]; b[i - 1] = b[ip - 1]; b[ip - 1] = t; }
    }
    for (int k = 1; k <= n; ++k) {
        if (b[k - 1] != real(0.0))
            for (int i = k + 1; i <= n; ++i) b[i - 1] = b[i - 1] - b[k - 1] * AT(a, lda, i, k);
    }
    for (int k = n; k >= 1; --k) {
        if (b[k - 1] != real(0.0)) {
            b[k - 1] = b[k - 1] / AT(a, lda, k, k);
            for (int i = 1; i <= k - 1; ++i) b[i - 1] = b[i - 1] - b[k - 1] * AT(a, lda, i, k);
        }
    }
}

void la_drscl(int n, real sa, real* x) {
    if (n <= 0) return;
    const real smlnum = pow2(-1022);
    const real bignum = real(1.0) / smlnum;
    real cden = sa, cnum = 1.0;
    for (;;) {
        real cden1 = cden * smlnum;
        real cnum1 = cnum / bignum;
        real mul;
        bool done;
        if (f_abs(cden1) > f_abs(cnum) && cnum != real(0.0)) {
            mul = smlnum; done = false; cden = cden1;
        } else if (f_abs(cnum1) > f_abs(cden)) {
            mul = bignum; done = false; cnum = cnum1;
        } else {
            mul = cnum / cden; done = true;
        }
        for (int i = 0; i < n; ++i) x[i] = mul * x[i];
        if (done) break;
    }
}

void la_dtrsv_unn(int n, const real* a, int lda, real* x) {
    for (int j = n; j >= 1; --j) {
        if (x[j - 1] != real(0.0)) {
            x[j - 1] = x[j - 1] / AT(a, lda, j, j);
            real temp = x[j - 1];
            for (int i = j - 1; i >= 1; --i) x[i - 1] = x[i - 1] - temp * AT(a, lda, i, j);
        }
    }
}

// ---- QRUPDATE pieces -------------------------------------------------------------------
// DQRTV1: rotations that reduce u to a multiple of e1, bottom-up.  cosines -> w(1:n-1),
// sines -> u(2:n), u(1) = +-|u|.
static void qr_dqrtv1(int n, real* u, real* w) {
    if (n <= 0) return;
    real rr = u[n - 1];
    for (int i = n - 1; i >= 1; --i) {
        real t;
        la_dlartg(u[i - 1], rr, &w[i - 1], &u[i], &t);
        rr = t;
    }
    u[0] = rr;
}

// DQRQH: apply those rotations to upper-trapezoidal R, column by column -> upper Hessenberg.
static void qr_dqrqh(int m, int n, real* r, int ldr, const real* c, const real* s) {
    if (m == 0 || m == 1 || n == 0) return;
    for (int i = 1; i <= n; ++i) {
        int ii = (m - 1 < i) ? m - 1 : i;
        real t = AT(r, ldr, ii + 1, i);
        for (int j = ii; j >= 1; --j) {
            AT(r, ldr, j + 1, i) = c[j - 1] * t - s[j - 1] * AT(r, ldr, j, i);
            t = c[j - 1] * AT(r, ldr, j, i) + s[j - 1] * t;
        }
        AT(r, ldr, 1, i) = t;
    }
}

// DQHQR: re-triangularise an upper Hessenberg R, generating rotation i from column i.
static void qr_dqhqr(int m, int n, real* r, int ldr, real* c, real* s) {
    if (m == 0 || m == 1 || n == 0) return;
    for (int i = 1; i <= n; ++i) {
        real t = AT(r, ldr, 1, i);
        int ii = (m < i) ? m : i;
        for (int j = 1; j <= ii - 1; ++j) {
            AT(r, ldr, j, i) = c[j - 1] * t + s[j - 1] * AT(r, ldr, j + 1, i);
            t = c[j - 1] * AT(r, ldr, j + 1, i) - s[j - 1] * t;
        }
        if (ii < m) {
            la_dlartg(t, AT(r, ldr, ii + 1, i), &c[i - 1], &s[i - 1], &AT(r, ldr, ii, i));
            AT(r, ldr, ii + 1, i) = 0.0;
        } else {
            AT(r, ldr, ii, i) = t;
        }
    }
}

// DROT on two columns.
static inline void qr_drot(int m, real* x, real* y, real c, real s) {
    for (int l = 0; l < m; ++l) {
        real t = c * x[l] + s * y[l];
        y[l] = c * y[l] - s * x[l];
        x[l] = t;
    }
}

// DQROT: apply n-1 rotations to the columns of Q, forward ('F') or backward ('B').
static void qr_dqrot(bool forward, int m, int n, real* q, int ldq, const real* c, const real* s) {
    if (m == 0 || n == 0 || n == 1) return;
    if (forward) {
        for (int i = 1; i <= n - 1; ++i) qr_drot(m, &AT(q, ldq, 1, i), &AT(q, ldq, 1, i + 1), c[i - 1], s[i - 1]);
    } else {
        for (int i = n - 1; i >= 1; --i) qr_drot(m, &AT(q, ldq, 1, i), &AT(q, ldq, 1, i + 1), c[i - 1], s[i - 1]);
    }
}

void la_dqr1up(int m, int n, real* q, int ldq, real* r, int ldr, const real* u, const real* v, real* w) {
    int k = m;   // full Q
    if (k == 0 || n == 0) return;
    // w(1:k) = Q^T u, each entry a sequential DDOT
    for (int i = 1; i <= k; ++i) {
        real s = 0.0;
        for (int l = 1; l <= m; ++l) s += AT(q, ldq, l, i) * u[l - 1];
        w[i - 1] = s;
    }
    qr_dqrtv1(k, w, w + k);                          // cos -> w(k+1..), sin -> w(2..k)
    qr_dqrqh(k, n, r, ldr, w + k, w + 1);            // R -> Hessenberg
    qr_dqrot(false, m, k, q, ldq, w + k, w + 1);     // Q <- Q G^T (backward)
    for (int j = 1; j <= n; ++j)                     // first row of R += w(1) v^T   (DAXPY)
        AT(r, ldr, 1, j) = AT(r, ldr, 1, j) + w[0] * v[j - 1];
    qr_dqhqr(k, n, r, ldr, w + k, w);                // back to triangular
    int nq = (k < n + 1) ? k : n + 1;
    qr_dqrot(true, m, nq, q, ldq, w + k, w);         // Q <- Q G^T (forward)
}

}  // namespace nlo
