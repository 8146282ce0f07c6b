// oracle/nl_polynomial.cpp — TEST INFRASTRUCTURE ONLY.  See nl_polynomial.h.
#include "nl_polynomial.h"

#include <cfloat>
#include <cmath>

#include "nl_lapack.h"

namespace nlo {

#define AT(a, ld, i, j) (a)[((long)(i) - 1) + ((long)(j) - 1) * (long)(ld)]

static inline real pow2(int e) { return real(std::ldexp(1.0, e)); }

void la_dlascl_g(real cfrom, real cto, int m, int n, real* a, int lda) {
    if (n == 0 || m == 0) return;
    const real smlnum = pow2(-1022);                 // dlamch('S')
    const real bignum = real(1.0) / smlnum;
    real cfromc = cfrom, ctoc = cto;
    for (;;) {
        real cfrom1 = cfromc * smlnum;
        real mul;
        bool done;
        if (cfrom1 == cfromc) {                      // cfromc is an infinity
            mul = ctoc / cfromc;
            done = true;
        } else {
            real cto1 = ctoc / bignum;
            if (cto1 == ctoc) {                      // ctoc is zero or an infinity
                mul = ctoc;
                done = true;
                cfromc = 1.0;
            } else if (f_abs(cfrom1) > f_abs(ctoc) && ctoc != real(0.0)) {
                mul = smlnum;
                done = false;
                cfromc = cfrom1;
            } else if (f_abs(cto1) > f_abs(cfromc)) {
                mul = bignum;
                done = false;
                ctoc = cto1;
            } else {
                mul = ctoc / cfromc;
                done = true;
                if (mul == real(1.0)) return;
            }
        }
        for (int j = 1; j <= n; ++j)
            for (int i = 1; i <= m; ++i) AT(a, lda, i, j) = AT(a, lda, i, j) * mul;
        if (done) break;
    }
}

real la_dlange_m(int m, int n, const real* a, int lda) {
    real value = 0.0;
    if (m == 0 || n == 0) return value;
    for (int j = 1; j <= n; ++j)
        for (int i = 1; i <= m; ++i) {
            real temp = f_abs(AT(a, lda, i, j));
            if (value < temp || f_isnan(temp)) value = temp;
        }
    return value;
}

int la_dgels(int m, int n, real* a, int lda, real* b, real* tau, real* work) {
    const real smlnum = pow2(-1022) / pow2(-52);     // dlamch('S') / dlamch('P')
    const real bignum = real(1.0) / smlnum;
    // scale A and b if their max entries are outside [smlnum, bignum]
    real anrm = la_dlange_m(m, n, a, lda);
    int iascl = 0;
    if (anrm > real(0.0) && anrm < smlnum) {
        la_dlascl_g(anrm, smlnum, m, n, a, lda);
        iascl = 1;
    } else if (anrm > bignum) {
        la_dlascl_g(anrm, bignum, m, n, a, lda);
        iascl = 2;
    } else if (anrm == real(0.0)) {
        for (int i = 1; i <= m; ++i) b[i - 1] = 0.0;   // DLASET on max(m, n) rows
        return 0;
    }
    real bnrm = la_dlange_m(m, 1, b, m);
    int ibscl = 0;
    if (bnrm > real(0.0) && bnrm < smlnum) {
        la_dlascl_g(bnrm, smlnum, m, 1, b, m);
        ibscl = 1;
    } else if (bnrm > bignum) {
        la_dlascl_g(bnrm, bignum, m, 1, b, m);
        ibscl = 2;
    }
    la_dgeqr2(m, n, a, lda, tau, work);              // DGEQRF, unblocked for min(m, n) < NB
    la_dorm2r_lt_vec(m, n, a, lda, tau, b);          // DORMQR('L','T'), unblocked for k <= NB
    for (int i = 1; i <= n; ++i)                     // DTRTRS: exact singularity check
        if (AT(a, lda, i, i) == real(0.0)) return i;
    for (int k = n; k >= 1; --k) {                   // DTRSM('L','U','N','N'), alpha = 1, one column
        if (b[k - 1] != real(0.0)) {
            b[k - 1] = b[k - 1] / AT(a, lda, k, k);
            for (int i = 1; i <= k - 1; ++i) b[i - 1] = b[i - 1] - b[k - 1] * AT(a, lda, i, k);
        }
    }
    if (iascl == 1) la_dlascl_g(anrm, smlnum, n, 1, b, m);
    else if (iascl == 2) la_dlascl_g(anrm, bignum, n, 1, b, m);
    if (ibscl == 1) la_dlascl_g(smlnum, bnrm, n, 1, b, m);
    else if (ibscl == 2) la_dlascl_g(bignum, bnrm, n, 1, b, m);
    return 0;
}

static int fit_common(int npts, int ncols, real* a, real* y, real* out) {
    real tau[64], work[64];
    int info = la_dgels(npts, ncols, a, npts, y, tau, work);
    for (int j = 0; j < ncols; ++j) out[j] = y[j];
    return info > 0 ? LA_INVALID_OPERATION_ERROR : 0;
}

int poly_fit(int npts, int order, const real* x, real* y, real* coeffs, real* a) {   // :146-199
    const int n = npts, ncols = order + 1;
    for (int j = 1; j <= n; ++j) {                   // :189-192
        AT(a, n, j, 1) = 1.0;
        AT(a, n, j, 2) = x[j - 1];
    }
    for (int j = 3; j <= ncols; ++j)                 // :193-195
        for (int i = 1; i <= n; ++i) AT(a, n, i, j) = AT(a, n, i, j - 1) * x[i - 1];
    return fit_common(n, ncols, a, y, coeffs);       // :198
}

int poly_fit_thru_zero(int npts, int order, const real* x, real* y, real* coeffs, real* a) {   // :202-253
    const int n = npts, ncols = order;
    for (int i = 1; i <= n; ++i) AT(a, n, i, 1) = x[i - 1];   // :245
    for (int j = 2; j <= ncols; ++j)                 // :246-248
        for (int i = 1; i <= n; ++i) AT(a, n, i, j) = AT(a, n, i, j - 1) * x[i - 1];
    coeffs[0] = 0.0;                                 // :251-252
    return fit_common(n, ncols, a, y, coeffs + 1);
}

real poly_eval(int order, const real* c, real x) {   // :256-283
    const int n = order + 1;
    if (order == -1) return real(0.0);
    if (order == 0) return c[0];
    real y = c[n - 1] * x + c[order - 1];
    for (int j = n - 2; j >= 1; --j) y = y * x + c[j - 1];
    return y;
}

}  // namespace nlo
