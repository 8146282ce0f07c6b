// oracle/nl_numerics.h — TEST INFRASTRUCTURE ONLY (see oracle/README.md).
//
// Scalar numerics of the Fortran run time that the reference relies on, restated so the
// CPU oracle performs the same IEEE-754 binary64 operations in the same order as a
// gfortran x86-64 build of /root/reference (no FMA contraction: compile with
// -ffp-contract=off).  Nothing here is shipped in the product library.
//
// What is restated:
//   * NORM2  -> libgfortran's one-pass scaled algorithm (scale/ssq with division)
//   * MAX/MIN on reals -> gfortran's NaN-aware compare-and-select expansion
//   * x**2, x**3 -> repeated multiplication (callers write x*x themselves)
//   * exp()  -> a software exponential built only from + - * / and exponent-field
//               arithmetic (argument reduction by ln2 split in hi/lo, degree-5 Remez
//               polynomial in r*r, the scheme used by fdlibm-class libms).  The CUDA
//               engine carries its own copy of the same scheme, so residuals that call
//               exp() are bit-identical on CPU and GPU.  nl_set_libm_exp(1) switches the
//               oracle to the host libm exp() so KATs can be pinned with either.
#ifndef NL_NUMERICS_H
#define NL_NUMERICS_H

#include <cmath>
#include <cstdint>
#include <cstring>

namespace nlo {

// ---------------------------------------------------------------------------------------
// `real`: plain double, or (build with -DNL_COUNT_FLOPS) a wrapper that counts every
// + - * / sqrt it executes.  The counting build gives the *algorithmic* FP64 operation
// count per system that bench.py's roofline uses (DESIGN.md "flop model").
// ---------------------------------------------------------------------------------------
#ifdef NL_COUNT_FLOPS
extern thread_local unsigned long long g_flops;
struct real {
    double v;
    real() = default;
    real(double d) : v(d) {}
    real(int i) : v((double)i) {}
    explicit operator double() const { return v; }
    explicit operator int() const { return (int)v; }
};
inline real operator+(real a, real b) { ++g_flops; return real(a.v + b.v); }
inline real operator-(real a, real b) { ++g_flops; return real(a.v - b.v); }
inline real operator*(real a, real b) { ++g_flops; return real(a.v * b.v); }
inline real operator/(real a, real b) { ++g_flops; return real(a.v / b.v); }
inline real operator-(real a) { return real(-a.v); }
inline real& operator+=(real& a, real b) { a = a + b; return a; }
inline real& operator-=(real& a, real b) { a = a - b; return a; }
inline real& operator*=(real& a, real b) { a = a * b; return a; }
inline real& operator/=(real& a, real b) { a = a / b; return a; }
inline bool operator<(real a, real b) { return a.v < b.v; }
inline bool operator>(real a, real b) { return a.v > b.v; }
inline bool operator<=(real a, real b) { return a.v <= b.v; }
inline bool operator>=(real a, real b) { return a.v >= b.v; }
inline bool operator==(real a, real b) { return a.v == b.v; }
inline bool operator!=(real a, real b) { return a.v != b.v; }
inline double dval(real a) { return a.v; }
inline real f_sqrt(real a) { ++g_flops; return real(std::sqrt(a.v)); }
#else
typedef double real;
inline double dval(real a) { return a; }
inline real f_sqrt(real a) { return std::sqrt(a); }
#endif

static_assert(sizeof(real) == sizeof(double), "real must be layout-compatible with double");

inline real f_abs(real a) { return real(std::fabs(dval(a))); }
inline bool f_isnan(real a) { return dval(a) != dval(a); }
// Fortran SIGN(a, b): |a| with the sign of b.
inline real f_sign(real a, real b) { return real(std::copysign(std::fabs(dval(a)), dval(b))); }

// gfortran expands MAX(a, b) on reals (no -ffinite-math-only, no fast fmax on SSE2) to
//   m = a; if (b > m || isnan(m)) m = b;
// and MIN likewise with '<'.  Ties keep the first argument; a NaN first argument is
// replaced by the second.
inline real f_max(real a, real b) { return (b > a || f_isnan(a)) ? b : a; }
inline real f_min(real a, real b) { return (b < a || f_isnan(a)) ? b : a; }

// STUDY SWITCH - not the reference's arithmetic (default 0 = off; nothing in the test-suite turns it on).
// Mode 1 evaluates every long sum (64 terms or more: the m-length norms and dot products of the LM path) in the order a
// warp-shuffle reduction would: lane l adds terms l, l+32, l+64, ..., then a butterfly adds the 32 partial sums; norms
// become sqrt of the plain sum of squares.  scripts/striped_sum_study.py uses it to measure how far iteration counts
// and results move when the summation order is not the reference's (VERDICT r1 #2a / SURVEY 7 option i).
void nl_set_sum_mode(int mode);
int nl_get_sum_mode();
template <class Term>
inline real f_striped_sum(int n, Term term) {
    real part[32];
    for (int l = 0; l < 32; ++l) part[l] = 0.0;
    for (int i = 0; i < n; ++i) part[i & 31] += term(i);
    for (int d = 16; d >= 1; d >>= 1)
        for (int l = 0; l < d; ++l) part[l] += part[l + d];
    return part[0];
}

// libgfortran norm2_r8: one pass, running scale and scaled sum of squares.
inline real f_norm2(const real* v, int n, int stride = 1) {
    if (n >= 64 && nl_get_sum_mode() == 1)
        return f_sqrt(f_striped_sum(n, [&](int i) { const real x = v[(long)i * stride]; return x * x; }));
    real scale = 1.0;
    real ssq = 0.0;
    for (int i = 0; i < n; ++i) {
        real x = v[(long)i * stride];
        if (x != real(0.0)) {
            real a = f_abs(x);
            if (scale < a) {
                real t = scale / a;
                ssq = real(1.0) + ssq * t * t;
                scale = a;
            } else {
                real t = a / scale;
                ssq += t * t;
            }
        }
    }
    return scale * f_sqrt(ssq);
}

// Sequential dot product, index order, as an inlined Fortran DOT_PRODUCT.
inline real f_dot(const real* a, const real* b, int n) {
    if (n >= 64 && nl_get_sum_mode() == 1) return f_striped_sum(n, [&](int i) { return a[i] * b[i]; });
    real s = 0.0;
    for (int i = 0; i < n; ++i) s += a[i] * b[i];
    return s;
}

// ---------------------------------------------------------------------------------------
// Software exp: basic IEEE operations only (bit-reproducible on any conforming target).
// ---------------------------------------------------------------------------------------
void nl_set_libm_exp(int on);
int nl_get_libm_exp();

inline double soft_exp_d(double x) {
    const double LN2_HI = 6.93147180369123816490e-01;   // upper bits of ln 2 (low 21 bits zero)
    const double LN2_LO = 1.90821492927058770002e-10;   // ln 2 - LN2_HI
    const double INV_LN2 = 1.44269504088896338700e+00;
    const double P1 = 1.66666666666666019037e-01;
    const double P2 = -2.77777777770155933842e-03;
    const double P3 = 6.61375632143793436117e-05;
    const double P4 = -1.65339022054652515390e-06;
    const double P5 = 4.13813679705723846039e-08;
    if (x != x) return x;
    if (x > 7.09782712893383973096e+02) return INFINITY;
    if (x < -7.45133219101941108420e+02) return 0.0;
    double ax = std::fabs(x);
    double hi = 0.0, lo = 0.0, r;
    int k = 0;
    if (ax > 0.34657359027997264) {            // |x| > ln2/2 : reduce
        if (ax < 1.0397207708399179) {         // |x| < 1.5 ln2 : k = +-1
            if (x > 0.0) { hi = x - LN2_HI; lo = LN2_LO; k = 1; }
            else         { hi = x + LN2_HI; lo = -LN2_LO; k = -1; }
        } else {
            k = (int)(INV_LN2 * x + (x > 0.0 ? 0.5 : -0.5));
            double t = (double)k;
            hi = x - t * LN2_HI;
            lo = t * LN2_LO;
        }
        r = hi - lo;
    } else if (ax < 3.7252902984619141e-09) {  // |x| < 2^-28
        return 1.0 + x;
    } else {
        r = x;
    }
    double t = r * r;
    double c = r - t * (P1 + t * (P2 + t * (P3 + t * (P4 + t * P5))));
    if (k == 0) return 1.0 - ((r * c) / (c - 2.0) - r);
    double y = 1.0 - ((lo - (r * c) / (2.0 - c)) - hi);
    // scale by 2^k through the exponent field (y is in [0.5, 2))
    uint64_t bits;
    if (k >= -1021) {
        if (k == 1024) return y * 2.0 * 8.98846567431158e+307;
        std::memcpy(&bits, &y, 8);
        bits += (uint64_t)((int64_t)k << 52);
        std::memcpy(&y, &bits, 8);
        return y;
    }
    std::memcpy(&bits, &y, 8);
    bits += (uint64_t)((int64_t)(k + 1000) << 52);
    std::memcpy(&y, &bits, 8);
    return y * 9.33263618503218878990e-302;    // 2^-1000
}

#ifdef NL_COUNT_FLOPS
// exp is charged as the 25 basic operations its reduced path executes.
inline real f_exp(real x) { g_flops += 25; return real(nl_get_libm_exp() ? std::exp(x.v) : soft_exp_d(x.v)); }
#else
inline real f_exp(real x) { return nl_get_libm_exp() ? std::exp(x) : soft_exp_d(x); }
#endif

}  // namespace nlo
#endif
