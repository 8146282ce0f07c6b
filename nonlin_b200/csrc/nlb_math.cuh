// nlb_math.cuh — FP64 device numerics shared by every kernel of the engine.
//
// The engine is compiled with -fmad=false: every a*b+c below is a DMUL followed by a DADD,
// in the order written, which is what a default gfortran x86-64 build of the reference
// executes.  IEEE division and square root are the CUDA
// defaults for double.  Nothing here calls a libdevice transcendental on the parity path:
// exp() is the software scheme below (argument reduction by a hi/lo split of ln 2, degree-5
// minimax polynomial in r*r), built from basic operations only so that it rounds the same
// way on any IEEE target.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#define NLB_DEV __device__ __forceinline__

namespace nlb {

// Fortran MAX/MIN on reals as gfortran expands them: keep the first argument on ties,
// replace a NaN first argument by the second.
NLB_DEV double nl_max(double a, double b) { return (b > a || a != a) ? b : a; }
NLB_DEV double nl_min(double a, double b) { return (b < a || a != a) ? b : a; }
// Fortran SIGN(a, b)
NLB_DEV double nl_sign(double a, double b) { return copysign(fabs(a), b); }

// Running state of Fortran NORM2 as libgfortran evaluates it: one pass, a running scale and
// a scaled sum of squares (the reference calls NORM2 at src/nonlin_least_squares.f90:213,
// 235,291,299,313,346,475,501,511,531,554,612,642,660 and src/nonlin_linesearch.f90:569).
struct Norm2 {
    double scale = 1.0, ssq = 0.0;
    NLB_DEV void add(double x) {
        if (x != 0.0) {
            // one division serves both cases (scale/a when a new maximum arrives, a/scale otherwise);
            // the two updates are the ones libgfortran performs, selected after the fact
            const double a = fabs(x);
            const bool up = scale < a;
            const double t = (up ? scale : a) / (up ? a : scale);
            const double tt = t * t;
            ssq = up ? (1.0 + ssq * t * t) : (ssq + tt);
            scale = up ? a : scale;
        }
    }
    NLB_DEV double value() const { return scale * sqrt(ssq); }
};

// Short register vectors: the plain two-branch form of the same recurrence is slightly faster there
// (measured on the 2x2 Broyden kernel), the select form wins in the long local-memory loops of LM.
template <int N>
NLB_DEV double norm2_vec(const double (&v)[N]) {
    double scale = 1.0, ssq = 0.0;
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const double x = v[i];
        if (x != 0.0) {
            const double a = fabs(x);
            if (scale < a) {
                const double t = scale / a;
                ssq = 1.0 + ssq * t * t;
                scale = a;
            } else {
                const double t = a / scale;
                ssq += t * t;
            }
        }
    }
    return scale * sqrt(ssq);
}

template <int N>
NLB_DEV double dot_vec(const double (&a)[N], const double (&b)[N]) {
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < N; ++i) s += a[i] * b[i];
    return s;
}

// Register-array access with a run-time index.  For short vectors a compare/select chain
// keeps the array in registers; longer ones fall back to (local-memory) indexing.
template <int N>
NLB_DEV double vget(const double (&v)[N], int idx) {
    if constexpr (N <= 8) {
        double r = v[0];
#pragma unroll
        for (int i = 1; i < N; ++i) r = (idx == i) ? v[i] : r;
        return r;
    } else {
        return v[idx];
    }
}
template <int N>
NLB_DEV void vset(double (&v)[N], int idx, double val) {
    if constexpr (N <= 8) {
#pragma unroll
        for (int i = 0; i < N; ++i) v[i] = (idx == i) ? val : v[i];
    } else {
        v[idx] = val;
    }
}
template <int N>
NLB_DEV int iget(const int (&v)[N], int idx) {
    if constexpr (N <= 8) {
        int r = v[0];
#pragma unroll
        for (int i = 1; i < N; ++i) r = (idx == i) ? v[i] : r;
        return r;
    } else {
        return v[idx];
    }
}
template <int N>
NLB_DEV void iset(int (&v)[N], int idx, int val) {
    if constexpr (N <= 8) {
#pragma unroll
        for (int i = 0; i < N; ++i) v[i] = (idx == i) ? val : v[i];
    } else {
        v[idx] = val;
    }
}

// ---- software exp --------------------------------------------------------------------
NLB_DEV double nl_exp(double x) {
    const double LN2_HI = 6.93147180369123816490e-01;
    const double LN2_LO = 1.90821492927058770002e-10;
    const double INV_LN2 = 1.44269504088896338700e+00;
    const double P1 = 1.66666666666666019037e-01;
    const double P2 = -2.77777777770155933842e-03;
    const double P3 = 6.61375632143793436117e-05;
    const double P4 = -1.65339022054652515390e-06;
    const double P5 = 4.13813679705723846039e-08;
    if (x != x) return x;
    if (x > 7.09782712893383973096e+02) return __longlong_as_double(0x7ff0000000000000LL);
    if (x < -7.45133219101941108420e+02) return 0.0;
    const double ax = fabs(x);
    double hi = 0.0, lo = 0.0, r;
    int k = 0;
    if (ax > 0.34657359027997264) {
        if (ax < 1.0397207708399179) {
            if (x > 0.0) { hi = x - LN2_HI; lo = LN2_LO; k = 1; }
            else { hi = x + LN2_HI; lo = -LN2_LO; k = -1; }
        } else {
            k = (int)(INV_LN2 * x + (x > 0.0 ? 0.5 : -0.5));
            const double t = (double)k;
            hi = x - t * LN2_HI;
            lo = t * LN2_LO;
        }
        r = hi - lo;
    } else if (ax < 3.7252902984619141e-09) {
        return 1.0 + x;
    } else {
        r = x;
    }
    const double t = r * r;
    const double c = r - t * (P1 + t * (P2 + t * (P3 + t * (P4 + t * P5))));
    // one division for both forms: (r*c)/(c-2) == -((r*c)/(2-c)) exactly (negation commutes with IEEE - and /)
    const double q = (r * c) / (2.0 - c);
    if (k == 0) return 1.0 - (-q - r);
    double y = 1.0 - ((lo - q) - hi);
    if (k >= -1021) {
        if (k == 1024) return y * 2.0 * 8.98846567431158e+307;
        return __longlong_as_double(__double_as_longlong(y) + ((long long)k << 52));
    }
    y = __longlong_as_double(__double_as_longlong(y) + ((long long)(k + 1000) << 52));
    return y * 9.33263618503218878990e-302;
}

// ---- LAPACK scalar kernels (Reference LAPACK 3.12.0 forms) -----------------------------
// Accumulator of DNRM2 (Blue's three-accumulator algorithm).
struct Dnrm2 {
    double asml = 0.0, amed = 0.0, abig = 0.0;
    bool notbig = true;
    NLB_DEV void add(double x) {
        const double tsml = 0x1p-511;
        const double tbig = 0x1p486;
        const double ssml = 0x1p537;
        const double sbig = 0x1p-538;
        const double ax = fabs(x);
        if (ax == 0.0) return;   // would add (0*ssml)^2 = +0 to asml: no effect on any accumulator
        if (ax > tbig) {
            const double t = ax * sbig;
            abig += t * t;
            notbig = false;
        } else if (ax < tsml) {
            if (notbig) {
                const double t = ax * ssml;
                asml += t * t;
            }
        } else {
            amed += ax * ax;
        }
    }
    NLB_DEV double value() {
        const double ssml = 0x1p537, sbig = 0x1p-538;
        const double maxn = 1.7976931348623157e+308;
        double scl, sumsq;
        if (abig > 0.0) {
            if (amed > 0.0 || amed > maxn || amed != amed) abig += (amed * sbig) * sbig;
            scl = 1.0 / sbig;
            sumsq = abig;
        } else if (asml > 0.0) {
            if (amed > 0.0 || amed > maxn || amed != amed) {
                amed = sqrt(amed);
                asml = sqrt(asml) / ssml;
                double ymin, ymax;
                if (asml > amed) { ymin = amed; ymax = asml; }
                else { ymin = asml; ymax = amed; }
                scl = 1.0;
                const double q = ymin / ymax;
                sumsq = (ymax * ymax) * (1.0 + q * q);
            } else {
                scl = 1.0 / ssml;
                sumsq = asml;
            }
        } else {
            scl = 1.0;
            sumsq = amed;
        }
        return scl * sqrt(sumsq);
    }
};

NLB_DEV double dlapy2(double x, double y) {
    if (x != x) return x;
    if (y != y) return y;
    const double xa = fabs(x), ya = fabs(y);
    const double w = nl_max(xa, ya), z = nl_min(xa, ya);
    if (z == 0.0 || w > 1.7976931348623157e+308) return w;
    const double q = z / w;
    return w * sqrt(1.0 + q * q);
}

// DLARTG's scaled branch (|f| or |g| outside [2^-511, 2^510.5]): rare, kept out of line.
static __device__ __noinline__ void dlartg_scaled(double f, double g, double& c, double& s, double& r) {
    const double safmin = 0x1p-1022, safmax = 0x1p1022;
    const double u = nl_min(safmax, nl_max(safmin, nl_max(fabs(f), fabs(g))));
    const double fs = f / u, gs = g / u;
    const double d = sqrt(fs * fs + gs * gs);
    c = fabs(fs) / d;
    r = nl_sign(d, f);
    s = gs / r;
    r = r * u;
}

// DLARTG: c*f + s*g = r, -s*f + c*g = 0
NLB_DEV void dlartg(double f, double g, double& c, double& s, double& r) {
    const double safmin = 0x1p-1022;
    const double safmax = 0x1p1022;
    const double rtmin = 0x1p-511;                   // sqrt(safmin)
    const double rtmax = 0x1.6a09e667f3bcdp+510;     // sqrt(safmax/2)
    const double f1 = fabs(f), g1 = fabs(g);
    if (g == 0.0) {
        c = 1.0; s = 0.0; r = f;
    } else if (f == 0.0) {
        c = 0.0; s = nl_sign(1.0, g); r = g1;
    } else if (f1 > rtmin && f1 < rtmax && g1 > rtmin && g1 < rtmax) {
        const double d = sqrt(f * f + g * g);
        c = f1 / d;
        r = nl_sign(d, f);
        s = g / r;
    } else {
        dlartg_scaled(f, g, c, s, r);
    }
}

// r of DLARTG alone (same branches, same operations, no c and s): for the sequential Givens chains.
NLB_DEV double dlartg_r(double f, double g) {
    const double safmin = 0x1p-1022;
    const double safmax = 0x1p1022;
    const double rtmin = 0x1p-511;
    const double rtmax = 0x1.6a09e667f3bcdp+510;
    const double f1 = fabs(f), g1 = fabs(g);
    if (g == 0.0) return f;
    if (f == 0.0) return g1;
    if (f1 > rtmin && f1 < rtmax && g1 > rtmin && g1 < rtmax) return nl_sign(sqrt(f * f + g * g), f);
    const double u = nl_min(safmax, nl_max(safmin, nl_max(f1, g1)));
    const double fs = f / u, gs = g / u;
    return nl_sign(sqrt(fs * fs + gs * gs), f) * u;
}

}  // namespace nlb
