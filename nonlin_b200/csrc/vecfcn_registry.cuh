// vecfcn_registry.cuh — the registered __device__ residual functions ("vecfcn",
// reference src/nonlin_multi_eqn_mult_var.f90:14-25) and optional Jacobians ("jacobianfcn",
// :27-38).  A device function pointer cannot come from Fortran, so the batch extension
// addresses residuals by id / name (nlb_vecfcn_lookup).  To register a new residual: add a
// functor here, add it to NLB_FOR_EACH_* below, rebuild.
//
// Every expression keeps the evaluation order of the Fortran source it mirrors (left to
// right; x**2 = x*x; x**3 = (x*x)*x) because with a forward-difference Jacobian one ulp in
// the residual moves the iteration counts.
#pragma once
#include "nlb_math.cuh"

namespace nlb {

// Per-system view of the batch inputs.  sys is SoA: value k of system b is sys[k*B + b].
struct SysCtx {
    const double* __restrict__ sys;      // already offset by b
    const double* __restrict__ shared;
    long long B;
    int m, n;
    NLB_DEV double sysv(int k) const { return __ldg(sys + (long long)k * B); }
    NLB_DEV double sharedv(int i) const { return __ldg(shared + i); }
};

enum FcnId {
    FCN_MISC_2FCN = 0,
    FCN_MISC_2FCN_A = 1,
    FCN_POORLY_SCALED = 2,
    FCN_POWELL_BADLY_SCALED = 3,
    FCN_LSQ_POLY_FIT = 4,
    FCN_POLAR = 5,
    FCN_POLAR_SCALED = 6,
    FCN_MISC_2FCN_01 = 7,
    FCN_RATIONAL_7_8 = 8,
    FCN_EXP_SUM_8 = 9,
    FCN_EXT_ROSENBROCK = 10,
    FCN_EXP_DECAY_4 = 11,
    FCN_COUNT = 12
};

// Column-major M x N Jacobian view used by the analytic Jacobians.
template <int M>
struct JacView {
    double* a;
    NLB_DEV double& operator()(int i, int j) const { return a[i + j * M]; }
};

// ---- fixed-size systems (thread-per-system kernels) ------------------------------------

// x**2 + y**2 = 34, x**2 - 2 y**2 = 7 : reference tests/nonlin_test_solve.f90:41-47 (fcn1),
// :67-72 (jac1); examples/example_problems.f90 misc_2fcn (README Example 1).
struct Misc2Fcn {
    static constexpr int ID = FCN_MISC_2FCN, M = 2, N = 2, SYS_LEN = 0, SHARED_LEN = 0;
    static constexpr bool HAS_JAC = true;
    NLB_DEV static void eval(const double (&x)[2], double (&f)[2], const SysCtx&) {
        f[0] = x[0] * x[0] + x[1] * x[1] - 34.0;
        f[1] = x[0] * x[0] - 2.0 * (x[1] * x[1]) - 7.0;
    }
    NLB_DEV static void jac(const double (&x)[2], JacView<2> J, const SysCtx&) {
        J(0, 0) = 2.0 * x[0];
        J(1, 0) = 2.0 * x[0];
        J(0, 1) = 2.0 * x[1];
        J(1, 1) = 2.0 * (-2.0 * x[1]);
    }
};

// the same system with its coefficient passed through args: tests/nonlin_test_solve.f90:49-60, :74-84
struct Misc2FcnA {
    static constexpr int ID = FCN_MISC_2FCN_A, M = 2, N = 2, SYS_LEN = 1, SHARED_LEN = 0;
    static constexpr bool HAS_JAC = true;
    NLB_DEV static void eval(const double (&x)[2], double (&f)[2], const SysCtx& c) {
        const double a = c.sysv(0);
        f[0] = x[0] * x[0] + x[1] * x[1] - 34.0;
        f[1] = x[0] * x[0] - a * (x[1] * x[1]) - 7.0;
    }
    NLB_DEV static void jac(const double (&x)[2], JacView<2> J, const SysCtx& c) {
        const double a = c.sysv(0);
        J(0, 0) = 2.0 * x[0];
        J(1, 0) = 2.0 * x[0];
        J(0, 1) = 2.0 * x[1];
        J(1, 1) = 2.0 * (-a * x[1]);
    }
};

// x2 - 10 = 0, x1 x2 - 5e4 = 0 : tests/nonlin_test_solve.f90:109-115 (fcn2)
struct PoorlyScaled2Fcn {
    static constexpr int ID = FCN_POORLY_SCALED, M = 2, N = 2, SYS_LEN = 0, SHARED_LEN = 0;
    static constexpr bool HAS_JAC = false;
    NLB_DEV static void eval(const double (&x)[2], double (&f)[2], const SysCtx&) {
        f[0] = x[1] - 10.0;
        f[1] = x[0] * x[1] - 5.0e4;
    }
    NLB_DEV static void jac(const double (&)[2], JacView<2>, const SysCtx&) {}
};

// Powell's badly scaled function: tests/powell_badly_scaled.f90:9-27
struct PowellBadlyScaled {
    static constexpr int ID = FCN_POWELL_BADLY_SCALED, M = 2, N = 2, SYS_LEN = 0, SHARED_LEN = 0;
    static constexpr bool HAS_JAC = true;
    NLB_DEV static void eval(const double (&x)[2], double (&f)[2], const SysCtx&) {
        f[0] = 1.0e4 * x[0] * x[1] - 1.0;
        f[1] = nl_exp(-x[0]) + nl_exp(-x[1]) - 1.0001;
    }
    NLB_DEV static void jac(const double (&x)[2], JacView<2> J, const SysCtx&) {
        J(0, 0) = 1.0e4 * x[1];
        J(1, 0) = -nl_exp(-x[0]);
        J(0, 1) = 1.0e4 * x[0];
        J(1, 1) = -nl_exp(-x[1]);
    }
};

// 2 x1 - x2 = exp(-x1), -x1 + 2 x2 = exp(-x2): examples/example_problems.f90 misc_2fcn_01(+_jac)
struct Misc2Fcn01 {
    static constexpr int ID = FCN_MISC_2FCN_01, M = 2, N = 2, SYS_LEN = 0, SHARED_LEN = 0;
    static constexpr bool HAS_JAC = true;
    NLB_DEV static void eval(const double (&x)[2], double (&f)[2], const SysCtx&) {
        f[0] = 2.0 * x[0] - x[1] - nl_exp(-x[0]);
        f[1] = -x[0] + 2.0 * x[1] - nl_exp(-x[1]);
    }
    NLB_DEV static void jac(const double (&x)[2], JacView<2> J, const SysCtx&) {
        J(0, 0) = nl_exp(-x[0]) + 2.0;
        J(1, 0) = -1.0;
        J(0, 1) = -1.0;
        J(1, 1) = nl_exp(-x[1]) + 2.0;
    }
};

// polar -> Cartesian maps of the reference's Jacobian tests (tests/nonlin_test_jacobian.f90
// fcn1/jac1 and fcn2/jac2).  They call cos/sin, so they are outside the bitwise-parity set;
// the reference itself checks them to 1e-4 only.
struct Polar {
    static constexpr int ID = FCN_POLAR, M = 2, N = 2, SYS_LEN = 0, SHARED_LEN = 0;
    static constexpr bool HAS_JAC = true;
    NLB_DEV static void eval(const double (&x)[2], double (&f)[2], const SysCtx&) {
        f[0] = x[0] * cos(x[1]);
        f[1] = x[0] * sin(x[1]);
    }
    NLB_DEV static void jac(const double (&x)[2], JacView<2> J, const SysCtx&) {
        J(0, 0) = cos(x[1]);
        J(1, 0) = sin(x[1]);
        J(0, 1) = -x[0] * sin(x[1]);
        J(1, 1) = x[0] * cos(x[1]);
    }
};
struct PolarScaled {
    static constexpr int ID = FCN_POLAR_SCALED, M = 2, N = 2, SYS_LEN = 1, SHARED_LEN = 0;
    static constexpr bool HAS_JAC = true;
    NLB_DEV static void eval(const double (&x)[2], double (&f)[2], const SysCtx& c) {
        const double y = c.sysv(0);
        f[0] = y * x[0] * cos(x[1]);
        f[1] = y * x[0] * sin(x[1]);
    }
    NLB_DEV static void jac(const double (&x)[2], JacView<2> J, const SysCtx& c) {
        const double y = c.sysv(0);
        J(0, 0) = y * cos(x[1]);
        J(1, 0) = y * sin(x[1]);
        J(0, 1) = -y * x[0] * sin(x[1]);
        J(1, 1) = y * x[0] * cos(x[1]);
    }
};

// abscissae of README Example 2 (examples/example_problems.f90 lsq_poly_fit_fcn): decimal
// literals, not i*0.1
__device__ __constant__ const double kPolyFitXp[21] = {0.0, 0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9, 1.0,
                                                       1.1, 1.2, 1.3, 1.4, 1.5, 1.6, 1.7, 1.8, 1.9, 2.0};

// f = x1 xp**3 + x2 xp**2 + x3 xp + x4 - yp, 21 points, per-system yp:
// tests/nonlin_test_solve.f90:133-159 (lsfcn1) = README Example 2
struct LsqPolyFit {
    static constexpr int ID = FCN_LSQ_POLY_FIT, M = 21, N = 4, SYS_LEN = 21, SHARED_LEN = 0;
    static constexpr bool HAS_JAC = false;
    NLB_DEV static void eval(const double (&x)[4], double (&f)[21], const SysCtx& c) {
#pragma unroll 1
        for (int i = 0; i < 21; ++i) {
            const double xp = kPolyFitXp[i];
            const double yp = c.sysv(i);
            f[i] = x[0] * ((xp * xp) * xp) + x[1] * (xp * xp) + x[2] * xp + x[3] - yp;
        }
    }
    NLB_DEV static void jac(const double (&)[4], JacView<21>, const SysCtx&) {}
    // row-wise form of eval() for the kernels that split a system's rows over several lanes (quad_lm.cuh):
    // the same expression, one observation at a time
    NLB_DEV static double abscissa(int i, const double*) { return kPolyFitXp[i]; }
    NLB_DEV static double row(const double (&x)[4], double xp, double yp) {
        return x[0] * ((xp * xp) * xp) + x[1] * (xp * xp) + x[2] * xp + x[3] - yp;
    }
};

// ---- run-time sized families (cooperative kernels) --------------------------------------
// Curve-fit models give the residual of one observation (t, y); the kernels own the loop
// over observations.

// rational 7/8 in Horner form, x = [p0..p7, q0..q7]   (BASELINE config 4 parity model)
struct Rational78 {
    static constexpr int ID = FCN_RATIONAL_7_8, N = 16;
    NLB_DEV static double residual(const double* x, double t, double y) {
        double num = x[7];
#pragma unroll
        for (int k = 6; k >= 0; --k) num = num * t + x[k];
        double den = x[15];
#pragma unroll
        for (int k = 14; k >= 8; --k) den = den * t + x[k];
        den = 1.0 + t * den;
        return num / den - y;
    }
    // residual(x with x[c] replaced by xc, t, y): the same operations in the same order, the replaced parameter chosen
    // by a select so that a rolled loop over c keeps x in registers (forward-difference columns of tall_lm.cuh)
    NLB_DEV static double residual_pert(const double* x, int c, double xc, double t, double y) {
        double num = (c == 7) ? xc : x[7];
#pragma unroll
        for (int k = 6; k >= 0; --k) num = num * t + ((c == k) ? xc : x[k]);
        double den = (c == 15) ? xc : x[15];
#pragma unroll
        for (int k = 14; k >= 8; --k) den = den * t + ((c == k) ? xc : x[k]);
        den = 1.0 + t * den;
        return num / den - y;
    }
    // The residual is a quotient of two Horner chains and a forward-difference column changes a coefficient of one of
    // them only: parts() gives the two chains of the unperturbed point, residual_pert_parts() re-evaluates the chain that
    // contains x[c] and reuses the other (the same operations on the same values: the same bits as residual_pert).
    static constexpr bool SPLIT = true;
    NLB_DEV static void parts(const double* x, double t, double& num, double& den) {
        num = x[7];
#pragma unroll
        for (int k = 6; k >= 0; --k) num = num * t + x[k];
        den = x[15];
#pragma unroll
        for (int k = 14; k >= 8; --k) den = den * t + x[k];
        den = 1.0 + t * den;
    }
    // PART = 0: x[c] is a numerator coefficient (c < SPLIT_AT), PART = 1: a denominator coefficient
    static constexpr int SPLIT_AT = 8;
    template <int PART>
    NLB_DEV static double residual_pert_part(const double* x, int c, double xc, double t, double y, double num0, double den0) {
        double num = num0, den = den0;
        if constexpr (PART == 0) {
            num = (c == 7) ? xc : x[7];
#pragma unroll
            for (int k = 6; k >= 0; --k) num = num * t + ((c == k) ? xc : x[k]);
        } else {
            den = (c == 15) ? xc : x[15];
#pragma unroll
            for (int k = 14; k >= 8; --k) den = den * t + ((c == k) ? xc : x[k]);
            den = 1.0 + t * den;
        }
        return num / den - y;
    }
};

// sum of 8 exponentials, x = [a0..a7, b0..b7]         (BASELINE config 4 throughput model)
struct ExpSum8 {
    static constexpr int ID = FCN_EXP_SUM_8, N = 16;
    static constexpr bool SPLIT = false;
    NLB_DEV static void parts(const double*, double, double&, double&) {}
    static constexpr int SPLIT_AT = 16;
    template <int PART>
    NLB_DEV static double residual_pert_part(const double* x, int c, double xc, double t, double y, double, double) {
        return residual_pert(x, c, xc, t, y);
    }
    NLB_DEV static double residual(const double* x, double t, double y) { return residual_pert(x, -1, 0.0, t, y); }
    // one exponential per trip of a rolled loop (eight inlined copies of nl_exp cost 20 k instructions per kernel);
    // the parameters are picked out of registers by select chains
    NLB_DEV static double residual_pert(const double* x, int c, double xc, double t, double y) {
        double s = 0.0;
#pragma unroll 1
        for (int k = 0; k < 8; ++k) {
            double a = x[0], b = x[8];
#pragma unroll
            for (int u = 1; u < 8; ++u) { a = (k == u) ? x[u] : a; b = (k == u) ? x[8 + u] : b; }
            a = (c == k) ? xc : a;
            b = (c == 8 + k) ? xc : b;
            s += a * nl_exp(-(b * t));
        }
        return s - y;
    }
};

// y = x1 exp(-x2 t) + x3 exp(-x4 t)                   (4-parameter curve fit)
struct ExpDecay4 {
    static constexpr int ID = FCN_EXP_DECAY_4, N = 4;
    NLB_DEV static double residual(const double* x, double t, double y) {
        return x[0] * nl_exp(-(x[1] * t)) + x[2] * nl_exp(-(x[3] * t)) - y;
    }
};

// extended Rosenbrock, f(2i-1) = 10 (x(2i) - x(2i-1)**2), f(2i) = 1 - x(2i-1)   (config 5)
struct ExtRosenbrock {
    static constexpr int ID = FCN_EXT_ROSENBROCK;
    // residual component i (0-based) from the vector x
    NLB_DEV static double component(const double* x, int i) {
        return (i & 1) ? (1.0 - x[i - 1]) : (10.0 * (x[i + 1] - x[i] * x[i]));
    }
};

struct FcnInfo {
    const char* name;
    int m, n;          // 0 = run-time sized
    int sys_len;       // -1 = m
    int shared_len;    // -1 = m
    int has_jac;
};

inline const FcnInfo* fcn_table() {
    static const FcnInfo t[FCN_COUNT] = {
        {"misc_2fcn", 2, 2, 0, 0, 1},
        {"misc_2fcn_a", 2, 2, 1, 0, 1},
        {"poorly_scaled_2fcn", 2, 2, 0, 0, 0},
        {"powell_badly_scaled", 2, 2, 0, 0, 1},
        {"lsq_poly_fit", 21, 4, 21, 0, 0},
        {"polar", 2, 2, 0, 0, 1},
        {"polar_scaled", 2, 2, 1, 0, 1},
        {"misc_2fcn_01", 2, 2, 0, 0, 1},
        {"rational_7_8", 0, 16, -1, -1, 0},
        {"exp_sum_8", 0, 16, -1, -1, 0},
        {"ext_rosenbrock", 0, 0, 0, 0, 0},
        {"exp_decay_4", 0, 4, -1, -1, 0},
    };
    return t;
}

}  // namespace nlb
