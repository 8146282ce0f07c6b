// Example residual plug-in (see nonlin_b200/csrc/nlb_plugin.cuh): two residuals the engine was not built with.
//   freudenstein_roth   the classic 2 x 2 test system (root (5, 4); also a local minimum of |f| near (11.41, -0.8968))
//   circle_line_a       x^2 + y^2 = a, x - y = 1 with the radius^2 `a` passed per system (the reference's `args`)
// Every expression is written with explicit + - * only, in a fixed order: the CPU oracle evaluates the same operations
// through a callback in tests/test_plugin.py and the results must agree bit for bit.
#include "nlb_plugin.cuh"

struct FreudensteinRoth {
    static constexpr int M = 2, N = 2, SYS_LEN = 0, SHARED_LEN = 0;
    static constexpr bool HAS_JAC = true;
    NLB_DEV static void eval(const double (&x)[2], double (&f)[2], const nlb::SysCtx&) {
        f[0] = -13.0 + x[0] + ((5.0 - x[1]) * x[1] - 2.0) * x[1];
        f[1] = -29.0 + x[0] + ((x[1] + 1.0) * x[1] - 14.0) * x[1];
    }
    NLB_DEV static void jac(const double (&x)[2], nlb::JacView<2> J, const nlb::SysCtx&) {
        J(0, 0) = 1.0;
        J(1, 0) = 1.0;
        J(0, 1) = (10.0 - 3.0 * x[1]) * x[1] - 2.0;
        J(1, 1) = (3.0 * x[1] + 2.0) * x[1] - 14.0;
    }
};

struct CircleLineA {
    static constexpr int M = 2, N = 2, SYS_LEN = 1, SHARED_LEN = 0;
    static constexpr bool HAS_JAC = false;
    NLB_DEV static void eval(const double (&x)[2], double (&f)[2], const nlb::SysCtx& c) {
        const double a = c.sysv(0);
        f[0] = x[0] * x[0] + x[1] * x[1] - a;
        f[1] = x[0] - x[1] - 1.0;
    }
    NLB_DEV static void jac(const double (&)[2], nlb::JacView<2>, const nlb::SysCtx&) {}
};

NLB_PLUGIN_BEGIN
    NLB_PLUGIN_VECFCN(FreudensteinRoth, "freudenstein_roth")
    NLB_PLUGIN_VECFCN(CircleLineA, "circle_line_a")
NLB_PLUGIN_END
