// oracle/nl_problems.cpp — TEST INFRASTRUCTURE ONLY.  Residual functions, see nl_problems.h.
#include "nl_problems.h"

#include <cstring>

namespace nlo {

#define J(i, j) jac[((i) - 1) + ((j) - 1) * (long)c->m]

const double NL_POLYFIT_XP[21] = {0.0, 0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9, 1.0,
                                  1.1, 1.2, 1.3, 1.4, 1.5, 1.6, 1.7, 1.8, 1.9, 2.0};
const double NL_POLYFIT_YP[21] = {1.216737514, 1.250032542, 1.305579195, 1.040182335, 1.751867738,
                                  1.109716707, 2.018141531, 1.992418729, 1.807916923, 2.078806005,
                                  2.698801324, 2.644662712, 3.412756702, 4.406137221, 4.567156645,
                                  4.999550779, 5.652854194, 6.784320119, 8.307936836, 8.395126494,
                                  10.30252404};

// x**2 + y**2 = 34 ; x**2 - 2 y**2 = 7          (tests/nonlin_test_solve.f90:41-47)
static void misc_2fcn(const real* x, real* f, const FcnCtx*) {
    f[0] = x[0] * x[0] + x[1] * x[1] - real(34.0);
    f[1] = x[0] * x[0] - real(2.0) * (x[1] * x[1]) - real(7.0);
}
// j = 2 * reshape([x1, x1, x2, -2 x2], [2,2])   (tests/nonlin_test_solve.f90:67-72)
static void misc_2fcn_jac(const real* x, real* jac, const FcnCtx* c) {
    J(1, 1) = real(2.0) * x[0];
    J(2, 1) = real(2.0) * x[0];
    J(1, 2) = real(2.0) * x[1];
    J(2, 2) = real(2.0) * (real(-2.0) * x[1]);
}

// same with the coefficient passed through args   (tests/nonlin_test_solve.f90:49-60, :74-84)
static void misc_2fcn_a(const real* x, real* f, const FcnCtx* c) {
    real a = c->sys[0];
    f[0] = x[0] * x[0] + x[1] * x[1] - real(34.0);
    f[1] = x[0] * x[0] - a * (x[1] * x[1]) - real(7.0);
}
static void misc_2fcn_a_jac(const real* x, real* jac, const FcnCtx* c) {
    real a = c->sys[0];
    J(1, 1) = real(2.0) * x[0];
    J(2, 1) = real(2.0) * x[0];
    J(1, 2) = real(2.0) * x[1];
    J(2, 2) = real(2.0) * (-a * x[1]);
}

// x2 - 10 = 0 ; x1 x2 - 5e4 = 0                   (tests/nonlin_test_solve.f90:109-115)
static void poorly_scaled(const real* x, real* f, const FcnCtx*) {
    f[0] = x[1] - real(10.0);
    f[1] = x[0] * x[1] - real(5.0e4);
}

// Powell's badly scaled function                  (tests/powell_badly_scaled.f90:9-27)
static void powell(const real* x, real* f, const FcnCtx*) {
    f[0] = real(1.0e4) * x[0] * x[1] - real(1.0);
    f[1] = f_exp(-x[0]) + f_exp(-x[1]) - real(1.0001);
}
static void powell_jac(const real* x, real* jac, const FcnCtx* c) {
    J(1, 1) = real(1.0e4) * x[1];
    J(2, 1) = -f_exp(-x[0]);
    J(1, 2) = real(1.0e4) * x[0];
    J(2, 2) = -f_exp(-x[1]);
}

// f = x1 xp**3 + x2 xp**2 + x3 xp + x4 - yp       (tests/nonlin_test_solve.f90:133-159)
static void lsq_poly_fit(const real* x, real* f, const FcnCtx* c) {
    for (int i = 0; i < 21; ++i) {
        real xp = NL_POLYFIT_XP[i];
        real yp = c->sys ? c->sys[i] : real(NL_POLYFIT_YP[i]);
        f[i] = x[0] * ((xp * xp) * xp) + x[1] * (xp * xp) + x[2] * xp + x[3] - yp;
    }
}

// polar -> Cartesian                              (tests/nonlin_test_jacobian.f90 fcn1/jac1, fcn2/jac2)
static void polar(const real* x, real* f, const FcnCtx*) {
    real r = x[0], th = x[1];
    f[0] = r * real(std::cos(dval(th)));
    f[1] = r * real(std::sin(dval(th)));
}
static void polar_jac(const real* x, real* jac, const FcnCtx* c) {
    real r = x[0], th = x[1];
    real cs = real(std::cos(dval(th))), sn = real(std::sin(dval(th)));
    J(1, 1) = cs;
    J(2, 1) = sn;
    J(1, 2) = -r * sn;
    J(2, 2) = r * cs;
}
static void polar_scaled(const real* x, real* f, const FcnCtx* c) {
    real r = x[0], th = x[1], y = c->sys[0];
    f[0] = y * r * real(std::cos(dval(th)));
    f[1] = y * r * real(std::sin(dval(th)));
}
static void polar_scaled_jac(const real* x, real* jac, const FcnCtx* c) {
    real r = x[0], th = x[1], y = c->sys[0];
    real cs = real(std::cos(dval(th))), sn = real(std::sin(dval(th)));
    J(1, 1) = y * cs;
    J(2, 1) = y * sn;
    J(1, 2) = -y * r * sn;
    J(2, 2) = y * r * cs;
}

// 2 x1 - x2 = exp(-x1) ; -x1 + 2 x2 = exp(-x2)    (examples/example_problems.f90 misc_2fcn_01)
static void misc_2fcn_01(const real* x, real* f, const FcnCtx*) {
    f[0] = real(2.0) * x[0] - x[1] - f_exp(-x[0]);
    f[1] = -x[0] + real(2.0) * x[1] - f_exp(-x[1]);
}
static void misc_2fcn_01_jac(const real* x, real* jac, const FcnCtx* c) {
    J(1, 1) = f_exp(-x[0]) + real(2.0);
    J(2, 1) = real(-1.0);
    J(1, 2) = real(-1.0);
    J(2, 2) = f_exp(-x[1]) + real(2.0);
}

// rational 7/8 model in Horner form, x = [p0..p7, q0..q7]:
//   y(t) = (p0 + p1 t + ... + p7 t^7) / (1 + t (q0 + q1 t + ... + q7 t^7))
// shared = t[m], sys = y[m]                        
static void rational_7_8(const real* x, real* f, const FcnCtx* c) {
    for (int i = 0; i < c->m; ++i) {
        real t = c->shared[i];
        real num = x[7];
        for (int k = 6; k >= 0; --k) num = num * t + x[k];
        real den = x[15];
        for (int k = 14; k >= 8; --k) den = den * t + x[k];
        den = real(1.0) + t * den;
        f[i] = num / den - c->sys[i];
    }
}

// sum of 8 exponentials, x = [a0..a7, b0..b7]: y(t) = sum_k a_k exp(-b_k t)
static void exp_sum_8(const real* x, real* f, const FcnCtx* c) {
    for (int i = 0; i < c->m; ++i) {
        real t = c->shared[i];
        real s = 0.0;
        for (int k = 0; k < 8; ++k) s += x[k] * f_exp(-(x[8 + k] * t));
        f[i] = s - c->sys[i];
    }
}

// extended Rosenbrock: f(2i-1) = 10 (x(2i) - x(2i-1)**2), f(2i) = 1 - x(2i-1)
static void ext_rosenbrock(const real* x, real* f, const FcnCtx* c) {
    for (int i = 0; i + 1 < c->n; i += 2) {
        f[i] = real(10.0) * (x[i + 1] - x[i] * x[i]);
        f[i + 1] = real(1.0) - x[i];
    }
}

// 4-parameter double exponential: y(t) = x1 exp(-x2 t) + x3 exp(-x4 t); shared = t[m], sys = y[m]
static void exp_decay_4(const real* x, real* f, const FcnCtx* c) {
    for (int i = 0; i < c->m; ++i) {
        real t = c->shared[i];
        f[i] = x[0] * f_exp(-(x[1] * t)) + x[2] * f_exp(-(x[3] * t)) - c->sys[i];
    }
}

static const Problem g_problems[NL_FCN_COUNT] = {
    {NL_FCN_MISC_2FCN, "misc_2fcn", 2, 2, 0, 0, misc_2fcn, misc_2fcn_jac},
    {NL_FCN_MISC_2FCN_A, "misc_2fcn_a", 2, 2, 1, 0, misc_2fcn_a, misc_2fcn_a_jac},
    {NL_FCN_POORLY_SCALED, "poorly_scaled_2fcn", 2, 2, 0, 0, poorly_scaled, nullptr},
    {NL_FCN_POWELL_BADLY_SCALED, "powell_badly_scaled", 2, 2, 0, 0, powell, powell_jac},
    {NL_FCN_LSQ_POLY_FIT, "lsq_poly_fit", 21, 4, 21, 0, lsq_poly_fit, nullptr},
    {NL_FCN_POLAR, "polar", 2, 2, 0, 0, polar, polar_jac},
    {NL_FCN_POLAR_SCALED, "polar_scaled", 2, 2, 1, 0, polar_scaled, polar_scaled_jac},
    {NL_FCN_MISC_2FCN_01, "misc_2fcn_01", 2, 2, 0, 0, misc_2fcn_01, misc_2fcn_01_jac},
    {NL_FCN_RATIONAL_7_8, "rational_7_8", 0, 16, -1, -1, rational_7_8, nullptr},
    {NL_FCN_EXP_SUM_8, "exp_sum_8", 0, 16, -1, -1, exp_sum_8, nullptr},
    {NL_FCN_EXT_ROSENBROCK, "ext_rosenbrock", 0, 0, 0, 0, ext_rosenbrock, nullptr},
    {NL_FCN_EXP_DECAY_4, "exp_decay_4", 0, 4, -1, -1, exp_decay_4, nullptr},
};

// ---- callback residuals (one trampoline per slot: vecfcn_t carries no user data) ---------------------------------
static callback_t g_cb[NL_MAX_CALLBACKS] = {nullptr, nullptr, nullptr, nullptr};
static Problem g_cb_problem[NL_MAX_CALLBACKS];
static char g_cb_name[NL_MAX_CALLBACKS][64];
static int g_cb_count = 0;

template <int K>
static void cb_trampoline(const real* x, real* f, const FcnCtx* c) {
    double xd[64], fd[64];
    for (int j = 0; j < c->n; ++j) xd[j] = dval(x[j]);
    const Problem& p = g_cb_problem[K];
    double sd[64], hd[64];
    const int ns = p.sys_len < 0 ? c->m : p.sys_len, nh = p.shared_len < 0 ? c->m : p.shared_len;
    for (int k = 0; k < ns && c->sys; ++k) sd[k] = dval(c->sys[k]);
    for (int k = 0; k < nh && c->shared; ++k) hd[k] = dval(c->shared[k]);
    g_cb[K](xd, fd, c->sys ? sd : nullptr, c->shared ? hd : nullptr, c->m, c->n);
    for (int i = 0; i < c->m; ++i) f[i] = real(fd[i]);
}

int nl_register_callback(const char* name, int m, int n, int sys_len, int shared_len, callback_t fcn) {
    static const vecfcn_t tramp[NL_MAX_CALLBACKS] = {cb_trampoline<0>, cb_trampoline<1>, cb_trampoline<2>, cb_trampoline<3>};
    if (!name || !fcn || m <= 0 || n <= 0 || m > 64 || n > 64 || sys_len > 64 || shared_len > 64) return -1;
    for (int k = 0; k < g_cb_count; ++k)
        if (std::strcmp(g_cb_name[k], name) == 0) { g_cb[k] = fcn; return NL_FCN_COUNT + k; }   // re-registration
    if (g_cb_count >= NL_MAX_CALLBACKS) return -1;
    const int k = g_cb_count++;
    std::strncpy(g_cb_name[k], name, sizeof(g_cb_name[k]) - 1);
    g_cb[k] = fcn;
    g_cb_problem[k] = Problem{NL_FCN_COUNT + k, g_cb_name[k], m, n, sys_len, shared_len, tramp[k], nullptr};
    return NL_FCN_COUNT + k;
}

const Problem* nl_problem(int id) {
    if (id >= NL_FCN_COUNT && id < NL_FCN_COUNT + g_cb_count) return &g_cb_problem[id - NL_FCN_COUNT];
    if (id < 0 || id >= NL_FCN_COUNT) return nullptr;
    return &g_problems[id];
}

const Problem* nl_problem_by_name(const char* name) {
    for (int i = 0; i < NL_FCN_COUNT; ++i)
        if (std::strcmp(g_problems[i].name, name) == 0) return &g_problems[i];
    for (int k = 0; k < g_cb_count; ++k)
        if (std::strcmp(g_cb_name[k], name) == 0) return &g_cb_problem[k];
    return nullptr;
}

}  // namespace nlo
