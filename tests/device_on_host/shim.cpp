// tests/device_on_host/shim.cpp — TEST INFRASTRUCTURE ONLY.
//
// Compiles the engine's thread-per-system DEVICE SOURCE (nonlin_b200/csrc/*.cuh) for the host with g++
// (-ffp-contract=off, no fast-math) and runs it one "lane" at a time, so that the CPU test suite can compare the
// very code the GPU executes with the oracle, bit for bit, without a GPU.  This is not a CPU path of the product:
// it is built and loaded only by tests/test_device_source_on_host.py, lives under tests/, and libnonlin_b200.so
// knows nothing about it.  What it cannot cover: the cooperative (CTA / warp) kernels, launch geometry, the
// shared-memory variants, and nvcc's code generation itself - the `-m gpu` tests remain the parity tests proper.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include <cuda_runtime.h>   // host pass: __device__, __forceinline__, __global__ ... become ignorable attributes
#ifndef __noinline__
#define __noinline__ __attribute__((noinline))
#endif
#ifndef __launch_bounds__
#define __launch_bounds__(...)
#endif

// ---- the handful of device intrinsics the thread-per-system sources use, for a single lane -------------------
static inline double __ldg(const double* p) { return *p; }
static inline double __longlong_as_double(long long v) { double d; std::memcpy(&d, &v, 8); return d; }
static inline long long __double_as_longlong(double d) { long long v; std::memcpy(&v, &d, 8); return v; }
static inline unsigned __activemask() { return 1u; }
static inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { unsigned long long o = *p; *p += v; return o; }
static inline unsigned long long __shfl_sync(unsigned, unsigned long long v, int) { return v; }
static inline void __syncwarp() {}
static inline int __any_sync(unsigned, int pred) { return pred; }
struct dh_dim3 { unsigned x, y, z; };
static dh_dim3 threadIdx = {0, 0, 0}, blockIdx = {0, 0, 0}, blockDim = {1, 1, 1}, gridDim = {1, 1, 1};
double pf_smem[1];          // polyfit_kernel's `extern __shared__` array (unused: the global-workspace variant runs)

#include "../../nonlin_b200/csrc/tps_lm.cuh"
#include "../../nonlin_b200/csrc/tps_newton_broyden.cuh"
#include "../../nonlin_b200/csrc/tps_cls.cuh"
#include "../../nonlin_b200/csrc/polyfit.cuh"
#include "../../nonlin_b200/csrc/scalar_solvers.cuh"

using namespace nlb;

static DevParams to_dev(const nlb_params* p) {
    DevParams d;
    d.max_fcn_evals = p->max_fcn_evals; d.fcn_tol = p->fcn_tol; d.var_tol = p->var_tol; d.grad_tol = p->grad_tol;
    d.lm_factor = p->lm_factor; d.jacobian_interval = p->jacobian_interval; d.use_line_search = p->use_line_search;
    d.ls_max_fcn_evals = p->ls_max_fcn_evals; d.ls_alpha = p->ls_alpha; d.ls_factor = p->ls_factor;
    d.use_analytic_jacobian = p->use_analytic_jacobian; d.max_iter_guard = p->max_iter_guard;
    return d;
}

static void put_ib(nlb_iteration_behavior* ib, long long b, const SolveStats& st) {
    if (!ib) return;
    nlb_iteration_behavior o;
    o.iter_count = st.iter; o.fcn_count = st.nfev; o.jacobian_count = st.njac; o.gradient_count = 0;
    o.converge_on_fcn = st.cf; o.converge_on_chng = st.cx; o.converge_on_zero_diff = st.cg;
    ib[b] = o;
}

// solver: 0 LM, 1 Newton (per-system form), 2 Broyden, 4 Newton (persistent refill form, one lane)
template <class F>
static int solve_fixed(int solver, const DevParams& p, long long B, double* x, double* fvec, const double* sys,
                       const double* shared, nlb_iteration_behavior* ib, int32_t* status) {
    constexpr int M = F::M, N = F::N;
    if (solver == 4) {
        if constexpr (M == N) {
            unsigned long long cursor = 0;
            tps_newton_refill<F>(p, B, B, &cursor, x, fvec, sys, shared, ib, status);
            return 0;
        } else {
            return 3;
        }
    }
    for (long long b = 0; b < B; ++b) {
        double xl[N], fl[M];
        for (int j = 0; j < N; ++j) xl[j] = x[j * B + b];
        SysCtx c{sys ? sys + b : nullptr, shared, B, M, N};
        SolveStats st;
        if (solver == 0) tps_lm_solve<F>(p, c, xl, fl, st);
        else if constexpr (M == N) {
            if (solver == 1) tps_newton_solve<F>(p, c, xl, fl, st);
            else tps_broyden_solve<F>(p, c, xl, fl, st);
        } else {
            return 3;
        }
        for (int j = 0; j < N; ++j) x[j * B + b] = xl[j];
        for (int i = 0; i < M; ++i) fvec[i * B + b] = fl[i];
        put_ib(ib, b, st);
        if (status) status[b] = st.status;
    }
    return 0;
}

#define DH_FIXED(X) X(Misc2Fcn) X(Misc2FcnA) X(PoorlyScaled2Fcn) X(PowellBadlyScaled) X(Misc2Fcn01) X(LsqPolyFit)

extern "C" {

int dh_solve(int solver, int fcn_id, long long B, const nlb_params* prm, double* x, double* fvec, const double* sys,
             const double* shared, nlb_iteration_behavior* ib, int32_t* status) {
    const DevParams p = to_dev(prm);
    switch (fcn_id) {
#define X(F) case F::ID: return solve_fixed<F>(solver, p, B, x, fvec, sys, shared, ib, status);
        DH_FIXED(X)
#undef X
    }
    return 2;
}

int dh_cls_solve(int fcn_id, long long B, const nlb_params* prm, double radius, double scaling, const double* lower,
                 const double* upper, double* x, double* fvec, const double* sys, const double* shared,
                 nlb_iteration_behavior* ib, int32_t* status) {
    const DevParams p = to_dev(prm);
    DevCls o;
    o.radius = radius > 0.0 ? radius : 1.0;
    o.scaling = scaling > 0.0 ? scaling : 1.0;
    const double huge = 1.7976931348623157e+308;
    for (int i = 0; i < CLS_MAX_N; ++i) { o.xl[i] = lower ? lower[i] : -huge; o.xu[i] = upper ? upper[i] : huge; }
    unsigned long long cursor = 0;
    switch (fcn_id) {
#define X(F) case F::ID: tps_cls_refill<F>(p, o, B, B, &cursor, x, fvec, sys, shared, ib, status); return 0;
        DH_FIXED(X)
#undef X
    }
    return 2;
}

int dh_polyfit(long long B, int npts, int order, int thru_zero, int x_shared, const double* x, const double* y,
               double* coeffs, int32_t* status) {
    const int nc = thru_zero ? order : order + 1;
    std::vector<double> work((size_t)npts * (nc + 1));
    switch (nc) {
#define X(NC) case NC: polyfit_kernel<NC, false>(B, npts, thru_zero, x_shared, x, y, coeffs, status, work.data()); return 0;
        X(1) X(2) X(3) X(4) X(5) X(6) X(7) X(8)
#undef X
    }
    return 2;
}

int dh_solve_1var(int solver, int fcn_id, long long B, const nlb_params_1var* prm, const double* lim1,
                  const double* lim2, double* x, double* f, const double* args, nlb_iteration_behavior* ib,
                  int32_t* status) {
    DevParams1 p;
    p.max_fcn_evals = prm->max_fcn_evals; p.fcn_tol = prm->fcn_tol; p.var_tol = prm->var_tol;
    p.diff_tol = prm->diff_tol; p.use_analytic_diff = prm->use_analytic_diff;
    for (long long b = 0; b < B; ++b) {
        blockIdx.x = (unsigned)b;      // the kernel derives its equation index from blockIdx * blockDim + threadIdx
        switch (fcn_id) {
#define X(F)                                                                                              \
    case F::ID:                                                                                           \
        if (solver == 0) solve_1var_kernel<F, 0>(p, B, lim1, lim2, x, f, args, ib, status);               \
        else solve_1var_kernel<F, 1>(p, B, lim1, lim2, x, f, args, ib, status);                           \
        break;
            X(CubicWallis) X(ExpMinusX) X(CubicArgs)
#undef X
            default: return 2;
        }
    }
    blockIdx.x = 0;
    return 0;
}

}  // extern "C"
