// tps_kernels.cuh — the generic thread-per-system kernels (one CUDA thread = one system) and the conversion of the
// C ABI's settings into their device form.  Shared by the engine (nlb_api.cu) and by residual plug-ins
// (nlb_plugin.cuh), which instantiate the same kernels for residuals compiled outside the engine.
#pragma once
#include "tps_lm.cuh"
#include "tps_newton_broyden.cuh"

namespace nlb {

enum Solver { SOLVER_LM = 0, SOLVER_NEWTON = 1, SOLVER_BROYDEN = 2, SOLVER_CLS = 3 };

inline DevParams to_dev(const nlb_params* p) {
    DevParams d;
    d.max_fcn_evals = p->max_fcn_evals;
    d.fcn_tol = p->fcn_tol;
    d.var_tol = p->var_tol;
    d.grad_tol = p->grad_tol;
    d.lm_factor = p->lm_factor;
    d.jacobian_interval = p->jacobian_interval;
    d.use_line_search = p->use_line_search;
    d.ls_max_fcn_evals = p->ls_max_fcn_evals;
    d.ls_alpha = p->ls_alpha;
    d.ls_factor = p->ls_factor;
    d.use_analytic_jacobian = p->use_analytic_jacobian;
    d.max_iter_guard = p->max_iter_guard;
    return d;
}

#ifndef NLB_TPS_BLOCK
#define NLB_TPS_BLOCK 128
#endif
constexpr int TPS_BLOCK = NLB_TPS_BLOCK;
// Minimum resident CTAs per SM asked of ptxas.  Measured on B200 (launch-bounds sweep): capping the
// 2x2 Broyden kernel at 80 registers (6 CTAs/SM) is 11 % faster than 110 registers (4 CTAs/SM); LM is best at
// 3 CTAs/SM (160 registers); Newton does not gain from a cap.
#ifndef NLB_LM_MIN_BLOCKS
#define NLB_LM_MIN_BLOCKS 3
#endif
template <int SOLVER>
constexpr int tps_min_blocks() {
    return SOLVER == 2 ? 6 : (SOLVER == 0 ? NLB_LM_MIN_BLOCKS : 1);
}

template <class F, int SOLVER>
__global__ void __launch_bounds__(TPS_BLOCK, tps_min_blocks<SOLVER>())
tps_solve_kernel(DevParams p, long long nsys, long long B, double* __restrict__ x, double* __restrict__ fvec,
                 const double* __restrict__ sys, const double* __restrict__ shared,
                 nlb_iteration_behavior* __restrict__ ib, int32_t* __restrict__ status) {
    // nsys systems starting at the (pre-offset) pointers; B is the SoA stride of the whole batch
    constexpr int M = F::M, N = F::N;
    const long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (b >= nsys) return;
    double xl[N], fl[M];
#pragma unroll
    for (int j = 0; j < N; ++j) xl[j] = x[j * B + b];
    SysCtx c{sys ? sys + b : nullptr, shared, B, M, N};
    SolveStats st;
    if constexpr (SOLVER == SOLVER_LM) tps_lm_solve<F>(p, c, xl, fl, st);
    else if constexpr (SOLVER == SOLVER_NEWTON) tps_newton_solve<F>(p, c, xl, fl, st);
    else tps_broyden_solve<F>(p, c, xl, fl, st);
#pragma unroll
    for (int j = 0; j < N; ++j) x[j * B + b] = xl[j];
#pragma unroll(M <= 8 ? M : 1)
    for (int i = 0; i < M; ++i) fvec[i * B + b] = fl[i];
    if (ib) {
        nlb_iteration_behavior o;
        o.iter_count = st.iter;
        o.fcn_count = st.nfev;
        o.jacobian_count = st.njac;
        o.gradient_count = 0;
        o.converge_on_fcn = st.cf;
        o.converge_on_chng = st.cx;
        o.converge_on_zero_diff = st.cg;
        ib[b] = o;
    }
    if (status) status[b] = st.status;
}

template <class F>
__global__ void __launch_bounds__(TPS_BLOCK)
tps_eval_kernel(long long B, const double* __restrict__ x, double* __restrict__ fvec,
                const double* __restrict__ sys, const double* __restrict__ shared) {
    constexpr int M = F::M, N = F::N;
    const long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (b >= B) return;
    double xl[N], fl[M];
#pragma unroll
    for (int j = 0; j < N; ++j) xl[j] = x[j * B + b];
    SysCtx c{sys ? sys + b : nullptr, shared, B, M, N};
    F::eval(xl, fl, c);
#pragma unroll(M <= 8 ? M : 1)
    for (int i = 0; i < M; ++i) fvec[i * B + b] = fl[i];
}

template <class F>
__global__ void __launch_bounds__(TPS_BLOCK)
tps_jacobian_kernel(int analytic, long long B, const double* __restrict__ x, double* __restrict__ jac,
                    const double* __restrict__ sys, const double* __restrict__ shared) {
    constexpr int M = F::M, N = F::N;
    const long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (b >= B) return;
    double xl[N], fl[M], wrk[M], jl[M * N];
#pragma unroll
    for (int j = 0; j < N; ++j) xl[j] = x[j * B + b];
    SysCtx c{sys ? sys + b : nullptr, shared, B, M, N};
    F::eval(xl, fl, c);                      // fv not supplied: evaluate f(x) first (multi_eqn:257-259)
    fd_jacobian<F>(xl, jl, fl, wrk, c, analytic != 0);
#pragma unroll(M * N <= 16 ? M * N : 1)
    for (int e = 0; e < M * N; ++e) jac[e * B + b] = jl[e];
}


}  // namespace nlb
