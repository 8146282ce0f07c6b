// nlb_plugin.cuh — residuals compiled OUTSIDE the engine: the batch analogue of vecfcn_helper%set_fcn
// (reference src/nonlin_multi_eqn_mult_var.f90:126-140), which takes any procedure at run time.  A __device__ function
// pointer cannot cross a library boundary, so a new residual is a small plug-in library instead of an edit of the engine:
//
//     // my_residual.cu
//     #include "nlb_plugin.cuh"
//     struct MyFcn {
//         static constexpr int M = 2, N = 2, SYS_LEN = 0, SHARED_LEN = 0;     // equations, unknowns, per-system / shared doubles
//         static constexpr bool HAS_JAC = false;
//         NLB_DEV static void eval(const double (&x)[N], double (&f)[M], const nlb::SysCtx& c) { ... }
//         NLB_DEV static void jac(const double (&x)[N], nlb::JacView<M> J, const nlb::SysCtx& c) {}
//     };
//     NLB_PLUGIN_BEGIN
//         NLB_PLUGIN_VECFCN(MyFcn, "my_fcn")
//     NLB_PLUGIN_END
//
//     nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false -O3 -std=c++17 -shared -Xcompiler -fPIC \
//          -I <repo>/nonlin_b200/csrc my_residual.cu -o libmy_residual.so
//
// and the host calls nlb_load_plugin("libmy_residual.so") once (Python: nonlin_b200.load_plugin), after which "my_fcn"
// resolves through nlb_vecfcn_lookup like a built-in residual and runs through the same thread-per-system kernels
// (least squares, Newton, quasi-Newton, residual and Jacobian evaluation), with the same -fmad=false arithmetic.
#pragma once
#include "tps_kernels.cuh"

namespace nlb {

template <class F>
int plugin_solve(int solver, const nlb_params* params, int64_t nsys, int64_t B, double* x, double* fvec, const double* sys,
                 const double* shared, nlb_iteration_behavior* ib, int32_t* status, void* stream) {
    if (nsys <= 0) return NLB_OK;
    const DevParams p = to_dev(params);
    cudaStream_t s = (cudaStream_t)stream;
    const unsigned grid = (unsigned)((nsys + TPS_BLOCK - 1) / TPS_BLOCK);
    if (solver == SOLVER_LM) {
        if constexpr (F::M >= F::N)
            tps_solve_kernel<F, SOLVER_LM><<<grid, TPS_BLOCK, 0, s>>>(p, nsys, B, x, fvec, sys, shared, ib, status);
        else
            return NLB_ERR_SIZE;
    } else {
        if constexpr (F::M == F::N) {
            if (solver == SOLVER_NEWTON)
                tps_solve_kernel<F, SOLVER_NEWTON><<<grid, TPS_BLOCK, 0, s>>>(p, nsys, B, x, fvec, sys, shared, ib, status);
            else if (solver == SOLVER_BROYDEN)
                tps_solve_kernel<F, SOLVER_BROYDEN><<<grid, TPS_BLOCK, 0, s>>>(p, nsys, B, x, fvec, sys, shared, ib, status);
            else
                return NLB_ERR_UNSUPPORTED;
        } else {
            return NLB_ERR_SIZE;
        }
    }
    return cudaGetLastError() == cudaSuccess ? NLB_OK : NLB_ERR_CUDA;
}

template <class F>
int plugin_eval(int what, int analytic, int64_t B, const double* x, double* out, const double* sys, const double* shared,
                void* stream) {
    if (B <= 0) return NLB_OK;
    cudaStream_t s = (cudaStream_t)stream;
    const unsigned grid = (unsigned)((B + TPS_BLOCK - 1) / TPS_BLOCK);
    if (what == 0) tps_eval_kernel<F><<<grid, TPS_BLOCK, 0, s>>>(B, x, out, sys, shared);
    else tps_jacobian_kernel<F><<<grid, TPS_BLOCK, 0, s>>>(analytic, B, x, out, sys, shared);
    return cudaGetLastError() == cudaSuccess ? NLB_OK : NLB_ERR_CUDA;
}

}  // namespace nlb

// The plug-in's one exported symbol: the engine passes its registration entry point (nlb_register_vecfcn).
#define NLB_PLUGIN_BEGIN extern "C" int nlb_plugin_register(nlb_register_vecfcn_fn reg) { int count = 0;
#define NLB_PLUGIN_VECFCN(F, NAME)                                                                                   \
    if (reg(NAME, F::M, F::N, F::SYS_LEN, F::SHARED_LEN, F::HAS_JAC ? 1 : 0, &nlb::plugin_solve<F>, &nlb::plugin_eval<F>) >= 0) \
        ++count;
#define NLB_PLUGIN_END return count; }
