// coop_kernels.cuh — warp- / CTA-per-system kernels for the run-time sized residual families
// (BASELINE configs 4 and 5).  Declarations; definitions in coop_kernels.cu.
#pragma once
#include "tps_common.cuh"

namespace nlb {

// Returns NLB_OK, NLB_ERR_UNSUPPORTED (no kernel for this combination) or NLB_ERR_CUDA.
int launch_coop_solve(int solver, int fcn_id, const DevParams& p, long long nsys, long long B, int m, int n, double* x,
                      double* fvec,
                      const double* sys, const double* shared, nlb_iteration_behavior* ib, int32_t* status,
                      cudaStream_t s, int64_t* launches);

// constrained_least_squares_solver on a curve-fit family (run-time m): csrc/cls_rt.cuh
struct DevCls;
int launch_coop_cls(int fcn_id, const DevParams& p, const DevCls& o, long long nsys, long long B, int m, int n, double* x,
                    double* fvec, const double* sys, const double* shared, nlb_iteration_behavior* ib, int32_t* status,
                    cudaStream_t s, int64_t* launches);

int launch_coop_eval(int fcn_id, long long B, int m, int n, const double* x, double* fvec, const double* sys,
                     const double* shared, cudaStream_t s);

int launch_coop_jacobian(int fcn_id, long long B, int m, int n, const double* x, double* jac, const double* sys,
                         const double* shared, cudaStream_t s);

}  // namespace nlb
