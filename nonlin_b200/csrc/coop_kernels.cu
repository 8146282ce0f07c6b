// coop_kernels.cu — warp- / CTA-per-system kernels (run-time sized residual families).
#include "coop_kernels.cuh"

namespace nlb {

int launch_coop_solve(int, int, const DevParams&, long long, int, int, double*, double*, const double*, const double*,
                      nlb_iteration_behavior*, int32_t*, cudaStream_t, int64_t*) {
    return NLB_ERR_UNSUPPORTED;
}

int launch_coop_eval(int, long long, int, int, const double*, double*, const double*, const double*, cudaStream_t) {
    return NLB_ERR_UNSUPPORTED;
}

}  // namespace nlb
