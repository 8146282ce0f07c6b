// coop_kernels.cu — launchers of the CTA-per-system kernels (run-time sized residual families).
#include "coop_kernels.cuh"

#include "launch_cfg.cuh"
#include "coop_broyden.cuh"
#include "coop_lm.cuh"
#include "coop_lm_cta.cuh"
#include "tall_lm.cuh"
#include "cls_rt.cuh"

#include <cstdlib>

namespace nlb {

namespace {

enum { SOLVER_LM = 0, SOLVER_NEWTON = 1, SOLVER_BROYDEN = 2 };
constexpr int WLM_MIN_M = 512;   // rows from which the CTA-per-system LM kernels are used
#ifndef NLB_TLM_SEG
#define NLB_TLM_SEG 64           // rows per TMA segment of tall_lm.cuh (= producer threads per CTA)
#endif
#ifndef NLB_TLM_STAGES
#define NLB_TLM_STAGES 3         // stages of its input ring (measured: 4 stages were no faster, and cost a resident CTA)
#endif

template <class F, int N>
int launch_broyden(const DevParams& p, long long nsys, long long B, double* x, double* fvec, const double* sys, const double* shared,
                   nlb_iteration_behavior* ib, int32_t* status, cudaStream_t s, int64_t* launches) {
    using S = CoopBroydenSmem<N>;
    KernelCfg cfg;
    if (kernel_cfg<coop_broyden_kernel<F, N>>(N, S::BYTES, &cfg) != cudaSuccess) return NLB_ERR_CUDA;
    coop_broyden_kernel<F, N><<<(unsigned)nsys, N, S::BYTES, s>>>(p, nsys, B, x, fvec, sys, shared, ib, status);
    ++*launches;
    return cudaGetLastError() == cudaSuccess ? NLB_OK : NLB_ERR_CUDA;
}

// one thread per (system, observation): residual of a curve-fit model
template <class F>
__global__ void curvefit_eval_kernel(long long B, int m, const double* __restrict__ x, double* __restrict__ fvec,
                                     const double* __restrict__ sys, const double* __restrict__ shared) {
    const long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (b >= B) return;
    double xl[F::N];
#pragma unroll
    for (int j = 0; j < F::N; ++j) xl[j] = x[j * B + b];
    for (int i = blockIdx.y; i < m; i += gridDim.y) fvec[(long long)i * B + b] = F::residual(xl, __ldg(shared + i), sys[(long long)i * B + b]);
}

__global__ void rosenbrock_eval_kernel(long long B, int n, const double* __restrict__ x, double* __restrict__ fvec) {
    const long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (b >= B) return;
    for (int i = blockIdx.y * 2; i + 1 < n; i += gridDim.y * 2) {
        const double x0 = x[(long long)i * B + b], x1 = x[(long long)(i + 1) * B + b];
        fvec[(long long)i * B + b] = 10.0 * (x1 - x0 * x0);
        fvec[(long long)(i + 1) * B + b] = 1.0 - x0;
    }
}

// vecfcn_helper%jacobian for the curve-fit families: thread (b, j) forms column j of system b,
// jac[(i + j*m)*B + b] = (f_i(x + h e_j) - f_i(x)) / h   (vfh_jac_fcn, multi_eqn:257-275)
template <class F>
__global__ void curvefit_jacobian_kernel(long long B, int m, const double* __restrict__ x, double* __restrict__ jac,
                                         const double* __restrict__ sys, const double* __restrict__ shared) {
    const long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const int j = blockIdx.y;
    if (b >= B) return;
    double x0[F::N], x1[F::N];
#pragma unroll
    for (int c = 0; c < F::N; ++c) x0[c] = x[c * B + b];
    double temp = 0.0;
#pragma unroll
    for (int c = 0; c < F::N; ++c) temp = (c == j) ? x0[c] : temp;
    double h = 0x1p-26 * fabs(temp);
    if (h == 0.0) h = 0x1p-26;
#pragma unroll
    for (int c = 0; c < F::N; ++c) x1[c] = (c == j) ? (temp + h) : x0[c];
    for (int i = 0; i < m; ++i) {
        const double t = __ldg(shared + i), y = sys[(long long)i * B + b];
        jac[((long long)i + (long long)j * m) * B + b] = (F::residual(x1, t, y) - F::residual(x0, t, y)) / h;
    }
}

__global__ void rosenbrock_jacobian_kernel(long long B, int n, const double* __restrict__ x, double* __restrict__ jac) {
    const long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const int j = blockIdx.y;
    if (b >= B) return;
    // column j only touches the two equations of its pair; every other entry is (f - f)/h = 0 exactly
    const int p0 = j & ~1;
    const double xa = x[(long long)p0 * B + b], xb = x[(long long)(p0 + 1) * B + b];
    const double temp = (j & 1) ? xb : xa;
    double h = 0x1p-26 * fabs(temp);
    if (h == 0.0) h = 0x1p-26;
    const double xa1 = (j & 1) ? xa : temp + h, xb1 = (j & 1) ? temp + h : xb;
    for (int i = 0; i < n; ++i) {
        double v = 0.0;
        if (i == p0) v = (10.0 * (xb1 - xa1 * xa1) - 10.0 * (xb - xa * xa)) / h;
        else if (i == p0 + 1) v = ((1.0 - xa1) - (1.0 - xa)) / h;
        else {
            // (f_i(x + h e_j) - f_i(x)) / h with f_i independent of x_j: the difference of two equal values
            v = 0.0 / h;
        }
        jac[((long long)i + (long long)j * n) * B + b] = v;
    }
}

template <class F, int N>
int launch_lm(const DevParams& p, long long ntot, long long B, int m, double* x, double* fvec, const double* sys, const double* shared,
              nlb_iteration_behavior* ib, int32_t* status, cudaStream_t s, int64_t* launches) {
    using S = CoopLmSmem<N>;
    KernelCfg cfg;
    if (kernel_cfg<coop_lm_kernel<F, N>>(32 * N, S::BYTES, &cfg) != cudaSuccess) return NLB_ERR_CUDA;
    // HBM workspace: (n + 3) * m doubles per resident lane, allocated stream-ordered.  Persistent CTAs: only as
    // many as are resident at once; their 32 lanes pull systems from a cursor, so the workspace is sized by the
    // resident lanes (148 SMs x CTAs per SM x 32), not by the batch.
    const size_t per_sys = (size_t)(N + 3) * (size_t)m * sizeof(double);
    long long grid = (ntot + 31) / 32;
    if (grid > (long long)cfg.num_sms * cfg.ctas_per_sm) grid = (long long)cfg.num_sms * cfg.ctas_per_sm;
    double* ws = nullptr;
    if (cudaMallocAsync((void**)&ws, per_sys * 32 * (size_t)grid + 64, s) != cudaSuccess) return NLB_ERR_CUDA;
    unsigned long long* cursor = reinterpret_cast<unsigned long long*>(ws + (size_t)(N + 3) * m * 32 * (size_t)grid);
    if (cudaMemsetAsync(cursor, 0, sizeof(unsigned long long), s) != cudaSuccess) { cudaFreeAsync(ws, s); return NLB_ERR_CUDA; }
    coop_lm_kernel<F, N><<<(unsigned)grid, 32 * N, S::BYTES, s>>>(p, B, 0, ntot, m, x, fvec, sys, shared, ib, status, ws, cursor);
    ++*launches;
    if (cudaGetLastError() != cudaSuccess) { cudaFreeAsync(ws, s); return NLB_ERR_CUDA; }
    if (cudaFreeAsync(ws, s) != cudaSuccess) return NLB_ERR_CUDA;
    return NLB_OK;
}

// Tall systems: CTA-per-system / warp-per-column kernel (coop_lm_cta.cuh), persistent CTAs + work queue.
template <class F, int N>
int launch_wlm(const DevParams& p, long long ntot, long long B, int m, double* x, double* fvec, const double* sys,
               const double* shared, nlb_iteration_behavior* ib, int32_t* status, cudaStream_t s, int64_t* launches) {
    using S = WlmSmem<N>;
    KernelCfg cfg;
    if (kernel_cfg<wlm_kernel<F, N>>(32 * N, S::BYTES, &cfg) != cudaSuccess) return NLB_ERR_CUDA;
    long long grid = ntot;
    if (grid > (long long)cfg.num_sms * cfg.ctas_per_sm) grid = (long long)cfg.num_sms * cfg.ctas_per_sm;
    const size_t per_cta = (size_t)(N + 3) * (size_t)m;          // doubles
    double* ws = nullptr;
    if (cudaMallocAsync((void**)&ws, (per_cta * (size_t)grid + 8) * sizeof(double), s) != cudaSuccess) return NLB_ERR_CUDA;
    unsigned long long* cursor = reinterpret_cast<unsigned long long*>(ws + per_cta * (size_t)grid);
    if (cudaMemsetAsync(cursor, 0, sizeof(unsigned long long), s) != cudaSuccess) { cudaFreeAsync(ws, s); return NLB_ERR_CUDA; }
    wlm_kernel<F, N><<<(unsigned)grid, 32 * N, S::BYTES, s>>>(p, B, ntot, m, x, fvec, sys, shared, ib, status, ws, cursor);
    ++*launches;
    if (cudaGetLastError() != cudaSuccess) { cudaFreeAsync(ws, s); return NLB_ERR_CUDA; }
    if (cudaFreeAsync(ws, s) != cudaSuccess) return NLB_ERR_CUDA;
    return NLB_OK;
}

// Tall systems, TMA-streamed (tall_lm.cuh): persistent CTAs of 32 + S threads, several per SM.
template <class F, int N, int S, int NST>
int launch_tlm(const DevParams& p, long long ntot, long long B, int m, double* x, double* fvec, const double* sys,
               const double* shared, nlb_iteration_behavior* ib, int32_t* status, cudaStream_t s, int64_t* launches) {
    using C = TlmCfg<N, S, NST>;
    KernelCfg cfg;
    if (kernel_cfg<tlm_kernel<F, N, S, NST>>(C::NT, C::BYTES, &cfg) != cudaSuccess) return NLB_ERR_CUDA;
    static const int per_sm_override = [] {
        const char* e = std::getenv("NLB_TLM_CTAS_PER_SM");      // tuning knob
        return e ? std::atoi(e) : 0;
    }();
    int per_sm = cfg.ctas_per_sm;
    if (per_sm_override > 0 && per_sm_override < per_sm) per_sm = per_sm_override;
    long long grid = ntot;
    if (grid > (long long)cfg.num_sms * per_sm) grid = (long long)cfg.num_sms * per_sm;
    const int MP = (m + S - 1) / S * S;
    const size_t per_cta = (size_t)(N + 3) * (size_t)MP;         // doubles
    double* ws = nullptr;
    if (cudaMallocAsync((void**)&ws, (per_cta * (size_t)grid + (size_t)MP + 8) * sizeof(double), s) != cudaSuccess)
        return NLB_ERR_CUDA;
    double* tpad = ws + per_cta * (size_t)grid;                  // abscissae padded to whole segments
    unsigned long long* cursor = reinterpret_cast<unsigned long long*>(tpad + MP);
    if (cudaMemsetAsync(tpad, 0, ((size_t)MP + 8) * sizeof(double), s) != cudaSuccess ||
        cudaMemcpyAsync(tpad, shared, (size_t)m * sizeof(double), cudaMemcpyDeviceToDevice, s) != cudaSuccess) {
        cudaFreeAsync(ws, s);
        return NLB_ERR_CUDA;
    }
    tlm_kernel<F, N, S, NST><<<(unsigned)grid, C::NT, C::BYTES, s>>>(p, B, ntot, m, MP, x, fvec, sys, tpad, ib, status, ws, cursor);
    ++*launches;
    if (cudaGetLastError() != cudaSuccess) { cudaFreeAsync(ws, s); return NLB_ERR_CUDA; }
    if (cudaFreeAsync(ws, s) != cudaSuccess) return NLB_ERR_CUDA;
    return NLB_OK;
}

// m >= WLM_MIN_M: the TMA-streamed kernel; NLB_TALL_LM=wlm selects the previous CTA-per-system kernel (re-measurement)
template <class F>
int launch_tall(const DevParams& p, long long ntot, long long B, int m, double* x, double* fvec, const double* sys,
                const double* shared, nlb_iteration_behavior* ib, int32_t* status, cudaStream_t s, int64_t* launches) {
    static const int use_wlm = [] {
        const char* e = std::getenv("NLB_TALL_LM");
        return (e && e[0] == 'w') ? 1 : 0;
    }();
    if (use_wlm) return launch_wlm<F, 16>(p, ntot, B, m, x, fvec, sys, shared, ib, status, s, launches);
    return launch_tlm<F, 16, NLB_TLM_SEG, NLB_TLM_STAGES>(p, ntot, B, m, x, fvec, sys, shared, ib, status, s, launches);
}

}  // namespace

int launch_coop_lm(int fcn_id, const DevParams& p, long long nsys, long long B, int m, int n, double* x, double* fvec, const double* sys,
                   const double* shared, nlb_iteration_behavior* ib, int32_t* status, cudaStream_t s, int64_t* launches) {
    if (m < n) return NLB_ERR_UNSUPPORTED;
    switch (fcn_id) {
        // tall fits: one CTA per system (a lone slow system costs ~1 ms per iteration instead of ~17 ms);
        // short ones: one lane per system (more systems in flight)
        case FCN_RATIONAL_7_8:
            if (m >= WLM_MIN_M) return launch_tall<Rational78>(p, nsys, B, m, x, fvec, sys, shared, ib, status, s, launches);
            return launch_lm<Rational78, 16>(p, nsys, B, m, x, fvec, sys, shared, ib, status, s, launches);
        case FCN_EXP_SUM_8:
            if (m >= WLM_MIN_M) return launch_tall<ExpSum8>(p, nsys, B, m, x, fvec, sys, shared, ib, status, s, launches);
            return launch_lm<ExpSum8, 16>(p, nsys, B, m, x, fvec, sys, shared, ib, status, s, launches);
        case FCN_EXP_DECAY_4: return launch_lm<ExpDecay4, 4>(p, nsys, B, m, x, fvec, sys, shared, ib, status, s, launches);
        default: return NLB_ERR_UNSUPPORTED;
    }
}

int launch_coop_solve(int solver, int fcn_id, const DevParams& p, long long nsys, long long B, int m, int n, double* x,
                      double* fvec,
                      const double* sys, const double* shared, nlb_iteration_behavior* ib, int32_t* status,
                      cudaStream_t s, int64_t* launches) {
    if (solver == SOLVER_BROYDEN && fcn_id == FCN_EXT_ROSENBROCK) {
        switch (n) {
            case 8: return launch_broyden<ExtRosenbrockCoop, 8>(p, nsys, B, x, fvec, sys, shared, ib, status, s, launches);
            case 16: return launch_broyden<ExtRosenbrockCoop, 16>(p, nsys, B, x, fvec, sys, shared, ib, status, s, launches);
            case 32: return launch_broyden<ExtRosenbrockCoop, 32>(p, nsys, B, x, fvec, sys, shared, ib, status, s, launches);
            case 64: return launch_broyden<ExtRosenbrockCoop, 64>(p, nsys, B, x, fvec, sys, shared, ib, status, s, launches);
            default: return NLB_ERR_UNSUPPORTED;
        }
    }
    if (solver == SOLVER_LM) return launch_coop_lm(fcn_id, p, nsys, B, m, n, x, fvec, sys, shared, ib, status, s, launches);
    return NLB_ERR_UNSUPPORTED;
}

namespace {
template <class F>
int launch_cls_rt(const DevParams& p, const DevCls& o, long long nsys, long long B, int m, double* x, double* fvec,
                  const double* sys, const double* shared, nlb_iteration_behavior* ib, int32_t* status, cudaStream_t s,
                  int64_t* launches) {
    KernelCfg cfg;
    if (kernel_cfg<cls_rt_kernel<F>>(128, 0, &cfg) != cudaSuccess) return NLB_ERR_CUDA;
    const long long per = cls_rt_ws_doubles<F::N>(m) * (long long)sizeof(double);
    long long T = (long long)cfg.num_sms * cfg.ctas_per_sm * 128;
    const long long need = ((nsys + 127) / 128) * 128;
    if (T > need) T = need;
    const long long budget = 4LL << 30;          // workspace cap: fewer threads (grid-stride) rather than more memory
    while (T > 128 && T * per > budget) T -= 128;
    if (T * per > budget) return NLB_ERR_UNSUPPORTED;
    double* ws = nullptr;
    if (cudaMallocAsync((void**)&ws, (size_t)(T * per), s) != cudaSuccess) return NLB_ERR_CUDA;
    cls_rt_kernel<F><<<(unsigned)(T / 128), 128, 0, s>>>(p, o, nsys, B, m, x, fvec, sys, shared, ib, status, ws);
    ++*launches;
    const cudaError_t le = cudaGetLastError();
    if (cudaFreeAsync(ws, s) != cudaSuccess || le != cudaSuccess) return NLB_ERR_CUDA;
    return NLB_OK;
}
}  // namespace

int launch_coop_cls(int fcn_id, const DevParams& p, const DevCls& o, long long nsys, long long B, int m, int n, double* x,
                    double* fvec, const double* sys, const double* shared, nlb_iteration_behavior* ib, int32_t* status,
                    cudaStream_t s, int64_t* launches) {
    (void)n;
    switch (fcn_id) {
        case FCN_RATIONAL_7_8: return launch_cls_rt<Rational78>(p, o, nsys, B, m, x, fvec, sys, shared, ib, status, s, launches);
        case FCN_EXP_SUM_8: return launch_cls_rt<ExpSum8>(p, o, nsys, B, m, x, fvec, sys, shared, ib, status, s, launches);
        case FCN_EXP_DECAY_4: return launch_cls_rt<ExpDecay4>(p, o, nsys, B, m, x, fvec, sys, shared, ib, status, s, launches);
        default: return NLB_ERR_UNSUPPORTED;
    }
}

int launch_coop_eval(int fcn_id, long long B, int m, int n, const double* x, double* fvec, const double* sys,
                     const double* shared, cudaStream_t s) {
    const unsigned gx = (unsigned)((B + 127) / 128);
    const dim3 grid(gx, (unsigned)(m < 64 ? m : 64));
    switch (fcn_id) {
        case FCN_RATIONAL_7_8: curvefit_eval_kernel<Rational78><<<grid, 128, 0, s>>>(B, m, x, fvec, sys, shared); break;
        case FCN_EXP_SUM_8: curvefit_eval_kernel<ExpSum8><<<grid, 128, 0, s>>>(B, m, x, fvec, sys, shared); break;
        case FCN_EXP_DECAY_4: curvefit_eval_kernel<ExpDecay4><<<grid, 128, 0, s>>>(B, m, x, fvec, sys, shared); break;
        case FCN_EXT_ROSENBROCK: rosenbrock_eval_kernel<<<dim3(gx, (unsigned)(n / 2 < 32 ? n / 2 : 32)), 128, 0, s>>>(B, n, x, fvec); break;
        default: return NLB_ERR_UNSUPPORTED;
    }
    return NLB_OK;
}

int launch_coop_jacobian(int fcn_id, long long B, int m, int n, const double* x, double* jac, const double* sys,
                         const double* shared, cudaStream_t s) {
    const dim3 grid((unsigned)((B + 127) / 128), (unsigned)n);
    switch (fcn_id) {
        case FCN_RATIONAL_7_8: curvefit_jacobian_kernel<Rational78><<<grid, 128, 0, s>>>(B, m, x, jac, sys, shared); break;
        case FCN_EXP_SUM_8: curvefit_jacobian_kernel<ExpSum8><<<grid, 128, 0, s>>>(B, m, x, jac, sys, shared); break;
        case FCN_EXP_DECAY_4: curvefit_jacobian_kernel<ExpDecay4><<<grid, 128, 0, s>>>(B, m, x, jac, sys, shared); break;
        case FCN_EXT_ROSENBROCK: rosenbrock_jacobian_kernel<<<grid, 128, 0, s>>>(B, n, x, jac); break;
        default: return NLB_ERR_UNSUPPORTED;
    }
    return NLB_OK;
}

}  // namespace nlb

#ifdef NLB_TLM_TRACE
// debug build only (scripts/tlm_trace.py): copy the clock stamps of tall_lm.cuh's probes back
extern "C" int nlb_debug_tlm_trace(long long* stamps, int* counts) {
    if (cudaMemcpyFromSymbol(stamps, nlb::tlm_trace_buf, sizeof(long long) * 16 * 512) != cudaSuccess) return 1;
    (void)counts;
    return 0;
}
#endif

#ifdef NLB_CB_TRACE
extern "C" int nlb_debug_cb_trace(long long* stamps) {
    return cudaMemcpyFromSymbol(stamps, nlb::cb_trace_buf, sizeof(long long) * 64 * 32) != cudaSuccess;
}
#endif
