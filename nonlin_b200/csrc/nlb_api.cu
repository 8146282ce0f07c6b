// nlb_api.cu — the C ABI of the engine (include/nonlin_batch.h): handle, residual registry,
// launchers for the batch solvers, batch statistics and the FP64 peak probe.
//
// Compiled for sm_100a only, with -fmad=false (see nlb_math.cuh).  There is no CPU path in
// this library: without a CUDA device every computing entry point fails with
// NLB_ERR_NO_DEVICE.
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>

#include <cstdlib>
#include <deque>
#include <dlfcn.h>
#include <thread>
#include <vector>

#include "coop_kernels.cuh"
#include "launch_cfg.cuh"
#include "polyfit.cuh"
#include "quad_lm.cuh"
#include "scalar_solvers.cuh"
#include "tps_cls.cuh"
#include "tps_kernels.cuh"

using namespace nlb;

// ---------------------------------------------------------------------------------------
// handle
// ---------------------------------------------------------------------------------------
struct nlb_handle {
    int device = 0;
    cudaStream_t stream = nullptr;
    std::string last_error;
    int64_t launches = 0;
    // grow-only device staging for host-resident arguments
    static constexpr int NSLOT = 8;
    void* dbuf[NSLOT] = {nullptr};
    size_t dcap[NSLOT] = {0};
    int64_t* dstats = nullptr;
    int num_sms = 0;
    void* dwork = nullptr;                         // grow-only workspace of the polynomial-fit kernel
    size_t dwork_cap = 0;
    cudaEvent_t ev_work = nullptr;                 // last kernel that used dwork (orders users on different streams)
    bool work_used = false;
    // pipeline for host-resident batches: upload stream, two kernel streams (alternating, so the last wave of one
    // chunk's grid overlaps the first of the next), download stream; one event pair per chunk
    static constexpr int NPIPE = 4, MAXCHUNK = 16;
    cudaStream_t pipe[NPIPE] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev_in = nullptr, ev_out = nullptr;
    cudaEvent_t ev_up[MAXCHUNK] = {nullptr}, ev_done[MAXCHUNK] = {nullptr};
    std::mutex mu;
};

namespace {


bool make_chunk_events(nlb_handle* h) {
    for (int q = 0; q < nlb_handle::MAXCHUNK; ++q)
        if (cudaEventCreateWithFlags(&h->ev_up[q], cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&h->ev_done[q], cudaEventDisableTiming) != cudaSuccess)
            return false;
    return true;
}

int set_err(nlb_handle* h, int code, const char* what, cudaError_t ce = cudaSuccess) {
    if (h) {
        h->last_error = what;
        if (ce != cudaSuccess) {
            h->last_error += ": ";
            h->last_error += cudaGetErrorString(ce);
        }
    }
    return code;
}

// Make the handle's device current for the rest of the scope; the caller's device is restored on return.
#define NLB_DEVICE(h)                                                                        \
    DeviceGuard nlb_device_guard((h)->device);                                               \
    if (nlb_device_guard.err != cudaSuccess)                                                 \
        return set_err((h), NLB_ERR_NO_DEVICE, "cudaSetDevice", nlb_device_guard.err)

#define NLB_CUDA(h, call)                                                   \
    do {                                                                    \
        cudaError_t _e = (call);                                            \
        if (_e != cudaSuccess) return set_err((h), NLB_ERR_CUDA, #call, _e); \
    } while (0)

bool is_device_ptr(const void* p) {
    if (!p) return true;
    cudaPointerAttributes a;
    cudaError_t e = cudaPointerGetAttributes(&a, p);
    if (e != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

bool is_device_ptr_nonnull(const void* p) { return p != nullptr && is_device_ptr(p); }

// NVTX ranges around the phases of a host-buffer solve (H2D / kernel / D2H), resolved at run time from libnvToolsExt
// so that the library has no link-time dependency on it; no-ops when no profiler library is present.
struct NvtxApi {
    int (*push)(const char*) = nullptr;
    int (*pop)() = nullptr;
    NvtxApi() {
        for (const char* name : {"libnvToolsExt.so.1", "libnvToolsExt.so"}) {
            if (void* lib = dlopen(name, RTLD_LAZY | RTLD_GLOBAL)) {
                push = reinterpret_cast<int (*)(const char*)>(dlsym(lib, "nvtxRangePushA"));
                pop = reinterpret_cast<int (*)()>(dlsym(lib, "nvtxRangePop"));
                if (push && pop) return;
            }
        }
        push = nullptr; pop = nullptr;
    }
};
const NvtxApi& nvtx_api() { static const NvtxApi api; return api; }
void nvtx_push(const char* name) { if (nvtx_api().push) nvtx_api().push(name); }
void nvtx_pop() { if (nvtx_api().pop) nvtx_api().pop(); }

// One argument of a call: the caller's pointer and the device pointer the kernel uses.
struct Staged {
    void* user = nullptr;
    void* dev = nullptr;
    size_t bytes = 0;
    bool staged = false;
};

int stage_in(nlb_handle* h, int slot, const void* user, size_t bytes, bool copy_in, cudaStream_t s, Staged* out) {
    out->user = const_cast<void*>(user);
    out->bytes = bytes;
    if (!user || bytes == 0) { out->dev = const_cast<void*>(user); return NLB_OK; }
    if (is_device_ptr(user)) { out->dev = const_cast<void*>(user); return NLB_OK; }
    if (h->dcap[slot] < bytes) {
        if (h->dbuf[slot]) NLB_CUDA(h, cudaFree(h->dbuf[slot]));
        h->dbuf[slot] = nullptr;
        h->dcap[slot] = 0;
        NLB_CUDA(h, cudaMalloc(&h->dbuf[slot], bytes));
        h->dcap[slot] = bytes;
    }
    out->dev = h->dbuf[slot];
    out->staged = true;
    if (copy_in) NLB_CUDA(h, cudaMemcpyAsync(out->dev, user, bytes, cudaMemcpyHostToDevice, s));
    return NLB_OK;
}

int stage_out(nlb_handle* h, const Staged& a, cudaStream_t s) {
    if (a.staged && a.user) NLB_CUDA(h, cudaMemcpyAsync(a.user, a.dev, a.bytes, cudaMemcpyDeviceToHost, s));
    return NLB_OK;
}

// Levenberg-Marquardt with the m x n Jacobian in shared memory ([element][thread] tile, 8 M N bytes per thread): for
// residuals whose Jacobian does not fit the register file (21 x 4: 86 KB per CTA, two CTAs per SM).  The two m-vectors
// stay in local memory, where they fit L1 next to the tile.  Experimental: slower than the local-memory kernel on
// B200 (see launch_tps_lm_smem), not the default.
template <class F>
__global__ void __launch_bounds__(TPS_BLOCK, 2)
tps_lm_smem_kernel(DevParams p, long long nsys, long long B, double* __restrict__ x, double* __restrict__ fvec,
                   const double* __restrict__ sys, const double* __restrict__ shared,
                   nlb_iteration_behavior* __restrict__ ib, int32_t* __restrict__ status) {
    constexpr int M = F::M, N = F::N;
    extern __shared__ double lm_tile[];
    const long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (b >= nsys) return;
    double xl[N], fl[M];
#pragma unroll
    for (int j = 0; j < N; ++j) xl[j] = x[j * B + b];
    SysCtx c{sys ? sys + b : nullptr, shared, B, M, N};
    SolveStats st;
    StridedMat<TPS_BLOCK> jac{lm_tile + threadIdx.x};
    tps_lm_solve<F>(p, c, xl, fl, st, jac);
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

// constrained_least_squares_solver: persistent thread-per-system kernel (tps_cls_refill); the limits and the radius
// travel by value.
template <class F>
__global__ void __launch_bounds__(TPS_BLOCK)
tps_cls_kernel(DevParams p, DevCls o, long long nsys, long long B, unsigned long long* cursor, double* __restrict__ x,
               double* __restrict__ fvec, const double* __restrict__ sys, const double* __restrict__ shared,
               nlb_iteration_behavior* __restrict__ ib, int32_t* __restrict__ status) {
    tps_cls_refill<F>(p, o, nsys, B, cursor, x, fvec, sys, shared, ib, status);
}

// Persistent variant for Newton: grid = resident CTAs only, lanes pull systems from a cursor.
template <class F>
__global__ void __launch_bounds__(TPS_BLOCK)
tps_newton_refill_kernel(DevParams p, long long nsys, long long B, unsigned long long* cursor, double* __restrict__ x,
                         double* __restrict__ fvec, const double* __restrict__ sys, const double* __restrict__ shared,
                         nlb_iteration_behavior* __restrict__ ib, int32_t* __restrict__ status) {
    tps_newton_refill<F>(p, nsys, B, cursor, x, fvec, sys, shared, ib, status);
}

// ---------------------------------------------------------------------------------------
// batch statistics
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
stats_kernel(long long B, const nlb_iteration_behavior* __restrict__ ib, const int32_t* __restrict__ status,
             unsigned long long* __restrict__ out) {
    // grid-stride accumulation in registers, warp shuffle, one shared-memory pass per block, then one atomic per
    // block and statistic (HBM-bound: 32 B per system)
    unsigned long long v[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x; b < B; b += (long long)gridDim.x * blockDim.x) {
        const int st = status ? status[b] : 0;
        v[NLB_STAT_SYSTEMS] += 1;
        v[NLB_STAT_CONVERGED] += (st == 0);
        v[NLB_STAT_FAILED] += (st != 0);
        if (ib) {
            const int* r = reinterpret_cast<const int*>(ib + b);
            const int it = r[0], nf = r[1], nj = r[2], cf = r[4], cx = r[5], cg = r[6];
            v[NLB_STAT_CONVERGED_FCN] += (cf != 0);
            v[NLB_STAT_CONVERGED_CHNG] += (cx != 0);
            v[NLB_STAT_CONVERGED_ZERO_DIFF] += (cg != 0);
            v[NLB_STAT_SUM_ITER] += (unsigned long long)it;
            v[NLB_STAT_SUM_FCN] += (unsigned long long)nf;
            v[NLB_STAT_SUM_JAC] += (unsigned long long)nj;
            v[NLB_STAT_MAX_ITER] = max(v[NLB_STAT_MAX_ITER], (unsigned long long)it);
        }
    }
    __shared__ unsigned long long part[8][10];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 10; ++k) {
        unsigned long long s = v[k];
        if (k == NLB_STAT_MAX_ITER) {
            for (int off = 16; off > 0; off >>= 1) s = max(s, __shfl_down_sync(0xffffffffu, s, off));
        } else {
            for (int off = 16; off > 0; off >>= 1) s += __shfl_down_sync(0xffffffffu, s, off);
        }
        if (lane == 0) part[warp][k] = s;
    }
    __syncthreads();
    if (threadIdx.x < 10) {
        const int k = threadIdx.x;
        unsigned long long s = part[0][k];
        for (int w = 1; w < 8; ++w) s = (k == NLB_STAT_MAX_ITER) ? max(s, part[w][k]) : s + part[w][k];
        if (k == NLB_STAT_MAX_ITER) atomicMax(&out[k], s);
        else if (s) atomicAdd(&out[k], s);
    }
}

// ---------------------------------------------------------------------------------------
// FP64 peak probe: 8 independent chains per thread, no memory traffic in the loop
// ---------------------------------------------------------------------------------------
template <bool FMA>
__global__ void fp64_peak_kernel(double* out, int iters, double seed) {
    double a[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = seed + k + threadIdx.x * 1e-9;
    const double m = 1.0000000001, c = 1e-12;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (FMA) a[k] = __fma_rn(a[k], m, c);
            else { a[k] = __dmul_rn(a[k], m); a[k] = __dadd_rn(a[k], c); }
        }
    }
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += a[k];
    if (s == 12345.6789) out[0] = s;   // never true; keeps the chains alive
}

// Dependent-chain latencies of the FP64 operations the serial parts of the solvers are made of (one thread,
// clock64 around a chain): cycles per dependent DADD, per dependent IEEE division, per dependent sqrt, and per
// dependent shared-memory load + DADD (the replay loops of the cooperative kernels).
__global__ void fp64_latency_kernel(double* out, int n, double seed) {
    __shared__ double buf[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) buf[i] = 1e-9 * (i + 1);
    __syncthreads();
    if (threadIdx.x != 0) return;
    double a = seed;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) a = __dadd_rn(a, 1e-9);
    long long t1 = clock64();
    double b = a;
    for (int i = 0; i < n; ++i) b = b / 1.0000001;
    long long t2 = clock64();
    double c = b;
    for (int i = 0; i < n; ++i) c = sqrt(c + 2.0);
    long long t3 = clock64();
    double d = c;
    for (int i = 0; i < n; ++i) d = __dadd_rn(d, buf[i & 255]);
    long long t4 = clock64();
    out[0] = (double)(t1 - t0) / n;
    out[1] = (double)(t2 - t1) / n;
    out[2] = (double)(t3 - t2) / n;
    out[3] = (double)(t4 - t3) / n;
    out[4] = d;
}

// ---------------------------------------------------------------------------------------
// dispatch
// ---------------------------------------------------------------------------------------
template <class F, int SOLVER>
int launch_tps_solve(nlb_handle* h, const DevParams& p, long long nsys, long long B, double* x, double* fvec,
                     const double* sys, const double* shared, nlb_iteration_behavior* ib, int32_t* status,
                     cudaStream_t s) {
    if (nsys == 0) return NLB_OK;
    const unsigned grid = (unsigned)((nsys + TPS_BLOCK - 1) / TPS_BLOCK);
    tps_solve_kernel<F, SOLVER><<<grid, TPS_BLOCK, 0, s>>>(p, nsys, B, x, fvec, sys, shared, ib, status);
    ++h->launches;
    NLB_CUDA(h, cudaGetLastError());
    return NLB_OK;
}

// Work-queue cursor of a persistent kernel: allocated stream-ordered with the launch, zeroed, and released after the
// kernel on the same stream, so launches in flight on different streams never share one.
int cursor_acquire(nlb_handle* h, cudaStream_t s, unsigned long long** cursor) {
    NLB_CUDA(h, cudaMallocAsync((void**)cursor, sizeof(unsigned long long), s));
    NLB_CUDA(h, cudaMemsetAsync(*cursor, 0, sizeof(unsigned long long), s));
    return NLB_OK;
}

template <class F>
int launch_tps_newton_refill(nlb_handle* h, const DevParams& p, long long nsys, long long B, double* x, double* fvec,
                             const double* sys, const double* shared, nlb_iteration_behavior* ib, int32_t* status,
                             cudaStream_t s) {
    if (nsys == 0) return NLB_OK;
    KernelCfg cfg;
    NLB_CUDA(h, kernel_cfg<tps_newton_refill_kernel<F>>(TPS_BLOCK, 0, &cfg));
    long long grid = (nsys + TPS_BLOCK - 1) / TPS_BLOCK;
    const long long resident = (long long)cfg.num_sms * cfg.ctas_per_sm;
    if (grid > resident) grid = resident;
    unsigned long long* cursor = nullptr;
    int rc = cursor_acquire(h, s, &cursor);
    if (rc) return rc;
    tps_newton_refill_kernel<F><<<(unsigned)grid, TPS_BLOCK, 0, s>>>(p, nsys, B, cursor, x, fvec, sys, shared, ib, status);
    ++h->launches;
    const cudaError_t le = cudaGetLastError();
    NLB_CUDA(h, cudaFreeAsync(cursor, s));
    NLB_CUDA(h, le);
    return NLB_OK;
}

template <class F>
int launch_tps_lm_smem(nlb_handle* h, const DevParams& p, long long nsys, long long B, double* x, double* fvec,
                       const double* sys, const double* shared, nlb_iteration_behavior* ib, int32_t* status,
                       cudaStream_t s) {
    // Measured on B200 (C1, 2^20 fits): 8.85 ms with the Jacobian in shared memory (256 threads per SM) against
    // 6.98 ms with it in local memory (384 threads per SM) - the kernel is latency-bound and occupancy wins.  The
    // shared-memory variant stays selectable for re-measurement (NLB_LM_SMEM=1); it is bit-identical.
    static const int use_smem = std::getenv("NLB_LM_SMEM") ? 1 : 0;      // magic static: thread-safe
    if (!use_smem) return launch_tps_solve<F, SOLVER_LM>(h, p, nsys, B, x, fvec, sys, shared, ib, status, s);
    if (nsys == 0) return NLB_OK;
    constexpr size_t smem = sizeof(double) * F::M * F::N * TPS_BLOCK;
    KernelCfg cfg;
    NLB_CUDA(h, kernel_cfg<tps_lm_smem_kernel<F>>(TPS_BLOCK, smem, &cfg));
    const unsigned grid = (unsigned)((nsys + TPS_BLOCK - 1) / TPS_BLOCK);
    tps_lm_smem_kernel<F><<<grid, TPS_BLOCK, smem, s>>>(p, nsys, B, x, fvec, sys, shared, ib, status);
    ++h->launches;
    NLB_CUDA(h, cudaGetLastError());
    return NLB_OK;
}

// Levenberg-Marquardt for the small curve fits: four lanes per system, Jacobian in registers (quad_lm.cuh).
// Bit-identical to the thread-per-system kernel; NLB_LM_QUAD=0 / 1 forces one or the other (re-measurement).
#ifndef NLB_LM_QUAD_DEFAULT
#define NLB_LM_QUAD_DEFAULT 0      // measured on B200 (2^20 C1 fits): quad kernel v1 14.1 ms, thread-per-system 6.9 ms
#endif
template <class F>
int launch_qlm(nlb_handle* h, const DevParams& p, long long nsys, long long B, double* x, double* fvec,
               const double* sys, const double* shared, nlb_iteration_behavior* ib, int32_t* status, cudaStream_t s) {
    static const int use_quad = [] {
        const char* e = std::getenv("NLB_LM_QUAD");
        return e ? std::atoi(e) : NLB_LM_QUAD_DEFAULT;
    }();
    if (!use_quad) return launch_tps_lm_smem<F>(h, p, nsys, B, x, fvec, sys, shared, ib, status, s);
    if (nsys == 0) return NLB_OK;
    constexpr size_t smem = QlmSmem<F>::BYTES;
    KernelCfg cfg;
    NLB_CUDA(h, kernel_cfg<qlm_kernel<F>>(QLM_BLOCK, smem, &cfg));
    const unsigned grid = (unsigned)((nsys + QLM_QUADS - 1) / QLM_QUADS);
    qlm_kernel<F><<<grid, QLM_BLOCK, smem, s>>>(p, nsys, B, x, fvec, sys, shared, ib, status);
    ++h->launches;
    NLB_CUDA(h, cudaGetLastError());
    return NLB_OK;
}

#define NLB_SQUARE_FCNS(X) \
    X(Misc2Fcn) X(Misc2FcnA) X(PoorlyScaled2Fcn) X(PowellBadlyScaled) X(Misc2Fcn01) X(Polar) X(PolarScaled)
#define NLB_FIXED_FCNS(X) NLB_SQUARE_FCNS(X) X(LsqPolyFit)

template <int SOLVER>
int dispatch_tps(nlb_handle* h, int fcn_id, const DevParams& p, long long nsys, long long B, double* x, double* fvec,
                 const double* sys, const double* shared, nlb_iteration_behavior* ib, int32_t* status,
                 cudaStream_t s) {
    switch (fcn_id) {
#define X(F)                                                                                              \
    case F::ID:                                                                                           \
        if constexpr (SOLVER == SOLVER_NEWTON)                                                            \
            return launch_tps_newton_refill<F>(h, p, nsys, B, x, fvec, sys, shared, ib, status, s);       \
        else                                                                                              \
            return launch_tps_solve<F, SOLVER>(h, p, nsys, B, x, fvec, sys, shared, ib, status, s);
        NLB_SQUARE_FCNS(X)
#undef X
        case LsqPolyFit::ID:
            if constexpr (SOLVER == SOLVER_LM)
                return launch_qlm<LsqPolyFit>(h, p, nsys, B, x, fvec, sys, shared, ib, status, s);
            else
                return set_err(h, NLB_ERR_SIZE, "Newton / quasi-Newton need m == n");
        default: return set_err(h, NLB_ERR_UNSUPPORTED, "no thread-per-system kernel for this residual");
    }
}

template <class F>
int launch_cls(nlb_handle* h, const DevParams& p, const DevCls& o, long long nsys, long long B, double* x, double* fvec,
               const double* sys, const double* shared, nlb_iteration_behavior* ib, int32_t* status, cudaStream_t s) {
    KernelCfg cfg;
    NLB_CUDA(h, kernel_cfg<tps_cls_kernel<F>>(TPS_BLOCK, 0, &cfg));
    long long grid = (nsys + TPS_BLOCK - 1) / TPS_BLOCK;
    const long long resident = (long long)cfg.num_sms * cfg.ctas_per_sm;
    if (grid > resident) grid = resident;
    unsigned long long* cursor = nullptr;
    int rc = cursor_acquire(h, s, &cursor);
    if (rc) return rc;
    tps_cls_kernel<F><<<(unsigned)grid, TPS_BLOCK, 0, s>>>(p, o, nsys, B, cursor, x, fvec, sys, shared, ib, status);
    ++h->launches;
    const cudaError_t le = cudaGetLastError();
    NLB_CUDA(h, cudaFreeAsync(cursor, s));
    NLB_CUDA(h, le);
    return NLB_OK;
}

int dispatch_cls(nlb_handle* h, int fcn_id, const DevParams& p, const DevCls& o, long long nsys, long long B, double* x,
                 double* fvec, const double* sys, const double* shared, nlb_iteration_behavior* ib, int32_t* status,
                 cudaStream_t s) {
    if (nsys == 0) return NLB_OK;
    switch (fcn_id) {
#define X(F) \
    case F::ID: return launch_cls<F>(h, p, o, nsys, B, x, fvec, sys, shared, ib, status, s);
        NLB_FIXED_FCNS(X)
#undef X
        default: return set_err(h, NLB_ERR_UNSUPPORTED, "no constrained least-squares kernel for this residual");
    }
}

#define NLB_FCN1_LIST(X) X(SinxDivX) X(SinxDivXA) X(CubicWallis) X(ExpMinusX) X(CubicArgs)

int solve_1var_batch(nlb_handle* h, int solver, const nlb_params_1var* params, int fcn_id, int64_t B,
                     const double* lim1, const double* lim2, double* x, double* f, const double* args,
                     nlb_iteration_behavior* ib, int32_t* status, void* stream) {
    if (!h) return NLB_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::mutex> lock(h->mu);
    if (!params || B < 0 || (B > 0 && (!x || !lim1 || !lim2)))
        return set_err(h, NLB_ERR_INVALID_ARGUMENT, "null argument or B < 0");
    if (fcn_id < 0 || fcn_id >= FCN1_COUNT) return set_err(h, NLB_ERR_UNKNOWN_FCN, "unknown one-variable function id");
    const Fcn1Info& fi = fcn1_table()[fcn_id];
    if (fi.args_len > 0 && B > 0 && !args) return set_err(h, NLB_ERR_INVALID_ARGUMENT, "this function needs per-equation args");
    NLB_DEVICE(h);
    int rc = NLB_OK;
    if (B == 0) return NLB_OK;
    cudaStream_t s = stream ? (cudaStream_t)stream : h->stream;
    Staged a1, a2, ax, af, aa, aib, ast;
    if ((rc = stage_in(h, 0, x, sizeof(double) * (size_t)B, true, s, &ax))) return rc;
    if ((rc = stage_in(h, 1, f, sizeof(double) * (size_t)B, false, s, &af))) return rc;
    if ((rc = stage_in(h, 2, args, sizeof(double) * (size_t)fi.args_len * B, true, s, &aa))) return rc;
    if ((rc = stage_in(h, 3, lim1, sizeof(double) * (size_t)B, true, s, &a1))) return rc;
    if ((rc = stage_in(h, 6, lim2, sizeof(double) * (size_t)B, true, s, &a2))) return rc;
    if ((rc = stage_in(h, 4, ib, sizeof(nlb_iteration_behavior) * (size_t)B, false, s, &aib))) return rc;
    if ((rc = stage_in(h, 5, status, sizeof(int32_t) * (size_t)B, false, s, &ast))) return rc;
    DevParams1 p;
    p.max_fcn_evals = params->max_fcn_evals;
    p.fcn_tol = params->fcn_tol;
    p.var_tol = params->var_tol;
    p.diff_tol = params->diff_tol;
    p.use_analytic_diff = params->use_analytic_diff;
    const unsigned grid = (unsigned)((B + 127) / 128);
    switch (fcn_id) {
#define X(F)                                                                                                      \
    case F::ID:                                                                                                   \
        if (solver == 0)                                                                                          \
            solve_1var_kernel<F, 0><<<grid, 128, 0, s>>>(p, B, (const double*)a1.dev, (const double*)a2.dev,      \
                                                         (double*)ax.dev, (double*)af.dev, (const double*)aa.dev, \
                                                         (nlb_iteration_behavior*)aib.dev, (int32_t*)ast.dev);    \
        else                                                                                                      \
            solve_1var_kernel<F, 1><<<grid, 128, 0, s>>>(p, B, (const double*)a1.dev, (const double*)a2.dev,      \
                                                         (double*)ax.dev, (double*)af.dev, (const double*)aa.dev, \
                                                         (nlb_iteration_behavior*)aib.dev, (int32_t*)ast.dev);    \
        break;
        NLB_FCN1_LIST(X)
#undef X
    }
    ++h->launches;
    NLB_CUDA(h, cudaGetLastError());
    if ((rc = stage_out(h, ax, s))) return rc;
    if ((rc = stage_out(h, af, s))) return rc;
    if ((rc = stage_out(h, aib, s))) return rc;
    if ((rc = stage_out(h, ast, s))) return rc;
    if (a1.staged || a2.staged || ax.staged || af.staged || aa.staged || aib.staged || ast.staged)
        NLB_CUDA(h, cudaStreamSynchronize(s));
    return NLB_OK;
}

// residuals registered at run time by plug-ins (nlb_plugin.cuh): ids FCN_COUNT, FCN_COUNT + 1, ...
struct UserFcn {
    std::string name;
    int m, n, sys_len, shared_len, has_jac;
    nlb_user_solve_fn solve;
    nlb_user_eval_fn eval;
};
std::mutex g_user_mu;
std::deque<UserFcn> g_user;                      // deque: entries never move once registered

int fcn_total() {
    std::lock_guard<std::mutex> lock(g_user_mu);
    return FCN_COUNT + (int)g_user.size();
}
const UserFcn* user_fcn(int fcn_id) {
    std::lock_guard<std::mutex> lock(g_user_mu);
    const int k = fcn_id - FCN_COUNT;
    return (k >= 0 && k < (int)g_user.size()) ? &g_user[k] : nullptr;
}
bool fcn_info(int fcn_id, FcnInfo* fi) {
    if (fcn_id >= 0 && fcn_id < FCN_COUNT) { *fi = fcn_table()[fcn_id]; return true; }
    if (const UserFcn* u = user_fcn(fcn_id)) {
        *fi = FcnInfo{u->name.c_str(), u->m, u->n, u->sys_len, u->shared_len, u->has_jac};
        return true;
    }
    return false;
}

int check_sizes(nlb_handle* h, int fcn_id, int* m, int* n, int* sys_len, int* shared_len) {
    FcnInfo fi;
    if (!fcn_info(fcn_id, &fi)) return set_err(h, NLB_ERR_UNKNOWN_FCN, "unknown residual id");
    if (fi.m != 0) { if (*m != 0 && *m != fi.m) return set_err(h, NLB_ERR_SIZE, "m does not match the residual"); *m = fi.m; }
    if (fi.n != 0) { if (*n != 0 && *n != fi.n) return set_err(h, NLB_ERR_SIZE, "n does not match the residual"); *n = fi.n; }
    if (fcn_id == FCN_EXT_ROSENBROCK) {
        if (*m == 0) *m = *n;
        if (*m != *n || (*n & 1)) return set_err(h, NLB_ERR_SIZE, "ext_rosenbrock needs m == n, n even");
    }
    if (*m <= 0 || *n <= 0) return set_err(h, NLB_ERR_SIZE, "m and n must be positive");
    *sys_len = fi.sys_len < 0 ? *m : fi.sys_len;
    *shared_len = fi.shared_len < 0 ? *m : fi.shared_len;
    return NLB_OK;
}

// r0 / rcnt: solve only systems [r0, r0 + rcnt) of the batch (rcnt < 0: all of it).  A range call takes HOST buffers
// only and keeps a compact copy of its shard on the device (this is how nlb_solve_sharded spreads one batch over
// several GPUs).  want_stats: leave the shard's convergence statistics in h->dstats (device) before returning.
int solve_batch(nlb_handle* h, int solver, const nlb_params* params, int fcn_id, int64_t B, int m, int n, double* x,
                double* fvec, const double* sys, const double* shared, nlb_iteration_behavior* ib, int32_t* status,
                void* stream, const DevCls* cls = nullptr, int64_t r0 = 0, int64_t rcnt = -1, bool want_stats = false) {
    if (!h) return NLB_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::mutex> lock(h->mu);
    if (!params || B < 0 || (B > 0 && !x)) return set_err(h, NLB_ERR_INVALID_ARGUMENT, "null argument or B < 0");
    const bool ranged = rcnt >= 0;
    if (ranged && (r0 < 0 || r0 + rcnt > B)) return set_err(h, NLB_ERR_INVALID_ARGUMENT, "range outside the batch");
    int sys_len, shared_len;
    int rc = check_sizes(h, fcn_id, &m, &n, &sys_len, &shared_len);
    if (rc) return rc;
    const bool lsq = solver == SOLVER_LM || solver == SOLVER_CLS;
    if (lsq && n > m) return set_err(h, NLB_ERR_SIZE, "least squares needs m >= n");
    if (!lsq && n != m) return set_err(h, NLB_ERR_SIZE, "Newton / quasi-Newton need m == n");
    if (sys_len > 0 && B > 0 && !sys) return set_err(h, NLB_ERR_INVALID_ARGUMENT, "this residual needs per-system data");
    if (shared_len > 0 && B > 0 && !shared) return set_err(h, NLB_ERR_INVALID_ARGUMENT, "this residual needs shared data");
    NLB_DEVICE(h);
    const int64_t nb = ranged ? rcnt : B;          // systems solved by this call
    if (nb == 0) {
        if (want_stats) NLB_CUDA(h, cudaMemsetAsync(h->dstats, 0, sizeof(int64_t) * NLB_STAT_COUNT, h->stream));
        return NLB_OK;
    }
    cudaStream_t s = stream ? (cudaStream_t)stream : h->stream;
    if (ranged) {
        if (is_device_ptr_nonnull(x) || is_device_ptr_nonnull(fvec) || is_device_ptr_nonnull(sys) ||
            is_device_ptr_nonnull(ib) || is_device_ptr_nonnull(status))
            return set_err(h, NLB_ERR_INVALID_ARGUMENT, "a sharded solve takes host buffers");
        x += r0;
        if (fvec) fvec += r0;
        if (sys) sys += r0;
        if (ib) ib += r0;
        if (status) status += r0;
    }
    const int64_t Bh = B;                          // SoA stride of the caller's arrays
    const int64_t Bd = ranged ? nb : B;            // SoA stride of the device arrays
    // fvec is optional (NULL: the residuals stay in a device scratch buffer and are not copied back)
    const bool fvec_scratch = fvec == nullptr;
    if (fvec_scratch) {
        const size_t need = sizeof(double) * (size_t)m * (size_t)Bd;
        if (h->dcap[7] < need) {
            if (h->dbuf[7]) NLB_CUDA(h, cudaFree(h->dbuf[7]));
            h->dbuf[7] = nullptr; h->dcap[7] = 0;
            NLB_CUDA(h, cudaMalloc(&h->dbuf[7], need));
            h->dcap[7] = need;
        }
        fvec = (double*)h->dbuf[7];
    }

    // Arguments are SoA arrays of `rows` x B elements; host-resident ones get a device twin.
    struct Arg { Staged st; size_t rows, elem; bool in, out; };
    Arg a[5] = {
        {{}, (size_t)n, sizeof(double), true, true},                       // x
        {{}, (size_t)m, sizeof(double), false, !fvec_scratch},             // fvec
        {{}, (size_t)sys_len, sizeof(double), true, false},                // per-system data
        {{}, 1, sizeof(nlb_iteration_behavior), false, true},              // ib
        {{}, 1, sizeof(int32_t), false, true},                             // status
    };
    const void* user[5] = {x, fvec, sys, ib, status};
    const int slot[5] = {0, 1, 2, 4, 5};
    bool any_staged = false;
    for (int i = 0; i < 5; ++i) {
        if ((rc = stage_in(h, slot[i], user[i], a[i].rows * a[i].elem * (size_t)Bd, false, s, &a[i].st))) return rc;
        any_staged = any_staged || a[i].st.staged;
    }
    Staged ash;
    if ((rc = stage_in(h, 3, shared, sizeof(double) * (size_t)shared_len, true, s, &ash))) return rc;
    any_staged = any_staged || ash.staged;

    const DevParams p = to_dev(params);
    FcnInfo fi;
    fcn_info(fcn_id, &fi);
    const UserFcn* ufcn = user_fcn(fcn_id);
    auto launch = [&](long long b0, long long cnt, cudaStream_t st) -> int {
        double* dx = (double*)a[0].st.dev + b0;
        double* df = (double*)a[1].st.dev + b0;
        const double* ds = a[2].st.dev ? (const double*)a[2].st.dev + b0 : nullptr;
        nlb_iteration_behavior* dib = a[3].st.dev ? (nlb_iteration_behavior*)a[3].st.dev + b0 : nullptr;
        int32_t* dst = a[4].st.dev ? (int32_t*)a[4].st.dev + b0 : nullptr;
        const double* dsh = (const double*)ash.dev;
        int r;
        if (ufcn) {
            if (solver == SOLVER_CLS) return set_err(h, NLB_ERR_UNSUPPORTED, "constrained least squares: built-in residuals only");
            r = ufcn->solve(solver, params, cnt, Bd, dx, df, ds, dsh, dib, dst, (void*)st);
            ++h->launches;
            if (r == NLB_ERR_CUDA) return set_err(h, r, "plug-in kernel launch", cudaGetLastError());
            if (r != NLB_OK) return set_err(h, r, "the plug-in residual does not serve this solver");
        } else if (solver == SOLVER_CLS) {
            if (fi.m != 0 && fi.n != 0) {
                r = dispatch_cls(h, fcn_id, p, *cls, cnt, Bd, dx, df, ds, dsh, dib, dst, st);
            } else {
                r = launch_coop_cls(fcn_id, p, *cls, cnt, Bd, m, n, dx, df, ds, dsh, dib, dst, st, &h->launches);
                if (r == NLB_ERR_UNSUPPORTED)
                    return set_err(h, r, "constrained least squares: no kernel for this residual / size (workspace cap 4 GB)");
                if (r == NLB_ERR_CUDA) return set_err(h, r, "constrained least squares kernel launch", cudaGetLastError());
            }
        } else if (fi.m != 0 && fi.n != 0) {
            switch (solver) {
                case SOLVER_LM: r = dispatch_tps<SOLVER_LM>(h, fcn_id, p, cnt, Bd, dx, df, ds, dsh, dib, dst, st); break;
                case SOLVER_NEWTON: r = dispatch_tps<SOLVER_NEWTON>(h, fcn_id, p, cnt, Bd, dx, df, ds, dsh, dib, dst, st); break;
                default: r = dispatch_tps<SOLVER_BROYDEN>(h, fcn_id, p, cnt, Bd, dx, df, ds, dsh, dib, dst, st);
            }
        } else {
            r = launch_coop_solve(solver, fcn_id, p, cnt, Bd, m, n, dx, df, ds, dsh, dib, dst, st, &h->launches);
            if (r == NLB_ERR_UNSUPPORTED) return set_err(h, r, "no cooperative kernel for this (solver, residual, size)");
            if (r == NLB_ERR_CUDA) return set_err(h, r, "cooperative kernel launch", cudaGetLastError());
        }
        return r;
    };

    auto shard_stats = [&](cudaStream_t st) -> int {
        NLB_CUDA(h, cudaMemsetAsync(h->dstats, 0, sizeof(int64_t) * NLB_STAT_COUNT, st));
        unsigned grid = (unsigned)((nb + 255) / 256);
        const unsigned cap = (unsigned)(h->num_sms > 0 ? h->num_sms : 148) * 4u;
        if (grid > cap) grid = cap;
        stats_kernel<<<grid, 256, 0, st>>>(nb, (const nlb_iteration_behavior*)a[3].st.dev, (const int32_t*)a[4].st.dev,
                                           (unsigned long long*)h->dstats);
        ++h->launches;
        NLB_CUDA(h, cudaGetLastError());
        return NLB_OK;
    };

    if (!any_staged) {                                // all-device call: one asynchronous launch on the caller's stream
        if ((rc = launch(0, nb, s))) return rc;
        return want_stats ? shard_stats(s) : NLB_OK;
    }

    // Host-resident batch: split it into chunks and run them through an upload stream, two alternating kernel streams
    // and a download stream, so that both copy engines (PCIe is full duplex) and the SMs stay busy: chunk c's kernel
    // waits for its upload only, its download for its kernel only, the upload of chunk c+1 never queues behind a
    // download, and the last wave of one chunk's grid overlaps the first wave of the next.
    // Measured on B200, 2^20 systems, pinned buffers (scripts/time_e2e.py), chunks = 2 / 4 / 8 / 16:
    //   C1 (LM 21x4, 453 MB moved, kernel 6.9 ms)   11.2 / 8.9 / 7.8 / 7.8 ms   (two-stream pipeline before: 10.8 ms)
    //   C2 (Broyden 2x2, 101 MB, kernel 0.39 ms)    1.67 / 1.64 / 1.60 / - ms
    //   C3 (Newton refill kernel, 101 MB, 3.0 ms)   3.18 / 3.26 / 3.95 / - ms   (persistent grid: every launch has a tail)
    const bool refill = solver == SOLVER_NEWTON || solver == SOLVER_CLS;
    int nchunk = nb >= (1 << 18) ? (refill ? 2 : 8) : (nb >= (1 << 15) ? 2 : 1);
    static const int chunk_override = [] {
        const char* e = std::getenv("NLB_HOST_CHUNKS");      // tuning knob
        return e ? std::atoi(e) : 0;
    }();
    if (chunk_override > 0) nchunk = chunk_override;
    if (nchunk > nlb_handle::MAXCHUNK) nchunk = nlb_handle::MAXCHUNK;
    const long long chunk = (nb + nchunk - 1) / nchunk;
    cudaStream_t s_up = h->pipe[0], s_down = h->pipe[2];
    NLB_CUDA(h, cudaEventRecord(h->ev_in, s));
    for (int q = 0; q < nlb_handle::NPIPE; ++q) NLB_CUDA(h, cudaStreamWaitEvent(h->pipe[q], h->ev_in, 0));
    int c = 0;
    for (long long b0 = 0; b0 < nb; b0 += chunk, ++c) {
        const long long cnt = (nb - b0 < chunk) ? (nb - b0) : chunk;
        nvtx_push("nlb H2D");
        for (int i = 0; i < 5; ++i) {
            if (a[i].st.staged && a[i].in)
                NLB_CUDA(h, cudaMemcpy2DAsync((char*)a[i].st.dev + b0 * a[i].elem, (size_t)Bd * a[i].elem,
                                              (const char*)a[i].st.user + b0 * a[i].elem, (size_t)Bh * a[i].elem,
                                              (size_t)cnt * a[i].elem, a[i].rows, cudaMemcpyHostToDevice, s_up));
        }
        NLB_CUDA(h, cudaEventRecord(h->ev_up[c], s_up));
        nvtx_pop();
        nvtx_push("nlb solve kernel");
        cudaStream_t s_run = h->pipe[(c & 1) ? 3 : 1];
        NLB_CUDA(h, cudaStreamWaitEvent(s_run, h->ev_up[c], 0));
        rc = launch(b0, cnt, s_run);
        nvtx_pop();
        if (rc) return rc;
        NLB_CUDA(h, cudaEventRecord(h->ev_done[c], s_run));
        nvtx_push("nlb D2H");
        NLB_CUDA(h, cudaStreamWaitEvent(s_down, h->ev_done[c], 0));
        for (int i = 0; i < 5; ++i) {
            if (a[i].st.staged && a[i].out)
                NLB_CUDA(h, cudaMemcpy2DAsync((char*)a[i].st.user + b0 * a[i].elem, (size_t)Bh * a[i].elem,
                                              (const char*)a[i].st.dev + b0 * a[i].elem, (size_t)Bd * a[i].elem,
                                              (size_t)cnt * a[i].elem, a[i].rows, cudaMemcpyDeviceToHost, s_down));
        }
        nvtx_pop();
    }
    // the download stream finishes last (it waited for every kernel); the statistics only need the kernels
    NLB_CUDA(h, cudaEventRecord(h->ev_out, s_down));
    NLB_CUDA(h, cudaStreamWaitEvent(s, h->ev_done[c - 1], 0));
    if (c > 1) NLB_CUDA(h, cudaStreamWaitEvent(s, h->ev_done[c - 2], 0));
    if (want_stats && (rc = shard_stats(s))) return rc;
    NLB_CUDA(h, cudaStreamWaitEvent(s, h->ev_out, 0));
    NLB_CUDA(h, cudaStreamSynchronize(s));
    return NLB_OK;
}

}  // namespace

// ---------------------------------------------------------------------------------------
// extern "C"
// ---------------------------------------------------------------------------------------
extern "C" {

int nlb_create(nlb_handle** handle, int device) {
    if (!handle) return NLB_ERR_INVALID_ARGUMENT;
    *handle = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0 || device < 0 || device >= count) {
        cudaGetLastError();
        return NLB_ERR_NO_DEVICE;
    }
    DeviceGuard guard(device);
    if (guard.err != cudaSuccess) return NLB_ERR_NO_DEVICE;
    nlb_handle* h = new nlb_handle();
    h->device = device;
    if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&h->pipe[0], cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&h->pipe[1], cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&h->pipe[2], cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&h->pipe[3], cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_in, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_out, cudaEventDisableTiming) != cudaSuccess ||
        !make_chunk_events(h) ||
        cudaMalloc(&h->dstats, sizeof(int64_t) * NLB_STAT_COUNT) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_work, cudaEventDisableTiming) != cudaSuccess ||
        cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) {
        delete h;
        return NLB_ERR_CUDA;
    }
    // Work queues and workspaces are allocated stream-ordered per launch (cudaMallocAsync): keep freed blocks in the
    // device's pool instead of returning them to the driver at every synchronisation (the default threshold of 0 makes
    // each launch after a sync pay a fresh driver allocation: measured 10x on the host-buffer path of the Newton kernel).
    {
        cudaMemPool_t pool = nullptr;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess && pool) {
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        cudaGetLastError();
    }
    *handle = h;
    return NLB_OK;
}

int nlb_destroy(nlb_handle* h) {
    if (!h) return NLB_ERR_INVALID_ARGUMENT;
    DeviceGuard guard(h->device);
    for (int i = 0; i < nlb_handle::NSLOT; ++i)
        if (h->dbuf[i]) cudaFree(h->dbuf[i]);
    if (h->dwork) cudaFree(h->dwork);
    if (h->dstats) cudaFree(h->dstats);
    if (h->stream) cudaStreamDestroy(h->stream);
    for (int q = 0; q < nlb_handle::NPIPE; ++q)
        if (h->pipe[q]) cudaStreamDestroy(h->pipe[q]);
    for (int q = 0; q < nlb_handle::MAXCHUNK; ++q) {
        if (h->ev_up[q]) cudaEventDestroy(h->ev_up[q]);
        if (h->ev_done[q]) cudaEventDestroy(h->ev_done[q]);
    }
    if (h->ev_in) cudaEventDestroy(h->ev_in);
    if (h->ev_out) cudaEventDestroy(h->ev_out);
    if (h->ev_work) cudaEventDestroy(h->ev_work);
    delete h;
    return NLB_OK;
}

const char* nlb_last_error(const nlb_handle* h) { return h ? h->last_error.c_str() : "null handle"; }
int64_t nlb_kernel_launch_count(const nlb_handle* h) { return h ? h->launches : 0; }

void nlb_params_default(nlb_params* p) {
    if (!p) return;
    p->max_fcn_evals = 100;
    p->fcn_tol = 1.0e-8;
    p->var_tol = 1.0e-12;
    p->grad_tol = 1.0e-12;
    p->lm_factor = 100.0;
    p->jacobian_interval = 5;
    p->use_line_search = 1;
    p->ls_max_fcn_evals = 100;
    p->ls_alpha = 1.0e-4;
    p->ls_factor = 0.1;
    p->use_analytic_jacobian = 0;
    p->max_iter_guard = 100000;
}

int nlb_vecfcn_count(void) { return fcn_total(); }

int nlb_vecfcn_lookup(const char* name) {
    if (!name) return -1;
    for (int i = 0; i < FCN_COUNT; ++i)
        if (std::strcmp(fcn_table()[i].name, name) == 0) return i;
    std::lock_guard<std::mutex> lock(g_user_mu);
    for (size_t k = 0; k < g_user.size(); ++k)
        if (g_user[k].name == name) return FCN_COUNT + (int)k;
    return -1;
}

const char* nlb_vecfcn_name(int fcn_id) {
    if (fcn_id >= 0 && fcn_id < FCN_COUNT) return fcn_table()[fcn_id].name;
    const UserFcn* u = user_fcn(fcn_id);
    return u ? u->name.c_str() : nullptr;
}

int nlb_register_vecfcn(const char* name, int m, int n, int sys_len, int shared_len, int has_jacobian,
                        nlb_user_solve_fn solve, nlb_user_eval_fn eval) {
    if (!name || !*name || m <= 0 || n <= 0 || sys_len < 0 || shared_len < 0 || !solve) return -1;
    if (nlb_vecfcn_lookup(name) >= 0) return -1;                 // names are unique
    std::lock_guard<std::mutex> lock(g_user_mu);
    g_user.push_back(UserFcn{name, m, n, sys_len, shared_len, has_jacobian ? 1 : 0, solve, eval});
    return FCN_COUNT + (int)g_user.size() - 1;
}

int nlb_load_plugin(const char* path) {
    if (!path) return -1;
    void* lib = dlopen(path, RTLD_NOW | RTLD_LOCAL);
    if (!lib) return -1;
    typedef int (*entry_t)(nlb_register_vecfcn_fn);
    entry_t entry = reinterpret_cast<entry_t>(dlsym(lib, "nlb_plugin_register"));
    if (!entry) { dlclose(lib); return -1; }
    return entry(&nlb_register_vecfcn);                           // the library stays loaded: its kernels are in use
}

int nlb_vecfcn_info(int fcn_id, int* m, int* n, int* sys_len, int* shared_len, int* has_jacobian) {
    FcnInfo fi;
    if (!fcn_info(fcn_id, &fi)) return NLB_ERR_UNKNOWN_FCN;
    if (m) *m = fi.m;
    if (n) *n = fi.n;
    if (sys_len) *sys_len = fi.sys_len;
    if (shared_len) *shared_len = fi.shared_len;
    if (has_jacobian) *has_jacobian = fi.has_jac;
    return NLB_OK;
}

int nlb_least_squares_solve_batch(nlb_handle* h, const nlb_params* params, int fcn_id, int64_t B, int m, int n,
                                  double* x, double* fvec, const double* sys, const double* shared,
                                  nlb_iteration_behavior* ib, int32_t* status, void* stream) {
    return solve_batch(h, SOLVER_LM, params, fcn_id, B, m, n, x, fvec, sys, shared, ib, status, stream);
}

int nlb_newton_solve_batch(nlb_handle* h, const nlb_params* params, int fcn_id, int64_t B, int m, int n, double* x,
                           double* fvec, const double* sys, const double* shared, nlb_iteration_behavior* ib,
                           int32_t* status, void* stream) {
    return solve_batch(h, SOLVER_NEWTON, params, fcn_id, B, m, n, x, fvec, sys, shared, ib, status, stream);
}

int nlb_quasi_newton_solve_batch(nlb_handle* h, const nlb_params* params, int fcn_id, int64_t B, int m, int n,
                                 double* x, double* fvec, const double* sys, const double* shared,
                                 nlb_iteration_behavior* ib, int32_t* status, void* stream) {
    return solve_batch(h, SOLVER_BROYDEN, params, fcn_id, B, m, n, x, fvec, sys, shared, ib, status, stream);
}

void nlb_constrained_options_default(nlb_constrained_options* o) {
    if (!o) return;
    o->trust_region_radius = 1.0;
    o->step_scaling_factor = 1.0;
    o->lower = nullptr;
    o->upper = nullptr;
}

int nlb_constrained_least_squares_solve_batch(nlb_handle* h, const nlb_params* params,
                                              const nlb_constrained_options* options, int fcn_id, int64_t B, int m,
                                              int n, double* x, double* fvec, const double* sys, const double* shared,
                                              nlb_iteration_behavior* ib, int32_t* status, void* stream) {
    if (!h) return NLB_ERR_INVALID_ARGUMENT;
    if (!options) return set_err(h, NLB_ERR_INVALID_ARGUMENT, "null options");
    FcnInfo fi;
    if (!fcn_info(fcn_id, &fi)) return set_err(h, NLB_ERR_UNKNOWN_FCN, "unknown residual id");
    const int nv = fi.n != 0 ? fi.n : n;
    if (nv <= 0 || nv > CLS_MAX_N) return set_err(h, NLB_ERR_UNSUPPORTED, "constrained least squares: n <= 16");
    DevCls o;
    // cls_set_radius / cls_set_factor: a non-positive value selects 1 (least_squares:902-909, 927-934)
    o.radius = options->trust_region_radius > 0.0 ? options->trust_region_radius : 1.0;
    o.scaling = options->step_scaling_factor > 0.0 ? options->step_scaling_factor : 1.0;
    const double huge = 1.7976931348623157e+308;
    for (int i = 0; i < CLS_MAX_N; ++i) {
        o.xl[i] = (options->lower && i < nv) ? options->lower[i] : -huge;
        o.xu[i] = (options->upper && i < nv) ? options->upper[i] : huge;
    }
    return solve_batch(h, SOLVER_CLS, params, fcn_id, B, m, n, x, fvec, sys, shared, ib, status, stream, &o);
}

int nlb_polynomial_fit_batch(nlb_handle* h, int64_t B, int npts, int order, int thru_zero, int x_is_shared,
                             const double* x, const double* y, double* coeffs, int32_t* status, void* stream) {
    if (!h) return NLB_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::mutex> lock(h->mu);
    if (B < 0 || (B > 0 && (!x || !y || !coeffs))) return set_err(h, NLB_ERR_INVALID_ARGUMENT, "null argument or B < 0");
    // poly_fit: `order >= n .or. order < 1` -> error stop 4 (src/nonlin_polynomials.f90:166-169, 224-227)
    if (npts <= 0 || order < 1 || order >= npts) return set_err(h, NLB_ERR_SIZE, "need 1 <= order < npts");
    const int nc = thru_zero ? order : order + 1;
    if (nc > POLY_MAX_COLS) return set_err(h, NLB_ERR_UNSUPPORTED, "polynomial fit: at most 8 fitted coefficients");
    NLB_DEVICE(h);
    int rc = NLB_OK;
    if (B == 0) return NLB_OK;
    cudaStream_t s = stream ? (cudaStream_t)stream : h->stream;
    Staged ax, ay, ac, ast;
    const size_t xbytes = sizeof(double) * (size_t)npts * (x_is_shared ? 1 : (size_t)B);
    if ((rc = stage_in(h, x_is_shared ? 3 : 0, x, xbytes, true, s, &ax))) return rc;
    if ((rc = stage_in(h, 1, y, sizeof(double) * (size_t)npts * B, true, s, &ay))) return rc;
    if ((rc = stage_in(h, 2, coeffs, sizeof(double) * (size_t)(order + 1) * B, false, s, &ac))) return rc;
    if ((rc = stage_in(h, 5, status, sizeof(int32_t) * (size_t)B, false, s, &ast))) return rc;
    const size_t per_thread = sizeof(double) * (size_t)npts * (size_t)(nc + 1);
    const size_t smem_bytes = per_thread * 128;
    const size_t smem_max = 225 * 1024;                      // of the 227 KB a CTA may opt in to
    static const int force_global = std::getenv("NLB_POLYFIT_GLOBAL") ? 1 : 0;   // tuning knob
    if (smem_bytes <= smem_max && !force_global) {
        // shared-memory workspace: persistent grid of the CTAs that fit (each SM has 228 KB)
        const int ctas_per_sm = (int)((228 * 1024) / (smem_bytes + 1024)) > 0 ? (int)((228 * 1024) / (smem_bytes + 1024)) : 1;
        long long grid = (B + 127) / 128;
        const long long resident = (long long)h->num_sms * ctas_per_sm;
        if (grid > resident) grid = resident;
        switch (nc) {
#define X(NC)                                                                                                       \
    case NC:                                                                                                        \
        NLB_CUDA(h, cudaFuncSetAttribute(polyfit_kernel<NC, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,    \
                                         (int)smem_bytes));                                                        \
        polyfit_kernel<NC, true><<<(unsigned)grid, 128, smem_bytes, s>>>(                                           \
            B, npts, thru_zero != 0, x_is_shared != 0, (const double*)ax.dev, (const double*)ay.dev,                \
            (double*)ac.dev, (int32_t*)ast.dev, nullptr);                                                           \
        break;
            X(1) X(2) X(3) X(4) X(5) X(6) X(7) X(8)
#undef X
        }
    } else {
        // global workspace, persistent grid at full occupancy
        static const int threads_per_sm = [] {
            const char* e = std::getenv("NLB_POLYFIT_THREADS_PER_SM");   // tuning knob, multiple of 128
            return e ? std::atoi(e) : 0;
        }();
        long long T;
        if (threads_per_sm > 0) {
            T = (long long)h->num_sms * threads_per_sm;
        } else {
            // measured (B200, 100 x 6 and 512 x 6 fits): full occupancy beats keeping the workspace inside L2
            T = (long long)h->num_sms * 2048;
        }
        const long long need = ((B + 127) / 128) * 128;
        if (T > need) T = need;
        const size_t budget = (size_t)2 << 30;
        while (T > 128 && (size_t)T * per_thread > budget) T -= 128;
        if ((size_t)T * per_thread > budget) return set_err(h, NLB_ERR_UNSUPPORTED, "polynomial fit: npts too large");
        if (h->dwork_cap < (size_t)T * per_thread) {
            NLB_CUDA(h, cudaStreamSynchronize(s));
            if (h->dwork) NLB_CUDA(h, cudaFree(h->dwork));
            h->dwork = nullptr;
            h->dwork_cap = 0;
            NLB_CUDA(h, cudaMalloc(&h->dwork, (size_t)T * per_thread));
            h->dwork_cap = (size_t)T * per_thread;
        }
        const unsigned grid = (unsigned)(T / 128);
        // the workspace is shared by every call on this handle: a fit launched on another stream waits for the last one
        if (h->work_used) NLB_CUDA(h, cudaStreamWaitEvent(s, h->ev_work, 0));
        switch (nc) {
#define X(NC)                                                                                                       \
    case NC:                                                                                                        \
        polyfit_kernel<NC, false><<<grid, 128, 0, s>>>(B, npts, thru_zero != 0, x_is_shared != 0,                   \
                                                       (const double*)ax.dev, (const double*)ay.dev,                \
                                                       (double*)ac.dev, (int32_t*)ast.dev, (double*)h->dwork);      \
        break;
            X(1) X(2) X(3) X(4) X(5) X(6) X(7) X(8)
#undef X
        }
        NLB_CUDA(h, cudaEventRecord(h->ev_work, s));
        h->work_used = true;
    }
    ++h->launches;
    NLB_CUDA(h, cudaGetLastError());
    if ((rc = stage_out(h, ac, s))) return rc;
    if ((rc = stage_out(h, ast, s))) return rc;
    if (ax.staged || ay.staged || ac.staged || ast.staged) NLB_CUDA(h, cudaStreamSynchronize(s));
    return NLB_OK;
}

int nlb_polynomial_evaluate_batch(nlb_handle* h, int64_t B, int order, int npts, int x_is_shared,
                                  const double* coeffs, const double* x, double* y, void* stream) {
    if (!h) return NLB_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::mutex> lock(h->mu);
    if (B < 0 || npts < 0 || (B > 0 && npts > 0 && (!x || !y || !coeffs)))
        return set_err(h, NLB_ERR_INVALID_ARGUMENT, "null argument or negative size");
    if (order < 0 || order > POLY_MAX_COLS) return set_err(h, NLB_ERR_UNSUPPORTED, "polynomial evaluate: order 0..8");
    NLB_DEVICE(h);
    int rc = NLB_OK;
    if (B == 0 || npts == 0) return NLB_OK;
    cudaStream_t s = stream ? (cudaStream_t)stream : h->stream;
    Staged ax, ay, ac;
    const size_t xbytes = sizeof(double) * (size_t)npts * (x_is_shared ? 1 : (size_t)B);
    if ((rc = stage_in(h, x_is_shared ? 3 : 0, x, xbytes, true, s, &ax))) return rc;
    if ((rc = stage_in(h, 1, y, sizeof(double) * (size_t)npts * B, false, s, &ay))) return rc;
    if ((rc = stage_in(h, 2, coeffs, sizeof(double) * (size_t)(order + 1) * B, true, s, &ac))) return rc;
    const unsigned grid = (unsigned)((B + 127) / 128);
    polyval_kernel<<<grid, 128, 0, s>>>(B, order, npts, x_is_shared != 0, (const double*)ac.dev, (const double*)ax.dev,
                                        (double*)ay.dev);
    ++h->launches;
    NLB_CUDA(h, cudaGetLastError());
    if ((rc = stage_out(h, ay, s))) return rc;
    if (ax.staged || ay.staged || ac.staged) NLB_CUDA(h, cudaStreamSynchronize(s));
    return NLB_OK;
}

void nlb_params_1var_default(nlb_params_1var* p) {
    if (!p) return;
    p->max_fcn_evals = 100;
    p->fcn_tol = 1.0e-8;
    p->var_tol = 1.0e-12;
    p->diff_tol = 1.0e-12;
    p->use_analytic_diff = 0;
}

int nlb_fcn1var_count(void) { return FCN1_COUNT; }

int nlb_fcn1var_lookup(const char* name) {
    if (!name) return -1;
    for (int i = 0; i < FCN1_COUNT; ++i)
        if (std::strcmp(fcn1_table()[i].name, name) == 0) return i;
    return -1;
}

const char* nlb_fcn1var_name(int fcn_id) {
    if (fcn_id < 0 || fcn_id >= FCN1_COUNT) return nullptr;
    return fcn1_table()[fcn_id].name;
}

int nlb_fcn1var_info(int fcn_id, int* args_len, int* has_derivative) {
    if (fcn_id < 0 || fcn_id >= FCN1_COUNT) return NLB_ERR_UNKNOWN_FCN;
    if (args_len) *args_len = fcn1_table()[fcn_id].args_len;
    if (has_derivative) *has_derivative = fcn1_table()[fcn_id].has_diff;
    return NLB_OK;
}

int nlb_brent_solve_batch(nlb_handle* h, const nlb_params_1var* params, int fcn_id, int64_t B, const double* lim1,
                          const double* lim2, double* x, double* f, const double* args, nlb_iteration_behavior* ib,
                          int32_t* status, void* stream) {
    return solve_1var_batch(h, 0, params, fcn_id, B, lim1, lim2, x, f, args, ib, status, stream);
}

int nlb_newton_1var_solve_batch(nlb_handle* h, const nlb_params_1var* params, int fcn_id, int64_t B,
                                const double* lim1, const double* lim2, double* x, double* f, const double* args,
                                nlb_iteration_behavior* ib, int32_t* status, void* stream) {
    return solve_1var_batch(h, 1, params, fcn_id, B, lim1, lim2, x, f, args, ib, status, stream);
}

int nlb_vecfcn_eval_batch(nlb_handle* h, int fcn_id, int64_t B, int m, int n, const double* x, double* fvec,
                          const double* sys, const double* shared, void* stream) {
    if (!h) return NLB_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::mutex> lock(h->mu);
    if (B < 0 || (B > 0 && (!x || !fvec))) return set_err(h, NLB_ERR_INVALID_ARGUMENT, "null argument or B < 0");
    int sys_len, shared_len;
    int rc = check_sizes(h, fcn_id, &m, &n, &sys_len, &shared_len);
    if (rc) return rc;
    NLB_DEVICE(h);
    if (B == 0) return NLB_OK;
    cudaStream_t s = stream ? (cudaStream_t)stream : h->stream;
    Staged ax, af, as, ash;
    if ((rc = stage_in(h, 0, x, sizeof(double) * (size_t)n * B, true, s, &ax))) return rc;
    if ((rc = stage_in(h, 1, fvec, sizeof(double) * (size_t)m * B, false, s, &af))) return rc;
    if ((rc = stage_in(h, 2, sys, sizeof(double) * (size_t)sys_len * B, true, s, &as))) return rc;
    if ((rc = stage_in(h, 3, shared, sizeof(double) * (size_t)shared_len, true, s, &ash))) return rc;
    const unsigned grid = (unsigned)((B + TPS_BLOCK - 1) / TPS_BLOCK);
    switch (fcn_id) {
#define X(F)                                                                                                   \
    case F::ID:                                                                                                \
        tps_eval_kernel<F><<<grid, TPS_BLOCK, 0, s>>>(B, (const double*)ax.dev, (double*)af.dev,               \
                                                      (const double*)as.dev, (const double*)ash.dev);          \
        break;
        NLB_FIXED_FCNS(X)
#undef X
        default:
            if (const UserFcn* u = user_fcn(fcn_id)) {
                rc = u->eval ? u->eval(0, 0, B, (const double*)ax.dev, (double*)af.dev, (const double*)as.dev,
                                       (const double*)ash.dev, (void*)s)
                             : NLB_ERR_UNSUPPORTED;
            } else {
                rc = launch_coop_eval(fcn_id, B, m, n, (const double*)ax.dev, (double*)af.dev, (const double*)as.dev,
                                      (const double*)ash.dev, s);
            }
            if (rc) return set_err(h, rc, "no evaluation kernel for this residual");
    }
    ++h->launches;
    NLB_CUDA(h, cudaGetLastError());
    if ((rc = stage_out(h, af, s))) return rc;
    if (ax.staged || af.staged || as.staged || ash.staged) NLB_CUDA(h, cudaStreamSynchronize(s));
    return NLB_OK;
}

int nlb_jacobian_batch(nlb_handle* h, const nlb_params* params, int fcn_id, int64_t B, int m, int n, const double* x,
                       double* jac, const double* sys, const double* shared, void* stream) {
    if (!h) return NLB_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::mutex> lock(h->mu);
    if (!params || B < 0 || (B > 0 && (!x || !jac))) return set_err(h, NLB_ERR_INVALID_ARGUMENT, "null argument or B < 0");
    int sys_len, shared_len;
    int rc = check_sizes(h, fcn_id, &m, &n, &sys_len, &shared_len);
    if (rc) return rc;
    NLB_DEVICE(h);
    if (B == 0) return NLB_OK;
    cudaStream_t s = stream ? (cudaStream_t)stream : h->stream;
    Staged ax, aj, as, ash;
    if ((rc = stage_in(h, 0, x, sizeof(double) * (size_t)n * B, true, s, &ax))) return rc;
    if ((rc = stage_in(h, 1, jac, sizeof(double) * (size_t)m * n * B, false, s, &aj))) return rc;
    if ((rc = stage_in(h, 2, sys, sizeof(double) * (size_t)sys_len * B, true, s, &as))) return rc;
    if ((rc = stage_in(h, 3, shared, sizeof(double) * (size_t)shared_len, true, s, &ash))) return rc;
    const unsigned grid = (unsigned)((B + TPS_BLOCK - 1) / TPS_BLOCK);
    switch (fcn_id) {
#define X(F)                                                                                                    \
    case F::ID:                                                                                                 \
        tps_jacobian_kernel<F><<<grid, TPS_BLOCK, 0, s>>>(params->use_analytic_jacobian, B, (const double*)ax.dev, \
                                                          (double*)aj.dev, (const double*)as.dev,               \
                                                          (const double*)ash.dev);                              \
        break;
        NLB_FIXED_FCNS(X)
#undef X
        default:
            if (const UserFcn* u = user_fcn(fcn_id)) {
                rc = u->eval ? u->eval(1, params->use_analytic_jacobian, B, (const double*)ax.dev, (double*)aj.dev,
                                       (const double*)as.dev, (const double*)ash.dev, (void*)s)
                             : NLB_ERR_UNSUPPORTED;
            } else {
                rc = launch_coop_jacobian(fcn_id, B, m, n, (const double*)ax.dev, (double*)aj.dev, (const double*)as.dev,
                                          (const double*)ash.dev, s);
            }
            if (rc) return set_err(h, rc, "no Jacobian kernel for this residual");
    }
    ++h->launches;
    NLB_CUDA(h, cudaGetLastError());
    if ((rc = stage_out(h, aj, s))) return rc;
    if (ax.staged || aj.staged || as.staged || ash.staged) NLB_CUDA(h, cudaStreamSynchronize(s));
    return NLB_OK;
}

int nlb_reduce_stats(nlb_handle* h, int64_t B, const nlb_iteration_behavior* ib, const int32_t* status,
                     int64_t* stats, void* stream) {
    if (!h) return NLB_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::mutex> lock(h->mu);
    if (!stats || B < 0) return set_err(h, NLB_ERR_INVALID_ARGUMENT, "null stats or B < 0");
    NLB_DEVICE(h);
    int rc = NLB_OK;
    cudaStream_t s = stream ? (cudaStream_t)stream : h->stream;
    Staged aib, ast;
    if ((rc = stage_in(h, 4, ib, sizeof(nlb_iteration_behavior) * (size_t)B, true, s, &aib))) return rc;
    if ((rc = stage_in(h, 5, status, sizeof(int32_t) * (size_t)B, true, s, &ast))) return rc;
    const bool out_dev = is_device_ptr(stats);
    int64_t* d = out_dev ? stats : h->dstats;
    NLB_CUDA(h, cudaMemsetAsync(d, 0, sizeof(int64_t) * NLB_STAT_COUNT, s));
    if (B > 0) {
        unsigned grid = (unsigned)((B + 255) / 256);
        const unsigned cap = (unsigned)(h->num_sms > 0 ? h->num_sms : 148) * 4u;   // persistent-style: 4 CTAs per SM
        if (grid > cap) grid = cap;
        stats_kernel<<<grid, 256, 0, s>>>(B, (const nlb_iteration_behavior*)aib.dev, (const int32_t*)ast.dev,
                                          (unsigned long long*)d);
        ++h->launches;
        NLB_CUDA(h, cudaGetLastError());
    }
    if (!out_dev) {
        NLB_CUDA(h, cudaMemcpyAsync(stats, d, sizeof(int64_t) * NLB_STAT_COUNT, cudaMemcpyDeviceToHost, s));
        NLB_CUDA(h, cudaStreamSynchronize(s));
    }
    return NLB_OK;
}

int nlb_measure_fp64_latency(nlb_handle* h, double* cycles4) {
    if (!h || !cycles4) return NLB_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::mutex> lock(h->mu);
    NLB_DEVICE(h);
    int rc = NLB_OK;
    double* dout = nullptr;
    NLB_CUDA(h, cudaMalloc(&dout, 8 * sizeof(double)));
    for (int rep = 0; rep < 2; ++rep) {
        fp64_latency_kernel<<<1, 32, 0, h->stream>>>(dout, 20000, 1.0);
        ++h->launches;
    }
    double host[8];
    NLB_CUDA(h, cudaMemcpyAsync(host, dout, 5 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    NLB_CUDA(h, cudaStreamSynchronize(h->stream));
    cudaFree(dout);
    for (int i = 0; i < 4; ++i) cycles4[i] = host[i];
    return NLB_OK;
}

int nlb_measure_fp64_peak(nlb_handle* h, double* dfma_tflops, double* dadd_dmul_tflops) {
    if (!h) return NLB_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::mutex> lock(h->mu);
    NLB_DEVICE(h);
    int rc = NLB_OK;
    cudaDeviceProp prop;
    NLB_CUDA(h, cudaGetDeviceProperties(&prop, h->device));
    const int threads = 256, blocks = prop.multiProcessorCount * 8, iters = 20000;
    double* dout = (double*)h->dstats;
    cudaEvent_t e0, e1;
    NLB_CUDA(h, cudaEventCreate(&e0));
    NLB_CUDA(h, cudaEventCreate(&e1));
    double res[2] = {0.0, 0.0};
    for (int variant = 0; variant < 2; ++variant) {
        float best = 1e30f;
        for (int rep = 0; rep < 4; ++rep) {
            NLB_CUDA(h, cudaEventRecord(e0, h->stream));
            if (variant == 0) fp64_peak_kernel<true><<<blocks, threads, 0, h->stream>>>(dout, iters, 1.0);
            else fp64_peak_kernel<false><<<blocks, threads, 0, h->stream>>>(dout, iters, 1.0);
            ++h->launches;
            NLB_CUDA(h, cudaEventRecord(e1, h->stream));
            NLB_CUDA(h, cudaEventSynchronize(e1));
            float ms = 0.f;
            NLB_CUDA(h, cudaEventElapsedTime(&ms, e0, e1));
            if (rep > 0 && ms < best) best = ms;
        }
        const double flops = 2.0 * 8.0 * (double)iters * (double)threads * (double)blocks;
        res[variant] = flops / (best * 1e-3) / 1e12;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (dfma_tflops) *dfma_tflops = res[0];
    if (dadd_dmul_tflops) *dadd_dmul_tflops = res[1];
    return NLB_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------
// one batch over several GPUs of one process
// ---------------------------------------------------------------------------------------
namespace {

// NCCL is resolved at run time (libnccl.so.2: the copy torch has already loaded, or the system one), so that the
// engine has no link-time dependency on it and single-GPU hosts never touch it.
struct NcclApi {
    typedef void* comm_t;
    int (*CommInitAll)(comm_t*, int, const int*) = nullptr;
    int (*CommDestroy)(comm_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, comm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool ok = false;
    NcclApi() {
        void* lib = nullptr;
        for (const char* name : {"libnccl.so.2", "libnccl.so"})
            if ((lib = dlopen(name, RTLD_LAZY | RTLD_GLOBAL))) break;
        if (!lib) return;
        CommInitAll = reinterpret_cast<decltype(CommInitAll)>(dlsym(lib, "ncclCommInitAll"));
        CommDestroy = reinterpret_cast<decltype(CommDestroy)>(dlsym(lib, "ncclCommDestroy"));
        GroupStart = reinterpret_cast<decltype(GroupStart)>(dlsym(lib, "ncclGroupStart"));
        GroupEnd = reinterpret_cast<decltype(GroupEnd)>(dlsym(lib, "ncclGroupEnd"));
        AllReduce = reinterpret_cast<decltype(AllReduce)>(dlsym(lib, "ncclAllReduce"));
        GetErrorString = reinterpret_cast<decltype(GetErrorString)>(dlsym(lib, "ncclGetErrorString"));
        ok = CommInitAll && CommDestroy && GroupStart && GroupEnd && AllReduce;
    }
};
const NcclApi& nccl_api() { static const NcclApi api; return api; }
constexpr int kNcclInt64 = 4, kNcclSum = 0, kNcclMax = 2;      // ncclDataType_t / ncclRedOp_t values (nccl.h)

// communicators of one device list, created on first use and kept for the life of the process
struct CommSet {
    std::vector<int> devs;
    std::vector<NcclApi::comm_t> comms;
};
std::mutex g_comm_mu;
std::vector<CommSet*> g_comm_sets;

int comms_for(const std::vector<int>& devs, CommSet** out, std::string* err) {
    std::lock_guard<std::mutex> lock(g_comm_mu);
    for (CommSet* cs : g_comm_sets)
        if (cs->devs == devs) { *out = cs; return NLB_OK; }
    const NcclApi& api = nccl_api();
    if (!api.ok) { *err = "NCCL (libnccl.so.2) could not be loaded"; return NLB_ERR_UNSUPPORTED; }
    CommSet* cs = new CommSet();
    cs->devs = devs;
    cs->comms.resize(devs.size());
    const int r = api.CommInitAll(cs->comms.data(), (int)devs.size(), devs.data());
    if (r != 0) {
        *err = std::string("ncclCommInitAll: ") + (api.GetErrorString ? api.GetErrorString(r) : "error");
        delete cs;
        return NLB_ERR_CUDA;
    }
    g_comm_sets.push_back(cs);
    *out = cs;
    return NLB_OK;
}

}  // namespace

extern "C" int nlb_solve_sharded(nlb_handle* const* handles, int ndev, int solver, const nlb_params* params, int fcn_id,
                                 int64_t B, int m, int n, double* x, double* fvec, const double* sys, const double* shared,
                                 nlb_iteration_behavior* ib, int32_t* status, int64_t* stats) {
    if (!handles || ndev <= 0) return NLB_ERR_INVALID_ARGUMENT;
    for (int d = 0; d < ndev; ++d)
        if (!handles[d]) return NLB_ERR_INVALID_ARGUMENT;
    nlb_handle* h0 = handles[0];
    if (solver != NLB_SOLVER_LEAST_SQUARES && solver != NLB_SOLVER_NEWTON && solver != NLB_SOLVER_QUASI_NEWTON)
        return set_err(h0, NLB_ERR_INVALID_ARGUMENT, "solver: NLB_SOLVER_LEAST_SQUARES / _NEWTON / _QUASI_NEWTON");
    for (int d = 0; d < ndev; ++d)
        for (int e = 0; e < d; ++e)
            if (handles[d]->device == handles[e]->device)
                return set_err(h0, NLB_ERR_INVALID_ARGUMENT, "one handle per device: two handles share a device");
    // contiguous system ranges, sizes differing by at most one; one host thread per device
    std::vector<int> rc(ndev, NLB_OK);
    std::vector<std::thread> workers;
    const int64_t base = B / ndev, rem = B % ndev;
    for (int d = 0; d < ndev; ++d) {
        const int64_t lo = d * base + (d < rem ? d : rem);
        const int64_t cnt = base + (d < rem ? 1 : 0);
        workers.emplace_back([=, &rc] {
            rc[d] = solve_batch(handles[d], solver, params, fcn_id, B, m, n, x, fvec, sys, shared, ib, status, nullptr,
                                nullptr, lo, cnt, stats != nullptr);
        });
    }
    for (std::thread& t : workers) t.join();
    for (int d = 0; d < ndev; ++d)
        if (rc[d] != NLB_OK) {
            if (d != 0) set_err(h0, rc[d], handles[d]->last_error.c_str());
            return rc[d];
        }
    if (!stats) return NLB_OK;
    // the one collective of the path: convergence statistics, SUM (MAX for max_iter), NCCL over NVLink
    if (ndev > 1) {
        std::vector<int> devs(ndev);
        for (int d = 0; d < ndev; ++d) devs[d] = handles[d]->device;
        CommSet* cs = nullptr;
        std::string err;
        const int r = comms_for(devs, &cs, &err);
        if (r != NLB_OK) return set_err(h0, r, err.c_str());
        const NcclApi& api = nccl_api();
        int nr = api.GroupStart();
        for (int d = 0; d < ndev && nr == 0; ++d) {
            int64_t* v = handles[d]->dstats;
            nr = api.AllReduce(v, v, NLB_STAT_MAX_ITER, kNcclInt64, kNcclSum, cs->comms[d], handles[d]->stream);
            if (nr == 0)
                nr = api.AllReduce(v + NLB_STAT_MAX_ITER, v + NLB_STAT_MAX_ITER, 1, kNcclInt64, kNcclMax, cs->comms[d],
                                   handles[d]->stream);
        }
        const int ne = api.GroupEnd();
        if (nr == 0) nr = ne;
        if (nr != 0) return set_err(h0, NLB_ERR_CUDA, api.GetErrorString ? api.GetErrorString(nr) : "NCCL error");
    }
    {
        DeviceGuard guard(h0->device);
        if (guard.err != cudaSuccess) return set_err(h0, NLB_ERR_NO_DEVICE, "cudaSetDevice", guard.err);
        NLB_CUDA(h0, cudaMemcpyAsync(stats, h0->dstats, sizeof(int64_t) * NLB_STAT_COUNT, cudaMemcpyDeviceToHost, h0->stream));
        NLB_CUDA(h0, cudaStreamSynchronize(h0->stream));
    }
    for (int d = 1; d < ndev; ++d) {
        DeviceGuard guard(handles[d]->device);
        cudaStreamSynchronize(handles[d]->stream);
    }
    return NLB_OK;
}
