// coop_lm_cta.cuh — CTA-per-system Levenberg-Marquardt for TALL curve fits (BASELINE config 4:
// m = 4096, n = 16): one CTA of n warps owns one system at a time, warp w owns Jacobian column w.
//
// Behaviour reproduced: lss_solve / lmpar / lmfactor / lmsolve, reference
// src/nonlin_least_squares.f90:118-391 / 394-566 / 569-667 / 670-791, and vfh_jac_fcn,
// src/nonlin_multi_eqn_mult_var.f90:198-277 (the n-sized pieces are shared with coop_lm.cuh).
//
// Why a second mapping.  The lane-per-system kernel (coop_lm.cuh) streams each column with ONE
// thread, so one outer iteration of a lone system is ~75 dependent passes over 4096 rows (17 ms);
// a batch that contains a single slow system (the noise-free fits contain ~0.7 % that iterate to the
// evaluation budget) then waits seconds for it.  Here every pass over the m rows is done by a whole
// warp (coalesced 256 B loads of a column-major Jacobian) or the whole CTA, and only the strictly
// sequential part of each reduction is left to one lane:
//
//   * dot products:  the lanes form the products a(i)*b(i) of a 256-row segment in parallel and park
//     them in a warp-private shared-memory segment; lane 0 then adds them in index order.  Each
//     product and each addition is the one the reference performs, in the same order (no FMA), so
//     the sum is bit-identical; the chain costs one DADD per row from shared memory instead of a
//     dependent global round trip per 8 rows.
//   * NORM2: the running scale of libgfortran's recurrence is the prefix maximum of |x|, obtained per
//     32-row group with a warp max-scan; the lanes then form the quotients (all the divisions) in
//     parallel, lane 0 replays the ssq recurrence over them in order.
//
// n = 16 columns -> 16 dot chains run concurrently in 16 warps.  J (m x n, column-major), fvec and
// wa4 of the system being solved live in an HBM/L2 workspace owned by the CTA; everything n-sized is
// in shared memory.  CTAs are persistent and take systems from a work queue.
#pragma once
#include "coop_lm.cuh"

namespace nlb {

constexpr int WLM_SEG = 256;   // rows per staged segment of a warp

template <int N>
struct WlmSmem {
    static constexpr int NGRP = N * WLM_SEG / 32;     // 32-row groups the CTA staging buffer can hold
    static constexpr int NDV = 7 * N + N * N + 16 + N * WLM_SEG + NGRP;
    static constexpr int NIV = 2 * N + 16;
    static constexpr size_t BYTES = (size_t)NDV * sizeof(double) + (size_t)NIV * sizeof(int) + 16;
};

// sum_{i=i0}^{i1-1} prod(i), added in index order starting from 0.0; result in every lane.
template <class P>
NLB_DEV double wlm_seq_sum(int i0, int i1, double* stage, int lane, P prod) {
    double s = 0.0;
    for (int base = i0; base < i1; base += WLM_SEG) {
        const int cnt = (i1 - base < WLM_SEG) ? i1 - base : WLM_SEG;
#pragma unroll 4
        for (int u = lane; u < cnt; u += 32) stage[u] = prod(base + u);
        __syncwarp();
        if (lane == 0) {
            int t = 0;
            for (; t + 8 <= cnt; t += 8) {
                double q[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) q[u] = stage[t + u];
#pragma unroll
                for (int u = 0; u < 8; ++u) s += q[u];
            }
            for (; t < cnt; ++t) s += stage[t];
        }
        __syncwarp();
    }
    return __shfl_sync(0xffffffffu, s, 0);
}

// Continue libgfortran's NORM2 recurrence (scale, ssq) over load(i), i in [i0, i1), in index order.
// On entry and exit `scale` and `ssq` are the same in every lane.
template <class L>
NLB_DEV void wlm_norm2(double& scale, double& ssq, int i0, int i1, double* stage, int lane, L load) {
    for (int base = i0; base < i1; base += WLM_SEG) {
        const int cnt = (i1 - base < WLM_SEG) ? i1 - base : WLM_SEG;
        for (int u0 = 0; u0 < cnt; u0 += 32) {
            const bool valid = u0 + lane < cnt;
            const double x = valid ? load(base + u0 + lane) : 0.0;
            const double a = fabs(x);
            // inclusive prefix maximum of |x| over the lanes (= row order); NaN never raises the scale
            double pm = (a == a) ? a : 0.0;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const double t = __shfl_up_sync(0xffffffffu, pm, d);
                if (lane >= d && t > pm) pm = t;
            }
            double excl = __shfl_up_sync(0xffffffffu, pm, 1);
            if (lane == 0) excl = 0.0;
            const double sc = (scale < excl) ? excl : scale;          // running scale entering this element
            double q = 0.0;
            if (valid && x != 0.0) {
                const bool up = sc < a;
                const double t = (up ? sc : a) / (up ? a : sc);
                q = up ? -fmax(t, 4.9406564584124654e-324) : t;
            }
            if (valid) stage[u0 + lane] = q;
            const double last = __shfl_sync(0xffffffffu, pm, 31);
            if (scale < last) scale = last;
        }
        __syncwarp();
        if (lane == 0) {
            int t = 0;
            for (; t + 8 <= cnt; t += 8) {
                double q[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) q[u] = stage[t + u];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    if (q[u] < 0.0) { const double tt = -q[u]; ssq = 1.0 + ssq * tt * tt; }
                    else ssq = ssq + q[u] * q[u];
                }
            }
            for (; t < cnt; ++t) {
                const double q = stage[t];
                if (q < 0.0) { const double tt = -q; ssq = 1.0 + ssq * tt * tt; }
                else ssq = ssq + q * q;
            }
        }
        __syncwarp();
    }
    ssq = __shfl_sync(0xffffffffu, ssq, 0);
}


// NORM2 of load(i), i in [i0, i1), by the whole CTA (all threads must call; result valid in thread 0 after the
// last barrier inside).  Same recurrence, same order as libgfortran:
//   1. 32-row groups are dealt round-robin to the warps; each finds its group's max|x| (coalesced load + reduce)
//   2. an exclusive prefix maximum over the groups gives the running scale entering each group; inside a group a warp
//      max-scan gives it per element; the lanes then form the quotients (all divisions, in parallel) and store
//      t*t for ordinary elements, -t for an element that raises the scale, into the CTA's staging buffer
//   3. thread 0 replays ssq in index order: a plain add chain, with the rare scale-raising element handled apart
// Requires i1 - i0 <= N * WLM_SEG (the staging buffer); gmax: (N * WLM_SEG / 32) doubles of shared memory.
template <int N, class L>
NLB_DEV double wlm_cta_norm2(int i0, int i1, double* qbuf, double* gmax, int tid, L load) {
    const int lane = tid & 31, w = tid >> 5;
    const int L_ = i1 - i0;
    const int ngroups = (L_ + 31) / 32;
    for (int g = w; g < ngroups; g += N) {
        const int i = i0 + g * 32 + lane;
        const double a = (i < i1) ? fabs(load(i)) : 0.0;
        double pm = (a == a) ? a : 0.0;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            const double t = __shfl_xor_sync(0xffffffffu, pm, d);
            if (t > pm) pm = t;
        }
        if (lane == 0) gmax[g] = pm;
    }
    __syncthreads();
    for (int g = w; g < ngroups; g += N) {
        // running scale entering group g: max(1, maxima of the groups before it)
        double sc0 = 1.0;
        for (int h = lane; h < g; h += 32) { const double t = gmax[h]; if (t > sc0) sc0 = t; }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            const double t = __shfl_xor_sync(0xffffffffu, sc0, d);
            if (t > sc0) sc0 = t;
        }
        const int i = i0 + g * 32 + lane;
        const bool valid = i < i1;
        const double x = valid ? load(i) : 0.0;
        const double a = fabs(x);
        double pm = (a == a) ? a : 0.0;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const double t = __shfl_up_sync(0xffffffffu, pm, d);
            if (lane >= d && t > pm) pm = t;
        }
        double excl = __shfl_up_sync(0xffffffffu, pm, 1);
        if (lane == 0) excl = 0.0;
        const double sc = (sc0 < excl) ? excl : sc0;
        double q = 0.0;
        if (valid && x != 0.0) {
            const bool up = sc < a;
            const double t = (up ? sc : a) / (up ? a : sc);
            q = up ? -fmax(t, 4.9406564584124654e-324) : t * t;
        }
        if (valid) qbuf[g * 32 + lane] = q;
    }
    __syncthreads();
    double result = 0.0;
    if (tid == 0) {
        double scale = 1.0;
        for (int g = 0; g < ngroups; ++g) { const double t = gmax[g]; if (t > scale) scale = t; }
        double ssq = 0.0;
        int t = 0;
        for (; t + 8 <= L_; t += 8) {
            double q[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) q[u] = qbuf[t + u];
            double mn = q[0];
#pragma unroll
            for (int u = 1; u < 8; ++u) mn = fmin(mn, q[u]);
            if (mn >= 0.0) {
#pragma unroll
                for (int u = 0; u < 8; ++u) ssq = ssq + q[u];
            } else {
#pragma unroll 1
                for (int u = 0; u < 8; ++u) {
                    if (q[u] < 0.0) { const double tt = -q[u]; ssq = 1.0 + ssq * tt * tt; }
                    else ssq = ssq + q[u];
                }
            }
        }
        for (; t < L_; ++t) {
            const double q = qbuf[t];
            if (q < 0.0) { const double tt = -q; ssq = 1.0 + ssq * tt * tt; }
            else ssq = ssq + q;
        }
        result = scale * sqrt(ssq);
    }
    __syncthreads();
    return result;
}

enum { WS_FNORM = 0, WS_PAR, WS_XNORM, WS_DELTA, WS_GNORM, WS_AJNORM, WS_AJJ, WS_PNORM, WS_TEMP, WS_F1 };
enum { WI_ITER = 0, WI_NEVAL, WI_NJAC, WI_FLAG, WI_FCN, WI_XCN, WI_GCN, WI_PIVOT, WI_ACCEPT, WI_NEXT };
enum { WN_INNER = 0, WN_OUTER = 1, WN_DONE = 2 };

template <class F, int N>
__global__ void __launch_bounds__(32 * N, 2)
wlm_kernel(DevParams p, long long B, long long nsys, int m, double* __restrict__ xg, double* __restrict__ fg,
           const double* __restrict__ sys, const double* __restrict__ shared, nlb_iteration_behavior* __restrict__ ibg,
           int32_t* __restrict__ statusg, double* __restrict__ ws, unsigned long long* __restrict__ cursor) {
    static_assert(F::N == N, "residual / kernel size mismatch");
    constexpr int T = 32 * N;
    extern __shared__ double smem[];
    using V = SVec<1>;
    using Mt = SMat<N, 1>;
    using IV = SIVec<1>;
    const V x{smem}, diag{smem + N}, qtf{smem + 2 * N}, wa1{smem + 3 * N}, wa2{smem + 4 * N}, wa3{smem + 5 * N},
        w4h{smem + 6 * N}, sc{smem + 7 * N + N * N};
    const Mt R{smem + 7 * N};
    double* stage_all = smem + 7 * N + N * N + 16;
    double* gmax = stage_all + N * WLM_SEG;
    int* ibase = reinterpret_cast<int*>(gmax + WlmSmem<N>::NGRP);
    const IV ipvt{ibase}, pos{ibase + N}, si{ibase + 2 * N};
    unsigned long long* cur_s = reinterpret_cast<unsigned long long*>(ibase + 2 * N + 16);

    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    double* stage = stage_all + w * WLM_SEG;

    // workspace of this CTA: J column-major (m per column), fvec, wa4, y
    double* J = ws + (size_t)blockIdx.x * (size_t)(N + 3) * m;
    double* fv = J + (size_t)N * m;
    double* w4 = fv + m;
    double* yv = w4 + m;                      // the system's observations, copied once (the batch stores them strided)
    double* Jw = J + (size_t)w * m;           // this warp's column

    const double eps = 0x1p-52;
    const double ftol = p.fcn_tol, xtol = p.var_tol, gtol = p.grad_tol, fac = p.lm_factor;

    for (;;) {
        __syncthreads();
        if (tid == 0) *cur_s = atomicAdd(cursor, 1ull);
        __syncthreads();
        const long long b = (long long)*cur_s;
        if (b >= nsys) break;
        const double* ysys = sys + b;                                  // y(i) = ysys[i*B]

        // ---- x, fvec = F(x), fnorm ---------------------------------------------------------
        if (tid < N) x[tid] = xg[(long long)tid * B + b];
        for (int i = tid; i < m; i += T) yv[i] = ysys[(long long)i * B];
        if (tid == 0) {
            si[WI_ITER] = 1; si[WI_NEVAL] = 1; si[WI_NJAC] = 0; si[WI_FLAG] = 0;
            si[WI_FCN] = 0; si[WI_XCN] = 0; si[WI_GCN] = 0;
            sc[WS_PAR] = 0.0; sc[WS_XNORM] = 0.0; sc[WS_DELTA] = 0.0; sc[WS_GNORM] = 0.0;
        }
        __syncthreads();
        {
            double xl[N];
#pragma unroll
            for (int j = 0; j < N; ++j) xl[j] = x[j];
            for (int i = tid; i < m; i += T) fv[i] = F::residual(xl, __ldg(shared + i), yv[i]);
        }
        __syncthreads();
        const bool fits = m <= N * WLM_SEG;       // the CTA staging buffer holds a whole m-vector of quotients
        if (fits) {
            const double fn = wlm_cta_norm2<N>(0, m, stage_all, gmax, tid, [&](int i) { return fv[i]; });
            if (tid == 0) sc[WS_FNORM] = fn;
        } else if (w == 0) {
            double scale = 1.0, ssq = 0.0;
            wlm_norm2(scale, ssq, 0, m, stage, lane, [&](int i) { return fv[i]; });
            if (lane == 0) sc[WS_FNORM] = scale * sqrt(ssq);
        }
        __syncthreads();

        for (;;) {   // ---- outer iteration: Jacobian, QR, Q^T f ------------------------------------
            // forward-difference column w and its norm (vfh_jac_fcn :262-275, lmfactor :611-616)
            {
                double xl[N];
                const double temp = x[w];
                double h = 0x1p-26 * fabs(temp);
                if (h == 0.0) h = 0x1p-26;
#pragma unroll
                for (int j = 0; j < N; ++j) xl[j] = (j == w) ? (temp + h) : x[j];
                double scale = 1.0, ssq = 0.0;
                wlm_norm2(scale, ssq, 0, m, stage, lane, [&](int i) {
                    const double v = (F::residual(xl, __ldg(shared + i), yv[i]) - fv[i]) / h;
                    Jw[i] = v;
                    return v;
                });
                if (lane == 0) {
                    const double cn = scale * sqrt(ssq);
                    wa2[w] = cn; wa1[w] = cn; wa3[w] = cn;   // acnorm (physical), rdiag, wa (logical)
                    ipvt[w] = w; pos[w] = w;
                }
                if (tid == 0) si[WI_NJAC] = si[WI_NJAC] + 1;
            }
            __syncthreads();

            // pivoted Householder QR, one logical column per step (lmfactor :619-666)
            for (int j = 0; j < N; ++j) {
                if (tid == 0) {
                    int kmax = j;
                    double rmax = wa1[j];
                    for (int c = j + 1; c < N; ++c) {
                        const double rc = wa1[c];
                        if (rc > rmax) { rmax = rc; kmax = c; }
                    }
                    if (kmax != j) {
                        wa1[kmax] = wa1[j];
                        wa3[kmax] = wa3[j];
                        const int pj = ipvt[j], pk = ipvt[kmax];
                        ipvt[j] = pk; ipvt[kmax] = pj;
                        pos[pk] = j; pos[pj] = kmax;
                    }
                    si[WI_PIVOT] = ipvt[j];
                }
                __syncthreads();
                const int pc = si[WI_PIVOT];
                const double* Jp = J + (size_t)pc * m;
                if (fits) {
                    double ajnorm = wlm_cta_norm2<N>(j, m, stage_all, gmax, tid, [&](int i) { return Jp[i]; });
                    if (tid == 0) {
                        if (ajnorm != 0.0 && Jp[j] < 0.0) ajnorm = -ajnorm;
                        sc[WS_AJNORM] = ajnorm;
                    }
                } else if (w == pc) {
                    double scale = 1.0, ssq = 0.0;
                    wlm_norm2(scale, ssq, j, m, stage, lane, [&](int i) { return Jw[i]; });
                    if (lane == 0) {
                        double ajnorm = scale * sqrt(ssq);
                        if (ajnorm != 0.0 && Jw[j] < 0.0) ajnorm = -ajnorm;
                        sc[WS_AJNORM] = ajnorm;
                    }
                }
                __syncthreads();
                const double ajnorm = sc[WS_AJNORM];
                if (ajnorm != 0.0) {
                    double* Jpw = J + (size_t)pc * m;
                    for (int i = j + tid; i < m; i += T) {
                        double t = Jpw[i] / ajnorm;
                        if (i == j) { t = t + 1.0; sc[WS_AJJ] = t; }
                        Jpw[i] = t;
                    }
                }
                __syncthreads();
                const int mypos = pos[w];
                if (ajnorm != 0.0 && mypos > j) {
                    const double sm = wlm_seq_sum(j, m, stage, lane, [&](int i) { return Jp[i] * Jw[i]; });
                    double temp = sm / sc[WS_AJJ];
#pragma unroll 4
                    for (int i = j + lane; i < m; i += 32) Jw[i] = Jw[i] - temp * Jp[i];
                    __syncwarp();
                    // norm down-date (lmfactor :656-661): lane 0 decides, the warp follows
                    double rd = 0.0;
                    int recompute = 0;
                    if (lane == 0) {
                        rd = wa1[mypos];
                        if (rd != 0.0) {
                            temp = Jw[j] / rd;
                            rd = rd * sqrt(nl_max(0.0, 1.0 - temp * temp));
                            const double q = rd / wa3[mypos];
                            recompute = !(0.05 * (q * q) > eps);
                            if (!recompute) wa1[mypos] = rd;
                        }
                    }
                    recompute = __shfl_sync(0xffffffffu, recompute, 0);
                    if (recompute) {
                        double scale = 1.0, ssq = 0.0;
                        wlm_norm2(scale, ssq, j + 1, m, stage, lane, [&](int i) { return Jw[i]; });
                        if (lane == 0) {
                            rd = scale * sqrt(ssq);
                            wa1[mypos] = rd;
                            wa3[mypos] = rd;
                        }
                    }
                }
                __syncthreads();
                if (tid == 0) wa1[j] = -ajnorm;
            }
            __syncthreads();

            // wa4 = fvec; qtf = first n of Q^T fvec; R = top block (lss_solve :241-253)
            for (int i = tid; i < m; i += T) w4[i] = fv[i];
            __syncthreads();
            for (int j = 0; j < N; ++j) {
                const double* Jp = J + (size_t)ipvt[j] * m;
                const double ajj = Jp[j];
                if (ajj != 0.0) {
                    if (w == 0) {
                        const double sm = wlm_seq_sum(j, m, stage, lane, [&](int i) { return Jp[i] * w4[i]; });
                        if (lane == 0) sc[WS_TEMP] = -sm / ajj;
                    }
                    __syncthreads();
                    const double temp = sc[WS_TEMP];
                    for (int i = j + tid; i < m; i += T) w4[i] = w4[i] + Jp[i] * temp;
                    __syncthreads();
                }
                if (tid == 0) qtf[j] = w4[j];
            }
            {
                const double* Jp = J + (size_t)ipvt[w] * m;      // logical column w of R
                if (lane < w) R(lane, w) = Jp[lane];
                if (lane == 0) R(w, w) = wa1[w];
            }
            __syncthreads();

            // scaling, scaled gradient norm, gradient test (lss_solve :229-278)
            if (tid == 0) {
                const int iter = si[WI_ITER];
                const double fnorm = sc[WS_FNORM];
                if (iter == 1) {
                    for (int j = 0; j < N; ++j) {
                        const double a = wa2[j];
                        diag[j] = (a == 0.0) ? 1.0 : a;
                    }
                    for (int j = 0; j < N; ++j) wa3[j] = diag[j] * x[j];
                    const double xnorm = clm_norm2<N>(wa3);
                    double delta = fac * xnorm;
                    if (delta == 0.0) delta = fac;
                    sc[WS_XNORM] = xnorm;
                    sc[WS_DELTA] = delta;
                }
                double gnorm = 0.0;
                if (fnorm != 0.0) {
                    for (int j = 0; j < N; ++j) {
                        const double acn = wa2[ipvt[j]];
                        if (acn == 0.0) continue;
                        double sm = 0.0;
                        for (int i = 0; i <= j; ++i) sm += R(i, j) * (qtf[i] / fnorm);
                        gnorm = nl_max(gnorm, fabs(sm / acn));
                    }
                }
                sc[WS_GNORM] = gnorm;
                if (gnorm <= gtol) {
                    si[WI_GCN] = 1;
                    si[WI_NEXT] = WN_DONE;
                } else {
                    for (int j = 0; j < N; ++j) diag[j] = nl_max(diag[j], wa2[j]);
                    si[WI_NEXT] = WN_INNER;
                }
            }
            __syncthreads();
            if (si[WI_NEXT] == WN_DONE) break;

            for (;;) {   // ---- inner iteration: LM parameter, trial point, gain ratio -------------------
                if (w == 0) {
                    // every lane of warp 0 evaluates lmpar on the same shared operands (identical values, uniform
                    // control flow), so that the m-length tail of its dxnorm can be formed by the whole warp
                    double par = sc[WS_PAR];
                    double delta = sc[WS_DELTA];
                    clm_par<N>(R, ipvt, diag, qtf, delta, par, wa1, wa2, wa3, w4h, [&](double& scale, double& ssq) {
                        __syncwarp();
                        wlm_norm2(scale, ssq, N, m, stage, lane, [&](int i) { return w4[i]; });
                    });
                    __syncwarp();
                    if (lane == 0) {
                        for (int j = 0; j < N; ++j) {
                            const double pj = -wa1[j];
                            wa1[j] = pj;
                            wa2[j] = x[j] + pj;
                            wa3[j] = diag[j] * pj;
                        }
                        const double pnorm = clm_norm2<N>(wa3);
                        if (si[WI_ITER] == 1) delta = nl_min(delta, pnorm);
                        sc[WS_PAR] = par;
                        sc[WS_DELTA] = delta;
                        sc[WS_PNORM] = pnorm;
                    }
                }
                __syncthreads();
                {   // wa4 = F(x + p)
                    double xl[N];
#pragma unroll
                    for (int j = 0; j < N; ++j) xl[j] = wa2[j];
                    for (int i = tid; i < m; i += T) w4[i] = F::residual(xl, __ldg(shared + i), yv[i]);
                }
                __syncthreads();
                if (fits) {
                    const double f1n = wlm_cta_norm2<N>(0, m, stage_all, gmax, tid, [&](int i) { return w4[i]; });
                    if (tid == 0) sc[WS_F1] = f1n;
                } else if (w == 0) {
                    double scale = 1.0, ssq = 0.0;
                    wlm_norm2(scale, ssq, 0, m, stage, lane, [&](int i) { return w4[i]; });
                    if (lane == 0) sc[WS_F1] = scale * sqrt(ssq);
                }
                __syncthreads();
                if (tid == 0) {   // lss_solve :297-365
                    int iter = si[WI_ITER];
                    const int neval = si[WI_NEVAL] + 1;
                    si[WI_NEVAL] = neval;
                    double fnorm = sc[WS_FNORM], par = sc[WS_PAR], delta = sc[WS_DELTA], xnorm = sc[WS_XNORM];
                    const double pnorm = sc[WS_PNORM], gnorm = sc[WS_GNORM], fnorm1 = sc[WS_F1];
                    double actred = -1.0;
                    if (0.1 * fnorm1 < fnorm) { const double q = fnorm1 / fnorm; actred = 1.0 - q * q; }
                    double temp = 0.0;
                    for (int j = 0; j < N; ++j) {
                        wa3[j] = 0.0;
                        temp = wa1[ipvt[j]];
                        for (int i = 0; i <= j; ++i) wa3[i] = wa3[i] + R(i, j) * temp;
                    }
                    const double temp1 = clm_norm2<N>(wa3) / fnorm;
                    const double temp2 = (sqrt(par) * pnorm) / fnorm;
                    const double prered = temp1 * temp1 + temp2 * temp2 / 0.5;
                    const double dirder = -(temp1 * temp1 + temp2 * temp2);
                    double ratio = 0.0;
                    if (prered != 0.0) ratio = actred / prered;
                    if (ratio <= 0.25) {
                        if (actred >= 0.0) temp = 0.5;
                        if (actred < 0.0) temp = 0.5 * dirder / (dirder + 0.5 * actred);
                        if (0.1 * fnorm1 >= fnorm || temp < 0.1) temp = 0.1;
                        delta = temp * nl_min(delta, pnorm / 0.1);
                        par = par / temp;
                    } else if (!(par != 0.0 && ratio < 0.75)) {
                        delta = pnorm / 0.5;
                        par = 0.5 * par;
                    }
                    const bool accept = ratio >= 1.0e-4;
                    if (accept) {
                        for (int j = 0; j < N; ++j) {
                            const double xn = wa2[j];
                            x[j] = xn;
                            wa2[j] = diag[j] * xn;
                        }
                        xnorm = clm_norm2<N>(wa2);
                        fnorm = fnorm1;
                        ++iter;
                    }
                    bool fcnvrg = false, xcnvrg = false;
                    if (fabs(actred) <= ftol && prered <= ftol && 0.5 * ratio <= 1.0) fcnvrg = true;
                    if (delta <= xtol * xnorm) xcnvrg = true;
                    int flag = 0, next = WN_INNER;
                    if (fcnvrg || xcnvrg) {
                        next = WN_DONE;
                    } else {
                        if (neval >= p.max_fcn_evals) flag = NLB_CONVERGENCE_ERROR;
                        if (fabs(actred) <= eps && prered <= eps && 0.5 * ratio <= 1.0) flag = NLB_TOLERANCE_TOO_SMALL_ERROR;
                        if (delta <= eps * xnorm) flag = NLB_TOLERANCE_TOO_SMALL_ERROR;
                        if (gnorm <= eps) flag = NLB_TOLERANCE_TOO_SMALL_ERROR;
                        if (flag != 0) next = WN_DONE;
                        else if (accept) next = WN_OUTER;
                    }
                    si[WI_FCN] = fcnvrg; si[WI_XCN] = xcnvrg; si[WI_FLAG] = flag;
                    si[WI_ITER] = iter; si[WI_ACCEPT] = accept; si[WI_NEXT] = next;
                    sc[WS_FNORM] = fnorm; sc[WS_PAR] = par; sc[WS_DELTA] = delta; sc[WS_XNORM] = xnorm;
                }
                __syncthreads();
                if (si[WI_ACCEPT]) {
                    for (int i = tid; i < m; i += T) fv[i] = w4[i];
                }
                __syncthreads();
                if (si[WI_NEXT] != WN_INNER) break;
            }
            if (si[WI_NEXT] == WN_DONE) break;
        }

        // ---- results -----------------------------------------------------------------------------
        if (tid < N) xg[(long long)tid * B + b] = x[tid];
        for (int i = tid; i < m; i += T) fg[(long long)i * B + b] = fv[i];
        if (tid == 0) {
            if (ibg) {
                nlb_iteration_behavior o;
                o.iter_count = si[WI_ITER]; o.fcn_count = si[WI_NEVAL]; o.jacobian_count = si[WI_NJAC]; o.gradient_count = 0;
                o.converge_on_fcn = si[WI_FCN]; o.converge_on_chng = si[WI_XCN]; o.converge_on_zero_diff = si[WI_GCN];
                ibg[b] = o;
            }
            if (statusg) statusg[b] = si[WI_FLAG] != 0 ? NLB_CONVERGENCE_ERROR : NLB_NO_ERROR;
        }
    }
}

}  // namespace nlb
