// coop_lm.cuh — Levenberg-Marquardt for tall curve fits with run-time m (BASELINE config 4:
// m = 4096, n = 16; also the 4-parameter fits with m = 64).
//
// Behaviour reproduced: lss_solve / lmpar / lmfactor / lmsolve, reference
// src/nonlin_least_squares.f90:118-391 / 394-566 / 569-667 / 670-791, and the forward-difference
// Jacobian vfh_jac_fcn, src/nonlin_multi_eqn_mult_var.f90:198-277.
//
// Mapping.  A CTA owns 32 systems.  Thread (lane, k) = (system, Jacobian column): warp k holds
// column k of 32 different systems.  The m x n Jacobian (512 KB per system at C4 — it fits neither
// shared memory nor registers) lives in an HBM workspace laid out [group][row][column][lane], so a
// warp reads 256 contiguous bytes per row and the 16 warps of a CTA stream 4 KB-contiguous rows.
// The two m-vectors (fvec, wa4) use [group][row][lane].  Everything of size n or n x n (R, diag,
// qtf, the LM work vectors, pivots, per-system scalars and state) is in shared memory, [element]
// [lane], conflict-free.
//
// Parity.  Every m-length sum is walked sequentially by ONE thread in the reference's index
// order (thread k sums its own column: column norms, the Householder dot products, the FD
// differences), so the result is bit-identical to the CPU oracle; the parallelism comes from the
// n columns and from the 32 systems per warp, not from splitting a sum.  Column pivoting is a
// permutation table (ipvt / pos) instead of a physical swap — same values, no data movement.
// Row-wise elementwise passes (residual evaluation, wa4 += a(:,j)*temp, copies) are split over
// the n threads of a system.  The n-sized serial parts (lmpar, lmsolve, gain ratio) run on warp 0.
//
// Systems advance in lock step through the phases below (CTA barriers in uniform control flow);
// a system that is retrying a rejected step idles through the Jacobian/QR phases of its
// neighbours, a finished system idles until the whole CTA is done (refilling finished lanes from
// a work queue is the next step).
#pragma once
#include "tps_common.cuh"

namespace nlb {

// lane states: a lane holds one system at a time and takes the next one from a work queue when it is done
enum { CLM_NEED_JAC = 0, CLM_INNER = 1, CLM_DONE = 2, CLM_EMPTY = 3, CLM_NEW = 4, CLM_EXIT = 5 };

template <int N>
struct CoopLmSmem {
    static constexpr int NDV = 7 * N + N * N + 12;   // doubles per system
    static constexpr int NIV = 2 * N + 12;           // ints per system
    static constexpr size_t BYTES = 32 * ((size_t)NDV * sizeof(double) + (size_t)NIV * sizeof(int));
};

// L2 prefetch of a line that a later batch of the same sequential pass will load (turns an HBM round
// trip per batch into an L2 hit)
NLB_DEV void clm_prefetch(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
constexpr int CLM_PF = 48;   // rows ahead

// per-lane views into [element][32] shared arrays
template <int S>
struct SVec {
    double* p;
    NLB_DEV double& operator[](int i) const { return p[i * S]; }
};
template <int N, int S>
struct SMat {
    double* p;
    NLB_DEV double& operator()(int i, int j) const { return p[(i + j * N) * S]; }
};
template <int S>
struct SIVec {
    int* p;
    NLB_DEV int& operator[](int i) const { return p[i * S]; }
};
using LaneVec = SVec<32>;          // [element][lane] arrays of the lane-per-system kernel
template <int N>
using LaneMat = SMat<N, 32>;
using LaneIVec = SIVec<32>;

template <int N, class V>
NLB_DEV double clm_norm2(const V& v) {
    Norm2 acc;
#pragma unroll 1
    for (int i = 0; i < N; ++i) acc.add(v[i]);      // rolled: a dozen call sites x N inlined divisions otherwise
    return acc.value();
}


// Norm2 of v[i0*stride], ..., v[(i1-1)*stride] in index order; loads are issued eight at a time so that
// their latency overlaps (the accumulation itself stays strictly sequential).
NLB_DEV void clm_norm2_strided(Norm2& acc, const double* __restrict__ v, long long stride, int i0, int i1) {
    int i = i0;
    for (; i + 8 <= i1; i += 8) {
        double t[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) t[u] = v[(long long)(i + u) * stride];
#pragma unroll
        for (int u = 0; u < 8; ++u) acc.add(t[u]);
    }
    for (; i < i1; ++i) acc.add(v[(long long)i * stride]);
}

// libgfortran NORM2 of v[i0..i1) (element i at v[i*stride]) evaluated by the N threads of a system with the
// SAME result as the one-thread recurrence:
//   1. thread k finds max|v| of its contiguous chunk                       (parallel, exact)
//   2. the scale entering chunk k is max(1, maxima of the earlier chunks): exactly the running scale
//      of the sequential algorithm, which only ever grows to the largest magnitude seen so far
//   3. thread k forms, for each of its elements, the quotient the recurrence would form (scale/a when
//      the element raises the scale, a/scale otherwise) and stores it in tq, negated in the first case
//                                                                            (parallel: all divisions)
//   4. the leader replays ssq = 1 + ssq*t*t  /  ssq += t*t over tq in index order (sequential, cheap)
// All threads of the CTA must call this (three barriers); `on` = the lane takes part.
template <int N>
NLB_DEV double clm_coop_norm2(bool on, const double* __restrict__ v, long long stride, int i0, int i1,
                              double* __restrict__ tq, const LaneVec& cmax, int k, bool leader) {
    const int L = i1 - i0;
    const int len = (L + N - 1) / N;
    const int c0 = i0 + k * len;
    const int c1 = (c0 + len < i1) ? c0 + len : i1;
    if (on) {
        double mx = 0.0;
        int i = c0;
        for (; i + 4 <= c1; i += 4) {
            double t[4];
            if (i + CLM_PF + 4 <= c1) {
#pragma unroll
                for (int u = 0; u < 4; ++u) clm_prefetch(&v[(long long)(i + CLM_PF + u) * stride]);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) t[u] = fabs(v[(long long)(i + u) * stride]);
#pragma unroll
            for (int u = 0; u < 4; ++u) if (t[u] > mx) mx = t[u];
        }
        for (; i < c1; ++i) { const double a = fabs(v[(long long)i * stride]); if (a > mx) mx = a; }
        cmax[k] = mx;
    }
    __syncthreads();
    if (on) {
        double sc = 1.0;
        for (int c = 0; c < k; ++c) { const double a = cmax[c]; if (sc < a) sc = a; }
        int i = c0;
        for (; i + 4 <= c1; i += 4) {
            double x[4], q[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) x[u] = v[(long long)(i + u) * stride];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                q[u] = 0.0;
                if (x[u] != 0.0) {
                    const double a = fabs(x[u]);
                    const bool up = sc < a;
                    const double t = (up ? sc : a) / (up ? a : sc);
                    q[u] = up ? -fmax(t, 4.9406564584124654e-324) : t;
                    sc = up ? a : sc;
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) tq[(long long)(i + u) * 32] = q[u];
        }
        for (; i < c1; ++i) {
            const double x = v[(long long)i * stride];
            double q = 0.0;
            if (x != 0.0) {
                const double a = fabs(x);
                const bool up = sc < a;
                const double t = (up ? sc : a) / (up ? a : sc);
                q = up ? -fmax(t, 4.9406564584124654e-324) : t;
                sc = up ? a : sc;
            }
            tq[(long long)i * 32] = q;
        }
    }
    __syncthreads();
    double result = 0.0;
    if (on && leader) {
        double scale = 1.0;
        for (int c = 0; c < N; ++c) { const double a = cmax[c]; if (scale < a) scale = a; }
        double ssq = 0.0;
        int i = i0;
        for (; i + 8 <= i1; i += 8) {
            double q[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) q[u] = tq[(long long)(i + u) * 32];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                if (q[u] < 0.0) { const double t = -q[u]; ssq = 1.0 + ssq * t * t; }
                else ssq = ssq + q[u] * q[u];
            }
        }
        for (; i < i1; ++i) {
            const double q = tq[(long long)i * 32];
            if (q < 0.0) { const double t = -q; ssq = 1.0 + ssq * t * t; }
            else ssq = ssq + q * q;
        }
        result = scale * sqrt(ssq);
    }
    __syncthreads();
    return result;
}

// lmsolve on the n x n block held in shared memory (strict lower triangle = scratch for S^T).
template <int N, class Mt, class IV, class V>
NLB_DEV void clm_qrsolve(const Mt& r, const IV& ipvt, const V& diag, const V& qtb, const V& x, const V& sdiag,
                         const V& wa) {
    for (int j = 0; j < N; ++j) {
        for (int i = j; i < N; ++i) r(i, j) = r(j, i);
        x[j] = r(j, j);
        wa[j] = qtb[j];
    }
    for (int j = 0; j < N; ++j) {
        const double dl = diag[ipvt[j]];
        if (dl != 0.0) {
            for (int k = j; k < N; ++k) sdiag[k] = 0.0;
            sdiag[j] = dl;
            double qtbpj = 0.0;
            for (int k = j; k < N; ++k) {
                const double sk = sdiag[k];
                if (sk == 0.0) continue;
                double cs, sn;
                const double rkk = r(k, k);
                if (fabs(rkk) < fabs(sk)) {
                    const double ctan = rkk / sk;
                    sn = 0.5 / sqrt(0.25 + 0.25 * (ctan * ctan));
                    cs = sn * ctan;
                } else {
                    const double tn = sk / rkk;
                    cs = 0.5 / sqrt(0.25 + 0.25 * (tn * tn));
                    sn = cs * tn;
                }
                r(k, k) = cs * rkk + sn * sk;
                const double wk = wa[k];
                double temp = cs * wk + sn * qtbpj;
                qtbpj = -sn * wk + cs * qtbpj;
                wa[k] = temp;
                for (int i = k + 1; i < N; ++i) {
                    const double rik = r(i, k), si = sdiag[i];
                    temp = cs * rik + sn * si;
                    sdiag[i] = -sn * rik + cs * si;
                    r(i, k) = temp;
                }
            }
        }
        sdiag[j] = r(j, j);
        r(j, j) = x[j];
    }
    int nsing = N;
    for (int j = 0; j < N; ++j) {
        if (sdiag[j] == 0.0 && nsing == N) nsing = j;
        if (nsing < N) wa[j] = 0.0;
    }
    for (int j = nsing - 1; j >= 0; --j) {
        double sm = 0.0;
        for (int i = j + 1; i < nsing; ++i) sm += r(i, j) * wa[i];
        wa[j] = (wa[j] - sm) / sdiag[j];
    }
    for (int j = 0; j < N; ++j) x[ipvt[j]] = wa[j];
}

// lmpar.  wa2 is the reference's m-element work array: its first n entries are w4h (shared
// memory), entries n..m-1 are the tail of the HBM vector w4 (stride 32) — the in-loop dxnorm
// runs over all m of them (src/nonlin_least_squares.f90:531).
// `tail(scale, ssq)` continues the NORM2 recurrence over entries n..m-1 of the work array (how depends on the kernel).
template <int N, class Mt, class IV, class V, class Tail>
NLB_DEV void clm_par(const Mt& r, const IV& ipvt, const V& diag, const V& qtb, double delta, double& par, const V& x,
                     const V& sdiag, const V& wa1, const V& w4h, Tail tail) {
    const double dwarf = 0x1p-1022;
    int nsing = N;
    for (int j = 0; j < N; ++j) {
        wa1[j] = qtb[j];
        if (r(j, j) == 0.0 && nsing == N) nsing = j;
        if (nsing < N) wa1[j] = 0.0;
    }
    for (int j = nsing - 1; j >= 0; --j) {
        wa1[j] = wa1[j] / r(j, j);
        const double temp = wa1[j];
        for (int i = 0; i < j; ++i) wa1[i] = wa1[i] - r(i, j) * temp;
    }
    for (int j = 0; j < N; ++j) x[ipvt[j]] = wa1[j];

    int iter = 0;
    for (int j = 0; j < N; ++j) w4h[j] = diag[j] * x[j];
    double dxnorm = clm_norm2<N>(w4h);
    double fp = dxnorm - delta;
    if (fp <= 0.1 * delta) {
        par = 0.0;
        return;
    }
    double parl = 0.0;
    if (nsing == N) {
        for (int j = 0; j < N; ++j) {
            const int l = ipvt[j];
            wa1[j] = diag[l] * (w4h[l] / dxnorm);
        }
        for (int j = 0; j < N; ++j) {
            double sm = 0.0;
            for (int i = 0; i < j; ++i) sm += r(i, j) * wa1[i];
            wa1[j] = (wa1[j] - sm) / r(j, j);
        }
        const double temp = clm_norm2<N>(wa1);
        parl = ((fp / delta) / temp) / temp;
    }
    for (int j = 0; j < N; ++j) {
        double sm = 0.0;
        for (int i = 0; i <= j; ++i) sm += r(i, j) * qtb[i];
        wa1[j] = sm / diag[ipvt[j]];
    }
    const double gnorm = clm_norm2<N>(wa1);
    double paru = gnorm / delta;
    if (paru == 0.0) paru = dwarf / nl_min(delta, 0.1);
    par = nl_max(par, parl);
    par = nl_min(par, paru);
    if (par == 0.0) par = gnorm / dxnorm;

    for (;;) {
        ++iter;
        if (par == 0.0) par = nl_max(dwarf, 1.0e-3 * paru);
        double temp = sqrt(par);
        for (int j = 0; j < N; ++j) wa1[j] = temp * diag[j];
        clm_qrsolve<N>(r, ipvt, wa1, qtb, x, sdiag, w4h);
        for (int j = 0; j < N; ++j) w4h[j] = diag[j] * x[j];
        {
            Norm2 acc;
            for (int i = 0; i < N; ++i) acc.add(w4h[i]);
            tail(acc.scale, acc.ssq);
            dxnorm = acc.value();
        }
        temp = fp;
        fp = dxnorm - delta;
        if (fabs(fp) <= 0.1 * delta || (parl == 0.0 && fp <= temp && temp < 0.0) || iter == 10) break;
        for (int j = 0; j < N; ++j) {
            const int l = ipvt[j];
            wa1[j] = diag[l] * (w4h[l] / dxnorm);
        }
        for (int j = 0; j < N; ++j) {
            wa1[j] = wa1[j] / sdiag[j];
            temp = wa1[j];
            if (j + 1 < N)
                for (int i = 0; i < N; ++i) wa1[i] = wa1[i] - r(i, j) * temp;
        }
        temp = clm_norm2<N>(wa1);
        const double parc = ((fp / delta) / temp) / temp;
        if (fp > 0.0) parl = nl_max(parl, par);
        if (fp < 0.0) paru = nl_min(paru, par);
        par = nl_max(parl, par + parc);
    }
}

// scalar slots of a system in shared memory
enum { SC_FNORM = 0, SC_PAR, SC_XNORM, SC_DELTA, SC_GNORM, SC_AJNORM, SC_AJJ, SC_PNORM, SC_TEMP, SC_H, SC_NSC = 12 };
enum { SI_STATE = 0, SI_ITER, SI_NEVAL, SI_NJAC, SI_FLAG, SI_FCN, SI_XCN, SI_GCN, SI_PIVOT, SI_ACCEPT, SI_SYS, SI_NSI = 12 };

template <class F, int N>
__global__ void __launch_bounds__(32 * N)
coop_lm_kernel(DevParams p, long long B, long long b0, long long nsys, int m, double* __restrict__ xg,
               double* __restrict__ fg, const double* __restrict__ sys, const double* __restrict__ shared,
               nlb_iteration_behavior* __restrict__ ibg, int32_t* __restrict__ statusg, double* __restrict__ ws,
               unsigned long long* __restrict__ cursor) {
    static_assert(F::N == N, "residual / kernel size mismatch");
    extern __shared__ double smem[];
    const int lane = threadIdx.x & 31, k = threadIdx.x >> 5;

    // shared memory carve-up: [element][lane]
    double* sd = smem + lane;
    const LaneVec x{sd}, diag{sd + 32 * N}, qtf{sd + 64 * N}, wa1{sd + 96 * N}, wa2{sd + 128 * N}, wa3{sd + 160 * N},
        w4h{sd + 192 * N}, sc{sd + 32 * (7 * N + N * N)};
    const LaneMat<N> R{sd + 224 * N};
    int* si_base = reinterpret_cast<int*>(smem + 32 * CoopLmSmem<N>::NDV) + lane;
    const LaneIVec ipvt{si_base}, pos{si_base + 32 * N}, si{si_base + 64 * N};

    // HBM workspace of this group
    const long long gstride = (long long)(N + 3) * m * 32;
    double* J = ws + (long long)blockIdx.x * gstride + lane;        // J(i, c) = J[(i*N + c)*32]
    double* fv = J + (long long)m * N * 32;                         // fvec(i) = fv[i*32]
    double* w4 = fv + (long long)m * 32;                            // wa4(i)  = w4[i*32]
    double* tq = w4 + (long long)m * 32;                            // quotient scratch of the cooperative norms

    const double eps = 0x1p-52;
    const double ftol = p.fcn_tol, xtol = p.var_tol, gtol = p.grad_tol, fac = p.lm_factor;

    if (k == 0) { si[SI_STATE] = CLM_EMPTY; si[SI_SYS] = 0; si[SI_ACCEPT] = 0; }
    __syncthreads();

    for (;;) {
        // ---- retire finished systems, take new ones from the queue ---------------------------
        int st = si[SI_STATE];
        if (st == CLM_DONE) {
            const long long bo = b0 + si[SI_SYS];
            xg[(long long)k * B + bo] = x[k];
            for (int i = k; i < m; i += N) fg[(long long)i * B + bo] = fv[(long long)i * 32];
            if (k == 0) {
                if (ibg) {
                    nlb_iteration_behavior o;
                    o.iter_count = si[SI_ITER]; o.fcn_count = si[SI_NEVAL]; o.jacobian_count = si[SI_NJAC]; o.gradient_count = 0;
                    o.converge_on_fcn = si[SI_FCN]; o.converge_on_chng = si[SI_XCN]; o.converge_on_zero_diff = si[SI_GCN];
                    ibg[bo] = o;
                }
                if (statusg) statusg[bo] = si[SI_FLAG] != 0 ? NLB_CONVERGENCE_ERROR : NLB_NO_ERROR;
            }
        }
        __syncthreads();
        if (k == 0 && (st == CLM_DONE || st == CLM_EMPTY)) {
            // one atomic per warp for all lanes that need a system
            const unsigned mask = __activemask();
            const int leader = __ffs(mask) - 1;
            unsigned long long base = 0;
            if (lane == leader) base = atomicAdd(cursor, (unsigned long long)__popc(mask));
            base = __shfl_sync(mask, base, leader);
            const long long idx = (long long)(base + __popc(mask & ((1u << lane) - 1u)));
            if (idx < nsys) {
                si[SI_SYS] = (int)idx;
                si[SI_STATE] = CLM_NEW;
                si[SI_ITER] = 1; si[SI_NEVAL] = 1; si[SI_NJAC] = 0; si[SI_FLAG] = 0;
                si[SI_FCN] = 0; si[SI_XCN] = 0; si[SI_GCN] = 0; si[SI_ACCEPT] = 0;
                sc[SC_PAR] = 0.0; sc[SC_XNORM] = 0.0; sc[SC_DELTA] = 0.0; sc[SC_GNORM] = 0.0;
            } else {
                si[SI_STATE] = CLM_EXIT;
            }
        }
        __syncthreads();
        st = si[SI_STATE];
        if (__syncthreads_and(st == CLM_EXIT)) break;
        const long long b = b0 + si[SI_SYS];
        const double* ysys = sys + b;                               // y(i) = ysys[i*B]

        // ---- new system: load x, fvec = F(x), fnorm ------------------------------------------
        if (st == CLM_NEW) x[k] = xg[(long long)k * B + b];
        __syncthreads();
        if (st == CLM_NEW) {
            double xl[N];
#pragma unroll
            for (int j = 0; j < N; ++j) xl[j] = x[j];
            for (int i = k; i < m; i += N) fv[(long long)i * 32] = F::residual(xl, __ldg(shared + i), ysys[(long long)i * B]);
        }
        __syncthreads();
        {
            const double fn = clm_coop_norm2<N>(st == CLM_NEW, fv, 32, 0, m, tq, w4h, k, k == 0);
            if (st == CLM_NEW && k == 0) {
                sc[SC_FNORM] = fn;
                si[SI_STATE] = CLM_NEED_JAC;
            }
        }
        __syncthreads();

        const int state0 = si[SI_STATE];
        const bool needjac = state0 == CLM_NEED_JAC;

        // ---- phase J: forward-difference column k, its norm (vfh_jac_fcn :262-275, lmfactor :611-616)
        if (needjac) {
            double xl[N];
            const double temp = x[k];
            double h = 0x1p-26 * fabs(temp);
            if (h == 0.0) h = 0x1p-26;
#pragma unroll
            for (int j = 0; j < N; ++j) xl[j] = (j == k) ? (temp + h) : x[j];
            Norm2 acc;
            {
                int i = 0;
                for (; i + 4 <= m; i += 4) {
                    double tt[4], yy[4], ff[4], v[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        tt[u] = __ldg(shared + i + u);
                        yy[u] = ysys[(long long)(i + u) * B];
                        ff[u] = fv[(long long)(i + u) * 32];
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) v[u] = (F::residual(xl, tt[u], yy[u]) - ff[u]) / h;
#pragma unroll
                    for (int u = 0; u < 4; ++u) J[((long long)(i + u) * N + k) * 32] = v[u];
#pragma unroll
                    for (int u = 0; u < 4; ++u) acc.add(v[u]);
                }
                for (; i < m; ++i) {
                    const double f1 = F::residual(xl, __ldg(shared + i), ysys[(long long)i * B]);
                    const double v = (f1 - fv[(long long)i * 32]) / h;
                    J[((long long)i * N + k) * 32] = v;
                    acc.add(v);
                }
            }
            const double cn = acc.value();
            wa2[k] = cn;      // acnorm (physical column)
            wa1[k] = cn;      // rdiag  (logical position)
            wa3[k] = cn;      // wa
            ipvt[k] = k;
            pos[k] = k;
            if (k == 0) si[SI_NJAC] = si[SI_NJAC] + 1;
        }
        __syncthreads();

        // ---- phase QR: pivoted Householder, one logical column per step (lmfactor :619-666)
        for (int j = 0; j < N; ++j) {
            if (needjac && k == 0) {
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
                si[SI_PIVOT] = ipvt[j];
            }
            __syncthreads();
            {
                // ajnorm = norm2(a(j:m, pivot)), all threads of the system cooperating (bit-identical to the scan)
                const int pc = si[SI_PIVOT];
                double ajnorm = clm_coop_norm2<N>(needjac, J + pc * 32, (long long)N * 32, j, m, tq, w4h, k, k == 0);
                if (needjac && k == 0) {
                    if (ajnorm != 0.0 && J[((long long)j * N + pc) * 32] < 0.0) ajnorm = -ajnorm;
                    sc[SC_AJNORM] = ajnorm;
                }
            }
            __syncthreads();
            if (needjac && sc[SC_AJNORM] != 0.0) {
                // a(j:m, pivot) /= ajnorm is elementwise: rows are split over the n threads of the system
                const double ajnorm = sc[SC_AJNORM];
                double* col = J + si[SI_PIVOT] * 32;
                const long long rs = (long long)N * 32;
                int i = j + k;
                for (; i + 3 * N < m; i += 4 * N) {
                    double t[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) t[u] = col[(i + u * N) * rs];
#pragma unroll
                    for (int u = 0; u < 4; ++u) t[u] = t[u] / ajnorm;
                    if (i == j) { t[0] = t[0] + 1.0; sc[SC_AJJ] = t[0]; }      // a(j,j) = a(j,j)/ajnorm + 1
#pragma unroll
                    for (int u = 0; u < 4; ++u) { col[(i + u * N) * rs] = t[u]; tq[(long long)(i + u * N) * 32] = t[u]; }
                }
                for (; i < m; i += N) {
                    double t = col[i * rs] / ajnorm;
                    if (i == j) { t = t + 1.0; sc[SC_AJJ] = t; }
                    col[i * rs] = t;
                    tq[(long long)i * 32] = t;      // lane-contiguous copy of the reflector for the passes below
                }
            }
            __syncthreads();
            if (needjac) {
                const int pc = si[SI_PIVOT];
                const double ajnorm = sc[SC_AJNORM];
                const int mypos = pos[k];
                if (mypos > j && ajnorm != 0.0) {
                    // The reflector is read from its lane-contiguous copy in tq: every warp of the CTA then loads
                    // the same fully coalesced 256 B per row (L1 hits), instead of 32 scattered sectors of J.
                    const double* __restrict__ cp = tq;              // reflector v(i) = cp[i * 32]
                    double* __restrict__ ck = J + k * 32;            // own column
                    const long long rs = (long long)N * 32;
                    (void)pc;
                    double sm = 0.0;
                    {
                        int i = j;
                        for (; i + 8 <= m; i += 8) {
                            double a[8], c[8];
                            if (i + CLM_PF + 8 <= m) {
#pragma unroll
                                for (int u = 0; u < 8; ++u) clm_prefetch(&ck[(i + CLM_PF + u) * rs]);
                            }
#pragma unroll
                            for (int u = 0; u < 8; ++u) { a[u] = cp[(long long)(i + u) * 32]; c[u] = ck[(i + u) * rs]; }
#pragma unroll
                            for (int u = 0; u < 8; ++u) sm += a[u] * c[u];
                        }
                        for (; i < m; ++i) sm += cp[(long long)i * 32] * ck[i * rs];
                    }
                    double temp = sm / sc[SC_AJJ];
                    {
                        int i = j;
                        for (; i + 8 <= m; i += 8) {
                            double a[8], c[8];
                            if (i + CLM_PF + 8 <= m) {
#pragma unroll
                                for (int u = 0; u < 8; ++u) clm_prefetch(&ck[(i + CLM_PF + u) * rs]);
                            }
#pragma unroll
                            for (int u = 0; u < 8; ++u) { a[u] = cp[(long long)(i + u) * 32]; c[u] = ck[(i + u) * rs]; }
#pragma unroll
                            for (int u = 0; u < 8; ++u) ck[(i + u) * rs] = c[u] - temp * a[u];
                        }
                        for (; i < m; ++i) ck[i * rs] = ck[i * rs] - temp * cp[(long long)i * 32];
                    }
                    double rd = wa1[mypos];
                    if (rd != 0.0) {
                        temp = J[((long long)j * N + k) * 32] / rd;
                        rd = rd * sqrt(nl_max(0.0, 1.0 - temp * temp));
                        wa1[mypos] = rd;
                        const double q = rd / wa3[mypos];
                        if (!(0.05 * (q * q) > eps)) {
                            Norm2 acc;
                            clm_norm2_strided(acc, J + k * 32, (long long)N * 32, j + 1, m);
                            rd = acc.value();
                            wa1[mypos] = rd;
                            wa3[mypos] = rd;
                        }
                    }
                }
            }
            __syncthreads();
            if (needjac && k == 0) wa1[j] = -sc[SC_AJNORM];
        }
        __syncthreads();

        // ---- phase QTF: wa4 = fvec; qtf = first n of Q^T fvec; R = top block (lss_solve :241-253)
        if (needjac)
        {
            int i = k;
            for (; i + 3 * N < m; i += 4 * N) {
                double t[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) t[u] = fv[(long long)(i + u * N) * 32];
#pragma unroll
                for (int u = 0; u < 4; ++u) w4[(long long)(i + u * N) * 32] = t[u];
            }
            for (; i < m; i += N) w4[(long long)i * 32] = fv[(long long)i * 32];
        }
        __syncthreads();
        for (int j = 0; j < N; ++j) {
            if (needjac && k == 0) {
                const int pc = ipvt[j];
                const double ajj = J[((long long)j * N + pc) * 32];
                double temp = 0.0;
                if (ajj != 0.0) {
                    double sm = 0.0;
                    {
                        const double* col = J + pc * 32;
                        const long long rs = (long long)N * 32;
                        int i = j;
                        for (; i + 8 <= m; i += 8) {
                            double a[8], c[8];
                            if (i + CLM_PF + 8 <= m) {
#pragma unroll
                                for (int u = 0; u < 8; ++u) { clm_prefetch(&col[(i + CLM_PF + u) * rs]); clm_prefetch(&w4[(long long)(i + CLM_PF + u) * 32]); }
                            }
#pragma unroll
                            for (int u = 0; u < 8; ++u) { a[u] = col[(i + u) * rs]; c[u] = w4[(long long)(i + u) * 32]; }
#pragma unroll
                            for (int u = 0; u < 8; ++u) sm += a[u] * c[u];
                        }
                        for (; i < m; ++i) sm += col[i * rs] * w4[(long long)i * 32];
                    }
                    temp = -sm / ajj;
                }
                sc[SC_TEMP] = temp;
                sc[SC_AJJ] = ajj;
            }
            __syncthreads();
            if (needjac && sc[SC_AJJ] != 0.0) {
                const int pc = ipvt[j];
                const double temp = sc[SC_TEMP];
                const double* col = J + pc * 32;
                const long long rs = (long long)N * 32;
                int i = j + k;
                for (; i + 3 * N < m; i += 4 * N) {
                    double a[4], c[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) { a[u] = col[(i + u * N) * rs]; c[u] = w4[(long long)(i + u * N) * 32]; }
#pragma unroll
                    for (int u = 0; u < 4; ++u) w4[(long long)(i + u * N) * 32] = c[u] + a[u] * temp;
                }
                for (; i < m; i += N) w4[(long long)i * 32] = w4[(long long)i * 32] + col[i * rs] * temp;
            }
            __syncthreads();
            if (needjac && k == 0) qtf[j] = w4[(long long)j * 32];
        }
        if (needjac) {
            // logical column k of R: strict upper part from J, diagonal = rdiag(k)
            const int pc = ipvt[k];
            for (int i = 0; i < k; ++i) R(i, k) = J[((long long)i * N + pc) * 32];
            R(k, k) = wa1[k];
        }
        __syncthreads();

        // ---- phases G + P on warp 0: scaling, gradient test, LM parameter, trial point --------
        if (k == 0 && (state0 == CLM_NEED_JAC || state0 == CLM_INNER)) {
            int state = state0;
            const int iter = si[SI_ITER];
            const double fnorm = sc[SC_FNORM];
            if (state == CLM_NEED_JAC) {
                if (iter == 1) {
                    for (int j = 0; j < N; ++j) {
                        const double a = wa2[j];
                        diag[j] = (a == 0.0) ? 1.0 : a;
                    }
                    for (int j = 0; j < N; ++j) wa3[j] = diag[j] * x[j];
                    const double xnorm = clm_norm2<N>(wa3);
                    double delta = fac * xnorm;
                    if (delta == 0.0) delta = fac;
                    sc[SC_XNORM] = xnorm;
                    sc[SC_DELTA] = delta;
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
                sc[SC_GNORM] = gnorm;
                if (gnorm <= gtol) {
                    si[SI_GCN] = 1;
                    state = CLM_DONE;
                } else {
                    for (int j = 0; j < N; ++j) diag[j] = nl_max(diag[j], wa2[j]);
                    state = CLM_INNER;
                }
            }
            if (state == CLM_INNER) {
                double par = sc[SC_PAR];
                double delta = sc[SC_DELTA];
                clm_par<N>(R, ipvt, diag, qtf, delta, par, wa1, wa2, wa3, w4h, [&](double& scale, double& ssq) {
                    Norm2 acc;
                    acc.scale = scale; acc.ssq = ssq;
                    clm_norm2_strided(acc, w4, 32, N, m);
                    scale = acc.scale; ssq = acc.ssq;
                });
                for (int j = 0; j < N; ++j) {
                    const double pj = -wa1[j];
                    wa1[j] = pj;
                    wa2[j] = x[j] + pj;
                    wa3[j] = diag[j] * pj;
                }
                const double pnorm = clm_norm2<N>(wa3);
                if (iter == 1) delta = nl_min(delta, pnorm);
                sc[SC_PAR] = par;
                sc[SC_DELTA] = delta;
                sc[SC_PNORM] = pnorm;
            }
            si[SI_STATE] = state;
        }
        __syncthreads();

        // ---- phase E: wa4 = F(x + p), rows split over the n threads of the system -------------
        const bool inner = si[SI_STATE] == CLM_INNER;
        if (inner) {
            double xl[N];
#pragma unroll
            for (int j = 0; j < N; ++j) xl[j] = wa2[j];
            int i = k;
            for (; i + 3 * N < m; i += 4 * N) {
                double tt[4], yy[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) { tt[u] = __ldg(shared + i + u * N); yy[u] = ysys[(long long)(i + u * N) * B]; }
#pragma unroll
                for (int u = 0; u < 4; ++u) w4[(long long)(i + u * N) * 32] = F::residual(xl, tt[u], yy[u]);
            }
            for (; i < m; i += N) w4[(long long)i * 32] = F::residual(xl, __ldg(shared + i), ysys[(long long)i * B]);
        }
        __syncthreads();

        {
            const double f1n = clm_coop_norm2<N>(inner, w4, 32, 0, m, tq, w4h, k, k == 0);
            if (inner && k == 0) sc[SC_TEMP] = f1n;
        }
        // ---- phase A on warp 0: gain ratio, step bound, acceptance, convergence (lss_solve :297-365)
        if (k == 0 && inner) {
            int iter = si[SI_ITER];
            const int neval = si[SI_NEVAL] + 1;
            si[SI_NEVAL] = neval;
            double fnorm = sc[SC_FNORM], par = sc[SC_PAR], delta = sc[SC_DELTA], xnorm = sc[SC_XNORM];
            const double pnorm = sc[SC_PNORM], gnorm = sc[SC_GNORM];
            const double fnorm1 = sc[SC_TEMP];
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
            si[SI_ACCEPT] = accept;
            bool fcnvrg = false, xcnvrg = false;
            if (fabs(actred) <= ftol && prered <= ftol && 0.5 * ratio <= 1.0) fcnvrg = true;
            if (delta <= xtol * xnorm) xcnvrg = true;
            int flag = 0;
            int state = CLM_INNER;
            if (fcnvrg || xcnvrg) {
                state = CLM_DONE;
            } else {
                if (neval >= p.max_fcn_evals) flag = NLB_CONVERGENCE_ERROR;
                if (fabs(actred) <= eps && prered <= eps && 0.5 * ratio <= 1.0) flag = NLB_TOLERANCE_TOO_SMALL_ERROR;
                if (delta <= eps * xnorm) flag = NLB_TOLERANCE_TOO_SMALL_ERROR;
                if (gnorm <= eps) flag = NLB_TOLERANCE_TOO_SMALL_ERROR;
                if (flag != 0) state = CLM_DONE;
                else if (accept) state = CLM_NEED_JAC;
            }
            si[SI_FCN] = fcnvrg; si[SI_XCN] = xcnvrg; si[SI_FLAG] = flag;
            si[SI_ITER] = iter; si[SI_STATE] = state;
            sc[SC_FNORM] = fnorm; sc[SC_PAR] = par; sc[SC_DELTA] = delta; sc[SC_XNORM] = xnorm;
        }
        __syncthreads();
        // ---- phase C: accepted step -> fvec = wa4 ---------------------------------------------
        if (inner && si[SI_ACCEPT]) {
            int i = k;
            for (; i + 3 * N < m; i += 4 * N) {
                double t[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) t[u] = w4[(long long)(i + u * N) * 32];
#pragma unroll
                for (int u = 0; u < 4; ++u) fv[(long long)(i + u * N) * 32] = t[u];
            }
            for (; i < m; i += N) fv[(long long)i * 32] = w4[(long long)i * 32];
        }
        __syncthreads();
        if (k == 0) si[SI_ACCEPT] = 0;
    }

}

int launch_coop_lm(int fcn_id, const DevParams& p, long long nsys, long long B, int m, int n, double* x, double* fvec, const double* sys,
                   const double* shared, nlb_iteration_behavior* ib, int32_t* status, cudaStream_t s, int64_t* launches);

}  // namespace nlb
