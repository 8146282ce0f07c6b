// quad_lm.cuh — Levenberg-Marquardt for small curve fits (BASELINE config 1: m = 21, n = 4) with FOUR LANES PER
// SYSTEM and the whole m x n Jacobian in registers.
//
// Behaviour reproduced: lss_solve / lmpar / lmfactor / lmsolve, reference
// src/nonlin_least_squares.f90:118-391 / 394-566 / 569-667 / 670-791, and the forward-difference
// Jacobian vfh_jac_fcn, src/nonlin_multi_eqn_mult_var.f90:198-277.
//
// Why.  The thread-per-system kernel (tps_lm.cuh) keeps the 21 x 4 Jacobian and two 21-vectors in 1008 B of
// per-thread local memory: FP64 pipe 16 % busy, 3x the algorithmic DRAM traffic (profiles/r1_final_ncu.md).
// Here a quad of lanes owns one system and each lane owns a contiguous BLOCK OF ROWS (6, 5, 5, 5 of the 21) of
// every column, of fvec and of wa4: 36 doubles per lane, all in registers, and every elementwise pass (residual
// rows, forward differences, reflector scaling and application: all the divisions) runs on the four lanes at once.
//
// Parity.  The reference's sums over the m rows run in index order, so they cannot be split over the lanes.
// Instead the lanes form the summands in parallel (products for a dot, quotients for NORM2), park them in a
// shared-memory tile, and the ordered additions are then done per CHAIN: at a Householder step the (up to four)
// dot products - trailing columns plus the right-hand side - are independent chains, lane c adds chain c.  Each
// product and each addition is the one the reference performs, in its order, without FMA: bit-identical.
// NORM2 (libgfortran's scaled recurrence) uses the fact that its running scale is the prefix maximum of |x|:
// a block-local scan plus a 3-step exchange gives every element the scale it meets, the lanes divide in
// parallel, and the ssq recurrence is replayed in order.
// The right-hand side rides along as a fifth column: lss_solve's "Q^T fvec" pass (:241-253) applies the same
// reflectors to wa4 = fvec, with temp = -sum/ajj and wa4 + a*temp; since negation commutes with IEEE
// multiplication, division and addition, wa4 - (sum/ajj)*a yields the same bits, so it is done inside lmfactor's
// column loop.  Rows n+1..m of the factored Jacobian are never read again by the reference.
//
// The n-sized serial parts (gradient test, lmpar, lmsolve, gain ratio) run on lane 0 of the quad on
// shared-memory vectors (the clm_* routines of coop_lm.cuh).
#pragma once
#include "coop_lm.cuh"

namespace nlb {

#ifndef NLB_QLM_MIN_CTAS
#define NLB_QLM_MIN_CTAS 3
#endif
constexpr int QLM_BLOCK = 128;              // 32 systems per CTA
constexpr int QLM_QUADS = QLM_BLOCK / 4;

template <int M>
struct QRows {
    static constexpr int RQ = (M + 3) / 4, BASE = M / 4, REM = M % 4;
    NLB_DEV static constexpr int start(int l) { return l * BASE + (l < REM ? l : REM); }
    NLB_DEV static constexpr int count(int l) { return BASE + (l < REM ? 1 : 0); }
    // owner lane of row i
    NLB_DEV static constexpr int owner(int i) {
        return (i < start(1)) ? 0 : (i < start(2)) ? 1 : (i < start(3)) ? 2 : 3;
    }
};

template <class F>
struct QlmSmem {
    static constexpr int N = F::N, M = F::M;
    static constexpr int NDV = 7 * N + N * N + 12;                 // n-sized state of a quad (doubles)
    static constexpr int NIV = N + 4;                              // ipvt + control words
    static constexpr int TILE = M * 32;                            // summand tile of one warp (doubles)
    static constexpr size_t BYTES = sizeof(double) * ((size_t)QLM_QUADS * NDV + (size_t)(QLM_BLOCK / 32) * TILE + M) +
                                    sizeof(int) * (size_t)QLM_QUADS * NIV;
};

// Tile address of summand (row i, chain c) of quad q; rows owned by lane `blk` are rotated by 4*blk inside their
// 32-double line so that the four lanes of a quad (which write different rows of the same chain) hit different banks,
// while a warp reading one row of four chains still reads 32 consecutive doubles.
NLB_DEV int qlm_addr(int i, int c, int q, int blk) { return i * 32 + ((c * 8 + q + 4 * blk) & 31); }

// NORM2 of a vector whose rows are spread over the four lanes of a quad (v0..v5 = this lane's rows, rows lo..M-1 take
// part), same bits as libgfortran's one-pass recurrence; formed redundantly by the four lanes.  Out of line: a dozen
// inlined copies of the 21-step replay made the first version of this kernel instruction-fetch bound.
template <int M>
__device__ __noinline__ double qlm_norm(double v0, double v1, double v2, double v3, double v4, double v5, int lo, int l4,
                                        int q, unsigned qmask, double* __restrict__ tile) {
    using Q = QRows<M>;
    static_assert(Q::RQ <= 6, "rows per lane");
    const double v[6] = {v0, v1, v2, v3, v4, v5};
    const int start = Q::start(l4), cnt = Q::count(l4);
    double am[6], pm[6];
    double run = 0.0;
#pragma unroll
    for (int r = 0; r < Q::RQ; ++r) {
        const bool on = r < cnt && (l4 > 0 || r >= lo);
        const double a = on ? fabs(v[r]) : 0.0;
        am[r] = a;
        pm[r] = run;                                           // exclusive prefix maximum inside the block
        if (a > run) run = a;                                  // NaN never raises the scale
    }
    double before = 1.0;                                       // maxima of the lanes before this one
#pragma unroll
    for (int d = 1; d < 4; ++d) {
        const double t = __shfl_up_sync(qmask, run, d, 4);
        if (l4 >= d && t > before) before = t;
    }
#pragma unroll
    for (int r = 0; r < Q::RQ; ++r) {
        const bool on = r < cnt && (l4 > 0 || r >= lo);
        if (on) {
            const double sc = (pm[r] > before) ? pm[r] : before;   // running scale this element meets
            const double xv = v[r];
            double t = 0.0;
            if (xv != 0.0) {
                const double a = am[r];
                const bool up = sc < a;
                const double qt = (up ? sc : a) / (up ? a : sc);
                t = up ? -fmax(qt, 4.9406564584124654e-324) : qt * qt;
            }
            tile[qlm_addr(start + r, 0, q, l4)] = t;
        }
    }
    double total = (run > before) ? run : before;
    total = __shfl_sync(qmask, total, (threadIdx.x & 28) + 3);     // lane 3 of the quad has seen every block
    __syncwarp(qmask);
    double ssq = 0.0;
#pragma unroll 1
    for (int blk = 0; blk < 4; ++blk) {
        const int i0 = Q::start(blk), i1 = i0 + Q::count(blk);
        const double* tp = tile + ((q + 4 * blk) & 31);
#pragma unroll 1
        for (int i = (i0 > lo ? i0 : lo); i < i1; ++i) {
            const double t = tp[i * 32];
            if (t < 0.0) { const double tt = -t; ssq = 1.0 + ssq * tt * tt; }
            else ssq = ssq + t;
        }
    }
    __syncwarp(qmask);
    return total * sqrt(ssq);
}

// ordered sum of chain c over rows lo..M-1 of the tile (one lane per chain)
template <int M>
__device__ __noinline__ double qlm_chain_sum(const double* __restrict__ tile, int c, int q, int lo) {
    using Q = QRows<M>;
    double sm = 0.0;
#pragma unroll 1
    for (int blk = 0; blk < 4; ++blk) {
        const int i0 = Q::start(blk), i1 = i0 + Q::count(blk);
        const double* tp = tile + ((c * 8 + q + 4 * blk) & 31);
#pragma unroll 1
        for (int i = (i0 > lo ? i0 : lo); i < i1; ++i) sm += tp[i * 32];
    }
    return sm;
}

// libgfortran NORM2 of chain c over all M rows of the tile (raw values; one lane per chain)
template <int M>
__device__ __noinline__ double qlm_chain_norm(const double* __restrict__ tile, int c, int q) {
    using Q = QRows<M>;
    Norm2 acc;
#pragma unroll 1
    for (int blk = 0; blk < 4; ++blk) {
        const int i0 = Q::start(blk), i1 = i0 + Q::count(blk);
        const double* tp = tile + ((c * 8 + q + 4 * blk) & 31);
#pragma unroll 1
        for (int i = i0; i < i1; ++i) acc.add(tp[i * 32]);
    }
    return acc.value();
}

// lmpar + trial step on lane 0 of a quad (the n-sized state lives in shared memory); out of line for the same reason
template <int N, int M, int QS>
__device__ __noinline__ void qlm_par_step(double* sd, int* ipv, const double* __restrict__ tile, int q, double delta,
                                          double* par_io) {
    using Q = QRows<M>;
    using V = SVec<QS>;
    const V x{sd}, diag{sd + QS * N}, qtf{sd + QS * 2 * N}, wa1{sd + QS * 3 * N}, wa2{sd + QS * 4 * N},
        wa3{sd + QS * 5 * N}, w4h{sd + QS * 6 * N};
    const SMat<N, QS> R{sd + QS * 7 * N};
    const SIVec<QS> ipvt_s{ipv};
    double par = *par_io;
    clm_par<N>(R, ipvt_s, diag, qtf, delta, par, wa1, wa2, wa3, w4h, [&](double& scale, double& ssq) {
        Norm2 acc;
        acc.scale = scale; acc.ssq = ssq;
#pragma unroll 1
        for (int blk = 0; blk < 4; ++blk) {
            const int i0 = Q::start(blk), i1 = i0 + Q::count(blk);
            const double* tp = tile + ((q + 4 * blk) & 31);
#pragma unroll 1
            for (int i = (i0 > N ? i0 : N); i < i1; ++i) acc.add(tp[i * 32]);
        }
        scale = acc.scale; ssq = acc.ssq;
    });
#pragma unroll 1
    for (int j = 0; j < N; ++j) {
        const double pj = -wa1[j];
        wa1[j] = pj;
        wa2[j] = x[j] + pj;
        wa3[j] = diag[j] * pj;
    }
    *par_io = par;
}

template <class F>
__global__ void __launch_bounds__(QLM_BLOCK, NLB_QLM_MIN_CTAS)
qlm_kernel(DevParams p, long long nsys, long long B, double* __restrict__ xg, double* __restrict__ fg,
           const double* __restrict__ sys, const double* __restrict__ shared, nlb_iteration_behavior* __restrict__ ibg,
           int32_t* __restrict__ statusg) {
    constexpr int M = F::M, N = F::N;
    static_assert(N == 4, "four lanes per system: n = 4");
    using Q = QRows<M>;
    constexpr int RQ = Q::RQ;
    static_assert(RQ <= 6 && Q::count(0) >= N, "the top n x n block must sit in lane 0");
    using S = QlmSmem<F>;
    extern __shared__ double smem[];

    const int tid = threadIdx.x, lane = tid & 31, l4 = lane & 3, q = lane >> 2, warp = tid >> 5;
    const int quad = tid >> 2;                                     // quad index inside the CTA
    const unsigned qmask = 0xFu << (lane & ~3);
    const int lane0 = lane & ~3;                                   // lane 0 of this quad
    const int start = Q::start(l4), cnt = Q::count(l4);

    // shared memory: n-sized state [element][quad], per-warp summand tiles, abscissae, ints [element][quad]
    double* sd = smem + quad;
    using V = SVec<QLM_QUADS>;
    using Mt = SMat<N, QLM_QUADS>;
    using IV = SIVec<QLM_QUADS>;
    const V x{sd}, diag{sd + QLM_QUADS * N}, qtf{sd + QLM_QUADS * 2 * N}, wa1{sd + QLM_QUADS * 3 * N},
        wa2{sd + QLM_QUADS * 4 * N}, wa3{sd + QLM_QUADS * 5 * N};
    const Mt R{sd + QLM_QUADS * 7 * N};
    double* tile = smem + QLM_QUADS * S::NDV + warp * S::TILE;
    double* absc = smem + QLM_QUADS * S::NDV + (QLM_BLOCK / 32) * S::TILE;
    int* si_base = reinterpret_cast<int*>(absc + M) + quad;
    const IV ipvt_s{si_base};

    for (int i = tid; i < M; i += QLM_BLOCK) absc[i] = F::abscissa(i, shared);
    __syncthreads();

    const long long b = (long long)blockIdx.x * QLM_QUADS + quad;
    if (b >= nsys) return;                                         // whole quads leave together
    const double* ysys = sys + b;

    const double eps = 0x1p-52;
    const double ftol = p.fcn_tol, xtol = p.var_tol, gtol = p.grad_tol, fac = p.lm_factor;
    const int maxeval = p.max_fcn_evals;

    // rows of this lane
    auto row_valid = [&](int r) { return r < cnt; };
    // row j (j < N: it sits in lane 0) of a lane-0 row block, j a run-time index
    auto pick = [&](const double (&v)[RQ], int j) {
        double t = v[0];
#pragma unroll
        for (int r = 1; r < N; ++r) t = (j == r) ? v[r] : t;
        return t;
    };
    auto qnorm = [&](const double (&v)[RQ], int lo) {
        return qlm_norm<M>(v[0], v[1], v[2], v[3], v[4], RQ > 5 ? v[RQ > 5 ? 5 : 0] : 0.0, lo, l4, q, qmask, tile);
    };

    // ---- state -----------------------------------------------------------------------------------------------
    // Column QUEUE: at Householder step j the registers a[s] hold the column at pivot position j + s (s = 0: the pivot
    // after the swap); finished columns leave the queue, so one loop body serves every step.
    double a[N][RQ];
    double fv[RQ], w4[RQ];
    double rd[N], wv[N];                                           // rdiag, wa of the queued columns
    int ip[N];                                                     // their original indices
    double xl[N];
#pragma unroll
    for (int j = 0; j < N; ++j) xl[j] = xg[(long long)j * B + b];
    if (l4 == 0) {
#pragma unroll
        for (int j = 0; j < N; ++j) x[j] = xl[j];
    }
#pragma unroll
    for (int r = 0; r < RQ; ++r) {
        fv[r] = 0.0;
        if (row_valid(r)) fv[r] = F::row(xl, absc[start + r], __ldg(ysys + (long long)(start + r) * B));
    }
    double fnorm = qnorm(fv, 0);
    int iter = 1, neval = 1, njac = 0, flag = 0;
    bool fcnvrg = false, xcnvrg = false, gcnvrg = false;
    double par = 0.0, xnorm = 0.0, delta = 0.0, gnorm = 0.0;      // meaningful in lane 0 (replicated arithmetic)

    for (;;) {   // ---- outer iteration -------------------------------------------------------------------------
        // forward-difference columns (vfh_jac_fcn :262-275): every lane its rows of every column
#pragma unroll 1
        for (int j = 0; j < N; ++j) {
            double temp = xl[0];
#pragma unroll
            for (int c = 1; c < N; ++c) temp = (c == j) ? xl[c] : temp;
            double h = 0x1p-26 * fabs(temp);
            if (h == 0.0) h = 0x1p-26;
            double xp[N];
#pragma unroll
            for (int c = 0; c < N; ++c) xp[c] = (c == j) ? (temp + h) : xl[c];
#pragma unroll
            for (int r = 0; r < RQ; ++r) {
                if (row_valid(r)) {
                    const int i = start + r;
                    const double f1 = F::row(xp, absc[i], __ldg(ysys + (long long)i * B));
                    const double val = (f1 - fv[r]) / h;
#pragma unroll
                    for (int c = 0; c < N; ++c) a[c][r] = (c == j) ? val : a[c][r];
                }
            }
        }
        ++njac;
        // column norms (lmfactor :611-616): four independent recurrences, lane c walks column c
#pragma unroll
        for (int c = 0; c < N; ++c)
#pragma unroll
            for (int r = 0; r < RQ; ++r)
                if (row_valid(r)) tile[qlm_addr(start + r, c, q, l4)] = a[c][r];
        __syncwarp(qmask);
        {
            const double cn = qlm_chain_norm<M>(tile, l4, q);
#pragma unroll
            for (int c = 0; c < N; ++c) {
                rd[c] = __shfl_sync(qmask, cn, lane0 + c);
                wv[c] = rd[c];
                ip[c] = c;
            }
            if (l4 == 0) {
#pragma unroll
                for (int c = 0; c < N; ++c) wa2[c] = rd[c];         // acnorm, by original column
            }
        }
        __syncwarp(qmask);
#pragma unroll
        for (int r = 0; r < RQ; ++r) w4[r] = fv[r];

        // pivoted Householder QR with the right-hand side riding along (lmfactor :619-666, lss_solve :241-253)
#pragma unroll 1
        for (int j = 0; j < N; ++j) {
            const int nq = N - j;                                  // columns in the queue
            int ks = 0;
            double rmax = rd[0];
#pragma unroll
            for (int s = 1; s < N; ++s)
                if (s < nq && rd[s] > rmax) { rmax = rd[s]; ks = s; }
#pragma unroll
            for (int s = 1; s < N; ++s) {
                if (ks == s) {
#pragma unroll
                    for (int r = 0; r < RQ; ++r) { const double t = a[0][r]; a[0][r] = a[s][r]; a[s][r] = t; }
                    { const double t = rd[0]; rd[0] = rd[s]; rd[s] = t; }
                    { const double t = wv[0]; wv[0] = wv[s]; wv[s] = t; }
                    { const int t = ip[0]; ip[0] = ip[s]; ip[s] = t; }
                }
            }
            if (l4 == 0) {
                // rows above the diagonal of the pivot column are final: column j of R
#pragma unroll
                for (int i = 0; i < N; ++i)
                    if (i < j) R(i, j) = a[0][i];
                ipvt_s[j] = ip[0];
            }
            double ajnorm = qnorm(a[0], j);
            if (ajnorm != 0.0) {
                const double ajj0 = __shfl_sync(qmask, pick(a[0], j), lane0);
                if (ajj0 < 0.0) ajnorm = -ajnorm;
                double v[RQ];
#pragma unroll
                for (int r = 0; r < RQ; ++r) {
                    const bool on = row_valid(r) && (l4 > 0 || r >= j);
                    v[r] = on ? a[0][r] / ajnorm : 0.0;
                    if (l4 == 0 && r == j) v[r] = v[r] + 1.0;
                }
                const double ajj = __shfl_sync(qmask, pick(v, j), lane0);
                // summands of the dot products: chain s = queued column s (s >= 1), chain 0 = the right-hand side
#pragma unroll
                for (int r = 0; r < RQ; ++r) {
                    const bool on = row_valid(r) && (l4 > 0 || r >= j);
                    if (on) {
                        const int i = start + r;
                        tile[qlm_addr(i, 0, q, l4)] = v[r] * w4[r];
#pragma unroll
                        for (int s = 1; s < N; ++s)
                            if (s < nq) tile[qlm_addr(i, s, q, l4)] = v[r] * a[s][r];
                    }
                }
                __syncwarp(qmask);
                double sm = 0.0;
                if (l4 < nq) sm = qlm_chain_sum<M>(tile, l4, q, j);
                const double tmine = sm / ajj;
                __syncwarp(qmask);
                double tc[N];
#pragma unroll
                for (int s = 0; s < N; ++s) tc[s] = __shfl_sync(qmask, tmine, lane0 + s);
#pragma unroll
                for (int r = 0; r < RQ; ++r) {
                    const bool on = row_valid(r) && (l4 > 0 || r >= j);
                    if (on) {
                        w4[r] = w4[r] - tc[0] * v[r];
#pragma unroll
                        for (int s = 1; s < N; ++s)
                            if (s < nq) a[s][r] = a[s][r] - tc[s] * v[r];
                    }
                }
                // norm down-dates (lmfactor :656-661), same arithmetic in the four lanes
#pragma unroll
                for (int s = 1; s < N; ++s) {
                    if (s >= nq || rd[s] == 0.0) continue;
                    const double ajc = __shfl_sync(qmask, pick(a[s], j), lane0);
                    const double temp = ajc / rd[s];
                    rd[s] = rd[s] * sqrt(nl_max(0.0, 1.0 - temp * temp));
                    const double qq = rd[s] / wv[s];
                    if (0.05 * (qq * qq) > eps) continue;
                    rd[s] = qnorm(a[s], j + 1);
                    wv[s] = rd[s];
                }
            }
            if (l4 == 0) {
                R(j, j) = -ajnorm;
                qtf[j] = pick(w4, j);
            }
            // the pivot leaves the queue
#pragma unroll
            for (int s = 0; s + 1 < N; ++s) {
#pragma unroll
                for (int r = 0; r < RQ; ++r) a[s][r] = a[s + 1][r];
                rd[s] = rd[s + 1]; wv[s] = wv[s + 1]; ip[s] = ip[s + 1];
            }
        }
        int next = 0;                                              // 0 = inner, 1 = done
        if (l4 == 0) {
            if (iter == 1) {
#pragma unroll 1
                for (int j = 0; j < N; ++j) {
                    const double c = wa2[j];
                    diag[j] = (c == 0.0) ? 1.0 : c;
                }
#pragma unroll 1
                for (int j = 0; j < N; ++j) wa3[j] = diag[j] * x[j];
                xnorm = clm_norm2<N>(wa3);
                delta = fac * xnorm;
                if (delta == 0.0) delta = fac;
            }
            gnorm = 0.0;
            if (fnorm != 0.0) {
#pragma unroll 1
                for (int j = 0; j < N; ++j) {
                    const double acn = wa2[ipvt_s[j]];
                    if (acn == 0.0) continue;
                    double sm = 0.0;
#pragma unroll 1
                    for (int i = 0; i <= j; ++i) sm += R(i, j) * (qtf[i] / fnorm);
                    gnorm = nl_max(gnorm, fabs(sm / acn));
                }
            }
            if (gnorm <= gtol) { gcnvrg = true; next = 1; }
            else {
#pragma unroll 1
                for (int j = 0; j < N; ++j) diag[j] = nl_max(diag[j], wa2[j]);
            }
        }
        next = __shfl_sync(qmask, next, lane0);
        if (next == 1) break;

        for (;;) {   // ---- inner iteration ---------------------------------------------------------------------
            // rows n..m-1 of wa4 (Q^T f tail, or the last trial residual) for lmpar's in-loop norm (:531)
#pragma unroll
            for (int r = 0; r < RQ; ++r)
                if (row_valid(r)) tile[qlm_addr(start + r, 0, q, l4)] = w4[r];
            __syncwarp(qmask);
            if (l4 == 0) qlm_par_step<N, M, QLM_QUADS>(sd, si_base, tile, q, delta, &par);
            __syncwarp(qmask);
            double xt[N];
#pragma unroll
            for (int j = 0; j < N; ++j) xt[j] = wa2[j];
#pragma unroll
            for (int r = 0; r < RQ; ++r)
                if (row_valid(r)) w4[r] = F::row(xt, absc[start + r], __ldg(ysys + (long long)(start + r) * B));
            ++neval;
            const double fnorm1 = qnorm(w4, 0);
            int code = 0;                                          // bit 0: accept, bits 1-2: 0 inner / 1 outer / 2 done
            if (l4 == 0) {
                const double pnorm = clm_norm2<N>(wa3);
                if (iter == 1) delta = nl_min(delta, pnorm);
                double actred = -1.0;
                if (0.1 * fnorm1 < fnorm) { const double qq = fnorm1 / fnorm; actred = 1.0 - qq * qq; }
                double temp = 0.0;
#pragma unroll 1
                for (int j = 0; j < N; ++j) {
                    wa3[j] = 0.0;
                    temp = wa1[ipvt_s[j]];
#pragma unroll 1
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
#pragma unroll 1
                    for (int j = 0; j < N; ++j) {
                        const double xn = wa2[j];
                        x[j] = xn;
                        wa2[j] = diag[j] * xn;
                    }
                    xnorm = clm_norm2<N>(wa2);
                    ++iter;
                }
                if (fabs(actred) <= ftol && prered <= ftol && 0.5 * ratio <= 1.0) fcnvrg = true;
                if (delta <= xtol * xnorm) xcnvrg = true;
                int nx = 0;
                if (fcnvrg || xcnvrg) {
                    nx = 2;
                } else {
                    if (neval >= maxeval) flag = NLB_CONVERGENCE_ERROR;
                    if (fabs(actred) <= eps && prered <= eps && 0.5 * ratio <= 1.0) flag = NLB_TOLERANCE_TOO_SMALL_ERROR;
                    if (delta <= eps * xnorm) flag = NLB_TOLERANCE_TOO_SMALL_ERROR;
                    if (gnorm <= eps) flag = NLB_TOLERANCE_TOO_SMALL_ERROR;
                    if (flag != 0) nx = 2;
                    else if (accept) nx = 1;
                }
                code = (accept ? 1 : 0) | (nx << 1);
            }
            code = __shfl_sync(qmask, code, lane0);
            if (code & 1) {
#pragma unroll
                for (int j = 0; j < N; ++j) xl[j] = xt[j];
#pragma unroll
                for (int r = 0; r < RQ; ++r) fv[r] = w4[r];
                fnorm = fnorm1;
            }
            next = code >> 1;
            if (next != 0) break;
        }
        if (next == 2) break;
    }

    // ---- results ---------------------------------------------------------------------------------------------
#pragma unroll
    for (int r = 0; r < RQ; ++r)
        if (row_valid(r)) fg[(long long)(start + r) * B + b] = fv[r];
    if (l4 == 0) {
#pragma unroll
        for (int j = 0; j < N; ++j) xg[(long long)j * B + b] = xl[j];
        if (ibg) {
            nlb_iteration_behavior o;
            o.iter_count = iter; o.fcn_count = neval; o.jacobian_count = njac; o.gradient_count = 0;
            o.converge_on_fcn = fcnvrg; o.converge_on_chng = xcnvrg; o.converge_on_zero_diff = gcnvrg;
            ibg[b] = o;
        }
        if (statusg) statusg[b] = flag != 0 ? NLB_CONVERGENCE_ERROR : NLB_NO_ERROR;
    }
}

}  // namespace nlb
