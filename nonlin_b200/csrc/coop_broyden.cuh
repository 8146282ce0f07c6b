// coop_broyden.cuh — CTA-per-system Broyden quasi-Newton for mid-size square systems
// (BASELINE config 5: extended Rosenbrock, n = 64, quasi_newton_solver + line search).
//
// Behaviour reproduced: qns_solve, reference src/nonlin_solve.f90:156-425, with its linalg
// calls (qr_factor, rank1_update, qr_rank1_update, mtx_mult, solve_triangular_system,
// recip_mult_array) as the Reference-LAPACK / QRUPDATE algorithms listed in tps_dense.cuh.
//
// Mapping.  One CTA of N threads owns one system for its whole solve.  Q (N x N, column-major, leading
// dimension N+1 so that both row-wise and column-wise sweeps are bank-conflict free), R (packed by columns
// with room for the sub-diagonal the rank-1 update creates: column j holds rows 0..j+1) and every vector
// live in shared memory; the Broyden matrix B lives in REGISTERS, thread t holding row t (B is only ever
// used row-wise - B*dx, the rank-1 update, the forward differences - except for B^T f, which goes through
// a 4-row staging buffer that aliases the update's work vectors).  54 KB of shared memory at N = 64: four
// CTAs per SM.  HBM is touched once to read x0 and once to write x, fvec, ib, status.
//
// The strictly sequential recurrences (the Givens chain that restores the triangle, the back substitution)
// run inside one warp at a time with the pivot scalars passed by shuffle instead of through shared memory
// and a CTA barrier per step; the other warp catches up from shared memory once per 32 steps.
//
// Parity.  Every sum keeps the index order of the reference's loop, so results are bit-identical
// to the CPU oracle.  That is affordable because the O(n^2) kernels parallelise over the
// *other* index: thread t owns output t (a row of B*dx, a column of B^T f / Q^T u / the
// Householder update, a row of the Q rotations, a column of the R sweeps) and walks its sum
// sequentially.  The O(n) reductions (dot products, norms, max-norm tests, the Givens chain of
// DQRTV1) are evaluated redundantly by every thread from the same shared-memory operands,
// which keeps all scalars and the control flow uniform without broadcasts.
#pragma once
#include "tps_common.cuh"
#include "nlb_div.cuh"

namespace nlb {

#ifdef NLB_CB_TRACE
// debug build only: clock stamps of thread 0 of CTA 0 at the phase boundaries of every iteration (scripts/cb_trace.py)
__device__ long long cb_trace_buf[64][32];
#define CB_TRACE(probe) do { if (blockIdx.x == 0 && tid == 0 && iter < 64) cb_trace_buf[iter][probe] = clock64(); } while (0)
#define CB_TRACE2(probe) do { if (blockIdx.x == 0 && tid == 0 && cb_it < 64) cb_trace_buf[cb_it][16 + (probe)] = clock64(); } while (0)
#else
#define CB_TRACE(probe) do { } while (0)
#define CB_TRACE2(probe) do { } while (0)
#endif

template <int N>
struct CoopBroydenSmem {
    static constexpr int LD = N + 1;
    static constexpr int MAT = N * LD;
    static constexpr int RPK = N * (N + 3) / 2;   // packed upper-Hessenberg storage of R
    static constexpr int NVEC = 10;
    static constexpr int STAGE = 8;     // rows of B published at a time for B^T f (aliases xold ... sn, all dead there)
    static constexpr int SLD = N + 1;   // their stride (conflict-free for the four owners writing side by side)
    static constexpr size_t BYTES = ((size_t)MAT + RPK + NVEC * (size_t)N + 8) * sizeof(double);
};

// offset of column j of the packed R (rows 0..j+1)
NLB_DEV constexpr int cb_ro(int j) { return j * (j + 3) / 2; }
template <int N>
NLB_DEV constexpr unsigned cb_mask() { return N >= 32 ? 0xffffffffu : ((1u << N) - 1u); }

// ---- redundant sequential reductions (identical in every thread) -----------------------
template <int N>
NLB_DEV double cb_dot(const double* a, const double* b) {
    double s = 0.0;
#pragma unroll 8
    for (int i = 0; i < N; ++i) s += a[i] * b[i];
    return s;
}
template <int N>
NLB_DEV double cb_norm2(const double* a) {
    Norm2 acc;
    for (int i = 0; i < N; ++i) acc.add(a[i]);
    return acc.value();
}
// NORM2 (libgfortran's one-pass recurrence) with the N divisions spread over the threads: thread i forms the quotient
// element i meets - its running scale is the largest of 1 and the magnitudes before it - and every thread then replays
// the ssq recurrence over the N quotients in order.  A quotient is stored as t*t for an ordinary element and as -t for
// one that raises the scale (same convention as tall_lm.cuh).  wk: N doubles of scratch.  Contains CTA barriers.
template <int N>
NLB_DEV double cb_norm2_par(const double* a, double* wk, int tid) {
    constexpr int NW = (N + 31) / 32;
    __shared__ double wtot[NW + 1];
    const unsigned mask = cb_mask<N>();
    const int lane = tid & 31;
    const double x = a[tid];
    const double ax = fabs(x);
    double inc = (ax == ax) ? ax : 0.0;               // a NaN never becomes the scale (scale < NaN is false)
#pragma unroll
    for (int d = 1; d < 32 && d < N; d <<= 1) {
        const double o = __shfl_up_sync(mask, inc, d);
        if (lane >= d) inc = (o > inc) ? o : inc;
    }
    double before = __shfl_up_sync(mask, inc, 1);
    if (lane == 0) before = 0.0;
    if constexpr (N > 32) {
        if (lane == 31) wtot[tid >> 5] = inc;
        __syncthreads();
        double carry = 0.0;
        for (int w = 0; w < (tid >> 5); ++w) carry = (wtot[w] > carry) ? wtot[w] : carry;
        before = (carry > before) ? carry : before;
        inc = (carry > inc) ? carry : inc;
    }
    const double sc = (before > 1.0) ? before : 1.0;  // the recurrence starts from scale = 1
    double qv = 0.0;
    if (x != 0.0) {
        const bool up = sc < ax;
        const double t = (up ? sc : ax) / (up ? ax : sc);
        qv = up ? -fmax(t, 4.9406564584124654e-324) : t * t;
    }
    wk[tid] = qv;
    if (tid == N - 1) wtot[NW] = (inc > 1.0) ? inc : 1.0;  // final scale
    __syncthreads();
    double ssq = 0.0;
#pragma unroll 8
    for (int i = 0; i < N; ++i) {
        const double v = wk[i];
        const double tt = -v;
        const double up = 1.0 + ssq * tt * tt;
        const double ord = ssq + v;
        ssq = (v < 0.0) ? up : ord;
    }
    const double scale = wtot[NW];
    __syncthreads();                                   // wk and wtot may be reused
    return scale * sqrt(ssq);
}

template <int N>
NLB_DEV double cb_maxabs(const double* a) {
    double t = 0.0;
    for (int i = 0; i < N; ++i) t = nl_max(fabs(a[i]), t);
    return t;
}

// DLARTG for the re-triangularisation chain: the ordinary branch forms its two quotients (c = |f| / d, s = g / r) with the
// branch-free division, so that they overlap instead of running one after the other; same values as dlartg().
NLB_DEV void cb_dlartg(double f, double g, double& c, double& s, double& r) {
    const double rtmin = 0x1p-511, rtmax = 0x1.6a09e667f3bcdp+510;
    const double f1 = fabs(f), g1 = fabs(g);
    if (g != 0.0 && f != 0.0 && f1 > rtmin && f1 < rtmax && g1 > rtmin && g1 < rtmax) {
        const double d = sqrt(f * f + g * g);
        r = nl_sign(d, f);
        bool ok1, ok2;
        const double cq = nl_div_try(f1, d, ok1);
        const double sq = nl_div_try(g, r, ok2);
        if (ok1 && ok2) { c = cq; s = sq; }
        else { c = f1 / d; s = g / r; }
    } else {
        dlartg(f, g, c, s, r);
    }
}

// DLARF('L'): H = I - tau v v^T with v = column i of a (a(i,i) == 1) applied to columns c > i.
// Thread c owns column c: sequential dot over the rows, then its own rank-1 update.
template <int N>
NLB_DEV void cb_reflect_trailing(double* a, int i, double tau, int tid) {
    constexpr int LD = N + 1;
    if (tau == 0.0) return;
    // DLARF's scan for the last non-zero row of v, done in parallel: thread r tests v(r); the index of
    // the highest set bit of the ballots is the same integer the sequential scan returns.
    int lastv;
    {
        const bool nz = tid >= i && a[tid + i * LD] != 0.0;
        if constexpr (N <= 32) {
            const unsigned m = __ballot_sync(0xffffffffu >> (32 - N), nz);
            lastv = m ? (32 - __clz(m)) - i : 0;
        } else {
            __shared__ unsigned ballots[N / 32];
            const unsigned m = __ballot_sync(0xffffffffu, nz);
            if ((tid & 31) == 0) ballots[tid >> 5] = m;
            __syncthreads();
            lastv = 0;
#pragma unroll
            for (int w = N / 32 - 1; w >= 0; --w) {
                const unsigned mw = ballots[w];
                if (lastv == 0 && mw) lastv = w * 32 + (32 - __clz(mw)) - i;
            }
            __syncthreads();
        }
    }
    if (lastv <= 0) return;
    if (tid > i) {
        const double* v = a + i + i * LD;
        double* c = a + i + tid * LD;
        double temp = 0.0;
        for (int r = 0; r < lastv; ++r) temp += c[r] * v[r];
        const double w = 0.0 + 1.0 * temp;
        if (w != 0.0) {
            const double t = (-tau) * w;
            for (int r = 0; r < lastv; ++r) c[r] = c[r] + v[r] * t;
        }
    }
}

// qr_factor(b, q = q, r = r): DGEQR2 on a copy of B, R = upper triangle, Q by DORG2R.
template <int N>
NLB_DEV void cb_qr_full(const double (&brow)[N], double* q, double* r, double* tau, double* sq, int tid) {
    constexpr int LD = N + 1;
#pragma unroll
    for (int j = 0; j < N; ++j) q[tid + j * LD] = brow[j];
    __syncthreads();
    for (int i = 0; i < N; ++i) {
        // DLARFG on column i (every thread evaluates the same scalars)
        double t = 0.0, beta = 0.0, sc = 0.0;
        bool scale = false;
        if (N - i > 1) {
            // DNRM2 of the sub-column.  Ordinary case - every element zero or inside Blue's middle range: the result is
            // sqrt of the ordered sum of squares, so thread rr squares element rr and every thread adds the squares in
            // order (leading zeros add nothing).  Otherwise: the three-accumulator scan, element by element.
            double xnorm;
            {
                const double ax = tid > i ? fabs(q[tid + i * LD]) : 0.0;
                const bool odd = ax > 0x1p486 || (ax < 0x1p-511 && ax != 0.0);
                sq[tid] = ax * ax;
                if (!__syncthreads_or(odd)) {
                    double amed = 0.0;
                    if constexpr (N % 8 == 0) {
#pragma unroll 1
                        for (int rr = (i + 1) & ~7; rr < N; rr += 8) {
#pragma unroll
                            for (int u = 0; u < 8; ++u) amed += sq[rr + u];
                        }
                    } else {
                        for (int rr = i + 1; rr < N; ++rr) amed += sq[rr];
                    }
                    xnorm = sqrt(amed);
                } else {
                    Dnrm2 acc;
                    for (int rr = i + 1; rr < N; ++rr) acc.add(q[rr + i * LD]);
                    xnorm = acc.value();
                }
            }
            if (xnorm != 0.0) {
                double alpha = q[i + i * LD];
                beta = -nl_sign(dlapy2(alpha, xnorm), alpha);
                const double safmin = 0x1p-969;
                int knt = 0;
                if (fabs(beta) < safmin) {               // rare: rescale until beta is representable
                    const double rsafmn = 1.0 / safmin;
                    do {
                        ++knt;
                        __syncthreads();
                        if (tid > i) q[tid + i * LD] = rsafmn * q[tid + i * LD];
                        __syncthreads();
                        beta = beta * rsafmn;
                        alpha = alpha * rsafmn;
                    } while (fabs(beta) < safmin && knt < 20);
                    Dnrm2 acc2;
                    for (int rr = i + 1; rr < N; ++rr) acc2.add(q[rr + i * LD]);
                    xnorm = acc2.value();
                    beta = -nl_sign(dlapy2(alpha, xnorm), alpha);
                }
                t = (beta - alpha) / beta;
                sc = 1.0 / (alpha - beta);
                scale = true;
                for (int k = 0; k < knt; ++k) beta = beta * safmin;
            }
        }
        __syncthreads();                                 // all reads of column i done
        if (scale) {
            if (tid > i) q[tid + i * LD] = sc * q[tid + i * LD];
            if (tid == i) q[i + i * LD] = beta;
        }
        if (tid == 0) tau[i] = t;
        __syncthreads();
        if (i < N - 1) {
            const double aii = q[i + i * LD];
            __syncthreads();
            if (tid == i) q[i + i * LD] = 1.0;
            __syncthreads();
            cb_reflect_trailing<N>(q, i, t, tid);
            __syncthreads();
            if (tid == i) q[i + i * LD] = aii;
            __syncthreads();
        }
    }
    for (int j = 0; j < N; ++j)
        if (tid <= j + 1) r[cb_ro(j) + tid] = (tid <= j) ? q[tid + j * LD] : 0.0;
    __syncthreads();
    // DORG2R
    for (int i = N - 1; i >= 0; --i) {
        const double t = tau[i];
        if (i < N - 1) {
            if (tid == i) q[i + i * LD] = 1.0;
            __syncthreads();
            cb_reflect_trailing<N>(q, i, t, tid);
            __syncthreads();
            if (tid > i) q[tid + i * LD] = (-t) * q[tid + i * LD];
        }
        if (tid == i) q[i + i * LD] = 1.0 - t;
        if (tid < i) q[tid + i * LD] = 0.0;
        __syncthreads();
    }
}

// DQR1UP (full Q): Q R + u v^T -> Q1 R1.  w, cs, sn: N-entry shared work vectors.  R packed (cb_ro).
template <int N>
NLB_DEV void cb_qr_rank1_update(double* q, double* r, const double* u, const double* v, double* w, double* cs,
                                double* sn, int tid, int cb_it = 64) {
    (void)cb_it;
    constexpr int LD = N + 1;
    constexpr int NWARP = (N + 31) / 32;
    const unsigned mask = cb_mask<N>();
    const int wid = tid >> 5;
    // w = Q^T u : thread t owns column t of Q
    {
        double s = 0.0;
        const double* qc = q + tid * LD;
#pragma unroll 8
        for (int l = 0; l < N; ++l) s += qc[l] * u[l];
        w[tid] = s;
    }
    __syncthreads();
    CB_TRACE2(0);
    // DQRTV1: the Givens chain that folds w into w(1), bottom-up (strictly sequential)
    // The chain only needs r of each rotation; it is evaluated (redundantly, uniformly) with the
    // division-free part of DLARTG, the partial r's are kept in sn[], and thread i then computes
    // c(i), s(i) of its own rotation from the same (f, g) pair - the same values, off the chain.
    double w0;
    {
        double rr = w[N - 1];
        for (int i = N - 2; i >= 0; --i) {
            if (tid == 0) sn[i] = rr;                      // g of rotation i
            rr = dlartg_r(w[i], rr);
        }
        w0 = rr;
    }
    __syncthreads();
    CB_TRACE2(1);
    if (tid < N - 1) {
        double c, s2, t;
        dlartg(w[tid], sn[tid], c, s2, t);
        cs[tid] = c;
        sn[tid] = s2;
    }
    __syncthreads();
    CB_TRACE2(2);
    double* rc = r + cb_ro(tid);
    // DQRQH: R -> upper Hessenberg, thread t owns column t
    {
        const int ii = (N - 2 < tid) ? N - 2 : tid;
        double t = rc[ii + 1];
        for (int j = ii; j >= 0; --j) {
            const double c = cs[j], s = sn[j], rj = rc[j];
            rc[j + 1] = c * t - s * rj;
            t = c * rj + s * t;
        }
        rc[0] = t;
    }
    CB_TRACE2(3);
    // DQROT('B'): Q <- Q G^T, last rotation first; thread t owns row t
    {
        double hi = q[tid + (N - 1) * LD];
        for (int i = N - 2; i >= 0; --i) {
            const double c = cs[i], s = sn[i], lo = q[tid + i * LD];
            const double t = c * lo + s * hi;
            q[tid + (i + 1) * LD] = c * hi - s * lo;
            hi = t;
        }
        q[tid] = hi;
    }
    __syncthreads();
    CB_TRACE2(4);
    // first row of R += w(1) v^T
    rc[0] = rc[0] + w0 * v[tid];
    // DQHQR: back to triangular.  Rotation j is generated from column j once rotations 0..j-1 have been applied to
    // it, then applied to the columns to its right: a chain of N-1 DLARTGs.  The warp that owns columns 32w..32w+31
    // runs its 32 links with (c, s) passed by shuffle; the warps to its right apply those 32 rotations from shared
    // memory afterwards (one CTA barrier per 32 links instead of one per link).
    {
        double t = rc[0];
#pragma unroll 1
        for (int wv = 0; wv < NWARP; ++wv) {
            const int j0 = 32 * wv;
            const int j1 = (j0 + 32 < N - 1) ? j0 + 32 : N - 1;
            if (wid == wv) {
#pragma unroll 1
                for (int j = j0; j < j1; ++j) {
                    const double rn = (tid >= j) ? rc[j + 1] : 0.0;
                    double c = 0.0, s = 0.0;
                    if (tid == j) {
                        double rjj;
                        cb_dlartg(t, rn, c, s, rjj);
                        cs[j] = c; sn[j] = s;
                        rc[j] = rjj;
                        rc[j + 1] = 0.0;
                    }
                    c = __shfl_sync(mask, c, j & 31);
                    s = __shfl_sync(mask, s, j & 31);
                    if (tid > j) {
                        rc[j] = c * t + s * rn;
                        t = c * rn - s * t;
                    }
                }
            }
            if (NWARP > 1) {
                __syncthreads();
                if (wid > wv) {
#pragma unroll 4
                    for (int j = j0; j < j1; ++j) {
                        const double c = cs[j], s = sn[j], rn = rc[j + 1];
                        rc[j] = c * t + s * rn;
                        t = c * rn - s * t;
                    }
                }
            }
        }
        if (tid == N - 1) rc[N - 1] = t;
        if (NWARP == 1) __syncthreads();
    }
    CB_TRACE2(5);
    // DQROT('F')
    {
        double lo = q[tid];
        for (int i = 0; i < N - 1; ++i) {
            const double c = cs[i], s = sn[i], hi = q[tid + (i + 1) * LD];
            q[tid + i * LD] = c * lo + s * hi;
            lo = c * hi - s * lo;
        }
        q[tid + (N - 1) * LD] = lo;
    }
    __syncthreads();
}

// Residual interface for this kernel: F::component(x, i, n, ctx) = f_i(x), x in shared memory.
struct ExtRosenbrockCoop {
    static constexpr int ID = FCN_EXT_ROSENBROCK;
    NLB_DEV static double component(const double* x, int i, int, const SysCtx&) { return ExtRosenbrock::component(x, i); }
};

template <class F, int N>
__global__ void __launch_bounds__(N)
coop_broyden_kernel(DevParams p, long long nsys, long long B, double* __restrict__ xg, double* __restrict__ fg,
                    const double* __restrict__ sys, const double* __restrict__ shared,
                    nlb_iteration_behavior* __restrict__ ibg, int32_t* __restrict__ statusg) {
    using S = CoopBroydenSmem<N>;
    constexpr int LD = S::LD;
    extern __shared__ double smem[];
    double* q = smem;
    double* r = q + S::MAT;   // packed, cb_ro(j) + i
    double* x = r + S::RPK;
    double* fvec = x + N;
    double* xold = fvec + N;
    double* fvold = xold + N;
    double* dx = fvold + N;
    double* df = dx + N;
    double* s = df + N;
    double* w = s + N;
    double* cs = w + N;
    double* sn = cs + N;
    double* stage = xold;     // STAGE x SLD staging rows for B^T f: xold, fvold, dx, df, s, w, cs, sn (+ 8 spare doubles)
                              // are all dead by then (the old state is about to be overwritten, the update is over)
    double* tau = w;          // Householder scalars: only live inside the refactorisation
    double* xp = s;           // perturbed copy of x for the forward differences: only live there too
    double brow[N];           // row `tid` of the Broyden matrix B
    const int tid = threadIdx.x;
    const long long b = blockIdx.x;
    if (b >= nsys) return;
    SysCtx c{sys ? sys + b : nullptr, shared, B, N, N};

    const double ftol = p.fcn_tol, xtol = p.var_tol, gtol = p.grad_tol;
    bool restart = true, xcnvrg = false, fcnvrg = false, gcnvrg = false;
    int neval = 0, iter = 0, njac = 0, flag = 0, jcount = 0, status = NLB_NO_ERROR;

    x[tid] = xg[(long long)tid * B + b];
    __syncthreads();
    fvec[tid] = F::component(x, tid, N, c);
    __syncthreads();
    double f = 0.5 * cb_dot<N>(fvec, fvec);
    ++neval;
    if (cb_maxabs<N>(fvec) < ftol) fcnvrg = true;

    if (!fcnvrg) {
        const double stpmax = 100.0 * nl_max(cb_norm2<N>(x), (double)N);
        double fold = f;
        for (;;) {
            ++iter;
            if (iter > p.max_iter_guard) { flag = 1; break; }
CB_TRACE(0);
                        if (restart) {
                // forward-difference Jacobian, column by column (vfh_jac_fcn :262-275)
                xp[tid] = x[tid];
                __syncthreads();
                const double eps = 0x1p-26;
#pragma unroll
                for (int j = 0; j < N; ++j) {
                    const double temp = x[j];
                    double h = eps * fabs(temp);
                    if (h == 0.0) h = eps;
                    if (tid == 0) xp[j] = temp + h;
                    __syncthreads();
                    const double f1 = F::component(xp, tid, N, c);
                    brow[j] = (f1 - fvec[tid]) / h;
                    __syncthreads();
                    if (tid == 0) xp[j] = temp;
                }
                ++njac;
                __syncthreads();
                CB_TRACE(1);
                cb_qr_full<N>(brow, q, r, tau, cs, tid);
                CB_TRACE(2);
                jcount = 0;
            } else {
                df[tid] = fvec[tid] - fvold[tid];
                dx[tid] = x[tid] - xold[tid];
                __syncthreads();
                const double x2 = cb_dot<N>(dx, dx);
                {   // s = df - matmul(b, dx): row t accumulates over j from zero
                    double acc = 0.0;
#pragma unroll
                    for (int j = 0; j < N; ++j) acc = acc + brow[j] * dx[j];
                    double sv = df[tid] - acc;
                    // recip_mult_array (DRSCL)
                    const double smlnum = 0x1p-1022, bignum = 1.0 / smlnum;
                    double cden = x2, cnum = 1.0;
                    for (;;) {
                        const double cden1 = cden * smlnum, cnum1 = cnum / bignum;
                        double mul;
                        bool done;
                        if (fabs(cden1) > fabs(cnum) && cnum != 0.0) { mul = smlnum; done = false; cden = cden1; }
                        else if (fabs(cnum1) > fabs(cden)) { mul = bignum; done = false; cnum = cnum1; }
                        else { mul = cnum / cden; done = true; }
                        sv = mul * sv;
                        if (done) break;
                    }
                    s[tid] = sv;
                    // rank1_update (DGER, alpha = 1): row t of B
#pragma unroll
                    for (int j = 0; j < N; ++j) {
                        const double dj = dx[j];
                        if (dj != 0.0) brow[j] = brow[j] + sv * (1.0 * dj);
                    }
                }
                __syncthreads();
                CB_TRACE(3);
                cb_qr_rank1_update<N>(q, r, s, dx, w, cs, sn, tid, iter);
                CB_TRACE(4);
                ++jcount;
            }

            CB_TRACE(5);
            // gradient B^T F -> dx ; save state ; -Q^T F -> df
            {
                double t1 = 0.0, t2 = 0.0;
                // column `tid` of B^T f = sum_i b(i, tid) f(i) in the order i = 0, 1, ...: rows of B are published
                // STAGE at a time from the registers of their owners, every thread then adds its column's terms
                constexpr int ST = S::STAGE < N ? S::STAGE : N;
                for (int i0 = 0; i0 < N; i0 += ST) {
                    if (tid >= i0 && tid < i0 + ST) {
                        double* dst = stage + (tid - i0) * S::SLD;
#pragma unroll
                        for (int j = 0; j < N; ++j) dst[j] = brow[j];
                    }
                    __syncthreads();
#pragma unroll
                    for (int u = 0; u < ST; ++u) t1 += stage[u * S::SLD + tid] * fvec[i0 + u];
                    __syncthreads();
                }
                const double* qc = q + tid * LD;
#pragma unroll 8
                for (int i = 0; i < N; ++i) t2 += qc[i] * fvec[i];
                __syncthreads();          // everyone is done reading dx / df of the update step
                dx[tid] = 0.0 + 1.0 * t1;
                df[tid] = 0.0 + (-1.0) * t2;
                xold[tid] = x[tid];
                fvold[tid] = fvec[tid];
            }
            fold = f;
            __syncthreads();
            CB_TRACE(6);
            // R step = -Q^T F (DTRSV upper, no-trans, non-unit): a chain of N divisions.  Thread i keeps entry i in a
            // register; the warp that owns rows 32w..32w+31 runs its 32 links with the pivot entry passed by shuffle
            // (every lane forms the same quotient), then publishes them, and the warps above it apply those 32 columns
            // to their rows from shared memory - one CTA barrier per 32 links.
            {
                constexpr int NWARP = (N + 31) / 32;
                __shared__ unsigned solved_nz[NWARP];
                const unsigned mask = cb_mask<N>();
                const int wid = tid >> 5;
                double v = df[tid];
                // the diagonal and its correctly rounded reciprocal, formed by all threads at once before the chain and
                // published (w, cs are free here): a link then costs one shuffle, five FMAs and the update
                const double dg = r[cb_ro(tid) + tid];
                w[tid] = dg;
                cs[tid] = 1.0 / dg;
                __syncthreads();
#pragma unroll 1
                for (int wv = NWARP - 1; wv >= 0; --wv) {
                    const int j0 = 32 * wv;
                    const int j1 = (j0 + 31 < N - 1) ? j0 + 31 : N - 1;
                    if (wid == wv) {
                        unsigned nz = 0;
#pragma unroll 1
                        for (int j = j1; j >= j0; --j) {
                            const double dj = w[j], yj = cs[j];
                            const double rij = r[cb_ro(j) + (tid < j ? tid : j)];
                            const double xj = __shfl_sync(mask, v, j & 31);
                            const bool live = xj != 0.0;               // the reference skips a zero entry altogether
                            const double t = nl_div_by_rcp(xj, dj, yj);     // == xj / r(j,j)
                            if (live && tid < j) v = v - t * rij;
                            if (live && tid == j) v = t;
                            nz |= (live ? 1u : 0u) << (j & 31);
                        }
                        df[tid] = v;
                        if ((tid & 31) == 0) solved_nz[wv] = nz;
                    }
                    if (wv > 0) {
                        __syncthreads();
                        if (wid < wv) {
                            const unsigned nz = solved_nz[wv];
#pragma unroll 4
                            for (int j = j1; j >= j0; --j) {
                                if ((nz >> (j & 31)) & 1u) v = v - df[j] * r[cb_ro(j) + tid];
                            }
                        }
                    }
                }
                __syncthreads();
            }

            CB_TRACE(7);
            double temp = cb_dot<N>(dx, df);
            if (temp >= 0.0) { restart = true; continue; }

            if (p.use_line_search) {
                temp = cb_dot<N>(df, df);
                __syncthreads();
                if (temp > stpmax) df[tid] = df[tid] * (stpmax / temp);
                __syncthreads();
                {   // limit_search_vector
                    const double mag = cb_norm2_par<N>(df, w, tid);
                    if (mag != 0.0 && mag > stpmax) df[tid] = (stpmax / mag) * df[tid];
                    __syncthreads();
                }
                CB_TRACE(8);
                // ls_search_mimo
                int ls_status = NLB_NO_ERROR, ls_eval = 0, niter = 0;
                const double slope = cb_dot<N>(dx, df);
                if (slope >= 0.0) {
                    ls_status = NLB_DIVERGENT_BEHAVIOR_ERROR;
                    f = 0.0;
                } else {
                    __syncthreads();
                    w[tid] = fabs(df[tid]) / nl_max(fabs(xold[tid]), 1.0);
                    __syncthreads();
                    double test = 0.0;
                    for (int i = 0; i < N; ++i) {
                        const double tt = w[i];
                        if (tt > test) test = tt;
                    }
                    const double alamin = 0x1p-51 / test;
                    double alam = 1.0, alam1 = 0.0, f1 = 0.0, tmplam = 0.0;
                    for (;;) {
                        x[tid] = xold[tid] + alam * df[tid];
                        __syncthreads();
                        fvec[tid] = F::component(x, tid, N, c);
                        __syncthreads();
                        f = 0.5 * cb_dot<N>(fvec, fvec);
                        ++ls_eval;
                        ++niter;
                        if (alam < alamin) {
                            bool same = true;
                            for (int i = 0; i < N; ++i) same = same && ((x[i] - xold[i]) == 0.0);
                            if (same) { ls_status = NLB_CONVERGENCE_ERROR; break; }
                            __syncthreads();
                            x[tid] = xold[tid];
                            __syncthreads();
                            break;
                        } else if (f <= fold + p.ls_alpha * alam * slope) {
                            break;
                        } else {
                            tmplam = backtrack_min(niter, fold, f, f1, alam, alam1, slope);
                        }
                        alam1 = alam;
                        f1 = f;
                        alam = nl_max(tmplam, p.ls_factor * alam);
                        if (ls_eval >= p.ls_max_fcn_evals) { ls_status = NLB_CONVERGENCE_ERROR; break; }
                        __syncthreads();
                    }
                }
                neval += ls_eval;
                if (ls_status != NLB_NO_ERROR) { status = ls_status; break; }
            } else {
                x[tid] = x[tid] + df[tid];
                __syncthreads();
                fvec[tid] = F::component(x, tid, N, c);
                __syncthreads();
                f = 0.5 * cb_dot<N>(fvec, fvec);
                ++neval;
            }

            CB_TRACE(9);
            // test_convergence(x, xold, fvec, dx, lg = .false.)
            bool check = false;
            xcnvrg = false; fcnvrg = false; gcnvrg = false;
            if (cb_maxabs<N>(fvec) < ftol) { fcnvrg = true; check = true; }
            else {
                __syncthreads();
                w[tid] = fabs(x[tid] - xold[tid]) / nl_max(fabs(x[tid]), 1.0);
                __syncthreads();
                double xnorm = 0.0;
                for (int i = 0; i < N; ++i) xnorm = nl_max(w[i], xnorm);
                if (xnorm < xtol) { xcnvrg = true; check = true; }
            }
            (void)gtol;
            if (check) break;
            CB_TRACE(10);
            restart = (jcount >= p.jacobian_interval);
            if (neval >= p.max_fcn_evals) { flag = 1; break; }
            __syncthreads();
        }
    }
    __syncthreads();
    xg[(long long)tid * B + b] = x[tid];
    fg[(long long)tid * B + b] = fvec[tid];
    if (tid == 0) {
        if (ibg) {
            nlb_iteration_behavior o;
            o.iter_count = iter; o.fcn_count = neval; o.jacobian_count = njac; o.gradient_count = 0;
            o.converge_on_fcn = fcnvrg; o.converge_on_chng = xcnvrg; o.converge_on_zero_diff = gcnvrg;
            ibg[b] = o;
        }
        if (statusg) statusg[b] = status != NLB_NO_ERROR ? status : (flag != 0 ? NLB_CONVERGENCE_ERROR : NLB_NO_ERROR);
    }
}

}  // namespace nlb
