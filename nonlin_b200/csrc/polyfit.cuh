// polyfit.cuh — batched polynomial%fit / fit_thru_zero / evaluate: one CUDA thread per data set.
//
// Reference behaviour reproduced here (src/nonlin_polynomials.f90):
//   polyfit_kernel    poly_fit            :146-199   (Vandermonde columns 1, x, x*x, ... then solve_least_squares)
//                     poly_fit_thru_zero  :202-253   (columns x, x*x, ...; c0 = 0)
//   polyval_kernel    poly_eval_double    :256-283   (Horner from the highest coefficient)
// `solve_least_squares` (:198, :252) is linalg's wrapper of LAPACK DGELS (external package).  The kernel follows
// Reference LAPACK 3.12.0 DGELS for m >= n and one right-hand side: DLANGE('M') of A and y with DLASCL when a
// norm leaves [2^-970, 2^970], DGEQR2 (DLARFG + DLARF with its trailing-zero scans), the reflectors applied to y
// as DORM2R('L','T') does, DTRTRS's exact-zero diagonal test, DTRSM, and the scaling undone.
//
// Mapping.  The m x n Vandermonde matrix and the right-hand side (n + 1 columns of m rows) of one data set live in
// a workspace laid out [column][row][thread]: lane-contiguous, so every pass of the factorisation is conflict-free
// in shared memory and fully coalesced in global memory.  Small fits (128 threads' workspaces fit one CTA's shared
// memory: m (n + 1) <= 225 doubles) run out of shared memory; larger ones use a global workspace with the grid
// capped (persistent, grid-stride over data sets) so that the live workspace stays inside the 126 MB L2.  H(i) is applied to y in the same two passes that apply it to the trailing
// columns of A (LAPACK does it afterwards; v_i is final by then either way, so the arithmetic is identical), each
// column's dot product accumulating in row order.  Algorithmic HBM bytes per data set: 8 (m + m + n) (+ 4 status)
// with per-set abscissae, 8 (m + n) with shared ones.
#pragma once
#include "nlb_math.cuh"

namespace nlb {

constexpr int POLY_MAX_COLS = 8;
// rows fetched per batch in the factorisation passes, and the register cap of the global-workspace variant
#ifndef NLB_POLY_RB_SMEM
#define NLB_POLY_RB_SMEM 4
#endif
#ifndef NLB_POLY_RB_GLOBAL
#define NLB_POLY_RB_GLOBAL 2
#endif
#ifndef NLB_POLY_GLOBAL_MINB
#define NLB_POLY_GLOBAL_MINB 4
#endif
constexpr int LA_INVALID_OPERATION_ERROR = 107;   // linalg's code for a rank-deficient solve_least_squares

// DLASCL('G') multiplier sequence: calls apply(mul) one or more times so that the product of the multipliers is
// cto / cfrom without intermediate over/underflow.
template <class Fn>
NLB_DEV void dlascl_apply(double cfrom, double cto, Fn apply) {
    const double smlnum = 0x1p-1022;
    const double bignum = 1.0 / smlnum;
    double cfromc = cfrom, ctoc = cto;
    for (;;) {
        const double cfrom1 = cfromc * smlnum;
        double mul;
        bool done;
        if (cfrom1 == cfromc) {
            mul = ctoc / cfromc;
            done = true;
        } else {
            const double cto1 = ctoc / bignum;
            if (cto1 == ctoc) {
                mul = ctoc;
                done = true;
                cfromc = 1.0;
            } else if (fabs(cfrom1) > fabs(ctoc) && ctoc != 0.0) {
                mul = smlnum;
                done = false;
                cfromc = cfrom1;
            } else if (fabs(cto1) > fabs(cfromc)) {
                mul = bignum;
                done = false;
                ctoc = cto1;
            } else {
                mul = ctoc / cfromc;
                done = true;
                if (mul == 1.0) return;
            }
        }
        apply(mul);
        if (done) break;
    }
}

template <int NC, bool SMEM>
__global__ void __launch_bounds__(128, SMEM ? 1 : NLB_POLY_GLOBAL_MINB)
polyfit_kernel(long long B, int npts, int thru_zero, int x_shared, const double* __restrict__ x,
               const double* __restrict__ y, double* __restrict__ coeffs, int32_t* __restrict__ status,
               double* __restrict__ work) {
    // Workspace of this thread: column c, row r at W[(c * npts + r) * T].  SMEM: the CTA's dynamic shared memory
    // (T = 128 lanes, conflict-free); otherwise the global workspace (T = all threads of the grid, coalesced).
    extern __shared__ double pf_smem[];
    constexpr int RB = SMEM ? NLB_POLY_RB_SMEM : NLB_POLY_RB_GLOBAL;
    const long long GT = (long long)gridDim.x * blockDim.x;
    const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long T = SMEM ? (long long)blockDim.x : GT;
    double* const W = SMEM ? pf_smem + threadIdx.x : work + tid;
    const long long cs = (long long)npts * T;         // column stride; row stride is T
#define PF_AT(r, c) W[(long long)(c) * cs + (long long)(r) * T]
    const double smlnum = 0x1p-970, bignum = 0x1p970;  // dlamch('S') / dlamch('P') and its reciprocal

    for (long long b = tid; b < B; b += GT) {
        // Vandermonde columns and the right-hand side; DLANGE('M') of both on the way
        double anrm = 0.0, bnrm = 0.0;
        for (int r = 0; r < npts; ++r) {
            const double xv = x_shared ? x[r] : x[(long long)r * B + b];
            double v = thru_zero ? xv : 1.0;
            PF_AT(r, 0) = v;
            double t = fabs(v);
            if (anrm < t || t != t) anrm = t;
#pragma unroll
            for (int c = 1; c < NC; ++c) {
                v = v * xv;
                PF_AT(r, c) = v;
                t = fabs(v);
                if (anrm < t || t != t) anrm = t;
            }
            const double yv = y[(long long)r * B + b];
            PF_AT(r, NC) = yv;
            t = fabs(yv);
            if (bnrm < t || t != t) bnrm = t;
        }

        int iascl = 0, ibscl = 0;
        bool zero_matrix = false;
        if (anrm > 0.0 && anrm < smlnum) {
            iascl = 1;
        } else if (anrm > bignum) {
            iascl = 2;
        } else if (anrm == 0.0) {
            zero_matrix = true;
        }
        if (iascl)
            dlascl_apply(anrm, iascl == 1 ? smlnum : bignum, [&](double mul) {
                for (int c = 0; c < NC; ++c)
                    for (int r = 0; r < npts; ++r) PF_AT(r, c) = PF_AT(r, c) * mul;
            });
        if (!zero_matrix) {
            if (bnrm > 0.0 && bnrm < smlnum) ibscl = 1;
            else if (bnrm > bignum) ibscl = 2;
            if (ibscl)
                dlascl_apply(bnrm, ibscl == 1 ? smlnum : bignum, [&](double mul) {
                    for (int r = 0; r < npts; ++r) PF_AT(r, NC) = PF_AT(r, NC) * mul;
                });
        }

        double sol[NC];
        int info = 0;
        if (zero_matrix) {
#pragma unroll
            for (int k = 0; k < NC; ++k) sol[k] = 0.0;
        } else {
#pragma unroll 1
            for (int i = 0; i < NC; ++i) {
                // DLARFG on column i: alpha = a(i,i), x = a(i+1:m-1, i)
                double tau = 0.0;
                if (npts - i > 1) {
                    Dnrm2 acc;
                    {
                        int r = i + 1;
                        for (; r + RB <= npts; r += RB) {
                            double t[RB];
#pragma unroll
                            for (int k = 0; k < RB; ++k) t[k] = PF_AT(r + k, i);
#pragma unroll
                            for (int k = 0; k < RB; ++k) acc.add(t[k]);
                        }
                        for (; r < npts; ++r) acc.add(PF_AT(r, i));
                    }
                    double xnorm = acc.value();
                    if (xnorm != 0.0) {
                        double alpha = PF_AT(i, i);
                        double beta = -nl_sign(dlapy2(alpha, xnorm), alpha);
                        const double safmin = 0x1p-969;
                        int knt = 0;
                        if (fabs(beta) < safmin) {
                            const double rsafmn = 1.0 / safmin;
                            do {
                                ++knt;
                                for (int r = i + 1; r < npts; ++r) PF_AT(r, i) = rsafmn * PF_AT(r, i);
                                beta = beta * rsafmn;
                                alpha = alpha * rsafmn;
                            } while (fabs(beta) < safmin && knt < 20);
                            Dnrm2 acc2;
                            for (int r = i + 1; r < npts; ++r) acc2.add(PF_AT(r, i));
                            xnorm = acc2.value();
                            beta = -nl_sign(dlapy2(alpha, xnorm), alpha);
                        }
                        tau = (beta - alpha) / beta;
                        const double sc = 1.0 / (alpha - beta);
                        {
                            int r = i + 1;
                            for (; r + RB <= npts; r += RB) {
                                double t[RB];
#pragma unroll
                                for (int k = 0; k < RB; ++k) t[k] = PF_AT(r + k, i);
#pragma unroll
                                for (int k = 0; k < RB; ++k) PF_AT(r + k, i) = sc * t[k];
                            }
                            for (; r < npts; ++r) PF_AT(r, i) = sc * PF_AT(r, i);
                        }
                        for (int j = 0; j < knt; ++j) beta = beta * safmin;
                        PF_AT(i, i) = beta;
                    }
                }
                if (tau == 0.0) continue;              // H(i) = I
                // DLARF's scans: last non-zero row of v (v(i) = 1), last trailing column / whether y has a
                // non-zero entry in those rows
                int lastv = npts - i;
                while (lastv > 1 && PF_AT(i + lastv - 1, i) == 0.0) --lastv;
                int lastc = 0;
                for (int c = NC - 1; c > i && lastc == 0; --c)
                    for (int r = i; r < i + lastv; ++r)
                        if (PF_AT(r, c) != 0.0) { lastc = c - i; break; }
                bool ylive = false;
                for (int r = i; r < i + lastv && !ylive; ++r) ylive = PF_AT(r, NC) != 0.0;
                // w = C^T v for every live column (DGEMV 'T'), each sum in row order.  Rows are fetched RB at a
                // time (all loads of a batch in flight together) and then consumed in order.
                const int rend = i + lastv;
                bool live[NC + 1];
#pragma unroll
                for (int c = 0; c <= NC; ++c) live[c] = (c == NC) ? ylive : (c > i && c - i <= lastc);
                double w[NC + 1];
#pragma unroll
                for (int c = 0; c <= NC; ++c) w[c] = 0.0;
#pragma unroll
                for (int c = 1; c <= NC; ++c)
                    if (live[c]) w[c] += PF_AT(i, c) * 1.0;            // row i: v(i) = 1
                int r = i + 1;
                for (; r + RB <= rend; r += RB) {
                    double vv[RB], aa[RB][NC + 1];
#pragma unroll
                    for (int k = 0; k < RB; ++k) {
                        vv[k] = PF_AT(r + k, i);
#pragma unroll
                        for (int c = 1; c <= NC; ++c)
                            if (c > i) aa[k][c] = PF_AT(r + k, c);
                    }
#pragma unroll
                    for (int k = 0; k < RB; ++k) {
#pragma unroll
                        for (int c = 1; c <= NC; ++c)
                            if (live[c]) w[c] += aa[k][c] * vv[k];
                    }
                }
                for (; r < rend; ++r) {
                    const double v = PF_AT(r, i);
#pragma unroll
                    for (int c = 1; c <= NC; ++c)
                        if (live[c]) w[c] += PF_AT(r, c) * v;
                }
                // C := C - tau v w^T (DGER, zero entries of w skipped)
                bool any = false;
#pragma unroll
                for (int c = 1; c <= NC; ++c) {
                    const double wc = 0.0 + 1.0 * w[c];
                    w[c] = (live[c] && wc != 0.0) ? (-tau) * wc : 0.0;
                    any = any || w[c] != 0.0;
                }
                if (!any) continue;
#pragma unroll
                for (int c = 1; c <= NC; ++c)
                    if (w[c] != 0.0) PF_AT(i, c) = PF_AT(i, c) + 1.0 * w[c];
                r = i + 1;
                for (; r + RB <= rend; r += RB) {
                    double vv[RB], aa[RB][NC + 1];
#pragma unroll
                    for (int k = 0; k < RB; ++k) {
                        vv[k] = PF_AT(r + k, i);
#pragma unroll
                        for (int c = 1; c <= NC; ++c)
                            if (c > i) aa[k][c] = PF_AT(r + k, c);
                    }
#pragma unroll
                    for (int k = 0; k < RB; ++k) {
#pragma unroll
                        for (int c = 1; c <= NC; ++c)
                            if (w[c] != 0.0) PF_AT(r + k, c) = aa[k][c] + vv[k] * w[c];
                    }
                }
                for (; r < rend; ++r) {
                    const double v = PF_AT(r, i);
#pragma unroll
                    for (int c = 1; c <= NC; ++c)
                        if (w[c] != 0.0) PF_AT(r, c) = PF_AT(r, c) + v * w[c];
                }
            }
            // DTRTRS: singular if a diagonal entry of R is exactly zero; otherwise DTRSM on (Q^T y)(1:n)
#pragma unroll
            for (int k = 0; k < NC; ++k) {
                sol[k] = PF_AT(k, NC);
                if (info == 0 && PF_AT(k, k) == 0.0) info = k + 1;
            }
            if (info == 0) {
#pragma unroll
                for (int k = NC - 1; k >= 0; --k) {
                    if (sol[k] != 0.0) {
                        sol[k] = sol[k] / PF_AT(k, k);
#pragma unroll
                        for (int i = 0; i < k; ++i) sol[i] = sol[i] - sol[k] * PF_AT(i, k);
                    }
                }
                auto scale_sol = [&](double mul) {
#pragma unroll
                    for (int k = 0; k < NC; ++k) sol[k] = sol[k] * mul;
                };
                if (iascl) dlascl_apply(anrm, iascl == 1 ? smlnum : bignum, scale_sol);
                if (ibscl) dlascl_apply(ibscl == 1 ? smlnum : bignum, bnrm, scale_sol);
            }
        }
        if (thru_zero) coeffs[b] = 0.0;
#pragma unroll
        for (int k = 0; k < NC; ++k) coeffs[(long long)(k + (thru_zero ? 1 : 0)) * B + b] = sol[k];
        if (status) status[b] = info ? LA_INVALID_OPERATION_ERROR : 0;
    }
#undef PF_AT
}

// polynomial%evaluate: y(i, b) = p_b(x_i), Horner from the highest coefficient (order >= 1).
__global__ void __launch_bounds__(128)
polyval_kernel(long long B, int order, int npts, int x_shared, const double* __restrict__ coeffs,
               const double* __restrict__ x, double* __restrict__ y) {
    const long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (b >= B) return;
    double c[POLY_MAX_COLS + 1];
#pragma unroll
    for (int j = 0; j <= POLY_MAX_COLS; ++j) c[j] = (j <= order) ? coeffs[(long long)j * B + b] : 0.0;
    for (int i = 0; i < npts; ++i) {
        const double xv = x_shared ? x[i] : x[(long long)i * B + b];
        double v;
        if (order == 0) {
            v = c[0];
        } else {
            v = vget(c, order) * xv + vget(c, order - 1);
            for (int j = order - 2; j >= 0; --j) v = v * xv + vget(c, j);
        }
        y[(long long)i * B + b] = v;
    }
}

}  // namespace nlb
