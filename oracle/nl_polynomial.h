// oracle/nl_polynomial.h — TEST INFRASTRUCTURE ONLY (see oracle/README.md).
//
// CPU restatement of the reference's polynomial fit:
//   poly_fit           <- poly_fit            src/nonlin_polynomials.f90:146-199
//   poly_fit_thru_zero <- poly_fit_thru_zero  src/nonlin_polynomials.f90:202-253
//   poly_eval          <- poly_eval_double    src/nonlin_polynomials.f90:256-283
// Both fits build a Vandermonde matrix and call linalg's `solve_least_squares(a, y)`
// (:198, :252), which is in the external, un-vendored linalg package; it forwards to
// LAPACK DGELS.  la_dgels restates Reference LAPACK 3.12.0 DGELS for TRANS = 'N',
// m >= n, one right-hand side: DLANGE('M') of A and b, DLASCL when a norm is outside
// [smlnum, bignum], DGEQRF (unblocked DGEQR2 for n < 32), DORMQR('L','T') (unblocked
// DORM2R), DTRTRS (exact-zero diagonal check, then DTRSM), and the scaling undone.
//
// Pinned by README Example 3 (README.md:218-222: c0..c3 to the 10 printed digits, max
// residual 0.50636) and cross-checked against numpy's LAPACK in tests/test_oracle_polyfit.py.
// Bit-level parity of the linalg boundary itself is unpinned (vendor BLAS would differ).
#ifndef NL_POLYNOMIAL_H
#define NL_POLYNOMIAL_H

#include "nl_numerics.h"

namespace nlo {

// linalg's LA_INVALID_OPERATION_ERROR (value from linalg's published constants; not in the tree):
// what solve_least_squares reports when DGELS finds an exactly singular R.
enum { LA_INVALID_OPERATION_ERROR = 107 };

// DLASCL('G'): a := a * (cto / cfrom) without over/underflow, m-by-n, column-major.
void la_dlascl_g(real cfrom, real cto, int m, int n, real* a, int lda);
// DLANGE('M'): max |a(i,j)|, NaN-propagating.
real la_dlange_m(int m, int n, const real* a, int lda);
// DGELS('N'), m >= n, nrhs = 1.  a (m x n) is overwritten by its QR factors, b (m) by the
// solution in b(1:n).  work: n entries.  Returns LAPACK info (0, or i > 0: R(i,i) == 0).
int la_dgels(int m, int n, real* a, int lda, real* b, real* tau, real* work);

// coeffs has order+1 entries (c0 first).  y is overwritten, as in the reference (intent(inout)).
// a: npts x (order+1) scratch.  Returns 0 or LA_INVALID_OPERATION_ERROR.
int poly_fit(int npts, int order, const real* x, real* y, real* coeffs, real* a);
int poly_fit_thru_zero(int npts, int order, const real* x, real* y, real* coeffs, real* a);
real poly_eval(int order, const real* coeffs, real x);

}  // namespace nlo
#endif
