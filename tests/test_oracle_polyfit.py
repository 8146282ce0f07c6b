"""Pins the oracle's polynomial fit (poly_fit / poly_fit_thru_zero / poly_eval_double,
src/nonlin_polynomials.f90:146-283, with linalg's solve_least_squares restated as LAPACK DGELS) to the
reference's published output — README Example 3 (README.md:218-222), which is also tests/nonlin_test_poly.f90's
data — and cross-checks the DGELS restatement against numpy's real LAPACK.  CPU only."""
import numpy as np
import pytest

from nonlin_b200.workloads import POLYFIT_XP, POLYFIT_YP


def f12_10(v):
    s = "%.10f" % v
    return s.replace("-0.", "-.")  # Fortran drops the leading zero of negative fractions


def test_readme_example_3(oracle):
    c, st = oracle.polyfit_batch(POLYFIT_XP, POLYFIT_YP[:, None], 3)
    assert st[0] == 0
    assert [f12_10(v) for v in c[:, 0]] == ["1.1866141861", "0.4466136311", "-.1223204989", "1.0647628218"]
    yf = oracle.polyval_batch(c, POLYFIT_XP)
    assert "%.5f" % np.abs(yf[:, 0] - POLYFIT_YP).max() == "0.50636"
    # "the results are very similar to the output of Example 2" (README.md:223): LM on the same data
    x, _, _, _ = oracle.solve("lm", "lsq_poly_fit", [1.0] * 4)
    assert np.allclose(x[::-1], c[:, 0], rtol=0, atol=1e-5)


@pytest.mark.parametrize("order", [1, 2, 3, 5, 7])
@pytest.mark.parametrize("npts", [8, 21, 100])
def test_fit_matches_lapack_least_squares(oracle, order, npts):
    rng = np.random.default_rng(100 * order + npts)
    B = 16
    x = np.sort(rng.uniform(-1.5, 2.0, size=(npts, B)), axis=0)
    y = rng.standard_normal((npts, B))
    c, st = oracle.polyfit_batch(x, y, order)
    assert np.all(st == 0)
    for b in range(B):
        A = np.vander(x[:, b], order + 1, increasing=True)
        ref = np.linalg.lstsq(A, y[:, b], rcond=None)[0]
        assert np.allclose(c[:, b], ref, rtol=1e-7, atol=1e-9 * np.abs(ref).max())
    # shared abscissae = the same numbers passed once
    c1, _ = oracle.polyfit_batch(x[:, 0].copy(), y, order)
    c2, _ = oracle.polyfit_batch(np.ascontiguousarray(np.tile(x[:, :1], (1, B))), y, order)
    assert np.array_equal(c1, c2)


@pytest.mark.parametrize("order", [1, 3, 8])
def test_fit_thru_zero(oracle, order):
    rng = np.random.default_rng(order)
    npts, B = 30, 8
    x = rng.uniform(0.1, 2.0, size=(npts, B))
    y = rng.standard_normal((npts, B))
    c, st = oracle.polyfit_batch(x, y, order, thru_zero=True)
    assert np.all(st == 0) and np.all(c[0] == 0.0)
    for b in range(B):
        A = np.stack([x[:, b] ** k for k in range(1, order + 1)], axis=1)
        ref = np.linalg.lstsq(A, y[:, b], rcond=None)[0]
        assert np.allclose(c[1:, b], ref, rtol=1e-6, atol=1e-8 * np.abs(ref).max())


def test_exact_polynomial_is_recovered_and_evaluate_is_horner(oracle):
    x = np.linspace(-1, 1, 12)
    coef = np.array([[6.0], [1.0], [-4.0], [1.0]])            # tests/nonlin_test_poly.f90:61 (x^3 - 4x^2 + x + 6)
    y = oracle.polyval_batch(coef, x)
    assert np.array_equal(y[:, 0], ((coef[3, 0] * x + coef[2, 0]) * x + coef[1, 0]) * x + coef[0, 0])
    c, st = oracle.polyfit_batch(x, y, 3)
    assert st[0] == 0 and np.allclose(c, coef, atol=1e-12)
    for r in (2.0, 3.0, -1.0):                                # its roots
        assert abs(oracle.polyval_batch(coef, np.array([r]))[0, 0]) < 1e-12


def test_dgels_scaling_branches(oracle):
    # DGELS rescales when max|A| or max|y| leaves [2^-970, 2^970]; the answer must match a well-scaled solve
    x = np.linspace(0.5, 2.0, 15)
    rng = np.random.default_rng(3)
    y = rng.standard_normal((15, 4))
    c, _ = oracle.polyfit_batch(x, y, 2)
    cb, st = oracle.polyfit_batch(x, y * 2.0 ** 990, 2)       # bnrm > bignum
    assert np.all(st == 0) and np.allclose(cb / 2.0 ** 990, c, rtol=1e-13)
    cs, st = oracle.polyfit_batch(x, y * 2.0 ** -1000, 2)     # bnrm < smlnum
    assert np.all(st == 0) and np.allclose(cs * 2.0 ** 1000, c, rtol=1e-10)
    xs = x * 2.0 ** -990                                      # thru-zero: every column tiny -> anrm < smlnum
    ct, st = oracle.polyfit_batch(xs, y, 1, thru_zero=True)
    c1, _ = oracle.polyfit_batch(x, y, 1, thru_zero=True)
    assert np.all(st == 0) and np.allclose(ct[1] * 2.0 ** -990, c1[1], rtol=1e-12)
    xl = x * 2.0 ** 980                                       # anrm > bignum
    ct, st = oracle.polyfit_batch(xl, y, 1, thru_zero=True)
    assert np.all(st == 0) and np.allclose(ct[1] * 2.0 ** 980, c1[1], rtol=1e-12)


def test_rank_deficient_and_degenerate_inputs(oracle):
    y = np.arange(10.0)[:, None].copy()
    # all abscissae zero: column 2 of the Vandermonde matrix is zero -> R(2,2) == 0 -> linalg reports an error
    c, st = oracle.polyfit_batch(np.zeros(10), y, 2)
    assert st[0] == 107
    # through zero with x == 0: the whole matrix is zero -> DGELS returns the zero solution, no error
    c, st = oracle.polyfit_batch(np.zeros(10), y, 2, thru_zero=True)
    assert st[0] == 0 and np.all(c == 0.0)
    # order >= npts or < 1: the reference's `error stop 4`
    with pytest.raises(RuntimeError):
        oracle.polyfit_batch(np.arange(3.0), np.ones((3, 1)), 3)
    with pytest.raises(RuntimeError):
        oracle.polyfit_batch(np.arange(3.0), np.ones((3, 1)), 0)
    # order = npts - 1 interpolates
    x = np.array([0.0, 1.0, 2.0, 4.0])
    yy = np.array([[1.0], [3.0], [-2.0], [5.0]])
    c, st = oracle.polyfit_batch(x, yy, 3)
    assert st[0] == 0 and np.allclose(oracle.polyval_batch(c, x), yy, atol=1e-12)
