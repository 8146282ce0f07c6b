"""Pins the oracle's one-variable solvers (brent_solve / newt1var_solve, src/nonlin_solve.f90:643-1032) to what the
reference's tests assert (tests/nonlin_test_solve.f90:729-790, 898-970: the root pi of sin(x)/x and of a sin(x)/x on
[1.5, 5] to 1e-6, both solvers) and cross-checks Brent against scipy's brentq.  CPU only."""
import numpy as np
import pytest
from scipy.optimize import brentq


@pytest.mark.parametrize("solver", ["brent", "newton_1var"])
def test_reference_tests_root_is_pi(oracle, solver):
    x, f, ib, st = oracle.solve_1var_batch(solver, "sinx_div_x", [1.5], [5.0])         # test_brent_1 / test_newton_1var_1
    assert st[0] == 0 and abs(x[0] - np.pi) <= 1e-6 and abs(f[0]) < 1e-8
    x, f, ib, st = oracle.solve_1var_batch(solver, "sinx_div_x_a", [1.5], [5.0], args=np.array([[2.0]]))   # _2, a = 2
    assert st[0] == 0 and abs(x[0] - np.pi) <= 1e-6
    x2, _, _, _ = oracle.solve_1var_batch(solver, "sinx_div_x", [5.0], [1.5])          # limits in either order (:692-693)
    assert abs(x2[0] - np.pi) <= 1e-6


def test_brent_matches_scipy_brentq(oracle):
    rng = np.random.default_rng(4)
    B = 200
    a = np.stack([rng.uniform(-8, -1, B), rng.uniform(-2, 2, B), rng.uniform(-1, 1, B), rng.uniform(0.5, 2, B)])
    lo, hi = np.full(B, -1.0), np.full(B, 6.0)
    x, f, ib, st = oracle.solve_1var_batch("brent", "cubic_args", lo, hi, args=a,
                                           params=oracle.params1(fcn_tol=1e-13, max_fcn_evals=200))
    poly = lambda t, c: ((c[3] * t + c[2]) * t + c[1]) * t + c[0]
    checked = 0
    for b in range(B):
        if poly(lo[b], a[:, b]) * poly(hi[b], a[:, b]) < 0:
            ref = brentq(poly, lo[b], hi[b], args=(a[:, b],), xtol=1e-14, rtol=1e-14)
            if st[b] == 0 and abs(poly(ref + 1e-6, a[:, b]) - poly(ref - 1e-6, a[:, b])) > 1e-7:   # simple root
                assert abs(poly(x[b], a[:, b])) < 1e-11
                checked += 1
    assert checked > 50


@pytest.mark.parametrize("solver,analytic", [("brent", 0), ("newton_1var", 0), ("newton_1var", 1)])
def test_known_roots(oracle, solver, analytic):
    p = oracle.params1(use_analytic_diff=analytic)
    x, f, ib, st = oracle.solve_1var_batch(solver, "cubic_wallis", [1.0], [3.0], params=p)
    assert st[0] == 0 and abs(x[0] - 2.0945514815423265) < 1e-8                    # Wallis' cubic
    x, f, ib, st = oracle.solve_1var_batch(solver, "exp_minus_x", [0.0], [2.0], params=p)
    assert st[0] == 0 and abs(x[0] - 0.5671432904097838) < 1e-8                    # omega constant
    assert ib["gradient_count"][0] == 0
    if solver == "newton_1var":
        assert ib["jacobian_count"][0] == ib["fcn_count"][0] - 3                    # 2 end points + the extra `f` evaluation
    else:
        assert ib["jacobian_count"][0] == 0 and ib["fcn_count"][0] == ib["iter_count"][0] + 1


def test_quirks(oracle):
    # |lim1 - lim2| < epsilon -> NL_INVALID_INPUT_ERROR (:713, :899); Newton leaves x alone, Brent has zeroed it
    x, f, ib, st = oracle.solve_1var_batch("brent", "cubic_wallis", [2.0], [2.0], x0=[7.0])
    assert st[0] == 201 and x[0] == 0.0 and f[0] == 0.0 and ib["fcn_count"][0] == 0
    x, f, ib, st = oracle.solve_1var_batch("newton_1var", "cubic_wallis", [2.0], [2.0], x0=[7.0])
    assert st[0] == 201 and x[0] == 7.0
    # Brent out of budget: NL_CONVERGENCE_ERROR and x stays 0 (:691, only assigned on convergence)
    x, f, ib, st = oracle.solve_1var_batch("brent", "cubic_wallis", [1.0], [3.0], params=oracle.params1(max_fcn_evals=4))
    assert st[0] == 106 and x[0] == 0.0 and ib["fcn_count"][0] == 4 and f[0] != 0.0
    # Newton: a root at an end point returns at once with fcn_count = 2 and only converge_on_fcn set (:906-923)
    a = np.array([[-8.0], [0.0], [0.0], [1.0]])                                     # x^3 - 8, root 2
    x, f, ib, st = oracle.solve_1var_batch("newton_1var", "cubic_args", [2.0], [5.0], args=a)
    assert st[0] == 0 and x[0] == 2.0 and f[0] == 0.0
    assert (ib["iter_count"][0], ib["fcn_count"][0], ib["jacobian_count"][0], ib["converge_on_fcn"][0]) == (0, 2, 0, 1)
    # the optional `f`: requesting it costs one counted evaluation (:1011-1014)
    _, _, ib1, _ = oracle.solve_1var_batch("newton_1var", "cubic_wallis", [1.0], [3.0], want_f=True)
    _, fn, ib0, _ = oracle.solve_1var_batch("newton_1var", "cubic_wallis", [1.0], [3.0], want_f=False)
    assert fn is None and ib1["fcn_count"][0] == ib0["fcn_count"][0] + 1 and ib1["iter_count"][0] == ib0["iter_count"][0]
    # no sign change in the bracket: Brent still terminates (on the interval width) and reports what it has
    x, f, ib, st = oracle.solve_1var_batch("brent", "cubic_wallis", [3.0], [4.0])
    assert st[0] in (0, 106)
