"""Pins the CPU oracle to everything the reference publishes for this path:
README Example 1 counts and digits, README Example 2 digits, the roots asserted by the
reference's own test table, and the forward-difference Jacobian checks.  CPU only.
"""
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
KATS = json.load(open(os.path.join(HERE, "golden", "reference_kats.json")))


def fortran_e9_3(v):
    """Fortran E9.3 edit descriptor, e.g. 0.323E-11."""
    if v == 0:
        return "0.000E+00"
    e = int(np.floor(np.log10(abs(v)))) + 1
    mant = v / 10.0 ** e
    s = "%.3f" % mant
    if s.startswith("1.000"):
        e += 1
        s = "0.100"
    return "%sE%+03d" % (s, e)


def test_kat1_readme_example_1(oracle):
    k = KATS["kat1_readme_example_1"]
    p = oracle.params(**k["settings"])
    x, f, ib, st = oracle.solve(k["solver"], k["fcn"], k["x0"], params=p)
    assert st == 0
    assert ["%.5f" % v for v in x] == k["solution_F7.5"]
    assert [fortran_e9_3(v) for v in f] == k["residual_E9.3"]
    assert (ib["iter_count"], ib["fcn_count"], ib["jacobian_count"]) == (k["iter_count"], k["fcn_count"], k["jacobian_count"])
    assert ib["converge_on_fcn"] == 1


def test_kat2_readme_example_2(oracle):
    k = KATS["kat2_readme_example_2"]
    x, f, ib, st = oracle.solve(k["solver"], k["fcn"], k["x0"])
    assert st == 0

    def f12_10(v):
        s = "%.10f" % v
        return s.replace("-0.", "-.")  # Fortran drops the leading zero of negative fractions

    assert f12_10(x[3]) == k["c0_F12.10"]
    assert f12_10(x[2]) == k["c1_F12.10"]
    assert f12_10(x[1]) == k["c2_F12.10"]
    assert f12_10(x[0]) == k["c3_F12.10"]
    assert "%.5f" % np.abs(f).max() == k["max_residual_F7.5"]
    # the same call with the data passed as per-system `args` gives the same bits
    from nonlin_b200.workloads import POLYFIT_YP

    x2, f2, ib2, _ = oracle.solve(k["solver"], k["fcn"], k["x0"], sys=POLYFIT_YP)
    assert np.array_equal(x, x2) and np.array_equal(f, f2) and ib == ib2


@pytest.mark.parametrize("libm", [0, 1])
def test_kat3_powell(oracle, libm):
    k = KATS["kat3_powell_badly_scaled"]
    oracle.set_libm_exp(libm)
    try:
        # tests/nonlin_test_solve.f90:806-848  Newton + line search, analytic Jacobian
        x, f, ib, st = oracle.solve("newton", "powell_badly_scaled", k["x0"], params=oracle.params(use_analytic_jacobian=1))
        assert st == 0 and np.all(np.abs(x - k["solution"]) <= k["tol"])
        assert (ib["iter_count"], ib["fcn_count"], ib["jacobian_count"]) == (51, 95, 51)
        # :851-895  quasi-Newton, analytic Jacobian, line search off
        x, f, ib, st = oracle.solve("quasi_newton", "powell_badly_scaled", k["x0"],
                                    params=oracle.params(use_analytic_jacobian=1, use_line_search=0))
        assert st == 0 and np.all(np.abs(x - k["solution"]) <= k["tol"])
        # finite differences give the same counts
        x, f, ib, st = oracle.solve("newton", "powell_badly_scaled", k["x0"])
        assert st == 0 and (ib["iter_count"], ib["fcn_count"], ib["jacobian_count"]) == (51, 95, 51)
    finally:
        oracle.set_libm_exp(0)


# (iter, nfev, njac) measured by the survey's independent restatement
TABLE_COUNTS = {
    ("quasi_newton", 0): (10, 15, 2), ("quasi_newton", 1): (10, 14, 2),
    ("newton", 0): (6, 9, 6), ("newton", 1): (6, 9, 6),
    ("least_squares", 0): (8, 9, 8), ("least_squares", 1): (9, 10, 8),
}


@pytest.mark.parametrize("solver", ["quasi_newton", "newton", "least_squares"])
@pytest.mark.parametrize("analytic", [0, 1])
def test_kat4_fcn1_all_solvers(oracle, solver, analytic):
    k = KATS["kat4_fcn1"]
    for i, x0 in enumerate(k["x0"]):
        x, f, ib, st = oracle.solve(solver, "misc_2fcn", x0, params=oracle.params(use_analytic_jacobian=analytic))
        assert st == 0
        assert np.all(np.abs(np.abs(x) - k["abs_solution"]) <= k["tol"])
        assert (ib["iter_count"], ib["fcn_count"], ib["jacobian_count"]) == TABLE_COUNTS[(solver, i)]
        # same system with the coefficient passed through args (test_*_3)
        xa, fa, iba, sta = oracle.solve(solver, "misc_2fcn_a", x0, sys=[2.0], params=oracle.params(use_analytic_jacobian=analytic))
        assert sta == 0 and np.array_equal(x, xa) and iba == ib


def test_kat5_fcn2(oracle):
    k = KATS["kat5_fcn2"]
    for x0 in k["x0"]:
        for solver, counts in (("quasi_newton", (3, 4, 1)), ("newton", (2, 3, 2))):
            x, f, ib, st = oracle.solve(solver, "poorly_scaled_2fcn", x0, params=oracle.params(use_line_search=0))
            assert st == 0 and np.all(np.abs(np.abs(x) - k["abs_solution"]) <= k["tol"])
            assert (ib["iter_count"], ib["fcn_count"], ib["jacobian_count"]) == counts
        # LM needs more than the default 100 evaluations (tests/nonlin_test_solve.f90:603)
        x, f, ib, st = oracle.solve("least_squares", "poorly_scaled_2fcn", x0, params=oracle.params(max_fcn_evals=1000))
        assert st == 0 and np.all(np.abs(np.abs(x) - k["abs_solution"]) <= k["tol"])
        x, f, ib, st = oracle.solve("least_squares", "poorly_scaled_2fcn", x0)
        assert st == 106 and ib["fcn_count"] == 100   # would be `error stop NL_CONVERGENCE_ERROR`


def test_kat6_fd_jacobian(oracle):
    k = KATS["kat6_fd_jacobian"]
    for pt in k["points"]:
        r, th = pt
        exact = np.array([[np.cos(th), -r * np.sin(th)], [np.sin(th), r * np.cos(th)]])
        num = oracle.jacobian("polar", pt)
        assert np.all(np.abs(num - exact) <= k["tol"])
        num = oracle.jacobian("polar_scaled", pt, sys=[2.0])
        assert np.all(np.abs(num - 2.0 * exact) <= k["tol"])
        ana = oracle.jacobian("polar", pt, params=oracle.params(use_analytic_jacobian=1))
        assert np.allclose(ana, exact, atol=1e-15)


def test_status_codes_and_edges(oracle):
    # starting on the root: iter 0, one evaluation, no Jacobian (src/nonlin_solve.f90:257-266)
    for solver in ("newton", "quasi_newton"):
        x, f, ib, st = oracle.solve(solver, "misc_2fcn", [5.0, 3.0])
        assert st == 0 and (ib["iter_count"], ib["fcn_count"], ib["jacobian_count"]) == (0, 1, 0) and ib["converge_on_fcn"] == 1
    # evaluation budget exhausted -> NL_CONVERGENCE_ERROR with ib still filled
    x, f, ib, st = oracle.solve("newton", "powell_badly_scaled", [0.0, 1.0], params=oracle.params(max_fcn_evals=20))
    assert st == 106 and ib["fcn_count"] >= 20
    # singular Jacobian at the origin -> uphill / zero direction path must terminate
    x, f, ib, st = oracle.solve("newton", "misc_2fcn", [0.0, 0.0])
    assert st != 0


def test_soft_exp_close_to_libm(oracle):
    rng = np.random.default_rng(0)
    xs = np.concatenate([rng.uniform(-30, 30, 20000), rng.uniform(-1e-3, 1e-3, 1000), [0.0, -0.0, 1.0, -1.0, 700.0, -740.0]])
    got = np.array([oracle.lib.nlo_soft_exp(float(v)) for v in xs])
    ref = np.exp(xs)
    ulp = np.abs(got - ref) / np.spacing(ref)
    assert ulp.max() <= 1.0
    assert oracle.lib.nlo_soft_exp(800.0) == np.inf and oracle.lib.nlo_soft_exp(-800.0) == 0.0


def test_norm2_matches_definition(oracle):
    rng = np.random.default_rng(1)
    for n in (1, 2, 5, 21, 4096):
        v = rng.standard_normal(n) * 10.0 ** rng.integers(-3, 3)
        assert abs(oracle.lib.nlo_norm2(v, n) - np.linalg.norm(v)) <= 4e-16 * n * np.linalg.norm(v)
        assert abs(oracle.lib.nlo_dnrm2(v, n) - np.linalg.norm(v)) <= 4e-16 * n * np.linalg.norm(v)
    big = np.array([1e200, 1e200]); small = np.array([1e-200, 1e-200])
    assert np.isclose(oracle.lib.nlo_norm2(big, 2), np.sqrt(2) * 1e200)
    assert np.isclose(oracle.lib.nlo_dnrm2(big, 2), np.sqrt(2) * 1e200)
    assert np.isclose(oracle.lib.nlo_dnrm2(small, 2), np.sqrt(2) * 1e-200)
    assert oracle.lib.nlo_norm2(np.zeros(3), 3) == 0.0
