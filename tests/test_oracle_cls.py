"""Pins the oracle's restatement of constrained_least_squares_solver (cls_solve,
src/nonlin_least_squares.f90:938-1176) to what the reference's own tests assert for it
(tests/nonlin_test_solve.f90:973-1230): the roots of test_constrained_least_squares_1..4 and
the feasibility property of test_constrained_least_squares_bounds.  The reference publishes no
iteration counts or digits for this solver and its linear algebra is in the un-vendored linalg
package, so bit-level parity of this path is unpinned; these are tolerance-level anchors.  CPU only.
"""
import numpy as np
import pytest

BIG = np.finfo(np.float64).max


@pytest.mark.parametrize("x0", [(0.5, 0.5), (1.0, 1.0)])
def test_cls_1_analytic_jacobian_explicit_huge_limits(oracle, x0):
    # test_constrained_least_squares_1: fcn1 + jac1, limits set to +-huge(big), |x| = (5, 3) to 1e-6
    x, f, ib, st = oracle.cls_solve("misc_2fcn", x0, params=oracle.params(use_analytic_jacobian=1),
                                    lower=[-BIG, -BIG], upper=[BIG, BIG])
    assert st == 0
    assert np.all(np.abs(np.abs(x) - [5.0, 3.0]) <= 1e-6)
    # explicit +-huge limits and no limits are the same solve (cls_solve installs +-huge itself)
    x2, f2, ib2, st2 = oracle.cls_solve("misc_2fcn", x0, params=oracle.params(use_analytic_jacobian=1))
    assert np.array_equal(x, x2) and np.array_equal(f, f2) and ib == ib2 and st == st2


@pytest.mark.parametrize("x0", [(0.5, 0.5), (1.0, 1.0)])
def test_cls_2_poorly_scaled(oracle, x0):
    # test_constrained_least_squares_2: fcn2, set_max_fcn_evals(5000), |x| = (5000, 10) to 1e-6
    x, f, ib, st = oracle.cls_solve("poorly_scaled_2fcn", x0, params=oracle.params(max_fcn_evals=5000))
    assert st == 0
    assert np.all(np.abs(np.abs(x) - [5000.0, 10.0]) <= 1e-6)
    # with the default budget of 100 evaluations the reference would `error stop NL_CONVERGENCE_ERROR`
    _, _, ib100, st100 = oracle.cls_solve("poorly_scaled_2fcn", x0)
    assert st100 == 106 and ib100["fcn_count"] >= 100


def test_cls_3_agrees_with_levenberg_marquardt_on_the_cubic_fit(oracle):
    # test_constrained_least_squares_3: lsfcn1 (21 x 4) from [1,1,1,1]; |x - xc| <= 1e-5
    x, _, _, st = oracle.solve("lm", "lsq_poly_fit", [1.0] * 4)
    xc, _, ibc, stc = oracle.cls_solve("lsq_poly_fit", [1.0] * 4)
    assert st == 0 and stc == 0
    assert np.all(np.abs(x - xc) <= 1e-5)


@pytest.mark.parametrize("x0", [(0.5, 0.5), (1.0, 1.0)])
@pytest.mark.parametrize("analytic", [0, 1])
def test_cls_4_args(oracle, x0, analytic):
    # test_constrained_least_squares_4: fcn1 with args a = 2, with and without jac1
    x, f, ib, st = oracle.cls_solve("misc_2fcn_a", x0, sys=[2.0], params=oracle.params(use_analytic_jacobian=analytic))
    assert st == 0
    assert np.all(np.abs(np.abs(x) - [5.0, 3.0]) <= 1e-6)


def test_cls_bounds(oracle):
    # test_constrained_least_squares_bounds: start (1, 1) outside [4, 5.6] x [2, 3.6]; solution stays feasible
    low, high = np.array([4.0, 2.0]), np.array([5.6, 3.6])
    x, f, ib, st = oracle.cls_solve("misc_2fcn", [1.0, 1.0], lower=low, upper=high)
    assert np.all(x >= low) and np.all(x <= high)
    assert st == 0 and np.all(np.abs(x - [5.0, 3.0]) <= 1e-6)


def test_cls_active_bound_stops_at_the_face(oracle):
    # the unconstrained root (5, 3) is cut off: the solver must end on the face x1 = 4.5, inside the box
    low, high = np.array([0.0, 0.0]), np.array([4.5, 10.0])
    x, f, ib, st = oracle.cls_solve("misc_2fcn", [1.0, 1.0], lower=low, upper=high,
                                    params=oracle.params(max_fcn_evals=500))
    assert np.all(x >= low) and np.all(x <= high)
    assert abs(x[0] - 4.5) < 1e-6


def test_cls_non_finite_start_returns_quietly(oracle):
    # :1043-1045 — early `return`: no error, iteration_behavior stays zero.  +-Inf is clamped to +-huge by
    # apply_limits first and then fails the |x| == huge test of is_finite_array.
    for x0 in ([np.nan, 1.0], [np.inf, 1.0], [1.0, -np.inf]):
        x, f, ib, st = oracle.cls_solve("misc_2fcn", x0)
        assert st == 0 and all(v == 0 for v in ib.values())
    x, _, _, _ = oracle.cls_solve("misc_2fcn", [np.inf, 1.0])
    assert x[0] == BIG


def test_cls_underdetermined(oracle):
    import ctypes as C
    from oracle.nl_oracle import IB_DTYPE, ClsOptions

    # nvar > neqn -> NL_UNDERDEFINED_PROBLEM_ERROR (:1005); ext_rosenbrock accepts any (m, n) pair only when equal,
    # so drive the check through the polynomial residual family with m < n
    fid = oracle.fcn_id("exp_decay_4")
    o = ClsOptions()
    oracle.lib.nlo_cls_options_default(C.byref(o))
    x = np.ones(4); f = np.zeros(3); ib = np.zeros(1, dtype=IB_DTYPE)
    p = oracle.params()
    sysv = np.ones(3); shared = np.linspace(0, 1, 3)
    st = oracle.lib.nlo_cls_solve(fid, 3, 4, C.byref(p), C.byref(o), x.ctypes.data_as(C.c_void_p),
                                  f.ctypes.data_as(C.c_void_p), sysv.ctypes.data_as(C.c_void_p),
                                  shared.ctypes.data_as(C.c_void_p), ib.ctypes.data_as(C.c_void_p))
    assert st == 212


def test_cls_settings_change_the_path(oracle):
    base = oracle.cls_solve("misc_2fcn", [1.0, 1.0])
    small = oracle.cls_solve("misc_2fcn", [1.0, 1.0], trust_region_radius=0.05)
    assert small[3] == 0 and np.all(np.abs(np.abs(small[0]) - [5.0, 3.0]) <= 1e-6)
    assert small[2]["iter_count"] > base[2]["iter_count"]


def test_cls_batch_equals_single(oracle):
    rng = np.random.default_rng(5)
    B = 64
    x0 = rng.uniform(0.2, 8.0, size=(2, B))
    xb, fb, ibb, stb = oracle.cls_solve_batch("misc_2fcn", x0, lower=[0.0, 0.0], upper=[6.0, 6.0])
    for b in range(0, B, 7):
        x, f, ib, st = oracle.cls_solve("misc_2fcn", x0[:, b], lower=[0.0, 0.0], upper=[6.0, 6.0])
        assert np.array_equal(x, xb[:, b]) and np.array_equal(f, fb[:, b]) and st == stb[b]
        assert ib["iter_count"] == ibb["iter_count"][b] and ib["fcn_count"] == ibb["fcn_count"][b]


def test_cls_matches_scipy_bounded_least_squares_when_limits_are_inactive(oracle):
    """Independent cross-check (not an oracle): scipy's trust-region-reflective bounded least squares reaches the
    same minimiser on the noisy cubic fits when the limits do not bind.  With binding limits the reference's
    method only promises feasibility and descent (that is all its own bounds test asserts), which is checked too."""
    from scipy.optimize import least_squares

    from nonlin_b200 import workloads as W

    B = 24
    w = W.c1_lm_polyfit(B)
    xp = W.POLYFIT_XP

    def res(c, y):
        return c[0] * xp ** 3 + c[1] * xp ** 2 + c[2] * xp + c[3] - y

    x0 = np.full((4, B), 0.5)
    lo, hi = [-10.0] * 4, [10.0] * 4
    x, f, ib, st = oracle.cls_solve_batch("lsq_poly_fit", x0, sys=w["args"], lower=lo, upper=hi)
    assert np.all(st == 0)
    for b in range(B):
        r = least_squares(res, x0[:, b], bounds=(lo, hi), args=(w["args"][:, b],), xtol=1e-14, ftol=1e-14, gtol=1e-14)
        assert np.abs(x[:, b] - r.x).max() < 1e-5
        assert 0.5 * np.sum(f[:, b] ** 2) <= r.cost * (1 + 1e-10)
    lo, hi = np.array([0.0, -1.0, 0.0, 0.0]), np.array([1.0, 1.0, 2.0, 2.0])      # x1 <= 1 cuts off the optimum
    x, f, ib, st = oracle.cls_solve_batch("lsq_poly_fit", x0, sys=w["args"], lower=lo, upper=hi)
    f0 = np.stack([res(x0[:, b], w["args"][:, b]) for b in range(B)], axis=1)
    assert np.all(x >= lo[:, None]) and np.all(x <= hi[:, None])
    assert np.all(np.sum(f ** 2, axis=0) <= np.sum(f0 ** 2, axis=0))
