"""Cross-checks of the oracle's restated LAPACK / QRUPDATE / MINPACK-lineage pieces against
scipy's real LAPACK (cross-checks, not oracles: SURVEY.md §8c).  CPU only."""
import ctypes as C

import numpy as np
import pytest
import scipy.linalg as sla
import scipy.optimize as sopt


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.mark.parametrize("n", [2, 3, 8, 64])
def test_qr_full_matches_lapack(oracle, n):
    rng = np.random.default_rng(n)
    a = np.asfortranarray(rng.standard_normal((n, n)))
    q = np.zeros((n, n), order="F"); r = np.zeros((n, n), order="F")
    oracle.lib.nlo_dgeqr2_dorg2r(n, _p(a), _p(q), _p(r))
    assert np.allclose(q @ r, a, atol=1e-12)
    assert np.allclose(q.T @ q, np.eye(n), atol=1e-12)
    assert np.allclose(np.tril(r, -1), 0.0)
    q_ref, r_ref = sla.qr(a)   # dgeqrf + dorgqr: same Householder convention -> same signs
    assert np.allclose(q, q_ref, atol=1e-10) and np.allclose(r, r_ref, atol=1e-10)


@pytest.mark.parametrize("n", [2, 5, 64])
def test_qr_rank1_update(oracle, n):
    rng = np.random.default_rng(100 + n)
    a = np.asfortranarray(rng.standard_normal((n, n)))
    u = rng.standard_normal(n); v = rng.standard_normal(n)
    q = np.zeros((n, n), order="F"); r = np.zeros((n, n), order="F")
    oracle.lib.nlo_dgeqr2_dorg2r(n, _p(a), _p(q), _p(r))
    oracle.lib.nlo_dqr1up(n, _p(q), _p(r), _p(u), _p(v))
    assert np.allclose(q @ r, a + np.outer(u, v), atol=1e-11)
    assert np.allclose(q.T @ q, np.eye(n), atol=1e-12)
    assert np.allclose(np.tril(r, -1), 0.0, atol=0)


@pytest.mark.parametrize("n", [2, 4, 16])
def test_lu_solve(oracle, n):
    rng = np.random.default_rng(200 + n)
    a = np.asfortranarray(rng.standard_normal((n, n)))
    b = rng.standard_normal(n)
    lu = a.copy(order="F"); x = b.copy(); piv = np.zeros(n, dtype=np.int32)
    info = oracle.lib.nlo_dgesv(n, _p(lu), _p(piv), _p(x))
    assert info == 0
    assert np.allclose(a @ x, b, atol=1e-10)
    lu_ref, piv_ref = sla.lu_factor(a)
    assert np.array_equal(piv - 1, piv_ref)
    assert np.allclose(lu, lu_ref, atol=1e-12)


@pytest.mark.parametrize("m,n", [(2, 2), (21, 4), (64, 16)])
def test_lmfactor_is_a_pivoted_qr(oracle, m, n):
    rng = np.random.default_rng(300 + m)
    a = np.asfortranarray(rng.standard_normal((m, n)) * rng.uniform(0.1, 10, n))
    fa = a.copy(order="F"); ipvt = np.zeros(n, dtype=np.int32); rdiag = np.zeros(n); acnorm = np.zeros(n)
    oracle.lib.nlo_lmfactor(m, n, _p(fa), _p(ipvt), _p(rdiag), _p(acnorm))
    assert np.allclose(acnorm, np.linalg.norm(a, axis=0))
    _, r_ref, p_ref = sla.qr(a, pivoting=True, mode="economic")
    assert np.array_equal(ipvt - 1, p_ref)
    assert np.allclose(np.abs(rdiag), np.abs(np.diag(r_ref)), rtol=1e-12)
    # strict upper triangle of the factored matrix is R (up to the sign convention of each row)
    r = np.triu(fa[:n, :n], 1)
    assert np.allclose(np.abs(r), np.abs(np.triu(r_ref, 1)), atol=1e-11)


def test_lm_solution_agrees_with_minpack(oracle):
    """scipy.optimize.leastsq is true MINPACK: same minimiser, though not the same counts
    (the reference's lmpar departs from MINPACK)."""
    from nonlin_b200 import workloads as W

    w = W.lm_expdecay4(16)
    t = w["shared"]
    x, f, ib, st = oracle.solve_batch("least_squares", "exp_decay_4", w["x0"], m=w["m"], sys=w["args"], shared=t,
                                      params=oracle.params(max_fcn_evals=1000))
    assert np.all(st == 0)
    for b in range(16):
        y = w["args"][:, b]
        res = lambda p: p[0] * np.exp(-p[1] * t) + p[2] * np.exp(-p[3] * t) - y
        ref, _ = sopt.leastsq(res, w["x0"][:, b], xtol=1e-12, ftol=1e-12)
        assert abs(np.linalg.norm(res(x[:, b])) - np.linalg.norm(res(ref))) <= 1e-8


# ---- pieces behind constrained_least_squares_solver and polynomial%fit ---------------------------
def _f(a):
    return np.asfortranarray(a, dtype=np.float64)


@pytest.mark.parametrize("m,n", [(2, 2), (21, 4), (64, 8), (300, 7)])
def test_unpivoted_qr_solve_matches_lapack(oracle, m, n):
    """qr_factor(a, tau=, qr=) + solve_qr(qr, tau, b) restated as DGEQR2 + DORM2R('L','T') + DTRSV against LAPACK's
    DGEQRF + DORMQR + DTRTRS (blocked, possibly FMA: agreement to rounding, not bits)."""
    from scipy.linalg import lapack

    rng = np.random.default_rng(10 * m + n)
    a = _f(rng.standard_normal((m, n)))
    b = rng.standard_normal(m)
    qr, tau, _, info = lapack.dgeqrf(a)
    cq, _, info2 = lapack.dormqr("L", "T", qr, tau, b.reshape(-1, 1).copy(order="F"), max(1, 64 * n))
    xs, info3 = lapack.dtrtrs(qr[:n, :n], cq[:n], lower=0)
    assert info == 0 and info2 == 0 and info3 == 0
    a2, b2 = a.copy(order="F"), b.copy()
    oracle.lib.nlo_qr_solve(m, n, a2.ctypes.data_as(C.c_void_p), b2.ctypes.data_as(C.c_void_p))
    assert np.allclose(np.triu(a2[:n]), np.triu(qr[:n]), rtol=1e-12, atol=1e-13)      # same R, same signs
    assert np.allclose(b2[:n], xs[:, 0], rtol=1e-10, atol=1e-12)
    assert np.allclose(b2[:n], np.linalg.lstsq(a, b, rcond=None)[0], rtol=1e-9, atol=1e-11)


@pytest.mark.parametrize("m,n", [(21, 4), (100, 6), (9, 8)])
@pytest.mark.parametrize("scale_a,scale_b", [(1.0, 1.0), (2.0 ** -990, 1.0), (2.0 ** 985, 1.0), (1.0, 2.0 ** 990), (1.0, 2.0 ** -1000)])
def test_dgels_matches_lapack_including_scaling_branches(oracle, m, n, scale_a, scale_b):
    from scipy.linalg import lapack

    rng = np.random.default_rng(m + 100 * n)
    t = np.linspace(0.2, 1.5, m)
    a = _f(np.vander(t, n, increasing=True) * scale_a)
    b = rng.standard_normal(m) * scale_b
    lqr, x, info = lapack.dgels(a, b)
    assert info == 0
    a2, b2 = a.copy(order="F"), b.copy()
    assert oracle.lib.nlo_dgels(m, n, a2.ctypes.data_as(C.c_void_p), b2.ctypes.data_as(C.c_void_p)) == 0
    ref = x[:n]
    assert np.allclose(b2[:n], ref, rtol=1e-7, atol=1e-9 * np.abs(ref).max())


def test_dgemv_forms(oracle):
    rng = np.random.default_rng(3)
    a = _f(rng.standard_normal((21, 4)))
    x4, x21 = rng.standard_normal(4), rng.standard_normal(21)
    y = np.zeros(21)
    oracle.lib.nlo_dgemv(0, 21, 4, a.ctypes.data_as(C.c_void_p), x4.ctypes.data_as(C.c_void_p), y.ctypes.data_as(C.c_void_p))
    assert np.allclose(y, a @ x4, rtol=1e-14, atol=1e-14)
    yt = np.zeros(4)
    oracle.lib.nlo_dgemv(1, 21, 4, a.ctypes.data_as(C.c_void_p), x21.ctypes.data_as(C.c_void_p), yt.ctypes.data_as(C.c_void_p))
    assert np.allclose(yt, a.T @ x21, rtol=1e-14, atol=1e-14)
