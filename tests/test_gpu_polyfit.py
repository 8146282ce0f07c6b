"""GPU parity tests of the batched polynomial fit: polyfit_kernel / polyval_kernel through the
Python mirror of `polynomial` over the C ABI, against the CPU oracle (bit for bit), against README Example 3, and
through size-independent properties at full batch size."""
import numpy as np
import pytest

from nonlin_b200.workloads import POLYFIT_XP, POLYFIT_YP

pytestmark = pytest.mark.gpu


def f12_10(v):
    return ("%.10f" % v).replace("-0.", "-.")


def test_readme_example_3_on_the_engine(engine):
    import nonlin_b200 as nb

    B = 257
    y = np.ascontiguousarray(np.tile(POLYFIT_YP[:, None], (1, B)))
    p = nb.polynomial()
    st = p.fit(POLYFIT_XP, y, 3)
    assert np.all(st == 0) and p.order() == 3
    for b in (0, 100, B - 1):
        assert [f12_10(p.get(i)[b]) for i in (1, 2, 3, 4)] == ["1.1866141861", "0.4466136311", "-.1223204989", "1.0647628218"]
    yf = p.evaluate(POLYFIT_XP)
    assert "%.5f" % np.abs(yf[:, 0] - POLYFIT_YP).max() == "0.50636"
    assert np.array_equal(y[:, 0], POLYFIT_YP)                 # y is not overwritten


@pytest.mark.parametrize("order", [1, 2, 3, 4, 5, 6, 7])
@pytest.mark.parametrize("npts,shared", [(8, False), (21, True), (100, False)])
def test_fit_parity_vs_oracle(engine, oracle, order, npts, shared):
    import nonlin_b200 as nb

    rng = np.random.default_rng(1000 * order + npts)
    B = 1500
    x = rng.uniform(-1.5, 2.0, size=(npts,) if shared else (npts, B))
    y = rng.standard_normal((npts, B))
    p = nb.polynomial()
    st = p.fit(x, y, order)
    co, sto = oracle.polyfit_batch(x, y, order)
    assert np.array_equal(st, sto) and np.all(st == 0)
    assert np.array_equal(p.get_all(), co)
    xe = rng.uniform(-2.0, 2.0, size=(7,) if shared else (7, B))
    assert np.array_equal(p.evaluate(xe), oracle.polyval_batch(co, xe))


@pytest.mark.parametrize("order", [1, 2, 3, 5, 8])
def test_fit_thru_zero_parity_vs_oracle(engine, oracle, order):
    import nonlin_b200 as nb

    rng = np.random.default_rng(order)
    npts, B = 33, 1000
    x = rng.uniform(0.1, 2.0, size=(npts, B))
    y = rng.standard_normal((npts, B))
    p = nb.polynomial()
    st = p.fit_thru_zero(x, y, order)
    co, sto = oracle.polyfit_batch(x, y, order, thru_zero=True)
    assert np.array_equal(st, sto) and np.array_equal(p.get_all(), co) and np.all(p.get(1) == 0.0)
    assert p.order() == order


def test_scaling_branches_and_degenerate_sets_parity(engine, oracle):
    import nonlin_b200 as nb

    rng = np.random.default_rng(9)
    npts, B = 15, 64
    x = np.ascontiguousarray(np.tile(np.linspace(0.5, 2.0, npts)[:, None], (1, B)))
    y = rng.standard_normal((npts, B))
    y[:, 0::8] *= 2.0 ** 990            # bnrm > bignum
    y[:, 1::8] *= 2.0 ** -1000          # bnrm < smlnum
    y[:, 2::8] = 0.0                    # zero right-hand side
    y[3, 3::8] = np.nan                 # NaN propagates through DLANGE and the factorisation
    x[:, 4::8] = 0.0                    # rank deficient: status 107
    x[:, 5::8] = 1.0                    # all abscissae equal
    p = nb.polynomial()
    st = p.fit(x, y, 2)
    co, sto = oracle.polyfit_batch(x, y, 2)
    assert np.array_equal(st, sto) and np.array_equal(p.get_all(), co, equal_nan=True)
    assert np.all(st[4::8] == nb.LA_INVALID_OPERATION_ERROR)
    # through zero: tiny / huge abscissae exercise the scaling of A, x == 0 the zero-matrix return
    xs = x.copy()
    xs[:, 0::8] *= 2.0 ** -990
    xs[:, 1::8] *= 2.0 ** 980
    st = p.fit_thru_zero(xs, y, 1)
    co, sto = oracle.polyfit_batch(xs, y, 1, thru_zero=True)
    assert np.array_equal(st, sto) and np.array_equal(p.get_all(), co, equal_nan=True)
    assert np.all(p.get_all()[:, 4::8] == 0.0) and np.all(st[4::8] == 0)


def test_interpolating_order_and_api_errors(engine, oracle):
    import nonlin_b200 as nb

    x = np.array([0.0, 1.0, 2.0, 4.0])
    y = np.array([[1.0, 0.0], [3.0, 1.0], [-2.0, 8.0], [5.0, 64.0]])
    p = nb.polynomial()
    st = p.fit(x, y, 3)                                        # order = npts - 1
    co, _ = oracle.polyfit_batch(x, y, 3)
    assert np.all(st == 0) and np.array_equal(p.get_all(), co)
    assert np.allclose(p.evaluate(x), y, atol=1e-11)
    for bad in (0, 4, 9):                                      # order < 1, order >= npts: `error stop 4`
        with pytest.raises(nb.NonlinError) as e:
            p.fit(x, y, bad)
        assert e.value.code == nb.NLB_ERR_SIZE
    with pytest.raises(nb.NonlinError) as e:                   # more than 8 fitted coefficients
        p.fit(np.linspace(0, 1, 30), np.ones((30, 2)), 8)
    assert e.value.code == nb.NLB_ERR_UNSUPPORTED
    with pytest.raises(nb.NonlinError):                        # size(y) /= size(x): `error stop 3`
        p.fit(np.linspace(0, 1, 5), np.ones((4, 2)), 2)


def test_device_tensors_equal_host_arrays(engine):
    import torch

    import nonlin_b200 as nb

    rng = np.random.default_rng(17)
    npts, B = 21, 5000
    x = rng.uniform(0.0, 2.0, size=(npts, B))
    y = rng.standard_normal((npts, B))
    ph = nb.polynomial()
    sth = ph.fit(x, y, 3)
    pd = nb.polynomial()
    xd, yd = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()
    std = pd.fit(xd, yd, 3)
    ev = pd.evaluate(xd)
    torch.cuda.synchronize()
    assert np.array_equal(pd.get_all().cpu().numpy(), ph.get_all()) and np.array_equal(std.cpu().numpy(), sth)
    assert np.array_equal(ev.cpu().numpy(), ph.evaluate(x))


def test_full_size_properties(engine, oracle):
    """2^20 data sets of README Example 2's shape: sampled bitwise parity plus the normal equations
    A^T (A c - y) = 0, which hold for any least-squares solution independently of the batch size."""
    import nonlin_b200 as nb
    from nonlin_b200 import workloads as W

    B = 1 << 20
    w = W.WORKLOADS["C1"](B)
    y = w["args"]
    p = nb.polynomial()
    st = p.fit(POLYFIT_XP, y, 3)
    assert np.all(st == 0)
    c = p.get_all()
    idx = np.arange(0, B, 997)
    co, _ = oracle.polyfit_batch(POLYFIT_XP, np.ascontiguousarray(y[:, idx]), 3)
    assert np.array_equal(c[:, idx], co)
    A = np.vander(POLYFIT_XP, 4, increasing=True)
    r = A @ c - y
    assert np.abs(A.T @ r).max() < 1e-10 * np.abs(y).max() * 21
    # the LM solver on the same data reaches the same coefficients (README: "very similar to Example 2")
    obj = nb.vecfcn_helper(); obj.set_fcn("lsq_poly_fit", 21, 4)
    xs = np.ones((4, 4096))
    nb.least_squares_solver().solve(obj, xs, args=np.ascontiguousarray(y[:, :4096]))
    assert np.abs(xs[::-1] - c[:, :4096]).max() < 1e-4


@pytest.mark.parametrize("npts", [44, 45, 46])
def test_storage_variant_boundary(engine, oracle, npts):
    """npts (order + 2) = 225 is the last fit that runs from shared memory (225 KB per CTA); one more point switches to
    the global workspace.  Both sides of the boundary must give the oracle's bits."""
    import nonlin_b200 as nb

    rng = np.random.default_rng(npts)
    B = 700
    x = rng.uniform(-1.0, 1.5, size=(npts, B))
    y = rng.standard_normal((npts, B))
    p = nb.polynomial()
    st = p.fit(x, y, 4)
    co, sto = oracle.polyfit_batch(x, y, 4)
    assert np.array_equal(st, sto) and np.array_equal(p.get_all(), co)
