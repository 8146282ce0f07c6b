"""GPU parity tests of the one-variable solvers: solve_1var_kernel through the Python mirror
of brent_solver / newton_1var_solver over the C ABI, against the CPU oracle (bit for bit on the functions built from
+ - * / and the shared software exp; to tolerance on sin(x)/x, whose libm differs between CPU and GPU), and against
the roots the reference's own tests assert."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SOLVERS = {"brent": "brent_solver", "newton_1var": "newton_1var_solver"}


def run(nb, solver, fcn, lim1, lim2, args=None, analytic=False, x0=None, want_f=True, **settings):
    obj = nb.fcn1var_helper()
    obj.set_fcn(fcn)
    if analytic:
        obj.set_diff()
    s = getattr(nb, SOLVERS[solver])()
    for k, v in settings.items():
        getattr(s, "set_" + k)(v)
    B = len(lim1)
    x = np.zeros(B) if x0 is None else np.array(x0, dtype=np.float64)
    ib = nb.iteration_behavior(B)
    st = s.solve(obj, x, nb.value_pair(np.asarray(lim1, dtype=np.float64), np.asarray(lim2, dtype=np.float64)), ib=ib,
                 args=args, want_f=want_f)
    return x, s.last_f, ib, st


def oracle_settings(oracle, analytic, settings):
    names = {"max_fcn_evals": "max_fcn_evals", "fcn_tolerance": "fcn_tol", "var_tolerance": "var_tol", "diff_tolerance": "diff_tol"}
    return oracle.params1(use_analytic_diff=int(analytic), **{names[k]: v for k, v in settings.items()})


@pytest.mark.parametrize("solver", ["brent", "newton_1var"])
def test_reference_tests_root_is_pi(engine, solver):
    import nonlin_b200 as nb

    B = 300
    x, f, ib, st = run(nb, solver, "sinx_div_x", np.full(B, 1.5), np.full(B, 5.0))
    assert np.all(st == 0) and np.all(np.abs(x - np.pi) <= 1e-6) and np.all(np.abs(f) < 1e-8)
    x, f, ib, st = run(nb, solver, "sinx_div_x_a", np.full(B, 1.5), np.full(B, 5.0), args=np.full((1, B), 2.0))
    assert np.all(st == 0) and np.all(np.abs(x - np.pi) <= 1e-6)
    # scalar limits broadcast like the reference's single value_pair
    obj = nb.fcn1var_helper(); obj.set_fcn("sinx_div_x")
    xs = np.zeros(B)
    assert np.all(getattr(nb, SOLVERS[solver])().solve(obj, xs, nb.value_pair(1.5, 5.0)) == 0)
    assert np.all(np.abs(xs - np.pi) <= 1e-6)


CASES = [
    # solver, fcn, analytic, settings
    ("brent", "cubic_wallis", False, {}),
    ("brent", "exp_minus_x", False, {}),
    ("brent", "cubic_args", False, {}),
    ("brent", "cubic_args", False, {"fcn_tolerance": 1e-14, "var_tolerance": 1e-15, "max_fcn_evals": 60}),
    ("brent", "cubic_args", False, {"max_fcn_evals": 6}),                 # budget exhausted: status 106, x = 0
    ("newton_1var", "cubic_wallis", False, {}),
    ("newton_1var", "cubic_wallis", True, {}),
    ("newton_1var", "exp_minus_x", False, {}),
    ("newton_1var", "exp_minus_x", True, {}),
    ("newton_1var", "cubic_args", False, {}),
    ("newton_1var", "cubic_args", True, {"fcn_tolerance": 1e-15, "var_tolerance": 1e-9}),
    ("newton_1var", "cubic_args", True, {"diff_tolerance": 0.5}),        # converge_on_zero_diff fires
    ("newton_1var", "cubic_args", False, {"max_fcn_evals": 5}),
]


@pytest.mark.parametrize("solver,fcn,analytic,settings", CASES)
def test_parity_vs_oracle(engine, oracle, solver, fcn, analytic, settings):
    import nonlin_b200 as nb

    rng = np.random.default_rng(31)
    B = 20000
    if fcn == "cubic_args":
        args = np.stack([rng.uniform(-8, -1, B), rng.uniform(-2, 2, B), rng.uniform(-1, 1, B), rng.uniform(0.5, 2, B)])
        lo, hi = rng.uniform(-2.0, -0.5, B), rng.uniform(4.0, 7.0, B)
    elif fcn == "cubic_wallis":
        args, lo, hi = None, rng.uniform(0.0, 2.0, B), rng.uniform(2.2, 4.0, B)
    else:
        args, lo, hi = None, rng.uniform(-1.0, 0.5, B), rng.uniform(0.6, 3.0, B)
    swap = rng.random(B) < 0.3                       # limits in either order
    lim1, lim2 = np.where(swap, hi, lo), np.where(swap, lo, hi)
    lim2[::97] = lim1[::97]                          # degenerate bracket: NL_INVALID_INPUT_ERROR
    x0 = rng.standard_normal(B)
    x, f, ib, st = run(nb, solver, fcn, lim1, lim2, args=args, analytic=analytic, x0=x0, **settings)
    xo, fo, ibo, sto = oracle.solve_1var_batch(solver, fcn, lim1, lim2, x0=x0, args=args,
                                               params=oracle_settings(oracle, analytic, settings))
    assert np.array_equal(st, sto)
    assert np.array_equal(x, xo, equal_nan=True) and np.array_equal(f, fo, equal_nan=True)
    assert np.array_equal(ib.view(np.int32), ibo.view(np.int32))
    assert np.all(st[::97] == 201)
    if not settings:
        assert (st == 0).mean() > 0.9


def test_optional_f_absent_and_nan_limits(engine, oracle):
    import nonlin_b200 as nb

    B = 512
    rng = np.random.default_rng(5)
    lo, hi = rng.uniform(0.0, 2.0, B), rng.uniform(2.2, 4.0, B)
    lo[::50] = np.nan
    x, f, ib, st = run(nb, "newton_1var", "cubic_wallis", lo, hi, want_f=False)
    xo, fo, ibo, sto = oracle.solve_1var_batch("newton_1var", "cubic_wallis", lo, hi, want_f=False)
    assert f is None and fo is None
    assert np.array_equal(st, sto) and np.array_equal(x, xo, equal_nan=True)
    assert np.array_equal(ib.view(np.int32), ibo.view(np.int32))
    x, f, ib, st = run(nb, "brent", "cubic_wallis", lo, hi, max_fcn_evals=40)
    xo, fo, ibo, sto = oracle.solve_1var_batch("brent", "cubic_wallis", lo, hi, params=oracle.params1(max_fcn_evals=40))
    assert np.array_equal(st, sto) and np.array_equal(x, xo, equal_nan=True) and np.array_equal(f, fo, equal_nan=True)
    assert np.array_equal(ib.view(np.int32), ibo.view(np.int32))


def test_device_tensors_and_api_errors(engine, oracle):
    import torch

    import nonlin_b200 as nb

    B = 1 << 20
    rng = np.random.default_rng(8)
    args = np.stack([rng.uniform(-8, -1, B), rng.uniform(-2, 2, B), rng.uniform(-1, 1, B), rng.uniform(0.5, 2, B)])
    obj = nb.fcn1var_helper(); obj.set_fcn("cubic_args")
    s = nb.brent_solver()
    xd = torch.zeros(B, dtype=torch.float64, device="cuda")
    ad = torch.from_numpy(args).cuda()
    ibd = nb.iteration_behavior(B, like=xd)
    std = s.solve(obj, xd, nb.value_pair(-1.0, 6.0), ib=ibd, args=ad)
    torch.cuda.synchronize()
    x = xd.cpu().numpy()
    ok = std.cpu().numpy() == 0
    assert ok.mean() > 0.9
    res = ((args[3] * x + args[2]) * x + args[1]) * x + args[0]           # full-size property: the residual at the root
    assert np.all(np.abs(res[ok]) < 1e-6) and np.all(x[ok] >= -1.0) and np.all(x[ok] <= 6.0)
    idx = np.arange(0, B, 1013)
    xo, fo, ibo, sto = oracle.solve_1var_batch("brent", "cubic_args", np.full(idx.size, -1.0), np.full(idx.size, 6.0),
                                               args=np.ascontiguousarray(args[:, idx]))
    assert np.array_equal(x[idx], xo) and np.array_equal(s.last_f.cpu().numpy()[idx], fo)
    assert np.array_equal(ibd.cpu().numpy()[idx], ibo.view(np.int32).reshape(-1, 7))
    with pytest.raises(nb.NonlinError):                                    # args missing
        s.solve(obj, np.zeros(4), nb.value_pair(0.0, 1.0))
    with pytest.raises(nb.NonlinError):                                    # no function set
        s.solve(nb.fcn1var_helper(), np.zeros(4), nb.value_pair(0.0, 1.0))
    with pytest.raises(nb.NonlinError):                                    # unknown name
        nb.fcn1var_helper().set_fcn("nope")
    with pytest.raises(nb.NonlinError):                                    # no registered derivative
        o2 = nb.fcn1var_helper(); o2.set_fcn("sinx_div_x"); o2.set_diff()
