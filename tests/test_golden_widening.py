"""Committed golden vectors (tests/golden/oracle_batches.npz, made by tests/golden/make_golden.py) for the
SURVEY 8(f) paths - constrained least squares, polynomial fit, one-variable solvers.  The CPU test guards the
oracle against silent drift; the GPU tests compare the engine with the same committed bits."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from make_golden import scalar_case  # noqa: E402

G = np.load(os.path.join(HERE, "golden", "oracle_batches.npz"))


def test_oracle_reproduces_the_committed_vectors(oracle):
    from nonlin_b200 import workloads as W

    for name in ("CLS1", "CLS2"):
        B = G[name + "_status"].shape[0]
        w = W.WORKLOADS[name](B)
        x, f, ib, st = oracle.cls_solve_batch(w["fcn"], w["x0"], m=w["m"], sys=w["args"],
                                              lower=w["settings"]["set_lower_limits"],
                                              upper=w["settings"]["set_upper_limits"])
        assert np.array_equal(x, G[name + "_x"]) and np.array_equal(f, G[name + "_f"])
        assert np.array_equal(ib.view(np.int32).reshape(B, 7), G[name + "_ib"]) and np.array_equal(st, G[name + "_status"])
    w = W.WORKLOADS["C1"](G["POLY1_status"].shape[0])
    c, st = oracle.polyfit_batch(W.POLYFIT_XP, w["args"], 3)
    assert np.array_equal(c, G["POLY1_c"]) and np.array_equal(st, G["POLY1_status"])
    c, st = oracle.polyfit_batch(W.POLYFIT_XP[1:], np.ascontiguousarray(w["args"][1:]), 2, thru_zero=True)
    assert np.array_equal(c, G["POLY0_c"]) and np.array_equal(st, G["POLY0_status"])
    s1 = scalar_case(G["brent_status"].shape[0])
    for solver in ("brent", "newton_1var"):
        x, f, ib, st = oracle.solve_1var_batch(solver, "cubic_args", s1["lim1"], s1["lim2"], args=s1["args"])
        assert np.array_equal(x, G[solver + "_x"]) and np.array_equal(f, G[solver + "_f"])
        assert np.array_equal(ib.view(np.int32).reshape(-1, 7), G[solver + "_ib"]) and np.array_equal(st, G[solver + "_status"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["CLS1", "CLS2"])
def test_engine_constrained_least_squares_vs_golden(engine, name):
    import nonlin_b200 as nb
    from nonlin_b200 import workloads as W

    B = G[name + "_status"].shape[0]
    w = W.WORKLOADS[name](B)
    obj = nb.vecfcn_helper()
    obj.set_fcn(w["fcn"], w["m"], w["n"])
    s = nb.constrained_least_squares_solver()
    for k, v in w["settings"].items():
        getattr(s, k)(v)
    x = w["x0"].copy()
    f = np.zeros((w["m"], B))
    ib = nb.iteration_behavior(B)
    st = s.solve(obj, x, f, ib, args=w["args"])
    assert np.array_equal(x, G[name + "_x"]) and np.array_equal(f, G[name + "_f"])
    assert np.array_equal(ib.view(np.int32).reshape(B, 7), G[name + "_ib"]) and np.array_equal(st, G[name + "_status"])


@pytest.mark.gpu
def test_engine_polynomial_fit_vs_golden(engine):
    import nonlin_b200 as nb
    from nonlin_b200 import workloads as W

    w = W.WORKLOADS["C1"](G["POLY1_status"].shape[0])
    p = nb.polynomial()
    st = p.fit(W.POLYFIT_XP, w["args"], 3)
    assert np.array_equal(p.get_all(), G["POLY1_c"]) and np.array_equal(st, G["POLY1_status"])
    st = p.fit_thru_zero(np.ascontiguousarray(W.POLYFIT_XP[1:]), np.ascontiguousarray(w["args"][1:]), 2)
    assert np.array_equal(p.get_all(), G["POLY0_c"]) and np.array_equal(st, G["POLY0_status"])


@pytest.mark.gpu
@pytest.mark.parametrize("solver", ["brent", "newton_1var"])
def test_engine_one_variable_solvers_vs_golden(engine, solver):
    import nonlin_b200 as nb

    B = G[solver + "_status"].shape[0]
    s1 = scalar_case(B)
    obj = nb.fcn1var_helper()
    obj.set_fcn("cubic_args")
    s = {"brent": nb.brent_solver, "newton_1var": nb.newton_1var_solver}[solver]()
    x = np.zeros(B)
    ib = nb.iteration_behavior(B)
    st = s.solve(obj, x, nb.value_pair(s1["lim1"], s1["lim2"]), ib=ib, args=s1["args"])
    assert np.array_equal(x, G[solver + "_x"]) and np.array_equal(s.last_f, G[solver + "_f"])
    assert np.array_equal(ib.view(np.int32).reshape(B, 7), G[solver + "_ib"]) and np.array_equal(st, G[solver + "_status"])
