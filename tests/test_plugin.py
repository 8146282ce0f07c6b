"""Residuals compiled outside the engine (nonlin_b200/csrc/nlb_plugin.cuh, examples/plugin/): the batch analogue of
vecfcn_helper%set_fcn taking a new procedure at run time (reference src/nonlin_multi_eqn_mult_var.f90:126-140).

CPU part: the plug-in library builds against the engine's headers, loads, and registers its residuals without a GPU.
GPU part: the registered residuals run through the engine's solvers and agree bit for bit with the CPU oracle, which
evaluates the same expressions through a Python callback."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PLUGIN_DIR = os.path.join(ROOT, "examples", "plugin")
PLUGIN = os.path.join(PLUGIN_DIR, "libnlb_plugin_example.so")


def build_plugin(force=False):
    src = os.path.join(PLUGIN_DIR, "freudenstein_roth.cu")
    if force or not os.path.exists(PLUGIN) or os.path.getmtime(PLUGIN) < os.path.getmtime(src):
        cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false", "-shared",
               "-Xcompiler", "-fPIC", "-diag-suppress", "177", "-I", os.path.join(ROOT, "nonlin_b200", "csrc"), src, "-o", PLUGIN]
        if os.path.exists("/usr/bin/g++"):
            cmd[1:1] = ["-ccbin", "/usr/bin/g++"]
        subprocess.check_call(cmd)
    return PLUGIN


def fr(x, sysv, shv):
    return [-13.0 + x[0] + ((5.0 - x[1]) * x[1] - 2.0) * x[1], -29.0 + x[0] + ((x[1] + 1.0) * x[1] - 14.0) * x[1]]


def circle(x, sysv, shv):
    return [x[0] * x[0] + x[1] * x[1] - sysv[0], x[0] - x[1] - 1.0]


def test_plugin_builds_loads_and_registers_without_a_gpu():
    import nonlin_b200 as nb

    build_plugin()
    before = len(nb.vecfcn_names())
    n = nb.load_plugin(PLUGIN)
    names = nb.vecfcn_names()
    if n == 0:                      # already registered earlier in this process: names are unique
        assert "freudenstein_roth" in names
    else:
        assert n == 2 and len(names) == before + 2
    obj = nb.vecfcn_helper()
    obj.set_fcn("freudenstein_roth", 2, 2)
    assert obj.get_equation_count() == 2 and obj.get_variable_count() == 2
    obj2 = nb.vecfcn_helper()
    obj2.set_fcn("circle_line_a", 2, 2)
    assert obj2._info["sys_len"] == 1
    with pytest.raises(nb.NonlinError):
        nb.load_plugin(os.path.join(PLUGIN_DIR, "no_such_library.so"))


@pytest.mark.gpu
def test_plugin_residuals_match_the_oracle_bit_for_bit(engine, oracle):
    import nonlin_b200 as nb

    build_plugin()
    nb.load_plugin(PLUGIN)
    oracle.register_callback("freudenstein_roth", 2, 2, fr)
    oracle.register_callback("circle_line_a", 2, 2, circle, sys_len=1)
    rng = np.random.default_rng(11)
    B = 600
    cases = [("freudenstein_roth", np.array([[4.0], [4.5]]) + rng.uniform(-1.5, 1.5, (2, B)), None),
             ("circle_line_a", np.array([[2.0], [1.0]]) + rng.uniform(-0.5, 0.5, (2, B)), rng.uniform(3.0, 40.0, (1, B)))]
    for fcn, x0, args in cases:
        for solver, cls in (("newton", nb.newton_solver), ("quasi_newton", nb.quasi_newton_solver),
                            ("least_squares", nb.least_squares_solver)):
            obj = nb.vecfcn_helper()
            obj.set_fcn(fcn, 2, 2)
            s = cls()
            s.set_max_fcn_evals(300)
            x = np.ascontiguousarray(x0.copy())
            f = np.zeros((2, B))
            ib = nb.iteration_behavior(B)
            st = s.solve(obj, x, f, ib, args=args)
            xo, fo, ibo, sto = oracle.solve_batch(solver, fcn, x0, m=2, sys=args, params=oracle.params(max_fcn_evals=300),
                                                  nthreads=1)
            assert np.array_equal(st, sto), (fcn, solver)
            assert np.array_equal(x, xo) and np.array_equal(f, fo), (fcn, solver)
            assert np.array_equal(ib, ibo), (fcn, solver)
            assert (st == 0).mean() > 0.5, (fcn, solver)
    # residual and forward-difference Jacobian evaluation of a plug-in residual
    obj = nb.vecfcn_helper()
    obj.set_fcn("freudenstein_roth", 2, 2)
    xs = np.ascontiguousarray(rng.uniform(-3, 6, (2, 50)))
    fe = obj.fcn(xs)
    je = obj.jacobian(xs)
    for b in range(50):
        assert np.array_equal(fe[:, b], np.array(fr(list(xs[:, b]), None, None)))
        jo = oracle.jacobian("freudenstein_roth", xs[:, b], m=2)
        assert np.array_equal(je[:, :, b].T, jo)
