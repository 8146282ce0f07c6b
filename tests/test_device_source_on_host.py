"""The engine's thread-per-system DEVICE SOURCE compiled for the host (tests/device_on_host/shim.cpp, g++
-ffp-contract=off) against the oracle, bit for bit.  CPU only: lets the CPU suite catch an arithmetic change in
nonlin_b200/csrc/*.cuh without a GPU.  Test infrastructure - not a CPU path of the product (libnonlin_b200.so has
none); nvcc's code generation, launch geometry and the cooperative kernels are covered by the `-m gpu` tests only."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "device_on_host", "shim.cpp")
OUT_DIR = os.path.join(ROOT, "tests", "device_on_host", "_build")
CUDA_INC = "/usr/local/cuda/include"


@pytest.fixture(scope="module")
def dh():
    if not os.path.exists(os.path.join(CUDA_INC, "cuda_runtime.h")):
        pytest.skip("CUDA headers not found")
    os.makedirs(OUT_DIR, exist_ok=True)
    so = os.path.join(OUT_DIR, "libdevice_on_host.so")
    deps = [SRC] + [os.path.join(ROOT, "nonlin_b200", "csrc", f) for f in os.listdir(os.path.join(ROOT, "nonlin_b200", "csrc"))]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.check_call([cxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math",
                               "-Wno-attributes", "-Wno-unknown-pragmas", "-I" + CUDA_INC, "-o", so, SRC])
    lib = C.CDLL(so)
    for name in ("dh_solve", "dh_cls_solve", "dh_polyfit", "dh_solve_1var"):
        getattr(lib, name).restype = C.c_int
    return lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def engine_params(**kw):
    from nonlin_b200 import _lib

    p = _lib.nlb_params()
    _lib.load().nlb_params_default(C.byref(p))
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def run_dh_solve(dh, solver, w, **kw):
    from nonlin_b200 import _lib

    lib = _lib.load()
    fid = lib.nlb_vecfcn_lookup(w["fcn"].encode())
    B = w["x0"].shape[1]
    x = w["x0"].copy()
    f = np.zeros((w["m"], B))
    ib = np.zeros((B, 7), dtype=np.int32)
    st = np.zeros(B, dtype=np.int32)
    p = engine_params(**kw)
    rc = dh.dh_solve(solver, fid, C.c_longlong(B), C.byref(p), _p(x), _p(f), _p(w["args"]), _p(w["shared"]), _p(ib), _p(st))
    assert rc == 0
    return x, f, ib, st


def same(got, ref):
    x, f, ib, st = got
    xo, fo, ibo, sto = ref
    assert np.array_equal(st, sto)
    assert np.array_equal(x, xo, equal_nan=True) and np.array_equal(f, fo, equal_nan=True)
    assert np.array_equal(ib, ibo.view(np.int32).reshape(-1, 7))


@pytest.mark.parametrize("name,solver,code", [("C1", "least_squares", 0), ("C2", "quasi_newton", 2), ("C3", "newton", 1),
                                              ("C3", "newton", 4)])
def test_baseline_configs(dh, oracle, name, solver, code):
    from nonlin_b200 import workloads as W

    w = W.WORKLOADS[name](1500)
    kw = {"max_fcn_evals": w["settings"]["set_max_fcn_evals"]} if "set_max_fcn_evals" in w["settings"] else {}
    got = run_dh_solve(dh, code, w, **kw)
    ref = oracle.solve_batch(solver, w["fcn"], w["x0"], m=w["m"], sys=w["args"], shared=w["shared"], params=oracle.params(**kw))
    same(got, ref)
    assert (ref[3] == 0).mean() > 0.99


@pytest.mark.parametrize("fcn", ["misc_2fcn", "poorly_scaled_2fcn", "powell_badly_scaled", "misc_2fcn_01"])
@pytest.mark.parametrize("solver,code", [("least_squares", 0), ("newton", 1), ("quasi_newton", 2)])
@pytest.mark.parametrize("settings", [{}, {"use_line_search": 0, "max_fcn_evals": 300}, {"use_analytic_jacobian": 1},
                                      {"jacobian_interval": 2, "fcn_tol": 1e-12}])
def test_square_systems_all_solvers_and_settings(dh, oracle, fcn, solver, code, settings):
    rng = np.random.default_rng(7)
    B = 300
    w = {"fcn": fcn, "m": 2, "n": 2, "x0": np.ascontiguousarray(rng.uniform(0.2, 3.0, size=(2, B))), "args": None, "shared": None}
    got = run_dh_solve(dh, code, w, **settings)
    ref = oracle.solve_batch(solver, fcn, w["x0"], params=oracle.params(**settings))
    same(got, ref)


def test_non_finite_starts(dh, oracle):
    rng = np.random.default_rng(2)
    B = 128
    x0 = rng.uniform(0.5, 1.5, size=(2, B))
    x0[0, ::9] = np.nan
    x0[1, 4::11] = np.inf
    w = {"fcn": "misc_2fcn", "m": 2, "n": 2, "x0": np.ascontiguousarray(x0), "args": None, "shared": None}
    for solver, code in (("least_squares", 0), ("newton", 1), ("quasi_newton", 2)):
        same(run_dh_solve(dh, code, w), oracle.solve_batch(solver, "misc_2fcn", w["x0"]))


@pytest.mark.parametrize("name", ["CLS1", "CLS2"])
def test_constrained_least_squares(dh, oracle, name):
    from nonlin_b200 import _lib
    from nonlin_b200 import workloads as W

    w = W.WORKLOADS[name](600)
    lo = np.array(w["settings"]["set_lower_limits"], dtype=np.float64)
    hi = np.array(w["settings"]["set_upper_limits"], dtype=np.float64)
    fid = _lib.load().nlb_vecfcn_lookup(w["fcn"].encode())
    B = w["x0"].shape[1]
    for lower, upper, radius in ((lo, hi, 1.0), (None, None, 0.05), (lo * 0.1 if name == "CLS2" else lo, hi * 0.5, 1.0)):
        x = w["x0"].copy(); f = np.zeros((w["m"], B)); ib = np.zeros((B, 7), dtype=np.int32); st = np.zeros(B, dtype=np.int32)
        p = engine_params(max_fcn_evals=150)
        lo8 = None if lower is None else np.concatenate([lower, np.full(8 - lower.size, -np.finfo(float).max)])
        hi8 = None if upper is None else np.concatenate([upper, np.full(8 - upper.size, np.finfo(float).max)])
        rc = dh.dh_cls_solve(fid, C.c_longlong(B), C.byref(p), C.c_double(radius), C.c_double(1.0), _p(lo8), _p(hi8), _p(x),
                             _p(f), _p(w["args"]), None, _p(ib), _p(st))
        assert rc == 0
        ref = oracle.cls_solve_batch(w["fcn"], w["x0"], m=w["m"], sys=w["args"], lower=lower, upper=upper,
                                     trust_region_radius=radius, params=oracle.params(max_fcn_evals=150))
        same((x, f, ib, st), ref)


@pytest.mark.parametrize("order,thru_zero", [(1, 0), (3, 0), (7, 0), (2, 1), (8, 1)])
@pytest.mark.parametrize("npts,shared", [(12, True), (40, False)])
def test_polynomial_fit(dh, oracle, order, thru_zero, npts, shared):
    rng = np.random.default_rng(order * 10 + npts)
    B = 200
    x = rng.uniform(0.1, 2.0, size=(npts,) if shared else (npts, B))
    y = rng.standard_normal((npts, B))
    y[:, 3] *= 2.0 ** 990          # DGELS scaling branches
    y[:, 4] *= 2.0 ** -1000
    y[:, 5] = 0.0
    c = np.zeros((order + 1, B)); st = np.zeros(B, dtype=np.int32)
    rc = dh.dh_polyfit(C.c_longlong(B), npts, order, thru_zero, int(shared), _p(np.ascontiguousarray(x)), _p(y), _p(c), _p(st))
    assert rc == 0
    co, sto = oracle.polyfit_batch(x, y, order, thru_zero=bool(thru_zero))
    assert np.array_equal(st, sto) and np.array_equal(c, co)


@pytest.mark.parametrize("solver,code", [("brent", 0), ("newton_1var", 1)])
@pytest.mark.parametrize("fcn", ["cubic_wallis", "exp_minus_x", "cubic_args"])
@pytest.mark.parametrize("analytic", [0, 1])
def test_one_variable_solvers(dh, oracle, solver, code, fcn, analytic):
    from nonlin_b200 import _lib

    lib = _lib.load()
    rng = np.random.default_rng(3)
    B = 500
    if fcn == "cubic_args":
        args = np.stack([rng.uniform(-8, -1, B), rng.uniform(-2, 2, B), rng.uniform(-1, 1, B), rng.uniform(0.5, 2, B)])
        lo, hi = rng.uniform(-2.0, -0.5, B), rng.uniform(4.0, 7.0, B)
    elif fcn == "cubic_wallis":
        args, lo, hi = None, rng.uniform(0.0, 2.0, B), rng.uniform(2.2, 4.0, B)
    else:
        args, lo, hi = None, rng.uniform(-1.0, 0.5, B), rng.uniform(0.6, 3.0, B)
    hi[::41] = lo[::41]
    x0 = rng.standard_normal(B)
    p = _lib.nlb_params_1var()
    lib.nlb_params_1var_default(C.byref(p))
    p.use_analytic_diff = analytic
    x = x0.copy(); f = np.zeros(B); ib = np.zeros((B, 7), dtype=np.int32); st = np.zeros(B, dtype=np.int32)
    rc = dh.dh_solve_1var(code, lib.nlb_fcn1var_lookup(fcn.encode()), C.c_longlong(B), C.byref(p), _p(lo), _p(hi), _p(x), _p(f),
                          _p(args), _p(ib), _p(st))
    assert rc == 0
    xo, fo, ibo, sto = oracle.solve_1var_batch(solver, fcn, lo, hi, x0=x0, args=args,
                                               params=oracle.params1(use_analytic_diff=analytic))
    same((x, f, ib, st), (xo, fo, ibo, sto))
