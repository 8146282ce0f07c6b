"""The C-ABI library loads and exports every symbol include/nonlin_batch.h declares; struct
layouts, defaults and the residual registry agree with the header and with the oracle's
independent table.  No compute calls (CPU only)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = open(os.path.join(ROOT, "include", "nonlin_batch.h")).read()


def declared_functions():
    names = re.findall(r"^(?:int|void|const char\*|int64_t)\s+(nlb_\w+)\s*\(", HEADER, flags=re.M)
    assert len(names) >= 16
    return names


def test_library_exports_every_declared_symbol():
    from nonlin_b200 import _lib

    lib = C.CDLL(_lib.LIB_PATH)
    for name in declared_functions():
        assert hasattr(lib, name), "libnonlin_b200.so does not export %s" % name
    assert sorted(declared_functions()) == sorted(_lib.EXPORTS)


def test_every_entry_point_cites_the_reference():
    # each solver / helper entry point names the reference procedure and file:line it replaces
    for proc in ("lss_solve, src/nonlin_least_squares.f90:118-391", "ns_solve, src/nonlin_solve.f90:452-638",
                 "qns_solve, src/nonlin_solve.f90:156-425", "vfh_jac_fcn, src/nonlin_multi_eqn_mult_var.f90:198-277",
                 "vfh_fcn, src/nonlin_multi_eqn_mult_var.f90:178-195", "src/nonlin_types.f90:8-29",
                 "cls_solve, src/nonlin_least_squares.f90:938-1176", "poly_fit, src/nonlin_polynomials.f90:146-199",
                 "poly_eval_double, src/nonlin_polynomials.f90:256-283", "brent_solve, src/nonlin_solve.f90:643-835",
                 "newt1var_solve,\n * src/nonlin_solve.f90:840-1032"):
        assert proc in HEADER


def test_struct_layouts_and_defaults():
    from nonlin_b200 import _lib

    lib = _lib.load()
    assert _lib.IB_DTYPE.itemsize == 28                      # 4 x int32 + 3 x default LOGICAL
    p = _lib.nlb_params()
    lib.nlb_params_default(C.byref(p))
    # defaults of the reference's solver objects (SURVEY.md §5 "Config / flags")
    assert (p.max_fcn_evals, p.fcn_tol, p.var_tol, p.grad_tol) == (100, 1e-8, 1e-12, 1e-12)
    assert (p.lm_factor, p.jacobian_interval, p.use_line_search) == (100.0, 5, 1)
    assert (p.ls_max_fcn_evals, p.ls_alpha, p.ls_factor, p.use_analytic_jacobian) == (100, 1e-4, 0.1, 0)
    # status codes in the header and in the Python mirror agree
    for name, val in re.findall(r"#define NLB_(\w+_ERROR) (\d+)", HEADER):
        assert getattr(_lib, "NL_" + name) == int(val)


def test_constrained_options_struct(oracle):
    from nonlin_b200 import _lib
    from oracle.nl_oracle import ClsOptions

    lib = _lib.load()
    o = _lib.nlb_constrained_options()
    lib.nlb_constrained_options_default(C.byref(o))
    assert (o.trust_region_radius, o.step_scaling_factor, o.lower, o.upper) == (1.0, 1.0, None, None)
    assert [f[0] for f in ClsOptions._fields_] == [f[0] for f in _lib.nlb_constrained_options._fields_]
    assert C.sizeof(ClsOptions) == C.sizeof(_lib.nlb_constrained_options) == 32
    block = HEADER[HEADER.index("typedef struct nlb_constrained_options {"):HEADER.index("} nlb_constrained_options;")]
    pos = [block.index(f[0]) for f in _lib.nlb_constrained_options._fields_]
    assert pos == sorted(pos)


def test_params_1var_struct_and_registry(oracle):
    from nonlin_b200 import _lib
    from oracle.nl_oracle import Params1

    lib = _lib.load()
    p = _lib.nlb_params_1var()
    lib.nlb_params_1var_default(C.byref(p))
    # defaults of equation_solver_1var (src/nonlin_single_var.f90:44-54)
    assert (p.max_fcn_evals, p.fcn_tol, p.var_tol, p.diff_tol, p.use_analytic_diff) == (100, 1e-8, 1e-12, 1e-12, 0)
    assert [f[0] for f in Params1._fields_] == [f[0] for f in _lib.nlb_params_1var._fields_]
    assert C.sizeof(Params1) == C.sizeof(_lib.nlb_params_1var)
    for i in range(lib.nlb_fcn1var_count()):
        name = lib.nlb_fcn1var_name(i).decode()
        assert oracle.fcn1_id(name) == i and lib.nlb_fcn1var_lookup(name.encode()) == i
        a, d = C.c_int(), C.c_int()
        assert lib.nlb_fcn1var_info(i, C.byref(a), C.byref(d)) == 0
        assert oracle.fcn1_info(i) == {"args_len": a.value, "has_diff": d.value}
    assert lib.nlb_fcn1var_lookup(b"nope") == -1 and lib.nlb_fcn1var_name(99) is None


def test_params_struct_matches_oracle_struct(oracle):
    from nonlin_b200 import _lib
    from oracle.nl_oracle import Params

    assert [f[0] for f in Params._fields_] == [f[0] for f in _lib.nlb_params._fields_]
    assert C.sizeof(Params) == C.sizeof(_lib.nlb_params)


def test_registry_matches_oracle_table(oracle):
    """Two independently written tables (CUDA registry, oracle problems) describe the same residuals."""
    from nonlin_b200 import _lib

    lib = _lib.load()
    BUILTIN = 12                          # residuals compiled into the engine; plug-ins loaded by other tests come after
    assert lib.nlb_vecfcn_count() >= BUILTIN
    for fid in range(BUILTIN):
        name = lib.nlb_vecfcn_name(fid).decode()
        assert lib.nlb_vecfcn_lookup(name.encode()) == fid
        assert oracle.fcn_id(name) == fid
        vals = [C.c_int() for _ in range(5)]
        assert lib.nlb_vecfcn_info(fid, *[C.byref(v) for v in vals]) == 0
        info = oracle.fcn_info(fid)
        assert [v.value for v in vals] == [info["m"], info["n"], info["sys_len"], info["shared_len"], info["has_jac"]]
    assert lib.nlb_vecfcn_lookup(b"no_such_fcn") == -1
    assert lib.nlb_vecfcn_name(99) is None


def test_no_cpu_fallback_without_a_device():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import nonlin_b200 as nb

    with pytest.raises(nb.NonlinError) as e:
        nb.Engine(0)
    assert e.value.code == nb.NLB_ERR_NO_DEVICE
    obj = nb.vecfcn_helper(); obj.set_fcn("misc_2fcn", 2, 2)
    with pytest.raises(nb.NonlinError):
        nb.quasi_newton_solver().solve(obj, np.ones((2, 4)))


def test_product_does_not_touch_the_oracle():
    """Nothing under nonlin_b200/ or include/ may import, include or link anything under oracle/."""
    for base in ("nonlin_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            for fn in files:
                if fn.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".f90")):
                    src = open(os.path.join(dirpath, fn)).read()
                    assert "nl_oracle" not in src and "oracle/" not in src and "import oracle" not in src, fn


def _build_cxx_example(tmp_path):
    import subprocess

    exe = str(tmp_path / "cxx_host_example")
    libdir = os.path.join(ROOT, "nonlin_b200")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.check_call([cxx, "-std=c++17", "-O2", "-o", exe, os.path.join(ROOT, "tests", "cxx_host_example.cpp"),
                           "-L" + libdir, "-lnonlin_b200", "-Wl,-rpath," + libdir])
    return exe


def test_cxx_host_mirror_links_against_the_c_abi(tmp_path):
    """The header-only C++ mirror of the reference's solver objects compiles and links against the
    C ABI with a plain host compiler (no nvcc, no torch)."""
    import subprocess

    import torch

    exe = _build_cxx_example(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True)
    if torch.cuda.is_available():
        assert r.returncode == 0, r.stdout
    else:
        assert r.returncode == 2 and "no usable CUDA device" in r.stdout


def test_fortran_binding_declares_every_solver_entry_point():
    """fortran/nonlin_batch.f90 cannot be compiled here (no Fortran compiler); check at least that its
    bind(C) names exist in the library and that the interoperable types have the header's field order."""
    from nonlin_b200 import _lib

    src = open(os.path.join(ROOT, "fortran", "nonlin_batch.f90")).read()
    names = set(re.findall(r'bind\(C, name = "(nlb_\w+)"\)', src))
    assert {"nlb_create", "nlb_destroy", "nlb_least_squares_solve_batch", "nlb_newton_solve_batch",
            "nlb_quasi_newton_solve_batch", "nlb_constrained_least_squares_solve_batch", "nlb_jacobian_batch",
            "nlb_reduce_stats"} <= names
    lib = C.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(lib, n)
    fields = [f[0] for f in _lib.nlb_params._fields_]
    block = src[src.index("type, bind(C) :: nlb_params"):src.index("end type")]
    pos = [block.index(f) for f in fields]
    assert pos == sorted(pos)


def test_header_is_valid_c99_and_cxx():
    """The boundary is a C ABI: include/nonlin_batch.h must compile as plain C as well as C++."""
    import shutil
    import subprocess

    hdr = os.path.join(ROOT, "include", "nonlin_batch.h")
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else shutil.which("gcc")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else shutil.which("g++")
    subprocess.check_call([cc, "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-x", "c", hdr])
    subprocess.check_call([cxx, "-std=c++17", "-Wall", "-Werror", "-fsyntax-only", "-x", "c++", hdr])
