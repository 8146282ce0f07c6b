"""Host-side mirror of the reference's solver objects: same names, defaults, clamps and error
behaviour.  CPU only — no solve is launched."""
import numpy as np
import pytest

import nonlin_b200 as nb


def test_equation_solver_defaults_and_setters():
    for cls in (nb.least_squares_solver, nb.newton_solver, nb.quasi_newton_solver):
        s = cls()
        assert s.get_max_fcn_evals() == 100                      # src/nonlin_multi_eqn_mult_var.f90:69
        assert s.get_fcn_tolerance() == 1e-8                     # :71
        assert s.get_var_tolerance() == 1e-12                    # :73
        assert s.get_gradient_tolerance() == 1e-12               # :75
        assert s.get_print_status() is False                     # :77
        s.set_max_fcn_evals(1000); s.set_fcn_tolerance(1e-10); s.set_var_tolerance(1e-9); s.set_gradient_tolerance(1e-7)
        s.set_print_status(True)
        assert (s.get_max_fcn_evals(), s.get_fcn_tolerance(), s.get_var_tolerance(), s.get_gradient_tolerance()) == (1000, 1e-10, 1e-9, 1e-7)
        assert s.get_print_status() is True


def test_lm_step_scaling_factor_clamp():
    s = nb.least_squares_solver()
    assert s.get_step_scaling_factor() == 100.0                  # src/nonlin_least_squares.f90:25
    s.set_step_scaling_factor(0.01); assert s.get_step_scaling_factor() == 0.1      # :108-109
    s.set_step_scaling_factor(1e3); assert s.get_step_scaling_factor() == 100.0     # :110-111
    s.set_step_scaling_factor(5.0); assert s.get_step_scaling_factor() == 5.0


def test_constrained_least_squares_solver_settings():
    s = nb.constrained_least_squares_solver()
    assert isinstance(s, nb.constrained_equation_solver) and isinstance(s, nb.equation_solver)
    assert s.get_trust_region_radius() == 1.0                    # src/nonlin_least_squares.f90:64
    assert s.get_step_scaling_factor() == 1.0                    # :66
    s.set_trust_region_radius(0.0); assert s.get_trust_region_radius() == 1.0       # :904-905 (x <= 0 -> 1)
    s.set_trust_region_radius(-3.0); assert s.get_trust_region_radius() == 1.0
    s.set_trust_region_radius(2.5); assert s.get_trust_region_radius() == 2.5
    s.set_step_scaling_factor(-1.0); assert s.get_step_scaling_factor() == 1.0      # :929-930
    s.set_step_scaling_factor(0.5); assert s.get_step_scaling_factor() == 0.5
    assert s.get_lower_limits().size == 0 and s.get_upper_limits().size == 0        # :797-808 unallocated -> size 0
    s.set_lower_limits([4.0, 2.0]); s.set_upper_limits([5.6, 3.6])
    assert np.array_equal(s.get_lower_limits(), [4.0, 2.0]) and np.array_equal(s.get_upper_limits(), [5.6, 3.6])
    x = np.array([1.0, 9.0])
    assert np.array_equal(s.apply_limits(x), [4.0, 3.6])         # ces_apply_limits :858-883
    xb = np.array([[1.0, 5.0, 7.0], [2.5, -1.0, 3.7]])
    assert np.array_equal(s.apply_limits(xb), [[4.0, 5.0, 5.6], [2.5, 2.0, 3.6]])
    # limits of the wrong length are ignored by the solve (cls_solve :1014-1024 replaces them by +-huge)
    s.set_lower_limits([0.0, 0.0, 0.0])
    (opt,) = s._extra_args(2)
    assert opt._obj.lower is None and opt._obj.upper is not None
    assert opt._obj.trust_region_radius == 2.5 and opt._obj.step_scaling_factor == 0.5


def test_polynomial_object():
    p = nb.polynomial()
    assert p.order() == -1                                       # get_poly_order: unallocated -> -1 (:131-139)
    p.initialize(3, B=5)
    assert p.order() == 3 and p.get_all().shape == (4, 5) and np.all(p.get_all() == 0.0)
    p.set(2, 1.5)
    assert np.all(p.get(2) == 1.5) and np.all(p.get(1) == 0.0)  # 1-based, get(1) = c0
    q = nb.polynomial([6.0, 1.0, -4.0, 1.0])                     # polynomial(c) constructor, tests/nonlin_test_poly.f90:61
    assert q.order() == 3 and q.get_all().shape == (4, 1) and q.get(4)[0] == 1.0
    with pytest.raises(nb.NonlinError):
        nb.polynomial(-2)
    with pytest.raises(nb.NonlinError):                          # y must be (npts, B)
        p.fit(np.zeros(4), np.zeros(4), 2)
    with pytest.raises(nb.NonlinError):                          # size(x) /= size(y): error stop 3
        p.fit(np.zeros(5), np.zeros((4, 2)), 2)


def test_one_variable_solver_objects():
    for cls in (nb.brent_solver, nb.newton_1var_solver):
        s = cls()
        assert isinstance(s, nb.equation_solver_1var)
        assert s.get_max_fcn_evals() == 100                      # src/nonlin_single_var.f90:45
        assert s.get_fcn_tolerance() == 1e-8                     # :47
        assert s.get_var_tolerance() == 1e-12                    # :49
        assert s.get_diff_tolerance() == 1e-12                   # :51
        assert s.get_print_status() is False                     # :54
        s.set_max_fcn_evals(7); s.set_fcn_tolerance(1e-3); s.set_var_tolerance(1e-4); s.set_diff_tolerance(1e-5)
        assert (s.get_max_fcn_evals(), s.get_fcn_tolerance(), s.get_var_tolerance(), s.get_diff_tolerance()) == (7, 1e-3, 1e-4, 1e-5)
    obj = nb.fcn1var_helper()
    assert not obj.is_fcn_defined() and not obj.is_derivative_defined()
    obj.set_fcn("cubic_wallis")
    assert obj.is_fcn_defined() and not obj.is_derivative_defined()
    obj.set_diff()
    assert obj.is_derivative_defined()
    assert "sinx_div_x" in nb.fcn1var_names()
    lim = nb.value_pair(1.5, 5.0)                                # src/nonlin_types.f90:31-37
    assert (lim.x1, lim.x2) == (1.5, 5.0)
    with pytest.raises(nb.NonlinError):
        nb.brent_solver().solve(nb.fcn1var_helper(), np.zeros(3), lim)
    with pytest.raises(nb.NonlinError):
        nb.brent_solver().solve(obj, np.zeros((3, 2)), lim)


def test_quasi_newton_and_line_search_settings():
    s = nb.quasi_newton_solver()
    assert s.get_jacobian_interval() == 5                        # src/nonlin_solve.f90:51
    s.set_jacobian_interval(20); assert s.get_jacobian_interval() == 20
    assert s.get_use_line_search() is True and not s.is_line_search_defined()   # :30
    ls = nb.line_search()
    assert (ls.get_max_fcn_evals(), ls.get_scaling_factor(), ls.get_distance_factor()) == (100, 1e-4, 0.1)
    ls.set_distance_factor(-1.0); assert ls.get_distance_factor() == 0.1          # src/nonlin_linesearch.f90:142-143
    ls.set_distance_factor(2.0); assert ls.get_distance_factor() == 0.99          # :144-145
    ls.set_distance_factor(0.5); assert ls.get_distance_factor() == 0.5
    ls.set_max_fcn_evals(7); ls.set_scaling_factor(1e-3)
    s.set_line_search(ls)
    assert s.is_line_search_defined() and s.get_line_search() is not ls           # stored as a copy (:103-111)
    obj = nb.vecfcn_helper(); obj.set_fcn("misc_2fcn", 2, 2)
    p = s._params(obj)
    assert (p.jacobian_interval, p.ls_max_fcn_evals, p.ls_alpha, p.ls_factor, p.use_line_search) == (20, 7, 1e-3, 0.5, 1)
    s.set_use_line_search(False)
    assert s._params(obj).use_line_search == 0
    # solve lazily installs a default line search, like the reference (:229-233)
    s2 = nb.newton_solver()
    s2._params(obj)
    assert s2.is_line_search_defined()


def test_vecfcn_helper():
    obj = nb.vecfcn_helper()
    assert not obj.is_fcn_defined() and not obj.is_jacobian_defined()
    obj.set_fcn("lsq_poly_fit", 21, 4)
    assert obj.is_fcn_defined() and (obj.get_equation_count(), obj.get_variable_count()) == (21, 4)
    with pytest.raises(nb.NonlinError):
        obj.set_jacobian()                       # no analytic Jacobian registered for this residual
    with pytest.raises(nb.NonlinError):
        obj.set_fcn("lsq_poly_fit", 20, 4)       # size mismatch
    with pytest.raises(nb.NonlinError):
        obj.set_fcn("not_registered", 2, 2)
    obj.set_fcn("powell_badly_scaled", 2, 2); obj.set_jacobian()
    assert obj.is_jacobian_defined()
    obj.set_fcn("ext_rosenbrock", 64, 64)
    assert obj.get_variable_count() == 64
    obj.set_fcn("rational_7_8", 4096, 16)
    assert obj._info["sys_len"] == 4096 and obj._info["shared_len"] == 4096
    with pytest.raises(nb.NonlinError):
        obj.set_fcn("rational_7_8")              # run-time sized family needs m


def test_solve_argument_checks_happen_before_any_launch():
    obj = nb.vecfcn_helper()
    s = nb.newton_solver()
    with pytest.raises(nb.NonlinError):          # NL_UNDEFINED_FUNCTION_ERROR analogue
        s.solve(obj, np.ones((2, 3)))
    obj.set_fcn("misc_2fcn", 2, 2)
    with pytest.raises(nb.NonlinError):          # size(x) /= nvar -> error stop 3
        s.solve(obj, np.ones((3, 3)))
    with pytest.raises(nb.NonlinError):          # size(fvec) /= neqn -> error stop 4
        s.solve(obj, np.ones((2, 3)), np.ones((3, 3)))
    with pytest.raises(TypeError):
        s.solve(obj, np.ones((2, 3), dtype=np.float32))


def test_workloads_are_seeded_and_shaped():
    from nonlin_b200 import workloads as W

    for name, fn in W.WORKLOADS.items():
        kw = {"m": 128} if name == "C4" else {}
        a, b = fn(64, **kw), fn(64, **kw)
        assert np.array_equal(a["x0"], b["x0"]) and a["x0"].shape == (a["n"], 64)
        if a["args"] is not None:
            assert np.array_equal(a["args"], b["args"]) and a["args"].shape[1] == 64
        assert a["bytes_per_system"] > 0
    assert W.c1_lm_polyfit(1)["bytes_per_system"] == 432 and W.c2_broyden_2x2(1)["bytes_per_system"] == 80
    assert W.c5_broyden_rosenbrock(1)["bytes_per_system"] == 1568
