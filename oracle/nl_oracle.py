"""ctypes binding of the CPU oracle (oracle/libnl_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, by __graft_entry__.smoke() and by bench.py's
cpu_baseline / --impl reference legs.  Nothing under nonlin_b200/ imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

LM, NEWTON, BROYDEN = 0, 1, 2
SOLVERS = {"least_squares": LM, "lm": LM, "newton": NEWTON, "quasi_newton": BROYDEN, "broyden": BROYDEN}


class Params(C.Structure):
    _fields_ = [
        ("max_fcn_evals", C.c_int32),
        ("fcn_tol", C.c_double),
        ("var_tol", C.c_double),
        ("grad_tol", C.c_double),
        ("lm_factor", C.c_double),
        ("jacobian_interval", C.c_int32),
        ("use_line_search", C.c_int32),
        ("ls_max_fcn_evals", C.c_int32),
        ("ls_alpha", C.c_double),
        ("ls_factor", C.c_double),
        ("use_analytic_jacobian", C.c_int32),
        ("max_iter_guard", C.c_int32),
    ]


class Params1(C.Structure):
    """equation_solver_1var settings (brent_solver, newton_1var_solver)."""

    _fields_ = [
        ("max_fcn_evals", C.c_int32),
        ("fcn_tol", C.c_double),
        ("var_tol", C.c_double),
        ("diff_tol", C.c_double),
        ("use_analytic_diff", C.c_int32),
    ]


class ClsOptions(C.Structure):
    """constrained_least_squares_solver's own settings (radius, step scaling) and the limit arrays."""

    _fields_ = [
        ("trust_region_radius", C.c_double),
        ("step_scaling_factor", C.c_double),
        ("lower", C.c_void_p),
        ("upper", C.c_void_p),
    ]


IB_DTYPE = np.dtype(
    [
        ("iter_count", "<i4"),
        ("fcn_count", "<i4"),
        ("jacobian_count", "<i4"),
        ("gradient_count", "<i4"),
        ("converge_on_fcn", "<i4"),
        ("converge_on_chng", "<i4"),
        ("converge_on_zero_diff", "<i4"),
    ]
)


def build(force=False):
    """Compile the oracle with its Makefile (gcc only; no reference sources involved)."""
    so = os.path.join(_HERE, "libnl_oracle.so")
    if force or not os.path.exists(so) or not os.path.exists(os.path.join(_HERE, "libnl_oracle_count.so")):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return so


_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")


class Oracle:
    def __init__(self, counting=False):
        build()
        name = "libnl_oracle_count.so" if counting else "libnl_oracle.so"
        self.lib = lib = C.CDLL(os.path.join(_HERE, name))
        lib.nlo_fcn_lookup.argtypes = [C.c_char_p]
        lib.nlo_fcn_lookup.restype = C.c_int
        lib.nlo_params_default.argtypes = [C.POINTER(Params)]
        lib.nlo_soft_exp.argtypes = [C.c_double]
        lib.nlo_soft_exp.restype = C.c_double
        lib.nlo_norm2.argtypes = [_dp, C.c_int]
        lib.nlo_norm2.restype = C.c_double
        lib.nlo_dnrm2.argtypes = [_dp, C.c_int]
        lib.nlo_dnrm2.restype = C.c_double
        lib.nlo_flops_total.restype = C.c_ulonglong
        lib.nlo_solve.restype = C.c_int
        lib.nlo_solve_batch.restype = C.c_int
        lib.nlo_eval_fcn.restype = C.c_int
        lib.nlo_jacobian.restype = C.c_int
        lib.nlo_dgesv.restype = C.c_int
        lib.nlo_cls_solve.restype = C.c_int
        lib.nlo_polyfit_batch.restype = C.c_int
        lib.nlo_fcn1_lookup.argtypes = [C.c_char_p]
        lib.nlo_fcn1_lookup.restype = C.c_int
        lib.nlo_fcn1_eval.argtypes = [C.c_int, C.c_double, C.c_void_p]
        lib.nlo_fcn1_eval.restype = C.c_double
        lib.nlo_solve_1var_batch.restype = C.c_int
        lib.nlo_polyval_batch.restype = C.c_int
        lib.nlo_cls_solve_batch.restype = C.c_int

    # -- residuals supplied by a test as Python callbacks (checker for plug-in residuals) ---
    CALLBACK = C.CFUNCTYPE(None, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double),
                           C.c_int, C.c_int)
    _callbacks = {}

    def register_callback(self, name, m, n, fcn, sys_len=0, shared_len=0):
        """fcn(x, sys, shared) -> sequence of m floats, evaluated with Python floats (IEEE doubles, no contraction).
        Solve callback residuals with nthreads=1."""
        def tramp(xp, fp, sp, hp, mm, nn):
            x = [xp[j] for j in range(nn)]
            sysv = [sp[k] for k in range(sys_len)] if sys_len else None
            shv = [hp[k] for k in range(shared_len)] if shared_len else None
            out = fcn(x, sysv, shv)
            for i in range(mm):
                fp[i] = out[i]
        cb = Oracle.CALLBACK(tramp)
        self.lib.nlo_register_callback.restype = C.c_int
        self.lib.nlo_register_callback.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, Oracle.CALLBACK]
        fid = self.lib.nlo_register_callback(name.encode(), m, n, sys_len, shared_len, cb)
        if fid < 0:
            raise RuntimeError("nlo_register_callback failed")
        Oracle._callbacks[(id(self.lib), name)] = cb          # keep the trampoline alive
        return fid

    # -- registry -----------------------------------------------------------------------
    def fcn_id(self, name):
        i = self.lib.nlo_fcn_lookup(name.encode())
        if i < 0:
            raise KeyError(name)
        return i

    def fcn_info(self, fid):
        v = [C.c_int() for _ in range(5)]
        if self.lib.nlo_fcn_info(fid, *[C.byref(x) for x in v]) != 0:
            raise KeyError(fid)
        return dict(zip(["m", "n", "sys_len", "shared_len", "has_jac"], [x.value for x in v]))

    def params(self, **kw):
        p = Params()
        self.lib.nlo_params_default(C.byref(p))
        for k, v in kw.items():
            if not hasattr(p, k):
                raise AttributeError(k)
            setattr(p, k, v)
        return p

    def set_sum_mode(self, mode):
        """STUDY SWITCH (not the reference's arithmetic): 1 = long sums in warp-shuffle (striped + butterfly) order."""
        self.lib.nlo_set_sum_mode(int(mode))

    def set_libm_exp(self, on):
        self.lib.nlo_set_libm_exp(int(bool(on)))

    def flops_total(self):
        return int(self.lib.nlo_flops_total())

    def flops_reset(self):
        self.lib.nlo_flops_reset()

    # -- helpers ------------------------------------------------------------------------
    @staticmethod
    def _ptr(a):
        return a.ctypes.data_as(C.c_void_p) if a is not None else None

    def eval_fcn(self, fcn, x, m=0, sys=None, shared=None):
        fid = self.fcn_id(fcn) if isinstance(fcn, str) else fcn
        info = self.fcn_info(fid)
        x = np.ascontiguousarray(x, dtype=np.float64)
        n = info["n"] or x.size
        m = info["m"] or m or n
        f = np.empty(m)
        sys = None if sys is None else np.ascontiguousarray(sys, dtype=np.float64)
        shared = None if shared is None else np.ascontiguousarray(shared, dtype=np.float64)
        rc = self.lib.nlo_eval_fcn(fid, m, n, self._ptr(x), self._ptr(sys), self._ptr(shared), self._ptr(f))
        if rc:
            raise RuntimeError("nlo_eval_fcn -> %d" % rc)
        return f

    def jacobian(self, fcn, x, m=0, sys=None, shared=None, params=None):
        fid = self.fcn_id(fcn) if isinstance(fcn, str) else fcn
        info = self.fcn_info(fid)
        x = np.array(x, dtype=np.float64)
        n = info["n"] or x.size
        m = info["m"] or m or n
        jac = np.empty((n, m))  # column-major m x n
        p = params or self.params()
        sys = None if sys is None else np.ascontiguousarray(sys, dtype=np.float64)
        shared = None if shared is None else np.ascontiguousarray(shared, dtype=np.float64)
        rc = self.lib.nlo_jacobian(fid, m, n, C.byref(p), self._ptr(x), self._ptr(sys), self._ptr(shared), self._ptr(jac))
        if rc:
            raise RuntimeError("nlo_jacobian -> %d" % rc)
        return jac.T.copy()

    def solve(self, solver, fcn, x0, m=0, sys=None, shared=None, params=None):
        """One system. Returns (x, fvec, ib(dict), status)."""
        fid = self.fcn_id(fcn) if isinstance(fcn, str) else fcn
        info = self.fcn_info(fid)
        x = np.array(x0, dtype=np.float64)
        n = info["n"] or x.size
        m = info["m"] or m or n
        f = np.zeros(m)
        ib = np.zeros(1, dtype=IB_DTYPE)
        p = params or self.params()
        sys = None if sys is None else np.ascontiguousarray(sys, dtype=np.float64)
        shared = None if shared is None else np.ascontiguousarray(shared, dtype=np.float64)
        st = self.lib.nlo_solve(SOLVERS[solver] if isinstance(solver, str) else solver, fid, m, n, C.byref(p),
                                self._ptr(x), self._ptr(f), self._ptr(sys), self._ptr(shared), self._ptr(ib))
        return x, f, {k: int(ib[0][k]) for k in IB_DTYPE.names}, st

    def _cls_options(self, lower, upper, trust_region_radius, step_scaling_factor, n):
        o = ClsOptions()
        self.lib.nlo_cls_options_default(C.byref(o))
        if trust_region_radius is not None:
            o.trust_region_radius = trust_region_radius
        if step_scaling_factor is not None:
            o.step_scaling_factor = step_scaling_factor
        keep = []
        for name, v in (("lower", lower), ("upper", upper)):
            if v is not None:
                a = np.ascontiguousarray(v, dtype=np.float64)
                if a.size != n:
                    raise ValueError("%s must have n entries" % name)
                keep.append(a)
                setattr(o, name, a.ctypes.data)
        return o, keep

    def cls_solve(self, fcn, x0, m=0, sys=None, shared=None, params=None, lower=None, upper=None,
                  trust_region_radius=None, step_scaling_factor=None):
        """constrained_least_squares_solver, one system. Returns (x, fvec, ib(dict), status)."""
        fid = self.fcn_id(fcn) if isinstance(fcn, str) else fcn
        info = self.fcn_info(fid)
        x = np.array(x0, dtype=np.float64)
        n = info["n"] or x.size
        m = info["m"] or m or n
        f = np.zeros(m)
        ib = np.zeros(1, dtype=IB_DTYPE)
        p = params or self.params()
        o, keep = self._cls_options(lower, upper, trust_region_radius, step_scaling_factor, n)
        sys = None if sys is None else np.ascontiguousarray(sys, dtype=np.float64)
        shared = None if shared is None else np.ascontiguousarray(shared, dtype=np.float64)
        st = self.lib.nlo_cls_solve(fid, m, n, C.byref(p), C.byref(o), self._ptr(x), self._ptr(f), self._ptr(sys),
                                    self._ptr(shared), self._ptr(ib))
        return x, f, {k: int(ib[0][k]) for k in IB_DTYPE.names}, st

    def cls_solve_batch(self, fcn, x0, m=0, sys=None, shared=None, params=None, lower=None, upper=None,
                        trust_region_radius=None, step_scaling_factor=None, nthreads=0):
        """x0: (n, B) SoA. Returns (x (n,B), fvec (m,B), ib (B,) structured, status (B,))."""
        fid = self.fcn_id(fcn) if isinstance(fcn, str) else fcn
        info = self.fcn_info(fid)
        x = np.array(x0, dtype=np.float64, order="C")
        n, B = x.shape
        if info["n"] and info["n"] != n:
            raise ValueError("n mismatch")
        m = info["m"] or m or n
        f = np.zeros((m, B))
        ib = np.zeros(B, dtype=IB_DTYPE)
        status = np.zeros(B, dtype=np.int32)
        p = params or self.params()
        o, keep = self._cls_options(lower, upper, trust_region_radius, step_scaling_factor, n)
        sys = None if sys is None else np.ascontiguousarray(sys, dtype=np.float64)
        shared = None if shared is None else np.ascontiguousarray(shared, dtype=np.float64)
        rc = self.lib.nlo_cls_solve_batch(fid, C.c_long(B), m, n, C.byref(p), C.byref(o), self._ptr(x), self._ptr(f),
                                          self._ptr(sys), self._ptr(shared), self._ptr(ib), self._ptr(status),
                                          int(nthreads))
        if rc:
            raise RuntimeError("nlo_cls_solve_batch -> %d" % rc)
        return x, f, ib, status

    def params1(self, **kw):
        p = Params1()
        self.lib.nlo_params1_default(C.byref(p))
        for k, v in kw.items():
            if not hasattr(p, k):
                raise AttributeError(k)
            setattr(p, k, v)
        return p

    def fcn1_id(self, name):
        i = self.lib.nlo_fcn1_lookup(name.encode())
        if i < 0:
            raise KeyError(name)
        return i

    def fcn1_info(self, fid):
        a, d = C.c_int(), C.c_int()
        if self.lib.nlo_fcn1_info(fid, C.byref(a), C.byref(d)) != 0:
            raise KeyError(fid)
        return {"args_len": a.value, "has_diff": d.value}

    def solve_1var_batch(self, solver, fcn, lim1, lim2, x0=None, args=None, params=None, want_f=True, nthreads=0):
        """solver: "brent" | "newton_1var".  lim1, lim2: (B,) limits.  Returns (x, f or None, ib, status)."""
        fid = self.fcn1_id(fcn) if isinstance(fcn, str) else fcn
        lim1 = np.ascontiguousarray(lim1, dtype=np.float64)
        lim2 = np.ascontiguousarray(lim2, dtype=np.float64)
        B = lim1.size
        x = np.zeros(B) if x0 is None else np.array(x0, dtype=np.float64)
        f = np.zeros(B) if want_f else None
        ib = np.zeros(B, dtype=IB_DTYPE)
        status = np.zeros(B, dtype=np.int32)
        p = params or self.params1()
        args = None if args is None else np.ascontiguousarray(args, dtype=np.float64)
        rc = self.lib.nlo_solve_1var_batch({"brent": 0, "newton_1var": 1}[solver], fid, C.c_long(B), C.byref(p),
                                           self._ptr(lim1), self._ptr(lim2), self._ptr(x), self._ptr(f), self._ptr(args),
                                           self._ptr(ib), self._ptr(status), int(nthreads))
        if rc:
            raise RuntimeError("nlo_solve_1var_batch -> %d" % rc)
        return x, f, ib, status

    def polyfit_batch(self, x, y, order, thru_zero=False, nthreads=0):
        """polynomial%fit over B data sets. x: (npts,) shared or (npts, B); y: (npts, B).
        Returns (coeffs (order+1, B), status (B,))."""
        y = np.ascontiguousarray(y, dtype=np.float64)
        x = np.ascontiguousarray(x, dtype=np.float64)
        npts, B = y.shape
        shared = x.ndim == 1
        if x.shape[0] != npts or (not shared and x.shape != y.shape):
            raise ValueError("x / y shapes")
        c = np.zeros((order + 1, B))
        status = np.zeros(B, dtype=np.int32)
        rc = self.lib.nlo_polyfit_batch(C.c_long(B), npts, int(order), int(bool(thru_zero)), int(shared), self._ptr(x),
                                        self._ptr(y), self._ptr(c), self._ptr(status), int(nthreads))
        if rc:
            raise RuntimeError("nlo_polyfit_batch -> %d" % rc)
        return c, status

    def polyval_batch(self, coeffs, x):
        """polynomial%evaluate: coeffs (order+1, B), x (npts,) or (npts, B) -> (npts, B)."""
        coeffs = np.ascontiguousarray(coeffs, dtype=np.float64)
        x = np.ascontiguousarray(x, dtype=np.float64)
        B = coeffs.shape[1]
        y = np.zeros((x.shape[0], B))
        self.lib.nlo_polyval_batch(C.c_long(B), coeffs.shape[0] - 1, x.shape[0], int(x.ndim == 1), self._ptr(coeffs),
                                   self._ptr(x), self._ptr(y))
        return y

    def solve_batch(self, solver, fcn, x0, m=0, sys=None, shared=None, params=None, nthreads=0):
        """x0: (n, B) SoA. Returns (x (n,B), fvec (m,B), ib (B,) structured, status (B,))."""
        fid = self.fcn_id(fcn) if isinstance(fcn, str) else fcn
        info = self.fcn_info(fid)
        x = np.array(x0, dtype=np.float64, order="C")
        n, B = x.shape
        if info["n"] and info["n"] != n:
            raise ValueError("n mismatch")
        m = info["m"] or m or n
        f = np.zeros((m, B))
        ib = np.zeros(B, dtype=IB_DTYPE)
        status = np.zeros(B, dtype=np.int32)
        p = params or self.params()
        sys = None if sys is None else np.ascontiguousarray(sys, dtype=np.float64)
        shared = None if shared is None else np.ascontiguousarray(shared, dtype=np.float64)
        rc = self.lib.nlo_solve_batch(SOLVERS[solver] if isinstance(solver, str) else solver, fid, C.c_long(B), m, n,
                                      C.byref(p), self._ptr(x), self._ptr(f), self._ptr(sys), self._ptr(shared),
                                      self._ptr(ib), self._ptr(status), int(nthreads))
        if rc:
            raise RuntimeError("nlo_solve_batch -> %d" % rc)
        return x, f, ib, status
