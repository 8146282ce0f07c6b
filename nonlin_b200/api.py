"""Host-side mirror of the reference's solver interface for the batched hot path.

Same type names, setters and argument meaning as jchristopherson/nonlin (Appendix C of
SURVEY.md), with a batch `solve` that forwards to the C ABI in include/nonlin_batch.h:

    reference (Fortran)                               here
    -----------------------------------------------   -------------------------------------------
    type(vecfcn_helper) :: obj                        obj = vecfcn_helper()
    call obj%set_fcn(fcn, m, n)                       obj.set_fcn("misc_2fcn", m, n)   # registered name
    call obj%set_jacobian(jac)                        obj.set_jacobian()               # registered Jacobian
    type(quasi_newton_solver) :: solver               solver = quasi_newton_solver()
    call solver%set_fcn_tolerance(1d-8) ...           solver.set_fcn_tolerance(1e-8) ...
    call solver%solve(obj, x, f, ib, args)            status = solver.solve(obj, x, f, ib, args=...)

`x` is (n, B), `f` is (m, B), args is (sys_len, B): system index fastest (a Fortran x(B, n)).
Arrays are numpy (host; staged by the engine) or torch CUDA tensors (used in place,
asynchronously on torch's current stream).  Where the reference executes `error stop code`,
the per-system `status[b]` holds that code instead.

This module contains no numerical code: everything is computed by libnonlin_b200.so on the GPU.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import IB_DTYPE, NLB_STAT_COUNT, NLB_STAT_NAMES, nlb_params

_LIB = _lib.load()


class NonlinError(RuntimeError):
    """API-level failure (bad sizes, unknown residual, CUDA error, no device)."""

    def __init__(self, code, msg):
        super().__init__("nonlin_b200 error %d: %s" % (code, msg))
        self.code = code


# ---------------------------------------------------------------------------------------------
# engine handle (one per GPU)
# ---------------------------------------------------------------------------------------------
class Engine:
    """Owns an nlb_handle (stream + staging workspace) on one device."""

    def __init__(self, device=0):
        self._h = C.c_void_p()
        rc = _LIB.nlb_create(C.byref(self._h), int(device))
        if rc != _lib.NLB_OK:
            self._h = None
            raise NonlinError(rc, "nlb_create(device=%d) failed (no usable CUDA device? the engine has no CPU fallback)" % device)
        self.device = int(device)

    def close(self):
        if getattr(self, "_h", None):
            _LIB.nlb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc):
        if rc != _lib.NLB_OK:
            raise NonlinError(rc, _LIB.nlb_last_error(self._h).decode())

    @property
    def kernel_launches(self):
        return int(_LIB.nlb_kernel_launch_count(self._h))

    def measure_fp64_peak(self):
        a, b = C.c_double(), C.c_double()
        self.check(_LIB.nlb_measure_fp64_peak(self._h, C.byref(a), C.byref(b)))
        return {"dfma_tflops": a.value, "dadd_dmul_tflops": b.value}

    def measure_fp64_latency(self):
        """Cycles per dependent DADD / division / sqrt / shared-load+DADD (the serial-chain floors)."""
        v = (C.c_double * 4)()
        self.check(_LIB.nlb_measure_fp64_latency(self._h, v))
        return dict(zip(["dadd", "ddiv", "dsqrt", "lds_dadd"], [float(x) for x in v]))

    def reduce_stats(self, ib, status, B=None):
        """Batch convergence statistics -> dict (host)."""
        out = np.zeros(NLB_STAT_COUNT, dtype=np.int64)
        B = int(B if B is not None else _length(status if status is not None else ib))
        self.check(_LIB.nlb_reduce_stats(self._h, B, _ptr(ib), _ptr(status), _ptr(out), _stream_of(ib, status)))
        return {k: int(out[i]) for i, k in enumerate(NLB_STAT_NAMES)}

    def reduce_stats_device(self, ib, status, out, B):
        """Same, into a device int64[16] tensor, asynchronously (for the NCCL all-reduce)."""
        self.check(_LIB.nlb_reduce_stats(self._h, int(B), _ptr(ib), _ptr(status), _ptr(out), _stream_of(ib, status, out)))


_default_engines = {}


def load_plugin(path):
    """Load a residual plug-in library (built from nonlin_b200/csrc/nlb_plugin.cuh; see examples/plugin/) and register the
    residuals it carries - the batch analogue of handing vecfcn_helper%set_fcn a new procedure.  Returns how many."""
    n = _LIB.nlb_load_plugin(str(path).encode())
    if n < 0:
        raise NonlinError(_lib.NLB_ERR_INVALID_ARGUMENT, "cannot load residual plug-in %s" % path)
    return n


def default_engine(device=0):
    e = _default_engines.get(device)
    if e is None:
        e = _default_engines[device] = Engine(device)
    return e


# ---------------------------------------------------------------------------------------------
# pointer plumbing: numpy arrays and torch tensors
# ---------------------------------------------------------------------------------------------
def _is_torch(a):
    return type(a).__module__.startswith("torch")


def _ptr(a):
    if a is None:
        return None
    if _is_torch(a):
        if not a.is_contiguous():
            raise ValueError("tensors passed to the engine must be contiguous")
        return C.c_void_p(a.data_ptr())
    if not a.flags["C_CONTIGUOUS"]:
        raise ValueError("arrays passed to the engine must be C-contiguous")
    return C.c_void_p(a.ctypes.data)


def _length(a):
    return a.shape[0]


def _stream_of(*arrays):
    for a in arrays:
        if a is not None and _is_torch(a) and a.is_cuda:
            import torch

            # torch's default stream is the legacy NULL stream; NULL means "the handle's own stream" in
            # the C ABI, so name it explicitly (cudaStreamLegacy == (cudaStream_t)0x1)
            return C.c_void_p(torch.cuda.current_stream(a.device).cuda_stream or 1)
    return None


def _device_of(*arrays):
    """Device of the first CUDA tensor among the arguments; for host-only calls, torch's current
    device when torch is already in use (one process per GPU under torchrun), else device 0."""
    for a in arrays:
        if a is not None and _is_torch(a) and a.is_cuda:
            return a.device.index or 0
    import sys

    torch = sys.modules.get("torch")
    if torch is not None and torch.cuda.is_available() and torch.cuda.is_initialized():
        return torch.cuda.current_device()
    # one process per GPU (torchrun): host-array calls go to this rank's GPU, not to device 0
    import os

    lr = os.environ.get("LOCAL_RANK")
    if lr is not None and lr.isdigit():
        return int(lr)
    return None


def _check_i32(name, a, B, width=None):
    """`status` (int32, (B,)) and `ib` (IB_DTYPE (B,) or int32 (B, 7)) are written by the engine in full: a short or
    narrow array would be overrun silently, so refuse it (NLB_ERR_SIZE / TypeError) before the call."""
    if a is None:
        return
    if _is_torch(a):
        import torch

        ok = a.dtype == torch.int32
        shape_ok = tuple(a.shape) == ((B,) if width is None else (B, width))
    elif width is not None and a.dtype == IB_DTYPE:
        ok, shape_ok = True, tuple(a.shape) == (B,)
    else:
        ok = a.dtype == np.int32
        shape_ok = tuple(a.shape) == ((B,) if width is None else (B, width))
    if not ok:
        raise TypeError("%s must be int32%s" % (name, " (or the iteration_behavior record dtype)" if width else ""))
    if not shape_ok:
        raise NonlinError(_lib.NLB_ERR_SIZE, "%s has shape %s for a batch of %d" % (name, tuple(a.shape), B))


def _check_f64(name, a, shape):
    if _is_torch(a):
        import torch

        ok = a.dtype == torch.float64
    else:
        ok = a.dtype == np.float64
    if not ok:
        raise TypeError("%s must be float64" % name)
    if tuple(a.shape) != tuple(shape):
        # mirrors the reference's size checks (`error stop flag`, src/nonlin_least_squares.f90:190-196)
        raise NonlinError(_lib.NLB_ERR_SIZE, "%s has shape %s, expected %s" % (name, tuple(a.shape), tuple(shape)))


# ---------------------------------------------------------------------------------------------
# reference types
# ---------------------------------------------------------------------------------------------
def iteration_behavior(B, like=None):
    """Array of B `iteration_behavior` records (reference src/nonlin_types.f90:8-29)."""
    if like is not None and _is_torch(like) and like.is_cuda:
        import torch

        return torch.zeros((B, 7), dtype=torch.int32, device=like.device)
    return np.zeros(B, dtype=IB_DTYPE)


def ib_view(ib):
    """Structured numpy view of an iteration_behavior array (copies a CUDA tensor to host)."""
    if _is_torch(ib):
        return ib.detach().cpu().numpy().view(IB_DTYPE).reshape(-1)
    return ib


class vecfcn_helper:
    """Problem definition (reference src/nonlin_multi_eqn_mult_var.f90:41-65).

    set_fcn takes the *name* of a registered __device__ residual instead of a procedure
    pointer; set_jacobian() selects its registered analytic Jacobian (the reference's
    vfh_set_jac), otherwise forward differences are used (vfh_jac_fcn :244-276).
    """

    def __init__(self):
        self._fcn_id = -1
        self._m = 0
        self._n = 0
        self._use_jac = False
        self._shared = None
        self._info = None

    def set_fcn(self, fcn, nfcn=0, nvar=0):
        fid = _LIB.nlb_vecfcn_lookup(fcn.encode()) if isinstance(fcn, str) else int(fcn)
        vals = [C.c_int() for _ in range(5)]
        if fid < 0 or _LIB.nlb_vecfcn_info(fid, *[C.byref(v) for v in vals]) != 0:
            raise NonlinError(_lib.NLB_ERR_UNKNOWN_FCN, "residual %r is not registered" % (fcn,))
        m, n, sys_len, shared_len, has_jac = [v.value for v in vals]
        if (m and nfcn and m != nfcn) or (n and nvar and n != nvar):
            raise NonlinError(_lib.NLB_ERR_SIZE, "%r is registered as %dx%d, not %dx%d" % (fcn, m, n, nfcn, nvar))
        self._fcn_id = fid
        self._m = m or int(nfcn)
        self._n = n or int(nvar)
        if fid == _LIB.nlb_vecfcn_lookup(b"ext_rosenbrock") and not self._m:
            self._m = self._n
        if self._m <= 0 or self._n <= 0:
            raise NonlinError(_lib.NLB_ERR_SIZE, "this residual family needs explicit nfcn / nvar")
        self._info = {
            "sys_len": self._m if sys_len < 0 else sys_len,
            "shared_len": self._m if shared_len < 0 else shared_len,
            "has_jac": bool(has_jac),
        }
        self._use_jac = False

    def set_jacobian(self, enable=True):
        if enable and not (self._info and self._info["has_jac"]):
            raise NonlinError(_lib.NLB_ERR_UNSUPPORTED, "no analytic Jacobian is registered for this residual")
        self._use_jac = bool(enable)

    def set_shared_data(self, shared):
        """Data every system of the batch sees (e.g. the abscissae of a curve fit)."""
        self._shared = shared

    def is_fcn_defined(self):
        return self._fcn_id >= 0

    def is_jacobian_defined(self):
        return self._use_jac

    def get_equation_count(self):
        return self._m

    def get_variable_count(self):
        return self._n

    # vecfcn_helper%fcn over a batch
    def fcn(self, x, f=None, args=None, engine=None):
        eng = engine or default_engine(_device_of(x, f, args) or 0)
        B = x.shape[1]
        _check_f64("x", x, (self._n, B))
        if f is None:
            f = _empty_like(x, (self._m, B))
        _check_f64("f", f, (self._m, B))
        eng.check(_LIB.nlb_vecfcn_eval_batch(eng._h, self._fcn_id, B, self._m, self._n, _ptr(x), _ptr(f), _ptr(args),
                                             _ptr(self._shared), _stream_of(x, f, args)))
        return f

    # vecfcn_helper%jacobian over a batch: (n, m, B) = column-major m x n per system
    def jacobian(self, x, jac=None, args=None, engine=None):
        eng = engine or default_engine(_device_of(x, jac, args) or 0)
        B = x.shape[1]
        _check_f64("x", x, (self._n, B))
        if jac is None:
            jac = _empty_like(x, (self._n, self._m, B))
        _check_f64("jac", jac, (self._n, self._m, B))
        p = nlb_params()
        _LIB.nlb_params_default(C.byref(p))
        p.use_analytic_jacobian = int(self._use_jac)
        eng.check(_LIB.nlb_jacobian_batch(eng._h, C.byref(p), self._fcn_id, B, self._m, self._n, _ptr(x), _ptr(jac),
                                          _ptr(args), _ptr(self._shared), _stream_of(x, jac, args)))
        return jac


def _empty_like(a, shape, dtype="float64"):
    if _is_torch(a):
        import torch

        return torch.empty(shape, dtype=getattr(torch, dtype), device=a.device)
    return np.empty(shape, dtype=dtype)


class line_search:
    """Backtracking line-search settings (reference src/nonlin_linesearch.f90:18-65)."""

    def __init__(self):
        self._max_eval = 100
        self._alpha = 1.0e-4
        self._factor = 0.1

    def get_max_fcn_evals(self):
        return self._max_eval

    def set_max_fcn_evals(self, x):
        self._max_eval = int(x)

    def get_scaling_factor(self):
        return self._alpha

    def set_scaling_factor(self, x):
        self._alpha = float(x)

    def get_distance_factor(self):
        return self._factor

    def set_distance_factor(self, x):
        # src/nonlin_linesearch.f90:142-148
        x = float(x)
        if x <= 0.0:
            self._factor = 0.1
        elif x >= 1.0:
            self._factor = 0.99
        else:
            self._factor = x


class equation_solver:
    """Base class with the tolerances (reference src/nonlin_multi_eqn_mult_var.f90:67-91)."""

    _entry = None

    def __init__(self, engine=None):
        self._engine = engine
        self._max_eval = 100
        self._fcn_tol = 1.0e-8
        self._xtol = 1.0e-12
        self._gtol = 1.0e-12
        self._print_status = False

    def get_max_fcn_evals(self):
        return self._max_eval

    def set_max_fcn_evals(self, n):
        self._max_eval = int(n)

    def get_fcn_tolerance(self):
        return self._fcn_tol

    def set_fcn_tolerance(self, x):
        self._fcn_tol = float(x)

    def get_var_tolerance(self):
        return self._xtol

    def set_var_tolerance(self, x):
        self._xtol = float(x)

    def get_gradient_tolerance(self):
        return self._gtol

    def set_gradient_tolerance(self, x):
        self._gtol = float(x)

    def get_print_status(self):
        return self._print_status

    def set_print_status(self, x):
        # Per-iteration printing needs a host round trip per step; the batch engine keeps every
        # iteration on the device, so the flag is stored but nothing is printed.
        self._print_status = bool(x)

    def _params(self, fcn):
        p = nlb_params()
        _LIB.nlb_params_default(C.byref(p))
        p.max_fcn_evals = self._max_eval
        p.fcn_tol = self._fcn_tol
        p.var_tol = self._xtol
        p.grad_tol = self._gtol
        p.use_analytic_jacobian = int(fcn.is_jacobian_defined())
        return p

    def solve(self, fcn, x, fvec=None, ib=None, args=None, status=None, stream=None, want_fvec=True):
        """Solve the B systems in x (n, B) in place. Returns the per-system status array.

        want_fvec=False (with fvec=None): the residuals are not returned - for host buffers the call is PCIe-bound
        and fvec is most of what travels back.

        `stream` (a raw cudaStream_t integer) overrides the stream choice: by default CUDA tensors
        run on torch's current stream and host arrays on the engine's own stream.

        Mirrors `call solver%solve(fcn, x, fvec, ib, args)` (nonlin_solver interface,
        src/nonlin_multi_eqn_mult_var.f90:94-119).
        """
        if not fcn.is_fcn_defined():
            raise NonlinError(_lib.NLB_ERR_UNKNOWN_FCN, "no residual set (NL_UNDEFINED_FUNCTION_ERROR)")
        m, n = fcn.get_equation_count(), fcn.get_variable_count()
        if x.ndim != 2:
            raise NonlinError(_lib.NLB_ERR_SIZE, "x must be (n, B)")
        B = x.shape[1]
        _check_f64("x", x, (n, B))
        if fvec is None and want_fvec:
            fvec = _empty_like(x, (m, B))
        if fvec is not None:
            _check_f64("fvec", fvec, (m, B))
        if args is not None:
            _check_f64("args", args, (fcn._info["sys_len"], B))
        if status is None:
            if _is_torch(x) and x.is_cuda:
                import torch

                status = torch.zeros(B, dtype=torch.int32, device=x.device)
            else:
                status = np.zeros(B, dtype=np.int32)
        _check_i32("status", status, B)
        _check_i32("ib", ib, B, 7)
        eng = self._engine or default_engine(_device_of(x, fvec, args, ib, status) or 0)
        p = self._params(fcn)
        entry = getattr(_LIB, self._entry)
        eng.check(entry(eng._h, C.byref(p), *self._extra_args(n), fcn._fcn_id, B, m, n, _ptr(x), _ptr(fvec), _ptr(args),
                        _ptr(fcn._shared), _ptr(ib), _ptr(status),
                        C.c_void_p(stream) if stream is not None else _stream_of(x, fvec, args, ib, status)))
        self.last_fvec = fvec
        return status

    def _extra_args(self, n):
        """Solver-specific arguments that follow `params` in the C entry point (none for most solvers)."""
        return ()

    _sharded_kind = None

    def solve_sharded(self, engines, fcn, x, fvec=None, ib=None, args=None, status=None, want_stats=True):
        """One HOST batch over several GPUs of this process (nlb_solve_sharded): contiguous system ranges, one host
        thread and one engine handle per device, no data-path collective; the convergence statistics are combined with
        one NCCL all-reduce.  Returns (status, stats dict or None)."""
        if self._sharded_kind is None:
            raise NonlinError(_lib.NLB_ERR_UNSUPPORTED, "no sharded entry point for this solver")
        if not fcn.is_fcn_defined():
            raise NonlinError(_lib.NLB_ERR_UNKNOWN_FCN, "no residual set (NL_UNDEFINED_FUNCTION_ERROR)")
        m, n = fcn.get_equation_count(), fcn.get_variable_count()
        B = x.shape[1]
        for name, a, shape in (("x", x, (n, B)), ("fvec", fvec, (m, B)), ("args", args, (fcn._info["sys_len"], B))):
            if a is not None:
                if _is_torch(a) and a.is_cuda:
                    raise NonlinError(_lib.NLB_ERR_INVALID_ARGUMENT, "a sharded solve takes host buffers")
                _check_f64(name, a, shape)
        if status is None:
            status = np.zeros(B, dtype=np.int32)
        _check_i32("status", status, B)
        _check_i32("ib", ib, B, 7)
        hs = (C.c_void_p * len(engines))(*[e._h for e in engines])
        stats = np.zeros(NLB_STAT_COUNT, dtype=np.int64) if want_stats else None
        p = self._params(fcn)
        rc = _LIB.nlb_solve_sharded(hs, len(engines), self._sharded_kind, C.byref(p), fcn._fcn_id, B, m, n, _ptr(x), _ptr(fvec),
                                    _ptr(args), _ptr(fcn._shared), _ptr(ib), _ptr(status), _ptr(stats))
        engines[0].check(rc)
        return status, (None if stats is None else {k: int(stats[i]) for i, k in enumerate(NLB_STAT_NAMES)})


class least_squares_solver(equation_solver):
    """Levenberg-Marquardt (reference src/nonlin_least_squares.f90:20-31, lss_solve :118-391)."""

    _entry = "nlb_least_squares_solve_batch"
    _sharded_kind = 0      # NLB_SOLVER_LEAST_SQUARES

    def __init__(self, engine=None):
        super().__init__(engine)
        self._factor = 100.0

    def get_step_scaling_factor(self):
        return self._factor

    def set_step_scaling_factor(self, x):
        # clamp to [0.1, 100], src/nonlin_least_squares.f90:108-114
        x = float(x)
        self._factor = 0.1 if x < 0.1 else (100.0 if x > 100.0 else x)

    def _params(self, fcn):
        p = super()._params(fcn)
        p.lm_factor = self._factor
        return p


class constrained_equation_solver(equation_solver):
    """Adds the variable limits (reference src/nonlin_least_squares.f90:34-48, ces_* :794-883)."""

    def __init__(self, engine=None):
        super().__init__(engine)
        self._upper = None
        self._lower = None

    def get_upper_limits(self):
        return np.empty(0) if self._upper is None else self._upper.copy()

    def set_upper_limits(self, x):
        self._upper = np.array(x, dtype=np.float64).ravel()

    def get_lower_limits(self):
        return np.empty(0) if self._lower is None else self._lower.copy()

    def set_lower_limits(self, x):
        self._lower = np.array(x, dtype=np.float64).ravel()

    def apply_limits(self, x):
        """Clamp x (n,) or (n, B) host array into the limits, lower first (ces_apply_limits :858-883)."""
        if self._lower is not None:
            k = min(x.shape[0], self._lower.size)
            lo = self._lower[:k].reshape((k,) + (1,) * (x.ndim - 1))
            np.copyto(x[:k], lo, where=x[:k] < lo)
        if self._upper is not None:
            k = min(x.shape[0], self._upper.size)
            hi = self._upper[:k].reshape((k,) + (1,) * (x.ndim - 1))
            np.copyto(x[:k], hi, where=x[:k] > hi)
        return x


class constrained_least_squares_solver(constrained_equation_solver):
    """Bounded trust-region dogleg (reference src/nonlin_least_squares.f90:50-75, cls_solve :938-1176)."""

    _entry = "nlb_constrained_least_squares_solve_batch"
    _sharded_kind = None

    def __init__(self, engine=None):
        super().__init__(engine)
        self._delta = 1.0
        self._scaling = 1.0

    def get_trust_region_radius(self):
        return self._delta

    def set_trust_region_radius(self, x):
        # non-positive -> 1, src/nonlin_least_squares.f90:898-910
        x = float(x)
        self._delta = 1.0 if x <= 0.0 else x

    def get_step_scaling_factor(self):
        return self._scaling

    def set_step_scaling_factor(self, x):
        # non-positive -> 1, src/nonlin_least_squares.f90:923-935
        x = float(x)
        self._scaling = 1.0 if x <= 0.0 else x

    def _extra_args(self, n):
        o = _lib.nlb_constrained_options()
        _LIB.nlb_constrained_options_default(C.byref(o))
        o.trust_region_radius = self._delta
        o.step_scaling_factor = self._scaling
        # a limit array of the wrong length is replaced by -huge / +huge, as cls_solve does (:1014-1024)
        if self._lower is not None and self._lower.size == n:
            o.lower = self._lower.ctypes.data
        if self._upper is not None and self._upper.size == n:
            o.upper = self._upper.ctypes.data
        self._options = o            # keeps the struct (and through self the arrays) alive across the call
        return (C.byref(o),)


class line_search_solver(equation_solver):
    """Base of the line-searched solvers (reference src/nonlin_solve.f90:20-41)."""

    def __init__(self, engine=None):
        super().__init__(engine)
        self._line_search = None
        self._use_line_search = True

    def get_line_search(self):
        return self._line_search

    def set_line_search(self, ls):
        c = line_search()
        c._max_eval, c._alpha, c._factor = ls._max_eval, ls._alpha, ls._factor
        self._line_search = c

    def set_default_line_search(self):
        self.set_line_search(line_search())

    def is_line_search_defined(self):
        return self._line_search is not None

    def get_use_line_search(self):
        return self._use_line_search

    def set_use_line_search(self, x):
        self._use_line_search = bool(x)

    def _params(self, fcn):
        p = super()._params(fcn)
        p.use_line_search = int(self._use_line_search)
        # the reference lazily installs a default line search inside solve (src/nonlin_solve.f90:229-233)
        if self._use_line_search and self._line_search is None:
            self.set_default_line_search()
        if self._line_search is not None:
            p.ls_max_fcn_evals = self._line_search._max_eval
            p.ls_alpha = self._line_search._alpha
            p.ls_factor = self._line_search._factor
        return p


class quasi_newton_solver(line_search_solver):
    """Broyden's method with QR rank-1 updates (reference src/nonlin_solve.f90:43-58, qns_solve :156-425)."""

    _entry = "nlb_quasi_newton_solve_batch"
    _sharded_kind = 2      # NLB_SOLVER_QUASI_NEWTON

    def __init__(self, engine=None):
        super().__init__(engine)
        self._jdelta = 5

    def get_jacobian_interval(self):
        return self._jdelta

    def set_jacobian_interval(self, n):
        self._jdelta = int(n)

    def _params(self, fcn):
        p = super()._params(fcn)
        p.jacobian_interval = self._jdelta
        return p


class newton_solver(line_search_solver):
    """Newton's method with LU (reference src/nonlin_solve.f90:60-67, ns_solve :452-638)."""

    _entry = "nlb_newton_solve_batch"
    _sharded_kind = 1      # NLB_SOLVER_NEWTON


class value_pair:
    """A pair of values per equation: the search limits of the one-variable solvers (reference `value_pair`,
    src/nonlin_types.f90:31-37).  x1, x2: scalars (broadcast over the batch) or (B,) arrays."""

    def __init__(self, x1=0.0, x2=0.0):
        self.x1 = x1
        self.x2 = x2


class fcn1var_helper:
    """Names a registered one-variable function (reference `fcn1var_helper`, src/nonlin_single_var.f90:25-41)."""

    def __init__(self):
        self._fcn_id = -1
        self._info = None
        self._use_diff = False

    def set_fcn(self, fcn):
        fid = _LIB.nlb_fcn1var_lookup(fcn.encode()) if isinstance(fcn, str) else int(fcn)
        a, d = C.c_int(), C.c_int()
        if fid < 0 or _LIB.nlb_fcn1var_info(fid, C.byref(a), C.byref(d)) != _lib.NLB_OK:
            raise NonlinError(_lib.NLB_ERR_UNKNOWN_FCN, "one-variable function %r is not registered" % (fcn,))
        self._fcn_id = fid
        self._info = {"args_len": a.value, "has_diff": bool(d.value)}
        self._use_diff = False

    def set_diff(self, enable=True):
        """`call obj%set_diff(diff)`: use the registered derivative instead of the forward difference (:203-211)."""
        if enable and not (self._info and self._info["has_diff"]):
            raise NonlinError(_lib.NLB_ERR_UNSUPPORTED, "no derivative is registered for this function")
        self._use_diff = bool(enable)

    def is_fcn_defined(self):
        return self._fcn_id >= 0

    def is_derivative_defined(self):
        return self._use_diff


def fcn1var_names():
    return [_LIB.nlb_fcn1var_name(i).decode() for i in range(_LIB.nlb_fcn1var_count())]


class equation_solver_1var:
    """Base of the one-variable solvers (reference src/nonlin_single_var.f90:43-69)."""

    _entry = None

    def __init__(self, engine=None):
        self._engine = engine
        self._max_eval = 100
        self._fcn_tol = 1.0e-8
        self._xtol = 1.0e-12
        self._difftol = 1.0e-12
        self._print_status = False

    def get_max_fcn_evals(self):
        return self._max_eval

    def set_max_fcn_evals(self, n):
        self._max_eval = int(n)

    def get_fcn_tolerance(self):
        return self._fcn_tol

    def set_fcn_tolerance(self, x):
        self._fcn_tol = float(x)

    def get_var_tolerance(self):
        return self._xtol

    def set_var_tolerance(self, x):
        self._xtol = float(x)

    def get_diff_tolerance(self):
        return self._difftol

    def set_diff_tolerance(self, x):
        self._difftol = float(x)

    def get_print_status(self):
        return self._print_status

    def set_print_status(self, x):
        self._print_status = bool(x)      # stored, not acted on (device-resident iterations)

    def solve(self, fcn, x, lim, f=None, ib=None, args=None, status=None, stream=None, want_f=True):
        """`call solver%solve(fcn, x, lim, f, ib, args)` over the B equations of x (B,), in place.

        lim: a `value_pair` whose members are scalars or (B,) arrays.  f: (B,) output; with `want_f=False` the
        reference's optional `f` is treated as absent.  Returns the per-equation status array."""
        if not fcn.is_fcn_defined():
            raise NonlinError(_lib.NLB_ERR_UNKNOWN_FCN, "no function set (NL_UNDEFINED_FUNCTION_ERROR)")
        if x.ndim != 1:
            raise NonlinError(_lib.NLB_ERR_SIZE, "x must be (B,)")
        B = x.shape[0]
        _check_f64("x", x, (B,))

        def limits(v, name):
            if np.ndim(v) == 0:
                a = _empty_like(x, (B,))
                a[...] = float(v)
                return a
            _check_f64(name, v, (B,))
            return v

        l1, l2 = limits(lim.x1, "lim%x1"), limits(lim.x2, "lim%x2")
        if f is None and want_f:
            f = _empty_like(x, (B,))
        if f is not None:
            _check_f64("f", f, (B,))
        if args is not None:
            _check_f64("args", args, (fcn._info["args_len"], B))
        if status is None:
            status = _empty_like(x, (B,), dtype="int32")
            status[...] = 0
        _check_i32("status", status, B)
        _check_i32("ib", ib, B, 7)
        eng = self._engine or default_engine(_device_of(x, f, args, ib, status) or 0)
        p = _lib.nlb_params_1var()
        _LIB.nlb_params_1var_default(C.byref(p))
        p.max_fcn_evals, p.fcn_tol, p.var_tol, p.diff_tol = self._max_eval, self._fcn_tol, self._xtol, self._difftol
        p.use_analytic_diff = int(fcn.is_derivative_defined())
        entry = getattr(_LIB, self._entry)
        eng.check(entry(eng._h, C.byref(p), fcn._fcn_id, B, _ptr(l1), _ptr(l2), _ptr(x), _ptr(f), _ptr(args), _ptr(ib),
                        _ptr(status), C.c_void_p(stream) if stream is not None else _stream_of(x, f, args, ib, status)))
        self.last_f = f
        return status


class brent_solver(equation_solver_1var):
    """Brent's method (reference src/nonlin_solve.f90:69-76, brent_solve :643-835)."""

    _entry = "nlb_brent_solve_batch"


class newton_1var_solver(equation_solver_1var):
    """Newton's method safeguarded by bisection (reference src/nonlin_solve.f90:78-85, newt1var_solve :840-1032)."""

    _entry = "nlb_newton_1var_solve_batch"


class polynomial:
    """A batch of B polynomials of one order, c0 + c1 x + ... (reference `polynomial`,
    src/nonlin_polynomials.f90:20-71): `fit`, `fit_thru_zero`, `evaluate`, `order`, `get`, `get_all`, `set`,
    `initialize`.  Coefficients are an (order + 1, B) array, data-set index fastest; `get(i)` / `set(i, v)` use the
    reference's 1-based index (get(1) = c0).  Roots, companion matrix and polynomial arithmetic are host-side
    single-polynomial operations outside the batch path and are not mirrored."""

    def __init__(self, order=None, B=1, engine=None):
        self._engine = engine
        self._c = None
        if order is not None:
            self.initialize(order, B)

    def initialize(self, order, B=1):
        """`call p%initialize(order)` (init_poly :77-103) or, with an array, `p%initialize(c)` (init_poly_coeffs :106-128)."""
        if np.ndim(order) > 0:
            c = np.array(order, dtype=np.float64)
            self._c = np.ascontiguousarray(c.reshape(c.shape[0], -1))
            return
        if order < 0:
            raise NonlinError(_lib.NLB_ERR_INVALID_ARGUMENT, "order must be >= 0")   # reference: error stop
        self._c = np.zeros((int(order) + 1, int(B)))

    def order(self):
        return -1 if self._c is None else int(self._c.shape[0]) - 1

    def get(self, i):
        return self._c[i - 1]

    def get_all(self):
        return self._c

    def set(self, i, v):
        self._c[i - 1] = v

    def _fit(self, x, y, order, thru_zero, status, stream):
        if y.ndim != 2:
            raise NonlinError(_lib.NLB_ERR_SIZE, "y must be (npts, B)")
        npts, B = y.shape
        shared = x.ndim == 1
        _check_f64("y", y, (npts, B))
        _check_f64("x", x, (npts,) if shared else (npts, B))            # size(y) /= size(x): error stop 3
        c = _empty_like(y, (int(order) + 1, B))
        if status is None:
            status = _empty_like(y, (B,), dtype="int32")
        _check_i32("status", status, B)
        eng = self._engine or default_engine(_device_of(x, y) or 0)
        eng.check(_LIB.nlb_polynomial_fit_batch(eng._h, B, npts, int(order), int(thru_zero), int(shared), _ptr(x), _ptr(y),
                                                _ptr(c), _ptr(status),
                                                C.c_void_p(stream) if stream is not None else _stream_of(x, y)))
        self._c = c
        return status

    def fit(self, x, y, order, status=None, stream=None):
        """`call p%fit(x, y, order)` for B data sets: y (npts, B), x (npts,) shared or (npts, B).  Returns the
        per-set status (0, or LA_INVALID_OPERATION_ERROR for an exactly rank-deficient set).  y is not overwritten."""
        return self._fit(x, y, order, False, status, stream)

    def fit_thru_zero(self, x, y, order, status=None, stream=None):
        """`call p%fit_thru_zero(x, y, order)`: same with c0 forced to 0 (poly_fit_thru_zero :202-253)."""
        return self._fit(x, y, order, True, status, stream)

    def evaluate(self, x, stream=None):
        """`p%evaluate(x)`: (npts, B) values of the B polynomials at x (npts,) or (npts, B) (poly_eval_double :256-283)."""
        if self._c is None:
            raise NonlinError(_lib.NLB_ERR_INVALID_ARGUMENT, "polynomial is not initialised")
        B = self._c.shape[1]
        shared = x.ndim == 1
        npts = x.shape[0]
        _check_f64("x", x, (npts,) if shared else (npts, B))
        yv = _empty_like(self._c, (npts, B))
        eng = self._engine or default_engine(_device_of(x, self._c) or 0)
        eng.check(_LIB.nlb_polynomial_evaluate_batch(eng._h, B, self.order(), npts, int(shared), _ptr(self._c), _ptr(x),
                                                     _ptr(yv),
                                                     C.c_void_p(stream) if stream is not None else _stream_of(x, self._c)))
        return yv


def vecfcn_names():
    return [_LIB.nlb_vecfcn_name(i).decode() for i in range(_LIB.nlb_vecfcn_count())]
