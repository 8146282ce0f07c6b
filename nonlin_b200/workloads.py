"""Seeded synthetic batches for the BASELINE.json configurations.

Pure input generation (numpy, host): the same arrays are fed to the engine, to the CPU oracle
in the tests, and to bench.py.  Nothing here computes a solution.

Each workload is a dict:
    solver   "least_squares" | "newton" | "quasi_newton" | "constrained_least_squares"
    fcn      registered residual name
    m, n     system size
    x0       (n, B) float64 starting points, system index fastest
    args     (sys_len, B) per-system data or None
    shared   (shared_len,) shared data or None
    settings dict of solver setter -> value (only non-defaults)
    bytes_per_system  algorithmic HBM bytes per system (read once + write once)
"""
import numpy as np

# README Example 2 / tests/nonlin_test_solve.f90:133-159 data (decimal literals of the reference)
POLYFIT_XP = np.array([0.0, 0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9, 1.0,
                       1.1, 1.2, 1.3, 1.4, 1.5, 1.6, 1.7, 1.8, 1.9, 2.0])
POLYFIT_YP = np.array([1.216737514, 1.250032542, 1.305579195, 1.040182335, 1.751867738,
                       1.109716707, 2.018141531, 1.992418729, 1.807916923, 2.078806005,
                       2.698801324, 2.644662712, 3.412756702, 4.406137221, 4.567156645,
                       4.999550779, 5.652854194, 6.784320119, 8.307936836, 8.395126494,
                       10.30252404])


def _bytes(m, n, sys_len):
    # x in + x out + fvec out + per-system data + iteration_behavior (28) + status (4)
    return 8 * (n + n + m) + 8 * sys_len + 32


def c1_lm_polyfit(B, seed=1):
    """C1 batch: README Example 2 cubic fit (m=21, n=4), per-system y = yp + 0.05 N(0,1)."""
    rng = np.random.default_rng(seed)
    y = POLYFIT_YP[:, None] + 0.05 * rng.standard_normal((21, B))
    return dict(name="C1", solver="least_squares", fcn="lsq_poly_fit", m=21, n=4,
                x0=np.ones((4, B)), args=np.ascontiguousarray(y), shared=None, settings={},
                bytes_per_system=_bytes(21, 4, 21))


def c2_broyden_2x2(B, seed=2):
    """C2: README Example 1 system, x0 = (1,1) + U(-0.5,0.5)^2, quasi-Newton defaults."""
    rng = np.random.default_rng(seed)
    x0 = 1.0 + rng.uniform(-0.5, 0.5, size=(2, B))
    return dict(name="C2", solver="quasi_newton", fcn="misc_2fcn", m=2, n=2,
                x0=np.ascontiguousarray(x0), args=None, shared=None, settings={},
                bytes_per_system=_bytes(2, 2, 0))


def c3_newton_powell(B, seed=3):
    """C3: Powell badly scaled, x0 = (0,1) + U(-0.1,0.1)^2, Newton + line search, max evals 1000."""
    rng = np.random.default_rng(seed)
    x0 = np.array([[0.0], [1.0]]) + rng.uniform(-0.1, 0.1, size=(2, B))
    return dict(name="C3", solver="newton", fcn="powell_badly_scaled", m=2, n=2,
                x0=np.ascontiguousarray(x0), args=None, shared=None,
                settings={"set_max_fcn_evals": 1000},
                bytes_per_system=_bytes(2, 2, 0))


def _rational(p, q, t):
    num = np.zeros((p.shape[1], t.size))
    den = np.zeros_like(num)
    for k in range(7, -1, -1):
        num = num * t[None, :] + p[k][:, None]
        den = den * t[None, :] + q[k][:, None]
    return num / (1.0 + t[None, :] * den)


def c4_lm_rational(B, m=4096, seed=4, noise=0.0, with_y=True):
    """C4: LM curve fit m x 16, rational 7/8 model on t in [-1,1]; x0 = truth*(1 + 0.1 U(-1,1)).

    with_y=False leaves `args` (the m x B observations, 2 GB at the BASELINE batch) to the caller, who forms them
    from `truth` with the same Horner recurrence as `_rational` (bench.py does it on the device); x0 is then drawn
    from the stream position of the noise-free workload."""
    rng = np.random.default_rng(seed)
    t = np.linspace(-1.0, 1.0, m)
    p = rng.uniform(-1.0, 1.0, size=(8, B))
    q = rng.uniform(-0.1, 0.1, size=(8, B))
    truth = np.concatenate([p, q], axis=0)
    y = None
    if with_y:
        y = _rational(p, q, t).T  # (m, B)
        if noise:
            y = y + noise * rng.standard_normal(y.shape)
        y = np.ascontiguousarray(y)
    x0 = truth * (1.0 + 0.1 * rng.uniform(-1.0, 1.0, size=truth.shape))
    return dict(name="C4" if not noise else "C4N", solver="least_squares", fcn="rational_7_8", m=m, n=16,
                x0=np.ascontiguousarray(x0), args=y, shared=t, truth=truth, noise=noise,
                settings={"set_max_fcn_evals": 1000},
                bytes_per_system=_bytes(m, 16, m))


def c4n_lm_rational(B, m=4096, seed=4, with_y=True):
    """C4N: the C4 fits with Gaussian noise of sigma = 1e-3 on the observations (SURVEY.md 8d variant)."""
    return c4_lm_rational(B, m=m, seed=seed, noise=1e-3, with_y=with_y)


def c5_broyden_rosenbrock(B, n=64, seed=5):
    """C5: extended Rosenbrock, x0 = (-1.2, 1, ...)*(1 + 0.1 U(-1,1)), quasi-Newton + line search."""
    rng = np.random.default_rng(seed)
    base = np.tile(np.array([-1.2, 1.0]), n // 2)[:, None]
    x0 = base * (1.0 + 0.1 * rng.uniform(-1.0, 1.0, size=(n, B)))
    return dict(name="C5", solver="quasi_newton", fcn="ext_rosenbrock", m=n, n=n,
                x0=np.ascontiguousarray(x0), args=None, shared=None, settings={},
                bytes_per_system=_bytes(n, n, 0))


def lm_expdecay4(B, m=64, seed=6, noise=1e-2):
    """4-parameter double-exponential LM fit (SURVEY.md §6 probe): y = a1 e^{-b1 t} + a2 e^{-b2 t}."""
    rng = np.random.default_rng(seed)
    t = np.linspace(0.0, 4.0, m)
    a1 = rng.uniform(1.0, 3.0, B)
    b1 = rng.uniform(0.2, 0.6, B)
    a2 = rng.uniform(0.5, 1.5, B)
    b2 = rng.uniform(1.5, 3.0, B)
    truth = np.stack([a1, b1, a2, b2])
    y = a1[None] * np.exp(-b1[None] * t[:, None]) + a2[None] * np.exp(-b2[None] * t[:, None])
    y = y + noise * rng.standard_normal(y.shape)
    x0 = truth * (1.0 + 0.2 * rng.uniform(-1.0, 1.0, size=truth.shape))
    return dict(name="LM4", solver="least_squares", fcn="exp_decay_4", m=m, n=4,
                x0=np.ascontiguousarray(x0), args=np.ascontiguousarray(y), shared=t,
                settings={"set_max_fcn_evals": 1000},
                bytes_per_system=_bytes(m, 4, m))


def cls1_bounded_polyfit(B, seed=7):
    """CLS1: the C1 cubic fits through constrained_least_squares_solver, coefficients boxed to [-10, 10]
    (limits inactive at the solution, Coleman-Li scaling live), start 0.5."""
    w = c1_lm_polyfit(B, seed)
    w.update(name="CLS1", solver="constrained_least_squares", x0=np.full((4, B), 0.5),
             settings={"set_lower_limits": [-10.0] * 4, "set_upper_limits": [10.0] * 4})
    return w


def cls2_bounded_2x2(B, seed=8):
    """CLS2: README Example 1 system in the box [0, 6]^2, starts U(0.2, 5.8)^2 (root (5, 3) inside the box)."""
    rng = np.random.default_rng(seed)
    x0 = rng.uniform(0.2, 5.8, size=(2, B))
    return dict(name="CLS2", solver="constrained_least_squares", fcn="misc_2fcn", m=2, n=2,
                x0=np.ascontiguousarray(x0), args=None, shared=None,
                settings={"set_lower_limits": [0.0, 0.0], "set_upper_limits": [6.0, 6.0]},
                bytes_per_system=_bytes(2, 2, 0))


WORKLOADS = {
    "C1": c1_lm_polyfit,
    "C2": c2_broyden_2x2,
    "C3": c3_newton_powell,
    "C4": c4_lm_rational,
    "C4N": c4n_lm_rational,
    "C5": c5_broyden_rosenbrock,
    "LM4": lm_expdecay4,
    "CLS1": cls1_bounded_polyfit,
    "CLS2": cls2_bounded_2x2,
}
