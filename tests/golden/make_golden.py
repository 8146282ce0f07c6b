"""Regenerate tests/golden/oracle_batches.npz: outputs of the CPU oracle on small seeded batches.

    python tests/golden/make_golden.py

The GPU parity tests compare the engine with these committed vectors as well as with the oracle
run live, so a silent change of either side is caught.  (The reference itself is Fortran and
cannot be built in this image — no Fortran compiler — so these vectors come from the oracle,
which is pinned to the reference's published outputs by tests/test_oracle_kat.py.)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from nonlin_b200 import workloads as W  # noqa: E402
from oracle.nl_oracle import Oracle  # noqa: E402

CASES = {"C1": 256, "C2": 512, "C3": 512, "C5": 8, "LM4": 128, "C4": 4}
EXTRA = {"C4": {"m": 256}}


def oracle_params(o, w):
    kw = {}
    if "set_max_fcn_evals" in w["settings"]:
        kw["max_fcn_evals"] = w["settings"]["set_max_fcn_evals"]
    return o.params(**kw)


def scalar_case(B, seed=12):
    """Seeded cubics a0 + a1 x + a2 x^2 + a3 x^3 with a bracket [-1, 6] (shared with tests/test_golden_widening.py)."""
    rng = np.random.default_rng(seed)
    args = np.stack([rng.uniform(-8, -1, B), rng.uniform(-2, 2, B), rng.uniform(-1, 1, B), rng.uniform(0.5, 2, B)])
    return {"args": args, "lim1": np.full(B, -1.0), "lim2": np.full(B, 6.0)}


def main():
    o = Oracle()
    out = {}
    for name, B in CASES.items():
        w = W.WORKLOADS[name](B, **EXTRA.get(name, {}))
        x, f, ib, st = o.solve_batch(w["solver"], w["fcn"], w["x0"], m=w["m"], sys=w["args"], shared=w["shared"],
                                     params=oracle_params(o, w))
        out[name + "_x"] = x
        out[name + "_f"] = f
        out[name + "_ib"] = ib.view(np.int32).reshape(B, 7)
        out[name + "_status"] = st
        print(name, "B", B, "converged", (st == 0).mean(), "mean iter", ib["iter_count"].mean())
    # SURVEY 8(f) widening: constrained least squares, polynomial fit, one-variable solvers
    for name, B in (("CLS1", 256), ("CLS2", 512)):
        w = W.WORKLOADS[name](B)
        x, f, ib, st = o.cls_solve_batch(w["fcn"], w["x0"], m=w["m"], sys=w["args"],
                                         lower=w["settings"]["set_lower_limits"], upper=w["settings"]["set_upper_limits"])
        out[name + "_x"], out[name + "_f"], out[name + "_ib"], out[name + "_status"] = x, f, ib.view(np.int32).reshape(B, 7), st
        print(name, "B", B, "converged", (st == 0).mean(), "mean iter", ib["iter_count"].mean())
    w = W.WORKLOADS["C1"](256)
    c, st = o.polyfit_batch(W.POLYFIT_XP, w["args"], 3)
    out["POLY1_c"], out["POLY1_status"] = c, st
    c0, st0 = o.polyfit_batch(W.POLYFIT_XP[1:], np.ascontiguousarray(w["args"][1:]), 2, thru_zero=True)
    out["POLY0_c"], out["POLY0_status"] = c0, st0
    s1 = scalar_case(512)
    for solver in ("brent", "newton_1var"):
        x, f, ib, st = o.solve_1var_batch(solver, "cubic_args", s1["lim1"], s1["lim2"], args=s1["args"])
        out[solver + "_x"], out[solver + "_f"], out[solver + "_ib"], out[solver + "_status"] = x, f, ib.view(np.int32).reshape(-1, 7), st
        print(solver, "converged", (st == 0).mean(), "mean iter", ib["iter_count"].mean())
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_batches.npz"), **out)


if __name__ == "__main__":
    main()
