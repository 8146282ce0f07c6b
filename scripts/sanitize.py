"""Tiny batches through every cooperative / persistent kernel, for compute-sanitizer (memcheck, racecheck).
usage: compute-sanitizer --tool memcheck python scripts/sanitize.py [which ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import nonlin_b200 as nb
from nonlin_b200 import workloads as W

def run(w, maxeval=None, **kw):
    obj = nb.vecfcn_helper(); obj.set_fcn(w["fcn"], w["m"], w["n"])
    if w["shared"] is not None: obj.set_shared_data(w["shared"])
    s = {"least_squares": nb.least_squares_solver, "newton": nb.newton_solver, "quasi_newton": nb.quasi_newton_solver,
         "constrained_least_squares": nb.constrained_least_squares_solver}[w["solver"]]()
    for k, v in w["settings"].items(): getattr(s, k)(v)
    if maxeval: s.set_max_fcn_evals(maxeval)
    x = w["x0"].copy(); f = np.zeros((w["m"], x.shape[1])); ib = nb.iteration_behavior(x.shape[1])
    st = s.solve(obj, x, f, ib, args=w["args"])
    return int((st == 0).sum()), int(ib["iter_count"].sum())

cases = {
    "tlm": lambda: run(W.c4_lm_rational(6, m=576), 12),                 # tall_lm.cuh (TMA ring, mbarriers, chain warp)
    "tlm_noise": lambda: run(W.c4_lm_rational(5, m=640, noise=1e-3), 10),
    "wlm": None,                                                         # selected with NLB_TALL_LM=wlm (separate process)
    "coop_lm": lambda: run(W.lm_expdecay4(70, m=33), 30),                # coop_lm.cuh (lane per system, work queue)
    "coop_lm16": lambda: run(W.c4_lm_rational(40, m=64), 20),
    "broyden64": lambda: run(W.c5_broyden_rosenbrock(6, n=64)),          # coop_broyden.cuh
    "broyden16": lambda: run(W.c5_broyden_rosenbrock(9, n=16)),
    "newton_refill": lambda: run(W.c3_newton_powell(3000)),              # persistent refill kernels
    "cls_refill": lambda: run(W.cls2_bounded_2x2(3000)),
    "cls1": lambda: run(W.cls1_bounded_polyfit(1000)),
    "tps_lm": lambda: run(W.c1_lm_polyfit(1000)),
    "broyden2": lambda: run(W.c2_broyden_2x2(3000)),
}
def cls_rt():
    """constrained least squares on a run-time-m family (cls_rt.cuh: HBM workspace)"""
    w = W.lm_expdecay4(300, m=33)
    obj = nb.vecfcn_helper(); obj.set_fcn(w["fcn"], w["m"], w["n"]); obj.set_shared_data(w["shared"])
    s = nb.constrained_least_squares_solver(); s.set_max_fcn_evals(60); s.set_lower_limits([0.5] * 4)
    x = w["x0"].copy(); f = np.zeros((w["m"], 300)); ib = nb.iteration_behavior(300)
    st = s.solve(obj, x, f, ib, args=w["args"])
    return int((st == 0).sum()), int(ib["iter_count"].sum())
cases["cls_rt"] = cls_rt
def polyfit():
    p = nb.polynomial()
    w = W.c1_lm_polyfit(700)
    st = p.fit(W.POLYFIT_XP, w["args"], 3)                                # shared-memory workspace
    y2 = np.random.default_rng(1).standard_normal((300, 64)); x2 = np.linspace(0, 1, 300)
    st2 = nb.polynomial().fit(x2, y2, 5)                                  # global workspace
    return int((st == 0).sum()), int((st2 == 0).sum())
cases["polyfit"] = polyfit
which = sys.argv[1:] or [k for k, v in cases.items() if v is not None]
for k in which:
    print(k, cases[k](), flush=True)
print("sanitize.py done")
