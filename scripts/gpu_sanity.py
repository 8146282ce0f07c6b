"""Scratch GPU check: engine vs CPU oracle on the thread-per-system configs, plus rough timings."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import nonlin_b200 as nb
from nonlin_b200 import workloads as W
from oracle.nl_oracle import Oracle

o = Oracle()
eng = nb.default_engine(0)
print("fp64 peak:", eng.measure_fp64_peak())

def mk_solver(w):
    s = {"least_squares": nb.least_squares_solver, "newton": nb.newton_solver, "quasi_newton": nb.quasi_newton_solver}[w["solver"]]()
    for k, v in w["settings"].items():
        getattr(s, k)(v)
    return s

def oparams(w):
    kw = {}
    if "set_max_fcn_evals" in w["settings"]:
        kw["max_fcn_evals"] = w["settings"]["set_max_fcn_evals"]
    return o.params(**kw)

for name in ("C2", "C3", "C1"):
    B = 8192
    w = W.WORKLOADS[name](B)
    obj = nb.vecfcn_helper(); obj.set_fcn(w["fcn"], w["m"], w["n"])
    s = mk_solver(w)
    x = w["x0"].copy(); f = np.zeros((w["m"], B)); ib = nb.iteration_behavior(B)
    st = s.solve(obj, x, f, ib, args=w["args"])
    xo, fo, ibo, sto = o.solve_batch(w["solver"], w["fcn"], w["x0"], m=w["m"], sys=w["args"], params=oparams(w))
    same_x = np.all(x == xo, axis=0); same_f = np.all(f == fo, axis=0)
    cnt = (ib["iter_count"] == ibo["iter_count"]) & (ib["fcn_count"] == ibo["fcn_count"]) & (ib["jacobian_count"] == ibo["jacobian_count"])
    print(name, "bitwise x: %.4f  f: %.4f  counts: %.4f  status eq: %.4f  converged: %.4f  mean iter %.2f nfev %.2f njac %.2f max iter %d" % (
        same_x.mean(), same_f.mean(), cnt.mean(), (st == sto).mean(), (st == 0).mean(), ib["iter_count"].mean(), ib["fcn_count"].mean(), ib["jacobian_count"].mean(), ib["iter_count"].max()))
    if same_x.mean() < 1:
        bad = np.where(~same_x)[0][:3]
        for b in bad:
            print("  b", b, "x0", w["x0"][:, b], "gpu", x[:, b], "cpu", xo[:, b], ib[b], ibo[b], st[b], sto[b])

# timing, device-resident
for name, B in (("C2", 1 << 20), ("C3", 1 << 20), ("C1", 1 << 20)):
    w = W.WORKLOADS[name](B)
    obj = nb.vecfcn_helper(); obj.set_fcn(w["fcn"], w["m"], w["n"])
    s = mk_solver(w)
    x0 = torch.from_numpy(w["x0"]).cuda(); args = None if w["args"] is None else torch.from_numpy(w["args"]).cuda()
    f = torch.empty((w["m"], B), dtype=torch.float64, device="cuda"); ib = nb.iteration_behavior(B, like=x0)
    status = torch.zeros(B, dtype=torch.int32, device="cuda")
    for it in range(3):
        x = x0.clone()
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        s.solve(obj, x, f, ib, args=args, status=status)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
    stats = eng.reduce_stats(ib, status, B)
    print(name, "B", B, "ms %.3f" % ms, "systems/s %.3e" % (B / ms * 1e3), stats)
    t0 = time.time(); nsub = 1 << 15
    o.solve_batch(w["solver"], w["fcn"], w["x0"][:, :nsub].copy(), m=w["m"], sys=None if w["args"] is None else w["args"][:, :nsub].copy(), params=oparams(w))
    dt = time.time() - t0
    print("   cpu oracle: %d systems in %.3fs -> %.3e systems/s (%d threads)" % (nsub, dt, nsub / dt, os.cpu_count()))
