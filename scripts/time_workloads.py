"""Time workloads device-resident + CPU port. usage: time_workloads.py NAME:B[:m] ..."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import nonlin_b200 as nb
from nonlin_b200 import workloads as W
from oracle.nl_oracle import Oracle
o = Oracle(); eng = nb.default_engine(0)
for spec in sys.argv[1:]:
    parts = spec.split(":"); name = parts[0]; B = int(parts[1]); kw = {"m": int(parts[2])} if len(parts) > 2 else {}
    w = W.WORKLOADS[name](B, **kw)
    if os.environ.get("MAXEVAL"): w["settings"]["set_max_fcn_evals"] = int(os.environ["MAXEVAL"])
    obj = nb.vecfcn_helper(); obj.set_fcn(w["fcn"], w["m"], w["n"])
    if w["shared"] is not None: obj.set_shared_data(torch.from_numpy(w["shared"]).cuda())
    s = {"least_squares": nb.least_squares_solver, "newton": nb.newton_solver, "quasi_newton": nb.quasi_newton_solver,
         "constrained_least_squares": nb.constrained_least_squares_solver}[w["solver"]]()
    okw = {}
    for k, v in w["settings"].items():
        getattr(s, k)(v)
        if k == "set_max_fcn_evals": okw["max_fcn_evals"] = v
    x0 = torch.from_numpy(w["x0"]).cuda(); args = None if w["args"] is None else torch.from_numpy(w["args"]).cuda()
    f = torch.empty((w["m"], B), dtype=torch.float64, device="cuda"); ib = nb.iteration_behavior(B, like=x0); st = torch.zeros(B, dtype=torch.int32, device="cuda")
    best = 1e30
    for it in range(3):
        x = x0.clone(); torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); s.solve(obj, x, f, ib, args=args, status=st); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    stats = eng.reduce_stats(ib, st, B)
    nsub = min(B, int(os.environ.get("CPU_SUB", "512")))
    t0 = time.time()
    sub = dict(m=w["m"], sys=None if w["args"] is None else w["args"][:, :nsub].copy(), shared=w["shared"], params=o.params(**okw))
    if w["solver"] == "constrained_least_squares":
        o.cls_solve_batch(w["fcn"], w["x0"][:, :nsub].copy(), lower=w["settings"].get("set_lower_limits"), upper=w["settings"].get("set_upper_limits"), **sub)
    else:
        o.solve_batch(w["solver"], w["fcn"], w["x0"][:, :nsub].copy(), **sub)
    dt = time.time() - t0
    print("%s B=%d %s: gpu %.2f ms -> %.3e systems/s | cpu port %.3e systems/s (%d cores) | ratio %.1f | conv %d/%d mean iter %.1f nfev %.1f njac %.1f max iter %d" % (
        name, B, kw, best, B / best * 1e3, nsub / dt, os.cpu_count(), (B / best * 1e3) / (nsub / dt), stats["converged"], B,
        stats["sum_iter"] / B, stats["sum_fcn"] / B, stats["sum_jac"] / B, stats["max_iter"]), flush=True)
