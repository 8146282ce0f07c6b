"""Time constrained_least_squares_solver on a run-time-m curve-fit family (csrc/cls_rt.cuh) against the CPU port.
usage: time_cls_rt.py [exp_decay_4|rational_7_8] [m] [B]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import numpy as np, torch
import nonlin_b200 as nb
from nonlin_b200 import workloads as W
import nl_oracle
fcn = sys.argv[1] if len(sys.argv) > 1 else "exp_decay_4"
m = int(sys.argv[2]) if len(sys.argv) > 2 else 64
B = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 18
w = W.lm_expdecay4(B, m=m) if fcn == "exp_decay_4" else W.c4_lm_rational(B, m=m, noise=1e-3)
n = w["n"]
obj = nb.vecfcn_helper(); obj.set_fcn(fcn, m, n); obj.set_shared_data(torch.from_numpy(w["shared"]).cuda())
s = nb.constrained_least_squares_solver(); s.set_max_fcn_evals(200)
lo, up = [-10.0] * n, [10.0] * n
s.set_lower_limits(lo); s.set_upper_limits(up)
x0 = torch.from_numpy(w["x0"]).cuda(); args = torch.from_numpy(w["args"]).cuda()
f = torch.empty((m, B), dtype=torch.float64, device="cuda"); ib = nb.iteration_behavior(B, like=x0)
st = torch.zeros(B, dtype=torch.int32, device="cuda")
best = 1e30
for k in range(4):
    x = x0.clone(); torch.cuda.synchronize()
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record(); s.solve(obj, x, f, ib, args=args, status=st); b.record(); torch.cuda.synchronize()
    if k: best = min(best, a.elapsed_time(b))
stats = nb.default_engine(0).reduce_stats(ib, st, B)
print("%s m=%d n=%d B=%d: %.2f ms -> %.3e systems/s; converged %d, sum_iter %d" % (fcn, m, n, B, best, B / best * 1e3, stats["converged"], stats["sum_iter"]))
o = nl_oracle.Oracle(); Bs = min(B, 16384)
t = time.time()
xo, fo, ibo, sto = o.cls_solve_batch(fcn, w["x0"][:, :Bs], m=m, sys=w["args"][:, :Bs], shared=w["shared"], lower=lo, upper=up, params=o.params(max_fcn_evals=200))
dt = time.time() - t
print("CPU port, all cores, %d systems: %.3e systems/s; engine == port on the sample: %s" % (Bs, Bs / dt, bool(np.array_equal(x[:, :Bs].cpu().numpy(), xo) and np.array_equal(st[:Bs].cpu().numpy(), sto))))
