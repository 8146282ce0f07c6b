"""Run one workload's solve a few times (for ncu). usage: profile_solve.py C2 [B] [reps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import nonlin_b200 as nb
from nonlin_b200 import workloads as W
name = sys.argv[1]; B = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20; reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
kw = {"m": int(os.environ.get("NLB_M", "4096"))} if name == "C4" else {}
w = W.WORKLOADS[name](B, **kw)
obj = nb.vecfcn_helper(); obj.set_fcn(w["fcn"], w["m"], w["n"])
if w["shared"] is not None: obj.set_shared_data(torch.from_numpy(w["shared"]).cuda())
s = {"least_squares": nb.least_squares_solver, "newton": nb.newton_solver, "quasi_newton": nb.quasi_newton_solver,
     "constrained_least_squares": nb.constrained_least_squares_solver}[w["solver"]]()
for k, v in w["settings"].items(): getattr(s, k)(v)
x0 = torch.from_numpy(w["x0"]).cuda(); args = None if w["args"] is None else torch.from_numpy(w["args"]).cuda()
f = torch.empty((w["m"], B), dtype=torch.float64, device="cuda"); ib = nb.iteration_behavior(B, like=x0); st = torch.zeros(B, dtype=torch.int32, device="cuda")
for _ in range(reps):
    x = x0.clone(); s.solve(obj, x, f, ib, args=args, status=st)
torch.cuda.synchronize()
print(nb.default_engine(0).reduce_stats(ib, st, B))
