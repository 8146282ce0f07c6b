"""Quad LM kernel (NLB_LM_QUAD=1) against the oracle on C1 samples, then a timing of 2^20 fits."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import nonlin_b200 as nb
from nonlin_b200 import workloads as W
from oracle.nl_oracle import Oracle
o = Oracle()
w = W.c1_lm_polyfit(20001, seed=7)
# harder variants too: random starts far away (more iterations, lmpar active), tiny budget
for tag, x0, kw in (("C1", w["x0"], {}), ("far", w["x0"] * np.random.default_rng(3).uniform(-50, 50, w["x0"].shape), {}),
                    ("budget3", w["x0"] * 30.0, {"max_fcn_evals": 3})):
    obj = nb.vecfcn_helper(); obj.set_fcn(w["fcn"], 21, 4)
    s = nb.least_squares_solver()
    if kw: s.set_max_fcn_evals(kw["max_fcn_evals"])
    x = x0.copy(); f = np.zeros((21, x.shape[1])); ib = nb.iteration_behavior(x.shape[1])
    st = s.solve(obj, x, f, ib, args=w["args"])
    xo, fo, ibo, sto = o.solve_batch("least_squares", w["fcn"], x0, m=21, sys=w["args"], params=o.params(**kw))
    same = np.array_equal(x, xo) and np.array_equal(f, fo) and np.array_equal(ib, ibo) and np.array_equal(st, sto)
    print(tag, "bit-identical:", same, "mean iter %.2f nfev %.2f failed %d" % (ibo["iter_count"].mean(), ibo["fcn_count"].mean(), (sto != 0).sum()), flush=True)
B = 1 << 20
w = W.c1_lm_polyfit(B)
obj = nb.vecfcn_helper(); obj.set_fcn(w["fcn"], 21, 4)
s = nb.least_squares_solver()
x0 = torch.from_numpy(w["x0"]).cuda(); args = torch.from_numpy(w["args"]).cuda()
f = torch.empty((21, B), dtype=torch.float64, device="cuda"); ib = nb.iteration_behavior(B, like=x0); st = torch.zeros(B, dtype=torch.int32, device="cuda")
for rep in range(4):
    x = x0.clone(); torch.cuda.synchronize(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record(); s.solve(obj, x, f, ib, args=args, status=st); b.record(); torch.cuda.synchronize()
print("2^20 fits: %.3f ms -> %.3g fits/s" % (a.elapsed_time(b), B / a.elapsed_time(b) * 1e3))
