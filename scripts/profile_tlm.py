"""One short launch of the tall LM kernel for ncu: `nrep` copies of a slow C4 system, max_fcn_evals = 8."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import nonlin_b200 as nb
from nonlin_b200 import workloads as W
nrep = int(sys.argv[1]) if len(sys.argv) > 1 else 444
w = W.c4_lm_rational(64)
obj = nb.vecfcn_helper(); obj.set_fcn(w["fcn"], w["m"], w["n"]); obj.set_shared_data(torch.from_numpy(w["shared"]).cuda())
s = nb.least_squares_solver(); s.set_max_fcn_evals(int(os.environ.get("MAXEVAL", "8")))
idx = torch.from_numpy(np.resize(np.arange(64), nrep)).cuda()
x0 = torch.from_numpy(w["x0"]).cuda()[:, idx].contiguous(); args = torch.from_numpy(w["args"]).cuda()[:, idx].contiguous()
f = torch.empty((4096, nrep), dtype=torch.float64, device="cuda"); ib = nb.iteration_behavior(nrep, like=x0); st = torch.zeros(nrep, dtype=torch.int32, device="cuda")
for _ in range(2):
    x = x0.clone(); s.solve(obj, x, f, ib, args=args, status=st)
torch.cuda.synchronize()
print(nb.default_engine(0).reduce_stats(ib, st, nrep))
