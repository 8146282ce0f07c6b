import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import nonlin_b200 as nb
from nonlin_b200 import workloads as W
# find a straggler system with the oracle-free route: run engine on 2368 with maxeval 200, pick those with nfev == 200
B = 2368
w = W.c4_lm_rational(B)
obj = nb.vecfcn_helper(); obj.set_fcn(w["fcn"], w["m"], w["n"]); obj.set_shared_data(torch.from_numpy(w["shared"]).cuda())
s = nb.least_squares_solver(); s.set_max_fcn_evals(100)
x0 = torch.from_numpy(w["x0"]).cuda(); args = torch.from_numpy(w["args"]).cuda()
def run(xs, ar, maxeval):
    s.set_max_fcn_evals(maxeval)
    Bn = xs.shape[1]
    x = xs.clone(); f = torch.empty((4096, Bn), dtype=torch.float64, device="cuda"); ib = nb.iteration_behavior(Bn, like=x); st = torch.zeros(Bn, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize(); t0 = time.time()
    s.solve(obj, x, f, ib, args=ar, status=st); torch.cuda.synchronize()
    return time.time() - t0, nb.ib_view(ib)
dt, ib = run(x0, args, 100)
print("B=2368 maxeval=100: %.2fs" % dt, "njac hist", np.percentile(ib["jacobian_count"], [50, 90, 99, 100]))
slow = np.where(ib["fcn_count"] >= 100)[0]
print("stragglers:", len(slow))
for nsel in (1, 4, 32):
    idx = torch.from_numpy(slow[:nsel]).cuda()
    xs = x0[:, idx].contiguous(); ar = args[:, idx].contiguous()
    for me in (50, 100):
        dt, ibs = run(xs, ar, me)
        print("lone group of %d straggler(s), maxeval=%d: %.3fs -> %.2f ms per evaluation (njac %s)" % (nsel, me, dt, dt / me * 1e3, ibs["jacobian_count"][:4]))
# throughput regime: the same straggler replicated so that every CTA does identical work
for nrep in (148, 444, 592, 1184, 2368):
    idx = torch.from_numpy(np.resize(slow[:8], nrep)).cuda()
    xs = x0[:, idx].contiguous(); ar = args[:, idx].contiguous()
    dt, ibs = run(xs, ar, 50)
    nj = int(ibs["jacobian_count"].sum())
    print("%d replicated stragglers, maxeval=50: %.3fs -> %.1f us per outer iteration per system amortised, %.0f outer iterations/s (njac/system %d)"
          % (nrep, dt, dt / nj * 1e6, nj / dt, nj // nrep))
