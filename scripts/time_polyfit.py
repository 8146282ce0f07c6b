"""Time the batched polynomial fit, device-resident, vs the CPU port. usage: time_polyfit.py B npts order [shared]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import nonlin_b200 as nb
from oracle.nl_oracle import Oracle
B, npts, order = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
shared = len(sys.argv) > 4 and sys.argv[4] == "shared"
rng = np.random.default_rng(1)
x = np.linspace(0.0, 2.0, npts) if shared else np.sort(rng.uniform(0.0, 2.0, size=(npts, B)), axis=0)
y = rng.standard_normal((npts, B))
xd, yd = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()
p = nb.polynomial()
best = 1e30
for it in range(5):
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); st = p.fit(xd, yd, order); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
nsub = min(B, 65536)
o = Oracle()
t0 = time.time(); co, _ = o.polyfit_batch(x if shared else np.ascontiguousarray(x[:, :nsub]), np.ascontiguousarray(y[:, :nsub]), order); dt = time.time() - t0
ok = np.array_equal(p.get_all()[:, :nsub].cpu().numpy(), co)
bytes_alg = 8 * (npts * (1 if shared else 2) + order + 1) + 4
print("polyfit B=%d npts=%d order=%d shared=%s tpsm=%s: gpu %.3f ms -> %.3e fits/s (%.1f GB/s algorithmic) | cpu port %.3e fits/s (%d cores) | bitwise %s" % (
    B, npts, order, shared, os.environ.get("NLB_POLYFIT_THREADS_PER_SM", "default"), best, B / best * 1e3, B * bytes_alg / best / 1e6, nsub / dt, os.cpu_count(), ok), flush=True)
