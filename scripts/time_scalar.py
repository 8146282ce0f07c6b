"""Time the one-variable solvers, device-resident, vs the CPU port. usage: time_scalar.py [B]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import nonlin_b200 as nb
from oracle.nl_oracle import Oracle
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
rng = np.random.default_rng(8)
args = np.stack([rng.uniform(-8, -1, B), rng.uniform(-2, 2, B), rng.uniform(-1, 1, B), rng.uniform(0.5, 2, B)])
lo, hi = np.full(B, -1.0), np.full(B, 6.0)
o = Oracle()
for solver, cls in (("brent", nb.brent_solver), ("newton_1var", nb.newton_1var_solver)):
    for fcn, a in (("cubic_args", args), ("exp_minus_x", None)):
        obj = nb.fcn1var_helper(); obj.set_fcn(fcn)
        s = cls()
        l1 = lo if a is not None else np.full(B, 0.0) + rng.uniform(-1, 0.5, B)
        l2 = hi if a is not None else rng.uniform(0.6, 3.0, B)
        xd = torch.zeros(B, dtype=torch.float64, device="cuda"); fd = torch.zeros_like(xd)
        ad = None if a is None else torch.from_numpy(a).cuda()
        ibd = nb.iteration_behavior(B, like=xd); st = torch.zeros(B, dtype=torch.int32, device="cuda")
        lim = nb.value_pair(torch.from_numpy(l1).cuda(), torch.from_numpy(l2).cuda())
        best = 1e30
        for it in range(5):
            torch.cuda.synchronize()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); s.solve(obj, xd, lim, f=fd, ib=ibd, args=ad, status=st); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        nsub = min(B, 1 << 18)
        t0 = time.time()
        xo, fo, ibo, sto = o.solve_1var_batch(solver, fcn, l1[:nsub], l2[:nsub], args=None if a is None else np.ascontiguousarray(a[:, :nsub]))
        dt = time.time() - t0
        ok = np.array_equal(xd.cpu().numpy()[:nsub], xo)
        ib = ibd.cpu().numpy()
        print("%s %s B=%d: gpu %.3f ms -> %.3e equations/s | cpu port %.3e /s (%d cores) | converged %.4f mean iter %.1f nfev %.1f | bitwise %s" % (
            solver, fcn, B, best, B / best * 1e3, nsub / dt, os.cpu_count(), float((st == 0).float().mean()), ib[:, 0].mean(), ib[:, 1].mean(), ok), flush=True)
