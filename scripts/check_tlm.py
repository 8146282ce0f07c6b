"""Tall LM kernel (tall_lm.cuh) against the CPU oracle, then a timing.  usage: check_tlm.py [B_time]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import nonlin_b200 as nb
from nonlin_b200 import workloads as W
from oracle.nl_oracle import Oracle

o = Oracle()
def solve(w):
    obj = nb.vecfcn_helper(); obj.set_fcn(w["fcn"], w["m"], w["n"]); obj.set_shared_data(w["shared"])
    s = nb.least_squares_solver(); s.set_max_fcn_evals(1000)
    x = w["x0"].copy(); f = np.zeros((w["m"], x.shape[1])); ib = nb.iteration_behavior(x.shape[1])
    st = s.solve(obj, x, f, ib, args=w["args"])
    return x, f, ib, st
bad = 0
for (B, m, noise, fcn) in [(8, 512, 0.0, None), (40, 640, 0.0, None), (24, 1000, 1e-3, None), (64, 4096, 0.0, None), (64, 4096, 1e-3, None),
                           (16, 1536, 1e-2, "exp_sum_8")]:
    w = W.c4_lm_rational(B, m=m, noise=noise)
    if fcn:
        w["fcn"] = fcn
    t0 = time.time(); x, f, ib, st = solve(w); dt = time.time() - t0
    xo, fo, ibo, sto = o.solve_batch("least_squares", w["fcn"], w["x0"], m=m, sys=w["args"], shared=w["shared"], params=o.params(max_fcn_evals=1000))
    same = np.array_equal(x, xo) and np.array_equal(f, fo) and np.array_equal(ib, ibo) and np.array_equal(st, sto)
    nbad = int((~((x == xo).all(axis=0) & (f == fo).all(axis=0) & (ib == ibo))).sum())
    print("B=%d m=%d noise=%g %s: bit-identical=%s (%d systems differ) njac mean %.1f max %d  %.2fs" % (B, m, noise, w["fcn"], same, nbad, ibo["jacobian_count"].mean(), ibo["jacobian_count"].max(), dt), flush=True)
    bad += not same
Bt = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
for noise in (0.0, 1e-3):
    w = W.c4_lm_rational(Bt, noise=noise)
    obj = nb.vecfcn_helper(); obj.set_fcn(w["fcn"], w["m"], w["n"]); obj.set_shared_data(torch.from_numpy(w["shared"]).cuda())
    s = nb.least_squares_solver(); s.set_max_fcn_evals(1000)
    x0 = torch.from_numpy(w["x0"]).cuda(); args = torch.from_numpy(w["args"]).cuda()
    f = torch.empty((w["m"], Bt), dtype=torch.float64, device="cuda"); ib = nb.iteration_behavior(Bt, like=x0); st = torch.zeros(Bt, dtype=torch.int32, device="cuda")
    for rep in range(2):
        x = x0.clone(); torch.cuda.synchronize(); t0 = time.time()
        s.solve(obj, x, f, ib, args=args, status=st); torch.cuda.synchronize(); dt = time.time() - t0
    stats = nb.default_engine(0).reduce_stats(ib, st, Bt)
    print("timing B=%d noise=%g: %.3f s -> %.0f systems/s; sum_jac %d sum_fcn %d failed %d" % (Bt, noise, dt, Bt / dt, stats["sum_jac"], stats["sum_fcn"], stats["failed"]), flush=True)
sys.exit(1 if bad else 0)
