"""How far does an FMA build move the results?  Runs the BASELINE workloads through whichever engine library NLB_LIB
names (default: the parity build) and reports, against the CPU oracle on the same systems, the north_star statistics:
share of converged systems with x and f within 1e-10 relative, share with equal (iter, nfev, njac), share bit-identical,
plus the device time of the solve.  usage: [NLB_LIB=nonlin_b200/libnonlin_b200_fast.so] fast_build_stats.py OUT.json"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import nonlin_b200 as nb
from nonlin_b200 import workloads as W
from oracle.nl_oracle import Oracle
import bench

o = Oracle()
out = {"lib": os.environ.get("NLB_LIB", "nonlin_b200/libnonlin_b200.so")}
for name, B, nchk in (("C1", 1 << 20, 1 << 17), ("C2", 1 << 20, 1 << 17), ("C3", 1 << 20, 1 << 17), ("C5", 16384, 16384),
                      ("LM4", 1 << 18, 1 << 15), ("C4", 1184, 512)):
    w = W.WORKLOADS[name](B)
    obj = nb.vecfcn_helper(); obj.set_fcn(w["fcn"], w["m"], w["n"])
    if w["shared"] is not None: obj.set_shared_data(torch.from_numpy(w["shared"]).cuda())
    s = bench.make_solver(nb, w, nb.default_engine(0))
    x0 = torch.from_numpy(w["x0"]).cuda(); args = None if w["args"] is None else torch.from_numpy(w["args"]).cuda()
    f = torch.empty((w["m"], B), dtype=torch.float64, device="cuda"); ib = nb.iteration_behavior(B, like=x0)
    st = torch.zeros(B, dtype=torch.int32, device="cuda")
    best = 1e30
    for rep in range(3):
        x = x0.clone(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); s.solve(obj, x, f, ib, args=args, status=st); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    n = nchk
    sysd = None if w["args"] is None else np.ascontiguousarray(w["args"][:, :n])
    ref = bench.oracle_solve_batch(o, w, np.ascontiguousarray(w["x0"][:, :n]), sysd, params=bench.oracle_params(o, w))
    eng = (x[:, :n].cpu().numpy(), f[:, :n].cpu().numpy(), ib[:n].cpu().numpy().reshape(-1, 7), st[:n].cpu().numpy())
    rep = bench.parity_report(eng, ref)
    rep.update(ms=best, systems_per_s=B / best * 1e3, batch=B)
    out[name] = rep
    print(name, json.dumps(rep), flush=True)
json.dump(out, open(sys.argv[1], "w"), indent=1)
