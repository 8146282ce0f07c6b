import os, subprocess, sys, glob
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
libs = [None] + sorted(glob.glob(os.path.join(root, "build", "variants", "*.so")))
code = r'''
import sys, os; sys.path.insert(0, %r)
import torch, nonlin_b200 as nb
from nonlin_b200 import workloads as W
out = []
for name in sys.argv[1:]:
    B = 1 << 20
    w = W.WORKLOADS[name](B)
    obj = nb.vecfcn_helper(); obj.set_fcn(w["fcn"], w["m"], w["n"])
    s = {"least_squares": nb.least_squares_solver, "newton": nb.newton_solver, "quasi_newton": nb.quasi_newton_solver}[w["solver"]]()
    for k, v in w["settings"].items(): getattr(s, k)(v)
    x0 = torch.from_numpy(w["x0"]).cuda(); args = None if w["args"] is None else torch.from_numpy(w["args"]).cuda()
    f = torch.empty((w["m"], B), dtype=torch.float64, device="cuda"); ib = nb.iteration_behavior(B, like=x0); st = torch.zeros(B, dtype=torch.int32, device="cuda")
    best = 1e30
    for it in range(6):
        x = x0.clone(); torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); s.solve(obj, x, f, ib, args=args, status=st); e1.record(); torch.cuda.synchronize()
        if it >= 2: best = min(best, e0.elapsed_time(e1))
    out.append("%%s %%.3f ms" %% (name, best))
print(" | ".join(out))
''' % root
for lib in libs:
    env = dict(os.environ)
    if lib: env["NLB_LIB"] = lib
    r = subprocess.run([sys.executable, "-c", code] + sys.argv[1:], env=env, capture_output=True, text=True)
    print("%-22s %s %s" % (os.path.basename(lib) if lib else "default", r.stdout.strip(), r.stderr.strip()[-200:]), flush=True)
