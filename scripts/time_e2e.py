"""Time solver.solve() on pinned host buffers (the e2e leg of bench.py) for one workload. usage: time_e2e.py [C2] [B]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import nonlin_b200 as nb
from nonlin_b200 import workloads as W
name = sys.argv[1] if len(sys.argv) > 1 else "C2"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20
w = W.WORKLOADS[name](B)
obj = nb.vecfcn_helper(); obj.set_fcn(w["fcn"], w["m"], w["n"])
s = {"least_squares": nb.least_squares_solver, "newton": nb.newton_solver, "quasi_newton": nb.quasi_newton_solver}[w["solver"]]()
for k, v in w["settings"].items(): getattr(s, k)(v)
pin = lambda a: torch.from_numpy(a).pin_memory()
x0 = pin(w["x0"]); args = None if w["args"] is None else pin(w["args"])
x = torch.empty_like(x0).pin_memory(); f = torch.empty((w["m"], B), dtype=torch.float64).pin_memory()
ib = torch.zeros((B, 7), dtype=torch.int32).pin_memory(); st = torch.zeros(B, dtype=torch.int32).pin_memory()
best = 1e30
for k in range(8):
    x.copy_(x0); torch.cuda.synchronize()
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record(); s.solve(obj, x, f, ib, args=args, status=st, stream=1); b.record(); torch.cuda.synchronize()
    if k >= 2: best = min(best, a.elapsed_time(b))
d2h = x.numel() * 8 + f.numel() * 8 + ib.numel() * 4 + st.numel() * 4
print("%s chunks=%s: %.3f ms -> %.3e systems/s, D2H %.1f MB -> %.1f GB/s if D2H-bound" % (name, os.environ.get("NLB_HOST_CHUNKS", "default"), best, B / best * 1e3, d2h / 1e6, d2h / best / 1e6))
# raw pinned D2H rate for reference
d = torch.empty(d2h // 8, dtype=torch.float64, device="cuda"); h = torch.empty(d2h // 8, dtype=torch.float64).pin_memory()
torch.cuda.synchronize(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
h.copy_(d, non_blocking=True); torch.cuda.synchronize()
a.record(); h.copy_(d, non_blocking=True); b.record(); torch.cuda.synchronize()
print("raw pinned D2H of the same bytes: %.3f ms (%.1f GB/s)" % (a.elapsed_time(b), d2h / a.elapsed_time(b) / 1e6))
