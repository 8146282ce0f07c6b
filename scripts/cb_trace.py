"""Phase timeline of coop_broyden.cuh's CTA 0 (debug build -DNLB_CB_TRACE, loaded through NLB_LIB).
usage: NLB_LIB=build/libnlb_cbtrace.so python scripts/cb_trace.py [nsystems]"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import nonlin_b200 as nb
from nonlin_b200 import workloads as W
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
w = W.c5_broyden_rosenbrock(max(B, 1))
obj = nb.vecfcn_helper(); obj.set_fcn(w["fcn"], w["m"], w["n"])
s = nb.quasi_newton_solver()
x = w["x0"].copy(); f = np.zeros_like(x); ib = nb.iteration_behavior(x.shape[1])
st = s.solve(obj, x, f, ib)
print("system 0: iter %d nfev %d njac %d" % (ib["iter_count"][0], ib["fcn_count"][0], ib["jacobian_count"][0]))
lib = ctypes.CDLL(os.environ["NLB_LIB"])
stamps = np.zeros((64, 32), dtype=np.int64)
print("rc", lib.nlb_debug_cb_trace(stamps.ctypes.data_as(ctypes.c_void_p)))
names = {0: "iter start", 1: "FD jacobian done", 2: "qr_full done", 3: "B update done", 4: "rank1 update done", 5: "(pre B^T f)", 6: "B^T f, Q^T f done",
         7: "trsv done", 8: "ls setup done", 9: "line search done", 10: "conv test done"}
inner = {16: "Q^T u", 17: "r-chain", 18: "c,s", 19: "DQRQH", 20: "DQROT B", 21: "row0 + DQHQR"}
for it in (3, 4):
    row = stamps[it]
    print("iter %d rank-1 update: " % it + "  ".join("%s %d" % (inner[p], row[p] - (row[p - 1] if p > 16 else row[3])) for p in range(16, 22)) + "  DQROT F %d" % (row[4] - row[21]))
for it in range(1, 4):
    row = stamps[it]
    ev = sorted((int(row[p]), p) for p in names if row[p] != 0)
    if not ev: continue
    t0 = ev[0][0]
    print("iter %2d: total %7d | " % (it, ev[-1][0] - t0) + "  ".join("%s +%d" % (names[p], t - t0) for t, p in ev[1:]))
