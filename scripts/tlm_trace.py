"""Per-segment timeline of tall_lm.cuh's CTA 0 (debug build -DNLB_TLM_TRACE=<pass>, loaded through NLB_LIB).
usage: NLB_LIB=build/libnlb_trace.so python scripts/tlm_trace.py [nsystems]"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import nonlin_b200 as nb
from nonlin_b200 import workloads as W
nrep = int(sys.argv[1]) if len(sys.argv) > 1 else 1
w = W.c4_lm_rational(64)
obj = nb.vecfcn_helper(); obj.set_fcn(w["fcn"], w["m"], w["n"]); obj.set_shared_data(torch.from_numpy(w["shared"]).cuda())
s = nb.least_squares_solver(); s.set_max_fcn_evals(8)
idx = torch.from_numpy(np.resize(np.arange(64), nrep)).cuda()
x0 = torch.from_numpy(w["x0"]).cuda()[:, idx].contiguous(); args = torch.from_numpy(w["args"]).cuda()[:, idx].contiguous()
f = torch.empty((4096, nrep), dtype=torch.float64, device="cuda"); ib = nb.iteration_behavior(nrep, like=x0)
st = torch.zeros(nrep, dtype=torch.int32, device="cuda")
x = x0.clone(); s.solve(obj, x, f, ib, args=args, status=st); torch.cuda.synchronize()
lib = ctypes.CDLL(os.environ["NLB_LIB"])
stamps = np.zeros((16, 512), dtype=np.int64); counts = np.zeros(8, dtype=np.int32)
rc = lib.nlb_debug_tlm_trace(stamps.ctypes.data_as(ctypes.c_void_p), counts.ctypes.data_as(ctypes.c_void_p))
print("rc", rc)
names = ["P:full_in", "P:empty_p", "P:rowop", "P:bar1", "IO:done", "C:full_p", "C:summed", "probe7", "probe8", "probe9", "probe10", "probe11", "probe12", "probe13"]
n = 128
t0 = stamps[0, 0]
half = n // 2
for lo, hi, tag in [(2, half, "first traced pass"), (half + 2, n, "second traced pass")]:
    if hi - lo < 4: continue
    print("== %s: per-segment period and offsets relative to the producer's full_in stamp (cycles, median over %d segments)" % (tag, hi - lo))
    per = np.diff(stamps[0, lo:hi]); print("   period between segments: median %d  min %d  max %d" % (np.median(per), per.min(), per.max()))
    for p in range(1, 14):
        if not stamps[p, lo:hi].any(): continue
        d = stamps[p, lo:hi] - stamps[0, lo:hi]
        print("   %-10s +%6d (min %d max %d)" % (names[p], np.median(d), d.min(), d.max()))
print("raw first 12 segments of each probe (relative to first stamp):")
for p in range(0): print("  ", names[p], (stamps[p, :12] - t0).tolist())
