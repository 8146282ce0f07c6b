"""One C4 solve at the BASELINE batch (65 536 x m = 4096 x n = 16), observations formed on the device as bench.py does.
usage: profile_c4_full.py [C4|C4N] [B]   (for ncu: one launch of tlm_kernel)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
import nonlin_b200 as nb
name = sys.argv[1] if len(sys.argv) > 1 else "C4"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
W = bench.load_workloads()
eng = nb.default_engine(0)
dev = torch.device("cuda", 0)
w, y = bench.build_workload(W, torch, name, B, 1000, dev)
run = bench.DeviceRun(nb, torch, w, eng, 1, args_dev=y)
torch.cuda.synchronize(); t0 = time.time()
run.step(0)
torch.cuda.synchronize(); dt = time.time() - t0
print(name, B, "%.3f s -> %.0f systems/s" % (dt, B / dt), eng.reduce_stats(run.ib[0], run.status[0], B))
