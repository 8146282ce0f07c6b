"""Run the batched polynomial fit a few times (for ncu). usage: profile_polyfit.py [B] [npts] [order] [reps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import nonlin_b200 as nb
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
npts = int(sys.argv[2]) if len(sys.argv) > 2 else 21
order = int(sys.argv[3]) if len(sys.argv) > 3 else 3
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
rng = np.random.default_rng(1)
x = torch.from_numpy(np.linspace(0.0, 2.0, npts)).cuda()
y = torch.from_numpy(rng.standard_normal((npts, B))).cuda()
p = nb.polynomial()
for _ in range(reps):
    st = p.fit(x, y, order)
torch.cuda.synchronize()
print(int((st == 0).sum().item()))
