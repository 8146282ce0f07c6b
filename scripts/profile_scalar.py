"""Run the Brent solver a few times (for ncu). usage: profile_scalar.py [B]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import nonlin_b200 as nb
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
rng = np.random.default_rng(8)
args = torch.from_numpy(np.stack([rng.uniform(-8, -1, B), rng.uniform(-2, 2, B), rng.uniform(-1, 1, B), rng.uniform(0.5, 2, B)])).cuda()
obj = nb.fcn1var_helper(); obj.set_fcn("cubic_args")
x = torch.zeros(B, dtype=torch.float64, device="cuda")
lim = nb.value_pair(torch.full((B,), -1.0, dtype=torch.float64, device="cuda"), torch.full((B,), 6.0, dtype=torch.float64, device="cuda"))
for _ in range(3):
    st = nb.brent_solver().solve(obj, x, lim, args=args)
torch.cuda.synchronize()
print(int((st == 0).sum().item()))
