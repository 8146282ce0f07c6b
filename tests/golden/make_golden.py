"""Regenerate tests/golden/oracle_batches.npz: outputs of the CPU oracle on small seeded batches.

    python tests/golden/make_golden.py

The GPU parity tests compare the engine with these committed vectors as well as with the oracle
run live, so a silent change of either side is caught.  (The reference itself is Fortran and
cannot be built in this image — no Fortran compiler — so these vectors come from the oracle,
which is pinned to the reference's published outputs by tests/test_oracle_kat.py.)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from nonlin_b200 import workloads as W  # noqa: E402
from oracle.nl_oracle import Oracle  # noqa: E402

CASES = {"C1": 256, "C2": 512, "C3": 512, "C5": 8, "LM4": 128, "C4": 4}
EXTRA = {"C4": {"m": 256}}


def oracle_params(o, w):
    kw = {}
    if "set_max_fcn_evals" in w["settings"]:
        kw["max_fcn_evals"] = w["settings"]["set_max_fcn_evals"]
    return o.params(**kw)


def main():
    o = Oracle()
    out = {}
    for name, B in CASES.items():
        w = W.WORKLOADS[name](B, **EXTRA.get(name, {}))
        x, f, ib, st = o.solve_batch(w["solver"], w["fcn"], w["x0"], m=w["m"], sys=w["args"], shared=w["shared"],
                                     params=oracle_params(o, w))
        out[name + "_x"] = x
        out[name + "_f"] = f
        out[name + "_ib"] = ib.view(np.int32).reshape(B, 7)
        out[name + "_status"] = st
        print(name, "B", B, "converged", (st == 0).mean(), "mean iter", ib["iter_count"].mean())
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_batches.npz"), **out)


if __name__ == "__main__":
    main()
