"""VERDICT r1 #2a / SURVEY 7 option (i): would a summation order other than the reference's be admissible?

Runs the CPU port of the reference's LM solver twice on the same systems - sequential (Fortran-order) sums, then every
m-length sum in warp-shuffle order (32 striped partial sums + butterfly; norms as sqrt of the plain sum of squares) - and
reports how many systems keep their (iter, nfev, njac) triple, their status, and x / f within 1e-10 relative.
north_star's bar: counts equal on >= 99 % of the systems, x and f within 1e-10.   usage: striped_sum_study.py [B]"""
import importlib.util, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import nl_oracle
spec = importlib.util.spec_from_file_location("wl", os.path.join(ROOT, "nonlin_b200", "workloads.py"))
W = importlib.util.module_from_spec(spec); spec.loader.exec_module(W)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
o = nl_oracle.Oracle()
out = {}
for name, w in [("C4 (m=4096, n=16, noise-free)", W.c4_lm_rational(B)), ("C4N (sigma=1e-3)", W.c4_lm_rational(B, noise=1e-3)),
                ("LM4 (exp_decay_4, m=64)", W.lm_expdecay4(B * 16))]:
    prm = o.params(max_fcn_evals=1000)
    res = []
    for mode in (0, 1):
        o.set_sum_mode(mode)
        t = time.time()
        res.append(o.solve_batch(nl_oracle.LM, w["fcn"], w["x0"], m=w["m"], sys=w["args"], shared=w["shared"], params=prm))
        dt = time.time() - t
    o.set_sum_mode(0)
    (x0, f0, ib0, s0), (x1, f1, ib1, s1) = res
    counts = (ib0["iter_count"] == ib1["iter_count"]) & (ib0["fcn_count"] == ib1["fcn_count"]) & (ib0["jacobian_count"] == ib1["jacobian_count"])
    both = (s0 == 0) & (s1 == 0)
    relx = np.max(np.abs(x0 - x1) / np.maximum(np.abs(x0), 1e-300), axis=0)
    fn0, fn1 = np.linalg.norm(f0, axis=0), np.linalg.norm(f1, axis=0)
    relf = np.abs(fn0 - fn1) / np.maximum(fn0, 1e-300)
    r = {"systems": int(x0.shape[1]), "counts_equal": float(counts.mean()), "status_equal": float((s0 == s1).mean()),
         "converged_seq": float((s0 == 0).mean()), "converged_striped": float((s1 == 0).mean()),
         "x_within_1e-10_on_both_converged": float((relx[both] <= 1e-10).mean()) if both.any() else None,
         "fnorm_within_1e-10_on_both_converged": float((relf[both] <= 1e-10).mean()) if both.any() else None,
         "median_rel_dx": float(np.median(relx[both])) if both.any() else None,
         "sum_iter_seq": int(ib0["iter_count"].sum()), "sum_iter_striped": int(ib1["iter_count"].sum())}
    out[name] = r
    print(name, json.dumps(r), flush=True)
json.dump(out, open(os.path.join(ROOT, "profiles", "r2_striped_sum_study.json"), "w"), indent=1)
