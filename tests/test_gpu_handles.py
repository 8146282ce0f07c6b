"""More than one engine handle in one process (the shape a Fortran host uses: one handle per GPU).

The dynamic shared-memory opt-in and the occupancy of the cooperative kernels are per device; a second handle -
on another GPU when the box has one, else on the same GPU - must launch the > 48 KB kernels as the first does,
concurrently from two host threads, with bit-identical results, and no entry point may change the calling thread's
current CUDA device."""
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _solve(nb, eng, w):
    obj = nb.vecfcn_helper()
    obj.set_fcn(w["fcn"], w["m"], w["n"])
    if w["shared"] is not None:
        obj.set_shared_data(w["shared"])
    s = {"least_squares": nb.least_squares_solver, "quasi_newton": nb.quasi_newton_solver,
         "newton": nb.newton_solver}[w["solver"]](engine=eng)
    for k, v in w["settings"].items():
        getattr(s, k)(v)
    x = w["x0"].copy()
    f = np.zeros((w["m"], x.shape[1]))
    ib = nb.iteration_behavior(x.shape[1])
    st = s.solve(obj, x, f, ib, args=w["args"])
    return x, f, ib, st


def test_two_handles_one_process(engine):
    import torch

    import nonlin_b200 as nb
    from nonlin_b200 import workloads as W

    ndev = torch.cuda.device_count()
    second = nb.Engine(1 if ndev > 1 else 0)
    try:
        cases = [W.WORKLOADS["C5"](64, n=64), W.WORKLOADS["C4"](24, m=640), W.WORKLOADS["LM4"](96),
                 W.WORKLOADS["C1"](2048), W.WORKLOADS["C3"](4096)]
        ref = [_solve(nb, engine, w) for w in cases]
        torch.cuda.set_device(0)
        out = [None] * len(cases)
        errs = []

        def worker(eng, slot):
            try:
                slot[:] = [_solve(nb, eng, w) for w in cases]
            except Exception as ex:      # pragma: no cover
                errs.append(ex)

        a, b = [None] * len(cases), [None] * len(cases)
        ta = threading.Thread(target=worker, args=(engine, a))
        tb = threading.Thread(target=worker, args=(second, b))
        ta.start(); tb.start(); ta.join(); tb.join()
        assert not errs, errs
        for got in (a, b):
            for r, g in zip(ref, got):
                for u, v in zip(r, g):
                    assert np.array_equal(u, v)
        assert torch.cuda.current_device() == 0      # the C ABI restores the caller's device
    finally:
        second.close()


def test_sharded_solve_over_all_devices(engine, oracle):
    """nlb_solve_sharded: one host batch over every GPU of the box (one handle per device, one host thread each, NCCL
    all-reduce of the statistics when there is more than one) gives the bits of the single-device solve."""
    import torch

    import nonlin_b200 as nb
    from nonlin_b200 import workloads as W

    ndev = torch.cuda.device_count()
    engines = [engine] + [nb.Engine(d) for d in range(1, ndev)]
    try:
        for name, B in (("C2", 10007), ("C1", 4099), ("C3", 5000)):
            w = W.WORKLOADS[name](B)
            ref = _solve(nb, engine, w)
            obj = nb.vecfcn_helper()
            obj.set_fcn(w["fcn"], w["m"], w["n"])
            s = {"least_squares": nb.least_squares_solver, "quasi_newton": nb.quasi_newton_solver,
                 "newton": nb.newton_solver}[w["solver"]]()
            for k, v in w["settings"].items():
                getattr(s, k)(v)
            x = w["x0"].copy()
            f = np.zeros((w["m"], B))
            ib = nb.iteration_behavior(B)
            st, stats = s.solve_sharded(engines, obj, x, f, ib, args=w["args"])
            for u, v in zip(ref, (x, f, ib, st)):
                assert np.array_equal(u, v)
            assert stats["systems"] == B and stats["converged"] == int((st == 0).sum())
            assert stats["sum_iter"] == int(ib["iter_count"].sum()) and stats["max_iter"] == int(ib["iter_count"].max())
            # outputs that are not asked for are not produced
            x2 = w["x0"].copy()
            st2, none = s.solve_sharded(engines, obj, x2, None, None, args=w["args"], want_stats=False)
            assert none is None and np.array_equal(x2, x) and np.array_equal(st2, st)
    finally:
        for e in engines[1:]:
            e.close()


def test_optional_outputs_single_device(engine):
    import nonlin_b200 as nb
    from nonlin_b200 import workloads as W

    w = W.WORKLOADS["C2"](70001)
    ref = _solve(nb, engine, w)
    obj = nb.vecfcn_helper()
    obj.set_fcn(w["fcn"], w["m"], w["n"])
    s = nb.quasi_newton_solver(engine=engine)
    x = w["x0"].copy()
    st = s.solve(obj, x, want_fvec=False)
    assert np.array_equal(x, ref[0]) and np.array_equal(st, ref[3])
