#!/usr/bin/env python
"""bench.py — converged systems / second of the batched nonlinear-solve hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload C1] [--impl engine|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path over one synthetic batch: one solve kernel over the batch
(inputs already resident in HBM) followed by the statistics kernel that counts the converged
systems.  The headline workload is the one BASELINE.json's north_star quotes its target on:
2^20 four-parameter Levenberg-Marquardt curve fits (C1 = README Example 2 batched, m = 21, n = 4).
With N GPUs every rank solves its own 2^20-system shard (weak scaling, no data-path collective);
the per-rank statistics are combined with one 128-byte all-gather at the end of the job (inside
the timed region).  A strong-scaling run of one fixed 2^20 batch split over the ranks is reported
beside it (`roofline.strong_scaling`).

One JSON line is printed by rank 0 (schema: the bench contract in the task description).
  value      converged systems / s, all ranks, device-resident inputs, CUDA-event time (max over ranks)
  e2e        same metric through the public solver object with pinned HOST buffers: H2D of x0 (and
             per-system data) and D2H of x, fvec, iteration_behavior, status inside the timed region
  roofline   dominant kernel (the solve kernel) against the FP64 pipe: algorithmic FP64 operations
             per system (counted by the oracle's counting build on a sample of the same batch) x
             systems / CUDA-event time of that kernel inside the timed region, against the DFMA peak
             measured on this GPU; plus the HBM side for reference.
             roofline.per_config holds one block per BASELINE configuration at its BASELINE batch size
             (C2, C3, C4 at 65 536, its sigma = 1e-3 variant C4N, C5): value, roofline, cpu_baseline,
             e2e and parity against the CPU port each.
  cpu_baseline  the CPU oracle (a C++ port; the Fortran reference cannot be built in this image) on
             all host cores, bounded sample of the same batch
`--impl reference` times that CPU port alone and prints the same line with "impl": "reference".
It never imports the engine package (the workload generators are loaded by path).
"""
import argparse
import importlib.util
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "converged systems/sec"
UNIT = "systems/s"
HEADLINE = "C1"
# BASELINE.json batch sizes
BATCH = {"C1": 1 << 20, "C2": 1 << 20, "C3": 1 << 20, "C4": 65536, "C4N": 65536, "C5": 16384, "LM4": 1 << 20,
         "CLS1": 1 << 20, "CLS2": 1 << 20}
EXTRA_CONFIGS = ("C2", "C3", "C4", "C4N", "C5")
KERNEL = {"C1": "tps_solve_kernel<LsqPolyFit, LM>", "C2": "tps_solve_kernel<Misc2Fcn, Broyden>",
          "C3": "tps_newton_refill_kernel<PowellBadlyScaled>", "C4": "tlm_kernel<Rational78, 16>",
          "C4N": "tlm_kernel<Rational78, 16>", "C5": "coop_broyden_kernel<ExtRosenbrock, 64>",
          "LM4": "coop_lm_kernel<ExpDecay4, 4>", "CLS1": "tps_cls_kernel<LsqPolyFit>", "CLS2": "tps_cls_kernel<Misc2Fcn>"}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--workload", default=HEADLINE)
    ap.add_argument("--batch", type=int, default=0, help="systems per GPU of the headline workload (default: BASELINE size)")
    ap.add_argument("--configs", default=",".join(EXTRA_CONFIGS),
                    help="BASELINE configurations reported under roofline.per_config (N = 1 only)")
    ap.add_argument("--c4-batch", type=int, default=0, help="override the C4 / C4N batch (default 65536)")
    ap.add_argument("--no-extras", action="store_true", help="headline only")
    return ap.parse_args()


def load_workloads():
    """nonlin_b200/workloads.py loaded by path: pure numpy generators, usable without importing the engine package
    (whose __init__ loads libnonlin_b200.so) - the reference arm must not map the engine's library."""
    spec = importlib.util.spec_from_file_location("_nlb_workloads", os.path.join(ROOT, "nonlin_b200", "workloads.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def workload_label(name, B):
    desc = {
        "C1": "README Ex.2 cubic LM fit m=21 n=4, least_squares_solver, per-system y",
        "C2": "README Ex.1 2x2, quasi_newton_solver, perturbed x0",
        "C3": "Powell badly scaled 2x2, newton_solver + line search, max evals 1000",
        "C4": "LM curve fits m=4096 n=16, rational 7/8 model, noise-free, max evals 1000",
        "C4N": "LM curve fits m=4096 n=16, rational 7/8 model, sigma=1e-3, max evals 1000",
        "C5": "extended Rosenbrock n=64, quasi_newton_solver + line search",
        "LM4": "4-parameter double-exponential LM fits, m=64",
        "CLS1": "README Ex.2 cubic fit, constrained_least_squares_solver, limits [-10,10]",
        "CLS2": "README Ex.1 2x2, constrained_least_squares_solver, box [0,6]^2",
    }[name]
    return "%s: %d x %s" % (name, B, desc)


def make_config(name, B, world):
    """The `config` object of the JSON line; both arms emit exactly this for the same workload."""
    return {"workload": workload_label(name, B), "batch_per_gpu": B,
            "parallelism": "dp%d, contiguous system shards, no data-path collective" % world,
            "l2": "inputs larger than L2: fresh x0 copy per step, buffer rings >= 768 MiB",
            "arithmetic": "FP64 without FMA contraction (bit-identical to the CPU port)"}


# ---------------------------------------------------------------------------------------------
# clocks sampler (NVML in a thread; nvidia-smi would be too slow for a sub-second timed region)
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {
        0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
        0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
        0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting",
    }

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if r & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        if self.nv:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()

    def stop(self):
        self._stop.set()
        if self._thr:
            self._thr.join()
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ---------------------------------------------------------------------------------------------
# CPU oracle legs
# ---------------------------------------------------------------------------------------------
def oracle_params(o, w):
    kw = {}
    if "set_max_fcn_evals" in w["settings"]:
        kw["max_fcn_evals"] = w["settings"]["set_max_fcn_evals"]
    return o.params(**kw)


def oracle_solve_batch(o, w, x0, sysd, **kw):
    """The oracle's batch solve for the workload's solver (constrained least squares has its own entry)."""
    if w["solver"] == "constrained_least_squares":
        return o.cls_solve_batch(w["fcn"], x0, m=w["m"], sys=sysd, shared=w["shared"],
                                 lower=w["settings"].get("set_lower_limits"), upper=w["settings"].get("set_upper_limits"), **kw)
    return o.solve_batch(w["solver"], w["fcn"], x0, m=w["m"], sys=sysd, shared=w["shared"], **kw)


def parity_report(engine_out, oracle_out):
    """Engine results against the CPU port's on the same systems (the port as the checker): the north_star bar is
    x and f within 1e-10 relative on converged systems and equal iteration / evaluation / Jacobian counts on >= 99 %."""
    x, f, ib, st = engine_out
    xo, fo, ibo, sto = oracle_out
    n = st.shape[0]
    ibo = ibo.view(np.int32).reshape(n, 7)
    ok = sto == 0
    sx = np.maximum(np.abs(xo).max(axis=0), 1e-300)
    sf = np.maximum(np.abs(fo).max(axis=0), 1e-300)
    close = (np.abs(x - xo).max(axis=0) <= 1e-10 * sx) & (np.abs(f - fo).max(axis=0) <= 1e-10 * sf + 1e-300)
    bit = (x == xo).all(axis=0) & (f == fo).all(axis=0)
    counts = (ib[:, :3] == ibo[:, :3]).all(axis=1)
    return {"systems_checked": int(n), "status_equal": float((st == sto).mean()),
            "x_f_within_1e-10_on_converged": float(close[ok].mean()) if ok.any() else None,
            "x_f_bit_identical": float(bit.mean()), "iter_nfev_njac_equal": float(counts.mean()),
            "flags_equal": float((ib[:, 4:] == ibo[:, 4:]).all(axis=1).mean())}


def cpu_port_throughput(w, x0, sysd, min_seconds=4.0, max_reps=50, engine_out=None):
    """Time the CPU oracle (OpenMP, all host cores) on (x0, sysd) = the workload's batch or a slice of it.  With
    engine_out = (x, f, ib[n,7], status) of the engine on the same systems, the port's results double as the parity check."""
    from oracle.nl_oracle import Oracle

    o = Oracle()
    nsub = x0.shape[1]
    p = oracle_params(o, w)
    cores = os.cpu_count() or 1
    wn = min(nsub, 256 if w["m"] >= 512 else 4096)
    oracle_solve_batch(o, w, x0[:, :wn].copy(), None if sysd is None else sysd[:, :wn].copy(), params=p)
    reps, elapsed, conv = 0, 0.0, 0
    while (elapsed < min_seconds and reps < max_reps) or reps < 1:
        t0 = time.perf_counter()
        xo, fo, ibo, st = oracle_solve_batch(o, w, x0, sysd, params=p, nthreads=cores)
        elapsed += time.perf_counter() - t0
        conv += int((st == 0).sum())
        reps += 1
    out = {"value": conv / elapsed, "unit": UNIT, "cores": cores, "kind": "port",
           "sample": "%d systems x %d passes, OpenMP dynamic, one system per thread" % (nsub, reps)}
    if engine_out is not None:
        try:
            out["parity"] = parity_report(engine_out, (xo, fo, ibo, st))
        except Exception as ex:      # the parity report must never hide the measurement
            out["parity"] = {"error": repr(ex)}
    return out


def flops_per_system(w, x0, sysd, sample):
    """Algorithmic FP64 operations (+ - * / sqrt, exp = 25) per system: counting build of the oracle."""
    from oracle.nl_oracle import Oracle

    o = Oracle(counting=True)
    B = x0.shape[1]
    idx = np.linspace(0, B - 1, min(sample, B)).astype(np.int64)
    xs = np.ascontiguousarray(x0[:, idx])
    ss = None if sysd is None else np.ascontiguousarray(sysd[:, idx])
    o.flops_reset()
    oracle_solve_batch(o, w, xs, ss, params=oracle_params(o, w))
    return o.flops_total() / float(idx.size)


def reference_sample(name, B):
    """Systems per step of the CPU arm: the whole batch for the small systems, a bounded slice for C4 / C5."""
    return {"C4": 1024, "C4N": 1024}.get(name, B)


def run_reference(args):
    """--impl reference: the reference's CPU path.  The Fortran sources need gfortran + the external
    linalg package, neither of which exists in this image, so this arm times the C++ port (oracle/)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    W = load_workloads()
    from oracle.nl_oracle import Oracle

    name = args.workload
    B = args.batch or BATCH[name]
    nsub = min(B, reference_sample(name, B))
    w = W.WORKLOADS[name](nsub if nsub < B else B, seed=1000)
    o = Oracle()
    p = oracle_params(o, w)
    cores = os.cpu_count() or 1
    x0 = np.ascontiguousarray(w["x0"][:, :nsub])
    sysd = None if w["args"] is None else np.ascontiguousarray(w["args"][:, :nsub])
    for _ in range(args.warmup):
        oracle_solve_batch(o, w, x0, sysd, params=p, nthreads=cores)
    t0 = time.perf_counter()
    conv = 0
    for _ in range(args.steps):
        _, _, _, st = oracle_solve_batch(o, w, x0, sysd, params=p, nthreads=cores)
        conv += int((st == 0).sum())
    dt = time.perf_counter() - t0
    val = conv / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": make_config(name, B, args.gpus),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "%d systems per step, OpenMP over all host cores; C++ port (no Fortran toolchain here)" % nsub},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------
# engine arm
# ---------------------------------------------------------------------------------------------
def make_solver(nb, w, eng):
    s = {"least_squares": nb.least_squares_solver, "newton": nb.newton_solver, "quasi_newton": nb.quasi_newton_solver,
         "constrained_least_squares": nb.constrained_least_squares_solver}[w["solver"]](engine=eng)
    for k, v in w["settings"].items():
        getattr(s, k)(v)
    return s


def c4_observations_device(torch, w, dev, chunk=4096):
    """The m x B observations of a C4 workload built with with_y=False, formed on the device with the Horner
    recurrence of workloads._rational (separate multiply and add kernels: no contraction, same bits as numpy).
    Noise (C4N) comes from torch's generator with a fixed seed."""
    truth = torch.from_numpy(w["truth"]).to(dev)
    t = torch.from_numpy(w["shared"]).to(dev)
    m, B = w["m"], truth.shape[1]
    y = torch.empty((m, B), dtype=torch.float64, device=dev)
    gen = torch.Generator(device=dev)
    gen.manual_seed(4242)
    for c0 in range(0, B, chunk):
        c1 = min(B, c0 + chunk)
        p, q = truth[:8, c0:c1], truth[8:, c0:c1]
        num = torch.zeros((c1 - c0, m), dtype=torch.float64, device=dev)
        den = torch.zeros_like(num)
        for k in range(7, -1, -1):
            num = num * t[None, :]
            num = num + p[k][:, None]
            den = den * t[None, :]
            den = den + q[k][:, None]
        den = t[None, :] * den
        den = 1.0 + den
        out = num / den
        if w.get("noise"):
            out = out + w["noise"] * torch.randn(out.shape, dtype=torch.float64, device=dev, generator=gen)
        y[:, c0:c1] = out.T
    return y


class DeviceRun:
    """One workload resident on one GPU, ready to be stepped.

    Cache hygiene: every step reads a fresh copy of x0 and writes its own fvec / ib / status buffers, taken
    from rings whose total size is several times the 126 MB L2 (or single buffers that are), so no step finds its
    inputs or outputs in L2 ("inputs larger than L2"); nothing but engine kernels runs inside the timed region."""

    RING_BYTES = 768 << 20

    def __init__(self, nb, torch, w, eng, nsteps, args_dev=None):
        self.nb, self.torch, self.w, self.eng = nb, torch, w, eng
        self.B = B = w["x0"].shape[1]
        dev = torch.device("cuda", eng.device)
        self.obj = nb.vecfcn_helper(); self.obj.set_fcn(w["fcn"], w["m"], w["n"])
        if w["shared"] is not None:
            self.obj.set_shared_data(torch.from_numpy(w["shared"]).to(dev))
        self.solver = make_solver(nb, w, eng)
        self.x0 = torch.from_numpy(w["x0"]).to(dev)
        has_args = args_dev is not None or w["args"] is not None
        per_step = 8 * B * (w["n"] + w["m"]) + 32 * B + (8 * B * w["m"] if has_args and w["m"] > 2 else 0)
        big = per_step >= self.RING_BYTES
        self.nring = 1 if big else max(2, min(nsteps, -(-self.RING_BYTES // per_step)))
        # the solve is in place: one fresh copy of x0 per step, made before the timed region
        self.xs = [self.x0.clone() for _ in range(nsteps)]
        if args_dev is not None:
            self.args = [args_dev]
        else:
            self.args = None if w["args"] is None else [torch.from_numpy(w["args"]).to(dev) for _ in range(self.nring)]
        self.f = [torch.empty((w["m"], B), dtype=torch.float64, device=dev) for _ in range(self.nring)]
        self.ib = [nb.iteration_behavior(B, like=self.x0) for _ in range(self.nring)]
        self.status = [torch.zeros(B, dtype=torch.int32, device=dev) for _ in range(self.nring)]
        self.stats = torch.zeros(16, dtype=torch.int64, device=dev)

    def step(self, k):
        r = k % self.nring
        self.solver.solve(self.obj, self.xs[k], self.f[r], self.ib[r],
                          args=None if self.args is None else self.args[r % len(self.args)], status=self.status[r])
        return r

    def _full_step(self, k):
        r = self.step(k)
        self.eng.reduce_stats_device(self.ib[r], self.status[r], self.stats, self.B)

    def timed(self, steps, warmup, dist_on, sampler=None):
        """W warm-up steps, then K timed steps.  Returns (ms of the K steps, ms of their K solve kernels, launches,
        clocks, mode).  Kernels of a millisecond or more are launched eagerly with a CUDA-event pair around every solve
        kernel (so the kernel time is measured inside the timed region itself); shorter ones (launch-bound) are
        captured into two CUDA graphs - K full steps, and the K solve kernels alone - each replayed once between events."""
        torch = self.torch
        ev = lambda: torch.cuda.Event(enable_timing=True)
        cal0, cal1 = ev(), ev()
        for k in range(warmup):
            if k == warmup - 1:
                cal0.record()
            r = self.step(k)
            if k == warmup - 1:
                cal1.record()
            self.eng.reduce_stats_device(self.ib[r], self.status[r], self.stats, self.B)
        if dist_on:
            # the first collective on a communicator pays NCCL's lazy connection set-up: do it in the warm-up
            from nonlin_b200.distributed import allreduce_stats

            allreduce_stats(self.stats.clone())
        torch.cuda.synchronize()
        use_graph = warmup > 0 and cal0.elapsed_time(cal1) < 1.0
        start, end = ev(), ev()
        graph = graph_k = None
        l0 = self.eng.kernel_launches
        if use_graph:
            try:
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                graph, graph_k = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
                with torch.cuda.stream(side):
                    with torch.cuda.graph(graph, stream=side):
                        for k in range(steps):
                            self._full_step(warmup + k)
                launches = self.eng.kernel_launches - l0
                with torch.cuda.stream(side):
                    with torch.cuda.graph(graph_k, stream=side):
                        for k in range(steps):
                            self.step(warmup + k)
                torch.cuda.current_stream().wait_stream(side)
                torch.cuda.synchronize()
            except Exception as ex:      # pragma: no cover - depends on driver support
                sys.stderr.write("bench: CUDA graph capture unavailable (%r), timing eager launches\n" % (ex,))
                graph = graph_k = None
        if sampler:
            sampler.start()
        if dist_on:
            import torch.distributed as dist

            dist.barrier()
        torch.cuda.synchronize()
        if graph is not None:
            start.record()
            graph.replay()
        else:
            e0 = [ev() for _ in range(steps)]
            e1 = [ev() for _ in range(steps)]
            l0 = self.eng.kernel_launches
            start.record()
            for k in range(steps):
                e0[k].record()
                r = self.step(warmup + k)
                e1[k].record()                         # end of the dominant (solve) kernel
                self.eng.reduce_stats_device(self.ib[r], self.status[r], self.stats, self.B)
        if dist_on:
            # the one collective of the path: the final convergence-statistics reduction of the job
            from nonlin_b200.distributed import allreduce_stats

            allreduce_stats(self.stats)
        end.record()
        torch.cuda.synchronize()
        clocks = sampler.stop() if sampler else None
        if dist_on:
            import torch.distributed as dist

            dist.barrier()
        step_ms = start.elapsed_time(end)              # K steps back to back on the device
        if graph is not None:
            # the same K solve kernels alone (x0 copies restored first: the solve is in place)
            for k in range(steps):
                self.xs[warmup + k].copy_(self.x0)
            torch.cuda.synchronize()
            a, b = ev(), ev()
            a.record()
            graph_k.replay()
            b.record()
            torch.cuda.synchronize()
            solve_ms = a.elapsed_time(b)
            mode = "CUDA graphs (K steps; K solve kernels alone), each replayed once between events"
        else:
            launches = self.eng.kernel_launches - l0
            solve_ms = sum(a.elapsed_time(b) for a, b in zip(e0, e1))
            mode = "eager launches, CUDA-event pair around every solve kernel inside the timed region"
        return step_ms, min(solve_ms, step_ms), launches, clocks, mode


def e2e_run(nb, torch, w, eng, steps, warmup, args_host=None):
    """Public API with pinned HOST buffers: H2D + kernel + D2H per step, timed with CUDA events on
    the stream the engine is told to use."""
    B = w["x0"].shape[1]
    obj = nb.vecfcn_helper(); obj.set_fcn(w["fcn"], w["m"], w["n"])
    if w["shared"] is not None:
        obj.set_shared_data(w["shared"])
    solver = make_solver(nb, w, eng)
    pin = lambda a: torch.from_numpy(a).pin_memory()
    x0 = pin(w["x0"])
    if args_host is not None:
        args = args_host
    else:
        args = None if w["args"] is None else pin(w["args"])
    x = torch.empty_like(x0).pin_memory()
    f = torch.empty((w["m"], B), dtype=torch.float64).pin_memory()
    ib = torch.zeros((B, 7), dtype=torch.int32).pin_memory()
    st = torch.zeros(B, dtype=torch.int32).pin_memory()
    stream = torch.cuda.current_stream().cuda_stream or 1
    h2d = x0.numel() * 8 + (0 if args is None else args.numel() * 8) + (0 if w["shared"] is None else w["shared"].size * 8)
    d2h = x.numel() * 8 + f.numel() * 8 + ib.numel() * 4 + st.numel() * 4
    total_ms, conv = 0.0, 0
    for k in range(warmup + steps):
        x.copy_(x0)
        torch.cuda.synchronize()
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record()
        solver.solve(obj, x, f, ib, args=args, status=st, stream=stream)
        b.record(); torch.cuda.synchronize()
        if k >= warmup:
            total_ms += a.elapsed_time(b)
            conv += int((st == 0).sum())      # device->host result read: the status array
    return total_ms, conv, h2d, d2h


def hbm_peak():
    peaks_file = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_file):
        try:
            return float(json.load(open(peaks_file))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback"


def ncu_traffic(name):
    tf = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tf):
        try:
            return json.load(open(tf)).get(name)
        except Exception:
            return None
    return None


def roofline_block(name, w, B, fl, kernel_s, peak):
    achieved_tf = fl * B / kernel_s / 1e12
    hp, hsrc = hbm_peak()
    hbm_ach = w["bytes_per_system"] * B / kernel_s / 1e9
    return {
        "bound": "fp64", "kernel": KERNEL.get(name), "achieved": achieved_tf, "peak": peak["dfma_tflops"],
        "unit": "TFLOP/s", "frac": achieved_tf / peak["dfma_tflops"],
        "peak_source": "DFMA micro-kernel on this GPU at run time (MEASURED_PEAKS.json has no FP64 entry)",
        "peak_no_fma": peak["dadd_dmul_tflops"], "frac_of_no_fma_peak": achieved_tf / peak["dadd_dmul_tflops"],
        "flops_per_system": fl, "kernel_ms": kernel_s * 1e3,
        "traffic": ncu_traffic(name),
        "hbm": {"bound": "hbm", "achieved": hbm_ach, "peak": hp, "unit": "GB/s", "frac": hbm_ach / hp,
                "bytes_per_system": w["bytes_per_system"], "peak_source": hsrc},
    }


def build_workload(W, torch, name, B, seed, dev):
    """(workload dict, device-resident observations or None).  C4 / C4N at the BASELINE batch hold 2 GB of
    observations: they are formed on the device, the host keeps only the parameters."""
    if name in ("C4", "C4N"):
        w = W.WORKLOADS[name](B, seed=seed, with_y=False)
        return w, c4_observations_device(torch, w, dev)
    return W.WORKLOADS[name](B, seed=seed), None


def host_slice(w, args_dev, n):
    """x0 and per-system data of the first n systems on the host (for the CPU port)."""
    x0 = np.ascontiguousarray(w["x0"][:, :n])
    if args_dev is not None:
        return x0, np.ascontiguousarray(args_dev[:, :n].cpu().numpy())
    return x0, None if w["args"] is None else np.ascontiguousarray(w["args"][:, :n])


def config_block(nb, torch, W, eng, name, B, steps, warmup, peak, e2e_steps):
    """One BASELINE configuration at its BASELINE batch size: device-timed value, kernel roofline, CPU port on a
    bounded sample (doubling as the parity check), end-to-end through host buffers."""
    dev = torch.device("cuda", eng.device)
    w, args_dev = build_workload(W, torch, name, B, 1000, dev)
    run = DeviceRun(nb, torch, w, eng, steps + warmup, args_dev=args_dev)
    step_ms, solve_ms, launches, _, mode = run.timed(steps, warmup, False)
    conv = int(run.stats[1].item())
    nsub = min(B, {"C4": 1024, "C4N": 1024}.get(name, 1 << 18))
    x0s, syss = host_slice(w, args_dev, nsub)
    fl = flops_per_system(w, x0s, syss, 64 if w["m"] >= 512 else 1024)
    eng_out = (run.xs[0][:, :nsub].cpu().numpy(), run.f[0][:, :nsub].cpu().numpy(),
               run.ib[0][:nsub].cpu().numpy().reshape(-1, 7), run.status[0][:nsub].cpu().numpy())
    cpu = cpu_port_throughput(w, x0s, syss, min_seconds=2.0, max_reps=8, engine_out=eng_out)
    parity = cpu.pop("parity", None)
    blk = {"workload": workload_label(name, B), "batch": B, "steps": steps, "warmup": warmup,
           "value": conv * steps / (step_ms * 1e-3), "unit": UNIT, "ms_per_step": step_ms / steps,
           "converged_per_step": conv, "gpu_launches": launches, "timing": mode,
           "roofline": roofline_block(name, w, B, fl, solve_ms * 1e-3 / steps, peak),
           "cpu_baseline": cpu, "parity": parity}
    del run
    torch.cuda.empty_cache()
    try:
        args_host = None
        if args_dev is not None:
            args_host = torch.empty(args_dev.shape, dtype=torch.float64).pin_memory()
            args_host.copy_(args_dev)
            del args_dev
            torch.cuda.empty_cache()
        e_ms, e_conv, h2d, d2h = e2e_run(nb, torch, w, eng, e2e_steps, 1, args_host=args_host)
        blk["e2e"] = {"value": e_conv / (e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                      "steps": e2e_steps}
    except Exception as ex:
        blk["e2e"] = {"error": repr(ex)}
    return blk


def run_engine(args):
    import torch

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device. The engine has no CPU fallback (use --impl reference for the CPU port).")
    torch.cuda.set_device(local)
    dist_on = world > 1
    if dist_on:
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            del os.environ["NCCL_DEBUG"]               # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import nonlin_b200 as nb

    W = load_workloads()
    eng = nb.default_engine(local)
    dev = torch.device("cuda", local)
    name = args.workload
    B = args.batch or BATCH[name]
    w, args_dev = build_workload(W, torch, name, B, 1000 + rank, dev)   # every rank its own shard (weak scaling)
    peak = eng.measure_fp64_peak()                               # also spins the clocks up
    run = DeviceRun(nb, torch, w, eng, args.steps + args.warmup, args_dev=args_dev)
    sampler = ClockSampler(local) if rank == 0 else None
    step_ms, solve_ms, launches, clocks, mode = run.timed(args.steps, args.warmup, dist_on, sampler)
    stats = run.stats.clone()
    t = torch.tensor([step_ms, solve_ms], dtype=torch.float64, device="cuda")
    if dist_on:
        import torch.distributed as dist

        dist.all_reduce(t, op=dist.ReduceOp.MAX)                 # device time = max over ranks
    step_ms, solve_ms = float(t[0]), float(t[1])
    from nonlin_b200.distributed import shard_range, stats_dict

    sd = stats_dict(stats)                                       # already summed over ranks when dist_on
    total_systems = sd["systems"]
    converged = sd["converged"]
    value = converged * args.steps / (step_ms * 1e-3)

    # results of the first step of the batch (every step solves the same systems), for the parity report
    eng_out = None
    if world == 1:
        eng_out = (run.xs[0].cpu().numpy(), run.f[0].cpu().numpy(), run.ib[0].cpu().numpy().reshape(-1, 7),
                   run.status[0].cpu().numpy())
    del run
    torch.cuda.empty_cache()

    # strong scaling: ONE fixed batch of the BASELINE size split over the ranks by contiguous ranges
    strong = None
    if dist_on:
        import torch.distributed as dist

        wf = W.WORKLOADS[name](B, seed=1000)
        lo, hi = shard_range(B, rank, world)
        ws = dict(wf)
        ws["x0"] = np.ascontiguousarray(wf["x0"][:, lo:hi])
        ws["args"] = None if wf["args"] is None else np.ascontiguousarray(wf["args"][:, lo:hi])
        srun = DeviceRun(nb, torch, ws, eng, args.steps + args.warmup)
        s_ms, _, _, _, _ = srun.timed(args.steps, args.warmup, True)
        ts = torch.tensor([s_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(ts, op=dist.ReduceOp.MAX)
        s_conv = stats_dict(srun.stats)["converged"]
        strong = {"systems_total": B, "value": s_conv * args.steps / (float(ts[0]) * 1e-3), "unit": UNIT,
                  "ms_per_step": float(ts[0]) / args.steps, "scaling": "strong",
                  "note": "one %d-system batch split over %d GPUs by contiguous ranges" % (B, world)}
        del srun
        torch.cuda.empty_cache()

    # e2e through the public API with host buffers (each rank its shard, max over ranks)
    e2e_steps = max(3, min(args.steps, 10))
    args_host = None
    if args_dev is not None:
        args_host = torch.empty(args_dev.shape, dtype=torch.float64).pin_memory()
        args_host.copy_(args_dev)
    e_ms, e_conv, h2d, d2h = e2e_run(nb, torch, w, eng, e2e_steps, 3, args_host=args_host)
    te = torch.tensor([e_ms], dtype=torch.float64, device="cuda"); tc = torch.tensor([e_conv], dtype=torch.int64, device="cuda")
    if dist_on:
        import torch.distributed as dist

        dist.all_reduce(te, op=dist.ReduceOp.MAX); dist.all_reduce(tc, op=dist.ReduceOp.SUM)
    e2e_value = float(tc[0]) / (float(te[0]) * 1e-3)

    if rank != 0:
        if dist_on:
            import torch.distributed as dist

            dist.barrier(); dist.destroy_process_group()
        return

    # roofline of the dominant kernel (solve kernel), rank 0's launches inside the timed region
    nsub = min(B, reference_sample(name, B))
    x0s, syss = host_slice(w, args_dev, nsub)
    fl = flops_per_system(w, x0s, syss, 64 if w["m"] >= 512 else 4096)
    roofline = roofline_block(name, w, B, fl, solve_ms * 1e-3 / args.steps, peak)
    roofline["note"] = "parity build issues DMUL+DADD instead of DFMA (-fmad=false): its own ceiling is peak_no_fma"
    cpu = None
    if world == 1:
        eo = tuple(a[..., :nsub] if a.ndim == 2 and a.shape[1] != 7 else a[:nsub] for a in eng_out)
        cpu = cpu_port_throughput(w, x0s, syss, min_seconds=8.0, engine_out=eo)
    if strong is not None:
        roofline["strong_scaling"] = strong

    per_config = {}
    if not args.no_extras and world == 1:
        for cname in [c for c in args.configs.split(",") if c and c != name]:
            try:
                cb = BATCH[cname]
                if cname in ("C4", "C4N") and args.c4_batch:
                    cb = args.c4_batch
                heavy = cname in ("C4", "C4N")
                per_config[cname] = config_block(nb, torch, W, eng, cname, cb, 1 if heavy else 5, 1 if heavy else 3, peak,
                                                 1 if heavy else 3)
            except Exception as ex:   # an extra must never hide the headline
                per_config[cname] = {"error": repr(ex)}
            torch.cuda.empty_cache()
        roofline["per_config"] = per_config

    cfg = make_config(name, B, world)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": step_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": cfg,
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps, "how": "solver.solve() on pinned host buffers; CUDA events around H2D + kernel + D2H"},
        "gpu_launches": launches,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "stats": sd,
        "timing": mode,
        "systems_per_step_all_gpus": total_systems,
        "converged_per_step": converged,
    }
    print(json.dumps(line))
    if dist_on:
        import torch.distributed as dist

        dist.barrier(); dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_engine(args)


if __name__ == "__main__":
    main()
