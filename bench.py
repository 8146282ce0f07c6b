#!/usr/bin/env python
"""bench.py — converged systems / second of the batched nonlinear-solve hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload C2] [--impl engine|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path over one synthetic batch: one solve kernel over
B = 2^20 systems per GPU (inputs already resident in HBM) followed by the statistics kernel
that counts the converged systems.  The default workload is BASELINE.json configs[1]:
1M x README Example 1 (2x2, quasi_newton_solver, perturbed x0).  With N GPUs every rank solves
its own 2^20-system shard (weak scaling, no data-path collective); the per-rank statistics are
combined with one 128-byte all-gather at the end of the job (inside the timed region).

One JSON line is printed by rank 0 (schema: the bench contract in the task description).
  value      converged systems / s, all ranks, device-resident inputs, CUDA-event time (max over ranks)
  e2e        same metric through the public solver object with pinned HOST buffers: H2D of x0 (and
             per-system data) and D2H of x, fvec, iteration_behavior, status inside the timed region
  roofline   dominant kernel (the solve kernel) against the FP64 pipe: algorithmic FP64 operations
             per system (counted by the oracle's counting build on a sample of the same batch) x
             systems / CUDA-event time of that kernel, against the DFMA peak measured on this GPU;
             plus the HBM side (algorithmic bytes per system from SURVEY.md §8d) for reference
  cpu_baseline  the CPU oracle (a port; the Fortran reference cannot be built in this image) on all
             host cores, same batch
`--impl reference` times that CPU port alone and prints the same line with "impl": "reference".
"""
import argparse
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
BATCH_PER_GPU = 1 << 20


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--workload", default="C2")
    ap.add_argument("--batch", type=int, default=0, help="systems per GPU (default: 2^20; smaller for C4/C5)")
    ap.add_argument("--no-extras", action="store_true", help="skip the short runs of the other workloads")
    return ap.parse_args()


def default_batch(name):
    return {"C4": 2368, "C5": 16384, "LM4": BATCH_PER_GPU}.get(name, BATCH_PER_GPU)


def workload_label(w, B):
    desc = {
        "C1": "1M x README Example 2 cubic fit (m=21, n=4), least_squares_solver, per-system noisy y",
        "C2": "1M x README Example 1 (2x2), quasi_newton_solver, perturbed x0",
        "C3": "1M x Powell badly scaled (2x2), newton_solver + line search, max_fcn_evals=1000",
        "C4": "LM curve fits m=4096 x n=16 (rational 7/8 model)",
        "C5": "extended Rosenbrock n=64, quasi_newton_solver + line search",
        "LM4": "4-parameter double-exponential LM fits, m=64",
        "CLS1": "1M x README Example 2 cubic fit (m=21, n=4), constrained_least_squares_solver, limits [-10, 10]",
        "CLS2": "1M x README Example 1 (2x2), constrained_least_squares_solver, box [0, 6]^2, random starts",
    }[w["name"]]
    tag = "BASELINE config %s" % w["name"] if w["name"].startswith("C") and not w["name"].startswith("CLS") else (
        "SURVEY 8(f) widening %s" % w["name"] if w["name"].startswith("CLS") else "SURVEY 6 probe %s" % w["name"])
    return "%s [%s], B=%d per GPU" % (desc, tag, B)


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


def extra_parity(w, run, max_systems):
    """Parity of a DeviceRun's first-step results against the CPU port on the first max_systems systems."""
    from oracle.nl_oracle import Oracle

    o = Oracle()
    n = min(w["x0"].shape[1], max_systems)
    x0 = np.ascontiguousarray(w["x0"][:, :n])
    sysd = None if w["args"] is None else np.ascontiguousarray(w["args"][:, :n])
    ref = oracle_solve_batch(o, w, x0, sysd, params=oracle_params(o, w))
    eng = (run.xs[0][:, :n].cpu().numpy(), run.f[0][:, :n].cpu().numpy(), run.ib[0][:n].cpu().numpy().reshape(-1, 7),
           run.status[0][:n].cpu().numpy())
    return parity_report(eng, ref)


def cpu_port_throughput(w, min_seconds=4.0, max_reps=50, sample=None, engine_out=None):
    """Time the CPU oracle (OpenMP, all host cores) on the workload's batch (or a slice of it).  With engine_out =
    (x, f, ib[B,7], status) of the engine on the same batch, the port's results also serve as the parity check."""
    from oracle.nl_oracle import Oracle

    o = Oracle()
    B = w["x0"].shape[1]
    nsub = min(B, sample or B)
    x0 = np.ascontiguousarray(w["x0"][:, :nsub])
    sysd = None if w["args"] is None else np.ascontiguousarray(w["args"][:, :nsub])
    p = oracle_params(o, w)
    cores = os.cpu_count() or 1
    oracle_solve_batch(o, w, x0[:, : min(nsub, 4096)].copy(), None if sysd is None else sysd[:, : min(nsub, 4096)].copy(),
                       params=p)
    reps, elapsed, conv = 0, 0.0, 0
    while (elapsed < min_seconds and reps < max_reps) or reps < 2:
        t0 = time.perf_counter()
        xo, fo, ibo, st = oracle_solve_batch(o, w, x0, sysd, params=p, nthreads=cores)
        elapsed += time.perf_counter() - t0
        conv += int((st == 0).sum())
        reps += 1
    out = {"value": conv / elapsed, "unit": UNIT, "cores": cores, "kind": "port",
           "sample": "%d systems of the batch x %d passes, OpenMP schedule(dynamic), one system per thread at a time" % (nsub, reps)}
    if engine_out is not None:
        try:
            out["parity"] = parity_report(tuple(a[..., :nsub] if a.ndim == 2 and a.shape[1] != 7 else a[:nsub] for a in engine_out),
                                          (xo, fo, ibo, st))
        except Exception as ex:      # the parity report must never hide the measurement
            out["parity"] = {"error": repr(ex)}
    return out


def flops_per_system(w, sample=4096):
    """Algorithmic FP64 operations (+ - * / sqrt, exp = 25) per system: counting build of the oracle."""
    from oracle.nl_oracle import Oracle

    o = Oracle(counting=True)
    B = w["x0"].shape[1]
    idx = np.linspace(0, B - 1, min(sample, B)).astype(np.int64)
    x0 = np.ascontiguousarray(w["x0"][:, idx])
    sysd = None if w["args"] is None else np.ascontiguousarray(w["args"][:, idx])
    o.flops_reset()
    oracle_solve_batch(o, w, x0, sysd, params=oracle_params(o, w))
    return o.flops_total() / float(idx.size)


def polyfit_extra(nb, torch, B, npts=21, order=3, steps=5):
    """Batched polynomial%fit (README Example 3 shape: 21 shared abscissae, cubic) on device-resident data: a ring of
    y buffers larger than L2, CUDA events around `steps` fits; CPU port on a slice for comparison."""
    from nonlin_b200 import workloads as W
    from oracle.nl_oracle import Oracle

    w = W.WORKLOADS["C1"](B, seed=1000)
    x = torch.from_numpy(W.POLYFIT_XP).cuda()
    nring = max(2, int(np.ceil(768 * 2 ** 20 / (npts * B * 8))))
    ys = [torch.from_numpy(w["args"]).cuda() for _ in range(nring)]
    p = nb.polynomial()
    st = torch.zeros(B, dtype=torch.int32, device="cuda")
    for k in range(3):
        p.fit(x, ys[k % nring], order, status=st)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for k in range(steps):
        p.fit(x, ys[(3 + k) % nring], order, status=st)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    ok = int((st == 0).sum().item())
    nsub = min(B, 1 << 18)
    o = Oracle()
    ysub = np.ascontiguousarray(w["args"][:, :nsub])
    o.polyfit_batch(W.POLYFIT_XP, ysub[:, :4096].copy(), order)
    t0 = time.perf_counter()
    co, _ = o.polyfit_batch(W.POLYFIT_XP, ysub, order)
    dt = time.perf_counter() - t0
    same = bool(np.array_equal(p.get_all()[:, :nsub].cpu().numpy(), co))   # every ring slot holds the same data
    bytes_alg = 8 * (npts + order + 1) + 4
    return {"workload": "1M x polynomial%%fit, %d shared abscissae, order %d [SURVEY 8(f) widening POLY1], B=%d per GPU" % (npts, order, B),
            "value": ok / (ms * 1e-3), "unit": "fits/s", "ms_per_step": ms, "kernel": "polyfit_kernel<4, smem>",
            "hbm_gbps_algorithmic": B * bytes_alg / (ms * 1e-3) / 1e9, "bytes_per_fit": bytes_alg,
            "cpu_port": {"value": nsub / dt, "unit": "fits/s", "cores": os.cpu_count() or 1, "sample": "%d fits" % nsub},
            "bitwise_equal_to_cpu_port": same}


def run_reference(args):
    """--impl reference: the reference's CPU path.  The Fortran sources need gfortran + the external
    linalg package, neither of which exists in this image, so this arm times the C++ port (oracle/)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from nonlin_b200 import workloads as W

    B = args.batch or default_batch(args.workload)
    w = W.WORKLOADS[args.workload](B)
    from oracle.nl_oracle import Oracle

    o = Oracle()
    p = oracle_params(o, w)
    cores = os.cpu_count() or 1
    # bounded sample per step so that K steps finish within minutes
    nsub = B if args.workload in ("C1", "C2", "C3", "LM4", "CLS1", "CLS2") else min(B, 256)
    x0 = np.ascontiguousarray(w["x0"][:, :nsub]); sysd = None if w["args"] is None else np.ascontiguousarray(w["args"][:, :nsub])
    for _ in range(args.warmup):
        oracle_solve_batch(o, w, x0, sysd, params=p, nthreads=cores)
    t0 = time.perf_counter(); conv = 0
    for _ in range(args.steps):
        _, _, _, st = oracle_solve_batch(o, w, x0, sysd, params=p, nthreads=cores)
        conv += int((st == 0).sum())
    dt = time.perf_counter() - t0
    val = conv / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_label(w, B), "sample_per_step": nsub},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "%d systems per step, OpenMP over all host cores; C++ port of the Fortran path "
                                   "(no Fortran compiler / linalg in the image)" % nsub},
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


class DeviceRun:
    """One workload resident on one GPU, ready to be stepped.

    Cache hygiene: every step reads a fresh copy of x0 and writes its own fvec / ib / status buffers, taken
    from rings whose total size is several times the 126 MB L2, so no step finds its inputs or outputs in
    L2 ("inputs larger than L2"); nothing but engine kernels runs inside the timed region."""

    RING_BYTES = 768 << 20

    def __init__(self, nb, torch, w, eng, nsteps):
        self.nb, self.torch, self.w, self.eng = nb, torch, w, eng
        self.B = B = w["x0"].shape[1]
        dev = torch.device("cuda", eng.device)
        self.obj = nb.vecfcn_helper(); self.obj.set_fcn(w["fcn"], w["m"], w["n"])
        if w["shared"] is not None:
            self.obj.set_shared_data(torch.from_numpy(w["shared"]).to(dev))
        self.solver = make_solver(nb, w, eng)
        self.x0 = torch.from_numpy(w["x0"]).to(dev)
        per_step = 8 * B * (w["n"] + w["m"]) + 32 * B + (0 if w["args"] is None else 8 * B * w["args"].shape[0])
        self.nring = max(2, min(nsteps, -(-self.RING_BYTES // per_step)))
        # the solve is in place: one fresh copy of x0 per step, made before the timed region
        self.xs = [self.x0.clone() for _ in range(nsteps)]
        self.args = None if w["args"] is None else [torch.from_numpy(w["args"]).to(dev) for _ in range(self.nring)]
        self.f = [torch.empty((w["m"], B), dtype=torch.float64, device=dev) for _ in range(self.nring)]
        self.ib = [nb.iteration_behavior(B, like=self.x0) for _ in range(self.nring)]
        self.status = [torch.zeros(B, dtype=torch.int32, device=dev) for _ in range(self.nring)]
        self.stats = torch.zeros(16, dtype=torch.int64, device=dev)

    def step(self, k):
        r = k % self.nring
        self.solver.solve(self.obj, self.xs[k], self.f[r], self.ib[r], args=None if self.args is None else self.args[r],
                          status=self.status[r])
        return r

    def timed(self, steps, warmup, dist_on, sampler=None, use_graph=True):
        torch = self.torch
        for k in range(warmup):
            r = self.step(k)
            self.eng.reduce_stats_device(self.ib[r], self.status[r], self.stats, self.B)
        if dist_on:
            # the first collective on a communicator pays NCCL's lazy connection set-up: do it in the warm-up
            from nonlin_b200.distributed import allreduce_stats

            allreduce_stats(self.stats.clone())
        torch.cuda.synchronize()
        ev = lambda: torch.cuda.Event(enable_timing=True)
        e0 = [ev() for _ in range(steps)]
        es = [ev() for _ in range(steps)]
        start, end = ev(), ev()
        if sampler:
            sampler.start()
        # The K timed steps are captured into one CUDA graph (launch-bound inner loop: a 0.4 ms kernel per
        # step) so that the device runs them back to back regardless of host-side launch jitter; the graph is
        # replayed exactly once, on x0 copies no kernel has touched yet.  Falls back to eager launches.
        l0 = self.eng.kernel_launches
        graph = None
        if use_graph:
            try:
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.stream(side):
                    with torch.cuda.graph(graph, stream=side):
                        for k in range(steps):
                            r = self.step(warmup + k)
                            self.eng.reduce_stats_device(self.ib[r], self.status[r], self.stats, self.B)
                torch.cuda.current_stream().wait_stream(side)
                torch.cuda.synchronize()
            except Exception as ex:      # pragma: no cover - depends on driver support
                sys.stderr.write("bench: CUDA graph capture unavailable (%r), timing eager launches\n" % (ex,))
                graph = None
        launches = self.eng.kernel_launches - l0 if graph is not None else None
        if dist_on:
            import torch.distributed as dist

            dist.barrier()
        torch.cuda.synchronize()
        l0 = self.eng.kernel_launches
        start.record()
        if graph is not None:
            graph.replay()
        else:
            for k in range(steps):
                e0[k].record()
                r = self.step(warmup + k)
                es[k].record()                         # end of the dominant (solve) kernel
                self.eng.reduce_stats_device(self.ib[r], self.status[r], self.stats, self.B)
        if dist_on:
            # the one collective of the path: the final convergence-statistics reduction of the job
            from nonlin_b200.distributed import allreduce_stats

            allreduce_stats(self.stats)
        end.record()
        torch.cuda.synchronize()
        if launches is None:
            launches = self.eng.kernel_launches - l0
        clocks = sampler.stop() if sampler else None
        if dist_on:
            import torch.distributed as dist

            dist.barrier()
        step_ms = start.elapsed_time(end)              # K steps back to back on the device
        solve_ms = sum(a.elapsed_time(b) for a, b in zip(e0, es)) if graph is None else None
        return step_ms, solve_ms, launches, clocks

    def solve_kernel_ms(self, reps=5):
        """Average duration of the dominant (solve) kernel alone, CUDA events on the launching stream."""
        torch = self.torch
        tot = 0.0
        for rep in range(reps):
            i = rep % len(self.xs)                     # --steps / --warmup may be smaller than reps
            self.xs[i].copy_(self.x0)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            self.step(i)
            b.record()
            torch.cuda.synchronize()
            tot += a.elapsed_time(b)
        return tot / reps


def e2e_run(nb, torch, w, eng, steps, warmup):
    """Public API with pinned HOST buffers: H2D + kernel + D2H per step, timed with CUDA events on
    the stream the engine is told to use."""
    B = w["x0"].shape[1]
    obj = nb.vecfcn_helper(); obj.set_fcn(w["fcn"], w["m"], w["n"])
    if w["shared"] is not None:
        obj.set_shared_data(w["shared"])
    solver = make_solver(nb, w, eng)
    pin = lambda a: torch.from_numpy(a).pin_memory()
    x0 = pin(w["x0"]); args = None if w["args"] is None else pin(w["args"])
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
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"          # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import nonlin_b200 as nb
    from nonlin_b200 import workloads as W

    eng = nb.default_engine(local)
    B = args.batch or default_batch(args.workload)
    w = W.WORKLOADS[args.workload](B, seed=1000 + rank)          # every rank its own shard (weak scaling)
    peak = eng.measure_fp64_peak()                               # also spins the clocks up
    run = DeviceRun(nb, torch, w, eng, args.steps + args.warmup)
    sampler = ClockSampler(local) if rank == 0 else None
    step_ms, solve_ms, launches, clocks = run.timed(args.steps, args.warmup, dist_on, sampler)
    solve_ms = run.solve_kernel_ms() * args.steps if solve_ms is None else solve_ms
    stats = run.stats.clone()
    t = torch.tensor([step_ms, solve_ms], dtype=torch.float64, device="cuda")
    if dist_on:
        import torch.distributed as dist

        dist.all_reduce(t, op=dist.ReduceOp.MAX)                 # device time = max over ranks
    step_ms, solve_ms = float(t[0]), float(t[1])
    from nonlin_b200.distributed import stats_dict

    sd = stats_dict(stats)                                       # already summed over ranks when dist_on
    total_systems = sd["systems"]
    converged = sd["converged"]
    value = converged * args.steps / (step_ms * 1e-3)

    # e2e through the public API with host buffers (each rank its shard, max over ranks)
    e2e_steps = max(3, min(args.steps, 10))
    e_ms, e_conv, h2d, d2h = e2e_run(nb, torch, w, eng, e2e_steps, 3)
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

    # roofline of the dominant kernel (solve kernel), rank 0's launch
    fl = flops_per_system(w)
    kernel_s = solve_ms * 1e-3 / args.steps
    achieved_tf = fl * B / kernel_s / 1e12
    peaks_file = os.path.join(ROOT, "MEASURED_PEAKS.json")
    hbm_peak, hbm_src = 6650.0, "fallback"
    if os.path.exists(peaks_file):
        try:
            hbm_peak, hbm_src = float(json.load(open(peaks_file))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    hbm_ach = w["bytes_per_system"] * B / kernel_s / 1e9
    traffic = None
    tf = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tf):
        try:
            traffic = json.load(open(tf)).get(args.workload)
        except Exception:
            traffic = None
    roofline = {
        "bound": "fp64", "kernel": {"C1": "tps_solve_kernel<LsqPolyFit, LM>", "C2": "tps_solve_kernel<Misc2Fcn, Broyden>",
                                    "C3": "tps_newton_refill_kernel<PowellBadlyScaled>", "C4": "wlm_kernel<Rational78, 16>",
                                    "C5": "coop_broyden_kernel<ExtRosenbrock, 64>", "LM4": "coop_lm_kernel<ExpDecay4, 4>",
                                    "CLS1": "tps_cls_kernel<LsqPolyFit>", "CLS2": "tps_cls_kernel<Misc2Fcn>"}[w["name"]],
        "achieved": achieved_tf, "peak": peak["dfma_tflops"], "unit": "TFLOP/s", "frac": achieved_tf / peak["dfma_tflops"],
        "peak_source": "DFMA micro-kernel measured on this GPU at run time (MEASURED_PEAKS.json has no FP64 entry)",
        "peak_no_fma": peak["dadd_dmul_tflops"], "frac_of_no_fma_peak": achieved_tf / peak["dadd_dmul_tflops"],
        "flops_per_system": fl, "kernel_ms": kernel_s * 1e3,
        "note": "parity build issues DMUL+DADD instead of DFMA (-fmad=false), so its own ceiling is peak_no_fma",
        "traffic": traffic,
        "hbm": {"bound": "hbm", "achieved": hbm_ach, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_ach / hbm_peak,
                "bytes_per_system": w["bytes_per_system"], "peak_source": hbm_src},
    }
    # results of the first step of the batch (every step solves the same systems), for the parity report
    eng_out = (run.xs[0].cpu().numpy(), run.f[0].cpu().numpy(), run.ib[0].cpu().numpy().reshape(-1, 7),
               run.status[0].cpu().numpy()) if world == 1 else None
    cpu = cpu_port_throughput(w, engine_out=eng_out) if world == 1 else None

    extras = {}
    if not args.no_extras and world == 1:
        for name in ("C1", "C3", "C5", "LM4", "CLS1", "CLS2"):
            if name == args.workload:
                continue
            try:
                we = W.WORKLOADS[name](default_batch(name), seed=1000)
                r = DeviceRun(nb, torch, we, eng, 5 + 3)
                sm, km, _, _ = r.timed(5, 3, False)
                km = r.solve_kernel_ms(3) * 5 if km is None else km
                c = int(r.stats[1].item())
                fle = flops_per_system(we, 1024)
                extras[name] = {"workload": workload_label(we, default_batch(name)), "value": c * 5 / (sm * 1e-3), "unit": UNIT,
                                "ms_per_step": sm / 5, "fp64_tflops": fle * default_batch(name) / (km / 5 * 1e-3) / 1e12,
                                "flops_per_system": fle}
                try:                      # engine vs CPU port on (a slice of) the same batch
                    extras[name]["parity"] = extra_parity(we, r, 1 << 18)
                except Exception as ex:
                    extras[name]["parity"] = {"error": repr(ex)}
                del r
            except Exception as ex:   # an extra must never hide the headline
                extras[name] = {"error": repr(ex)}

        try:
            extras["POLY1"] = polyfit_extra(nb, torch, BATCH_PER_GPU)
        except Exception as ex:
            extras["POLY1"] = {"error": repr(ex)}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": step_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_label(w, B), "systems_per_step_all_gpus": total_systems,
                   "converged_per_step": converged, "parallelism": "dp%d (contiguous system shards, no data-path collective)" % world,
                   "l2": "inputs larger than L2: every step reads a fresh x0 copy and writes its own output buffers from rings of >= 768 MiB",
                   "timing": "K steps captured in one CUDA graph, replayed once between two CUDA events (max over ranks); solve kernel timed separately with events",
                   "arithmetic": "FP64, no FMA contraction (bit-identical to the CPU oracle)"},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps, "how": "solver.solve() on pinned host buffers; CUDA events around H2D + kernel + D2H"},
        "gpu_launches": launches,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "stats": sd,
        "other_workloads": extras,
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
