#!/usr/bin/env python
"""Benchmark of the LM inner loop (BASELINE.json metric) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload C2]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference          # the reference's CPU path

A *step* is one LM iteration of the global fit: Jacobian sweep over every
point of every data set + block-arrow solve + chi2 sweep(s) of the trial
step(s).  The workload at N = 1 is BASELINE config 2 at full size (64 data
sets x 1M points, 20 shared parameters, FP64); at N > 1 every rank owns 64
such data sets (weak scaling) and only the 20 x 20 shared normal block, its
right-hand side and chi2 are all-reduced (NCCL) per trial step.

value   Jacobian entries (structural non-zeros, all ranks) processed per second
        of the timed K iterations, data resident in HBM.
e2e     the same through the public API from pinned HOST buffers: create,
        addDataset x S, setPar, fit(K iterations) incl. the H2D upload, and the
        read-back of chi2 + parameters.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SETS_PER_GPU = {"C2": 64, "C3": 1000, "C4": 256, "C5": 100_000}
POINTS = {"C2": 1_000_000, "C3": 4096, "C4": 65536, "C5": 256}
DESCRIPTION = {
    "C2": "BASELINE config 2: 64 data sets x 1M points per GPU, 3 distorted Gaussians, 20 shared parameters",
    "C3": "BASELINE config 3: 1000 x 4096 per GPU, Gaussian + linear background, 4 local + 2 shared",
    "C4": "BASELINE config 4: 256 x 65536 per GPU, single-integral model + geodesic acceleration",
    "C5": "BASELINE config 5: 100000 x 256 per GPU, Lorentzian + double-integral term, 3 local + 4 shared",
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="C2", choices=list(SETS_PER_GPU))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--sets", type=int, default=0, help="data sets per GPU (default: full size)")
    ap.add_argument("--points", type=int, default=0, help="points per data set (default: full size)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-workloads", action="store_true",
                    help="skip the quick lines of configs 3-5 inside the default config-2 run")
    ap.add_argument("--repeats", type=int, default=5,
                    help="the timed K-step fit is repeated from the same start; the median is reported")
    return ap.parse_args()


# --------------------------------------------------------------- clock sampler

class ClockSampler:
    """nvidia-smi clocks and throttle reasons DURING the timed region."""

    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap,utilization.gpu")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.QUERY}",
                 "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def wait_ready(self, timeout: float = 10.0) -> None:
        """nvidia-smi takes 0.1-1 s to deliver its first sample (longer on an
        8-GPU box): block until it does, so the timed region is covered."""
        t_end = time.perf_counter() + timeout
        while self.proc is not None and not self.rows and time.perf_counter() < t_end:
            time.sleep(0.01)

    def stop(self, t_begin: float | None = None, t_end: float | None = None) -> dict:
        """Samples taken within [t_begin, t_end] (the timed region)."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        rows = [r for t, r in self.rows
                if (t_begin is None or t >= t_begin) and (t_end is None or t <= t_end + 0.03)]
        sm, smax, power, loaded, reasons = [], [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            if len(r) < 8:
                continue
            try:
                sm.append(float(r[0]))
                smax.append(float(r[1]))
                power.append(float(r[2]))
                loaded.append(float(r[7]) >= 50.0)
            except ValueError:
                continue
            for name, flag in zip(names, r[3:7]):
                if flag.lower().startswith("active"):
                    reasons.add(name)
        under_load = [c for c, l in zip(sm, loaded) if l] or sm
        return {"sm_mhz": float(np.median(under_load)) if under_load else None,
                "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None,
                "samples": len(sm), "samples_under_load": int(sum(loaded)),
                "reasons": sorted(reasons)}


# ------------------------------------------------------------ reference (CPU)

def reference_api(use_blas: bool = False):
    """The reference's own CPU implementation, compiled from its unmodified
    sources (oracle/_ref, kind 'reference'); else the oracle's CPU restatement
    (kind 'port').  Used only as the timed CPU baseline."""
    from gadfit_b200._capi import Api
    # the reference's optimised configuration (lapack.cpp + OpenBLAS) when it
    # could be linked and the problem is small enough for it (n <~ 150,
    # lapack.cpp:92); else its default in-tree linear algebra
    blas = os.path.join(ROOT, "oracle", "_ref", "libgadfit_ref_blas.so")
    fast = os.path.join(ROOT, "oracle", "_ref", "libgadfit_ref_fast.so")
    if use_blas and os.path.exists(blas):
        try:
            return Api(blas, "gadfit_ref_", product=False), "reference", "lapack.cpp + OpenBLAS 0.3.15"
        except OSError:
            pass
    if os.path.exists(fast):
        return Api(fast, "gadfit_ref_", product=False), "reference", "in-tree lapack_fallback.cpp"
    port = os.path.join(ROOT, "oracle", "_build", "libgadfit_oracle.so")
    return Api(port, "gadfit_oracle_", product=False), "port", "oracle restatement"


def run_cpu_baseline(workload: str, steps: int, warmup: int, budget_s: float = 20.0) -> dict:
    """Times the CPU reference on a bounded sample of the same workload: same
    model, parameters, settings; fewer data sets / points so that the run ends
    within ~budget_s seconds.  Reports Jacobian entries/s over whole LM
    iterations (same definition as the GPU value)."""
    from gadfit_b200 import workloads as W
    from gadfit_b200.lm_solver import LMsolver

    api, kind, linalg = reference_api(use_blas=(workload == "C2"))
    cores = os.cpu_count() or 1
    full_sets, full_points = SETS_PER_GPU[workload], POINTS[workload]
    # the dense reference is limited by n = p_g + S p_l
    sets = {"C2": 8, "C3": 16, "C4": 16, "C5": 32}[workload]
    points = min(full_points, {"C2": 125_000, "C3": 4096, "C4": 2048, "C5": 256}[workload])
    iters = max(1, min(steps, 3))

    def once(n_sets, n_points, n_iter):
        w = W.WORKLOADS[workload](n_sets=n_sets, n_points=n_points, iteration_limit=n_iter)
        s = LMsolver(w.model, _api=api)
        w.setup(s)
        s.settings.n_threads = cores
        t0 = time.perf_counter()
        s.fit(w.lambda0)
        dt = time.perf_counter() - t0
        tm = s.timings()
        done = max(s.iterations(), 1)
        return w, dt, tm, done

    # calibration on a small instance, then scale the points to the budget
    w, dt, tm, done = once(sets, max(points // 8, 64), 1)
    per_point_iter = dt / (w.n_total * done)
    target = int(budget_s / (per_point_iter * iters * sets))
    points = int(max(64, min(points, target)))
    w, dt, tm, done = once(sets, points, iters)
    value = w.nnz_jacobian * done / tm["main"][0] if tm["main"][0] > 0 else w.nnz_jacobian * done / dt
    jac_sweep = tm["jacobian"][0] / max(tm["jacobian"][1], 1)
    return {
        "value": value, "unit": "jacobian_entries/s", "cores": cores, "kind": kind,
        "sample": f"{workload} model/settings, {sets} data sets x {points} points, {done} LM iterations, "
                  f"all {cores} host threads (OpenMP), {linalg}, g++ -O3 -march=x86-64-v3",
        "lm_iterations_per_s": done / dt, "steps_run": done,
        "jacobian_sweep_entries_per_s": w.nnz_jacobian / jac_sweep if jac_sweep > 0 else None,
        "seconds": dt,
    }


def run_reference_full(workload: str, steps: int) -> dict:
    """The reference's own CPU implementation on the SAME config as the GPU arm
    (C2: 64 data sets x 1M points, dense J = 10.2 GB): `steps` LM iterations
    from the same start, all host threads.  Only config 2 fits the dense
    reference at full size."""
    from gadfit_b200 import workloads as W
    from gadfit_b200.lm_solver import LMsolver

    api, kind, linalg = reference_api(use_blas=True)
    cores = os.cpu_count() or 1
    w = W.WORKLOADS[workload](iteration_limit=steps)
    s = LMsolver(w.model, _api=api)
    w.setup(s)
    s.settings.n_threads = cores
    t0 = time.perf_counter()
    s.fit(w.lambda0)
    dt = time.perf_counter() - t0
    tm = s.timings()
    done = max(s.iterations(), 1)
    main = tm["main"][0] if tm["main"][0] > 0 else dt
    jac_sweep = tm["jacobian"][0] / max(tm["jacobian"][1], 1)
    return {
        "value": w.nnz_jacobian * done / main, "unit": "jacobian_entries/s", "cores": cores, "kind": kind,
        "sample": f"{workload} at FULL size ({w.n_sets} data sets x {w.n_points} points), {done} LM iterations "
                  f"from the GPU arm's start values, all {cores} host threads (OpenMP), {linalg}, "
                  "g++ -O3 -march=x86-64-v3",
        "lm_iterations_per_s": done / main, "steps_run": done, "seconds": dt, "seconds_main_loop": main,
        "jacobian_sweep_entries_per_s": w.nnz_jacobian / jac_sweep if jac_sweep > 0 else None,
        "final_chi2": s.trace()[-1]["new_chi2"] if s.trace() else None,
    }


def main_reference(args, emit):
    """`--impl reference`: the unmodified reference (oracle/_ref) on the host
    cores.  Config 2 runs at the GPU arm's full size for min(K, 2) iterations
    (about 20 s per iteration on 16 threads: the printed `steps` is what ran);
    configs 3-5 cannot be run by the dense reference at full size and are timed
    on a bounded sample, labelled as such."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t0 = time.perf_counter()
    full = args.workload == "C2" and not args.sets and not args.points
    if full:
        base = run_reference_full(args.workload, max(1, min(args.steps, 2)))
        steps_run = base["steps_run"]
        workload = DESCRIPTION[args.workload] + " -- same config as the GPU arm: " + base["sample"]
    else:
        base = run_cpu_baseline(args.workload, args.steps, args.warmup, budget_s=30.0)
        steps_run = base["steps_run"]
        workload = DESCRIPTION[args.workload] + " -- bounded sample: " + base["sample"]
    line = {
        "impl": "reference",
        "metric": "jacobian_entries_per_s", "value": base["value"], "unit": "jacobian_entries/s",
        "n_gpus": args.gpus, "steps": steps_run, "warmup": 0,
        "steps_requested": args.steps, "warmup_requested": args.warmup,
        "ms_per_step": 1e3 / base["lm_iterations_per_s"] if base["lm_iterations_per_s"] else None,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "same_config_as_gpu_arm": bool(full),
        "config": {"workload": workload},
        "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": base["value"], "unit": "jacobian_entries/s",
                "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "lm_iterations_per_s": base["lm_iterations_per_s"],
        "wall_s": time.perf_counter() - t0,
    }
    emit(json.dumps(line))


# ------------------------------------------------------------------ GPU arm

def bind_to_gpu_numa_node(index: int) -> dict:
    """Run this rank on the CPUs of the NUMA node its GPU hangs off, BEFORE the
    pinned host buffers are allocated: cudaHostAlloc places pages on the node of
    the calling thread, and an upload that crosses the socket interconnect runs
    at a fraction of the PCIe rate once 8 ranks upload at the same time (round 1:
    e2e weak scaling 70 % at 8 GPUs).  Best effort: containers may hide sysfs."""
    info = {"bound": False}
    try:
        import pynvml
        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(index)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        dom, rest = bus.split(":", 1)
        dev = f"{int(dom, 16):04x}:{rest.lower()}"
        base = f"/sys/bus/pci/devices/{dev}"
        node = int(open(f"{base}/numa_node").read())
        info.update(pci=dev, numa_node=node)
        if node < 0:
            return info
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0)
        use = cpus & allowed
        if use:
            os.sched_setaffinity(0, use)
            info.update(bound=True, cpus=len(use))
    except Exception as exc:  # noqa: BLE001 -- never take the bench down
        info["error"] = repr(exc)[:120]
    return info


def load_cost_table() -> dict:
    """Cost table v2 (profiles/cost_table.json, written by
    tools/cost_table.py from committed `ncu --set full` captures): FP64-pipe
    work each sweep EXECUTES per data point, counted in flop (2 per FP64
    instruction slot, 512 per warp-level DMMA)."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "cost_table.json")))
    except (OSError, ValueError):
        return {}


def main_b200(args, emit):
    import ctypes as _C

    import torch
    import torch.distributed as dist

    from gadfit_b200 import LMsolver, comm_unique_id, measure_fp64_peak
    from gadfit_b200 import workloads as W
    from gadfit_b200._capi import product_api as _api

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    all_cpus = os.sched_getaffinity(0)
    numa = bind_to_gpu_numa_node(int(os.environ.get("CUDA_VISIBLE_DEVICES", "").split(",")[local_rank])
                                 if os.environ.get("CUDA_VISIBLE_DEVICES", "").replace(",", "").isdigit()
                                 else local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def fresh_uid():
        """A NCCL unique id per communicator: rank 0 creates, everyone receives."""
        buf = torch.zeros(128, dtype=torch.uint8, device=device)
        if rank == 0:
            buf.copy_(torch.frombuffer(bytearray(comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(buf, 0)
        return bytes(buf.cpu().numpy().tobytes())

    # FP64 roofline denominator: DFMA vector loop and DMMA tensor-path loop share
    # one datapath on B200; the HIGHER of the two measured rates is the peak
    dfma_peak = measure_fp64_peak()
    _out = _C.c_double()
    _api().microbench(1, _C.byref(_out))
    dmma_peak = _out.value
    fp64_peak = max(dfma_peak, dmma_peak)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    cost = load_cost_table()

    def attach(model):
        """A solver bound to its GPU and (N > 1) to the communicator: one-time
        set-up, like creating a process group."""
        s = LMsolver(model)
        s.setDevice(local_rank)
        if world > 1:
            s.setComm(rank, world, fresh_uid())
        return s

    def measure(name, sets_per_gpu, n_points, steps, warmup, repeats, with_e2e, sampler=None):
        """One workload: device-resident arm (upload once, W warm-up iterations,
        then K timed iterations, `repeats` times from the same start, median)
        and, optionally, the end-to-end arm from pinned host buffers."""
        n_sets = sets_per_gpu * world
        owned = range(rank * sets_per_gpu, (rank + 1) * sets_per_gpu)
        pinned = []

        def alloc(n):
            t = torch.empty(n, dtype=torch.float64, pin_memory=(n * 8 <= (8 << 30)))
            pinned.append(t)
            return t.numpy()

        w = W.WORKLOADS[name](n_sets=n_sets, n_points=n_points, alloc=alloc, owned=owned,
                              iteration_limit=steps)

        def describe(s, limit):
            """The calls a user of the reference makes: addDataset, setPar, settings."""
            w.setup(s, owned=owned if world > 1 else None, pointers=False)
            s.settings.iteration_limit = limit
            return s

        solver = describe(attach(w.model), warmup)
        solver.upload()
        barrier()
        if warmup > 0:
            solver.fit(w.lambda0)                      # W untimed warm-up iterations

        def reset_start():
            # the timed fit does the same work as the reference would from the
            # same initial guess
            for k in w.shared:
                solver.setPar(k, float(w.start[0, k]), True)
            for s_ in (owned if world > 1 else range(n_sets)):
                for k in w.local:
                    solver.setPar(k, float(w.start[s_, k]), True, s_)

        solver.settings.iteration_limit = steps
        if sampler is not None:
            sampler.wait_ready()
        t_begin = time.perf_counter()
        runs = []
        # Latency-bound shape (config 3, ~0.15 ms per iteration): the twelve
        # CUDA-event records per trial step of the phase timers are a visible
        # share of the step, so the timed repeats run without them and ONE
        # extra, instrumented repeat supplies the phase and K1 times.
        separate = name == "C3"
        for rep in range(max(repeats, 1) + (1 if separate else 0)):
            instrumented = not separate or rep == max(repeats, 1)
            solver.setPhaseTimers(instrumented)
            reset_start()
            launches0 = solver.launches()
            barrier()
            t0 = time.perf_counter()
            solver.fit(w.lambda0)
            torch.cuda.synchronize()
            t_local = time.perf_counter() - t0
            barrier()
            k1_sec, k1_calls = solver.jacobianKernelTime()
            runs.append(dict(seconds=max_over_ranks(t_local), launches=solver.launches() - launches0,
                             k1_sec=max_over_ranks(k1_sec), k1_calls=k1_calls, tm=solver.timings(),
                             done=solver.iterations(), instrumented=instrumented))
        t_end = time.perf_counter()
        probe = runs.pop() if separate else None           # the instrumented repeat
        runs.sort(key=lambda r: r["seconds"])
        med = dict(runs[len(runs) // 2])                  # median repeat
        if probe is not None:
            med.update(k1_sec=probe["k1_sec"], k1_calls=probe["k1_calls"], tm=probe["tm"],
                       seconds_instrumented=probe["seconds"])
        med["phase_timers"] = ("separate instrumented repeat (%.4f ms per step with the event records)"
                               % (probe["seconds"] / max(probe["done"], 1) * 1e3)) if separate else "inside the timed repeats"
        med["all_seconds"] = [r["seconds"] for r in runs]
        med["device_main"] = max_over_ranks(med["tm"]["main"][0])
        med["trials"] = len(solver.trace())
        accepted = [t for t in solver.trace() if t["decision"] == 0]
        med["chi2"] = accepted[-1]["new_chi2"] if accepted else float("nan")
        med["comm"] = solver.commStats() if world > 1 else None
        med["region"] = (t_begin, t_end)
        med["w"] = w
        med["points_per_gpu"] = sets_per_gpu * n_points
        med["value"] = w.nnz_jacobian * med["done"] / med["seconds"]
        del solver
        e2e = None
        if with_e2e:
            # ---- end-to-end arm: public API from pinned host buffers, H2D inside
            s2 = attach(w.model)          # communicator creation is one-time set-up, not timed
            barrier()
            t0 = time.perf_counter()
            describe(s2, steps)           # addDataset x S, setPar, settings (host pointers)
            s2.fit(w.lambda0)             # H2D upload of x, y, sigma + K LM iterations
            out = [s2.chi2()] + [s2.getParValue(k, owned[0]) for k in range(w.n_pars)]
            torch.cuda.synchronize()
            t_e2e = time.perf_counter() - t0
            barrier()
            t_e2e = max_over_ranks(t_e2e)
            done2 = max(s2.iterations(), 1)
            e2e = {
                "value": w.nnz_jacobian * done2 / t_e2e, "unit": "jacobian_entries/s",
                "h2d_bytes_per_step": int(24 * med["points_per_gpu"] / done2),
                "d2h_bytes_per_step": int(8 * (sets_per_gpu * w.n_pars + 1) / done2)
                + 80 * max(len(s2.trace()), 1) // done2,
                "seconds": t_e2e, "steps": done2,
                "note": "one fit() of K iterations from pinned host buffers per GPU: the H2D copy of x, y, "
                        "sigma (24 B/point) happens once per fit and is amortised over the K steps; creating "
                        "the communicator (N > 1, one-time) is outside the timed region",
                "lm_iterations_per_s": done2 / t_e2e,
                "chi2": out[0],
            }
            del s2
        if e2e is not None:
            # same start, same data, same iterations: the end-to-end fit (pipelined
            # upload, sigma inverted in place by the first sweep) must land on the
            # chi2 of the device-resident one
            ref = med["chi2"]
            e2e["chi2_rel_diff_to_device_resident_fit"] = (
                abs(e2e["chi2"] - ref) / abs(ref) if ref == ref and ref != 0.0 else None)
        med["e2e"] = e2e
        return med

    def roofline_of(name, m, full_size):
        """Roofline of the dominant kernel (Jacobian sweep), per launch."""
        w, k1_sec = m["w"], m["k1_sec"]
        pts = m["points_per_gpu"]
        v1 = w.k1_flops_per_point()
        entry = (cost.get(name) or {}).get("k_jacobian") or {}
        v2 = entry.get("fp64_flop_per_point")
        # Config 2: the survey's algorithmic count (cost table v1) is BELOW what the
        # kernel executes (padded DMMA tiles), so it is the honest numerator.  The
        # quadrature configs: v1 credits 4 F_val for a gradient the static-sparsity
        # forward mode gets far cheaper -- v1 would read > 1; the numerator is then
        # the FP64 work the kernel executes (cost table v2, from ncu).
        use_v2 = v2 is not None and v2 < v1
        flops_pt = v2 if use_v2 else v1
        flops_source = (
            "cost table v2: FP64 instruction slots the kernel executes per point (%s), 2 flop each + 512 per DMMA"
            % entry.get("source", "profiles/cost_table.json")
            if use_v2 else
            "cost table v1 (SURVEY.md 8(d)): 4 F_val + (2 + p) + 2 [p(p+1)/2 + p + 1], algorithmic")
        achieved = flops_pt * pts / k1_sec / 1e12 if k1_sec > 0 else None
        bytes_per_launch = 24 * pts
        traffic = entry.get("dram_bytes_per_launch") if full_size else None
        return {
            # the contract's "tensor" bound: here the FP64 tensor path (DMMA m8n8k4)
            # and the FP64 FMA pipe, which share one datapath -- not bf16
            "bound": "tensor", "pipe": "fp64",
            # csrc/model_ops.cuh: two or more 8-column blocks -> the warp-specialised sweep
            "kernel": "k_jacobian_ws" if w.n_pars + 1 > 8 else "k_jacobian (software-pipelined)",
            "achieved": achieved, "peak": fp64_peak / 1e12, "unit": "TFLOP/s",
            "frac": achieved / (fp64_peak / 1e12) if achieved else None,
            "flops_source": flops_source,
            "peak_source": "measured live on this GPU: the higher of a register-resident DFMA loop "
                           f"({dfma_peak / 1e12:.2f}) and a DMMA m8n8k4 loop ({dmma_peak / 1e12:.2f} TFLOP/s), "
                           "which share one datapath (csrc/microbench.cu); nominal 148 SM x 64 lanes x 2 x "
                           "1.965 GHz = 37.2",
            "frac_of_dfma_peak": achieved / (dfma_peak / 1e12) if achieved else None,
            "flops_per_point": flops_pt, "flops_per_point_v1": v1, "flops_per_point_v2": v2,
            "frac_v1": v1 * pts / k1_sec / fp64_peak if k1_sec > 0 else None,
            "points_per_launch": pts, "seconds_per_launch": k1_sec, "launches_timed": m["k1_calls"],
            "hbm": {"achieved": bytes_per_launch / k1_sec / 1e9 if k1_sec > 0 else None,
                    "peak": hbm_peak, "unit": "GB/s",
                    "frac": bytes_per_launch / k1_sec / 1e9 / hbm_peak if k1_sec > 0 else None,
                    "peak_source": "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback 6650"},
            "traffic": traffic,
            "traffic_source": entry.get("source") if traffic else None,
            "algorithmic_bytes": bytes_per_launch,
        }

    def multi_rank_parity():
        """N > 1: small global fits solved by ALL ranks together (data sets
        sharded, shared block / Schur sums / chi2 all-reduced over peer memory;
        one point-split case through NCCL) and by rank 0 ALONE on one GPU:
        identical accept/reject sequence and lambda, chi2 and parameters to
        1e-10 -- the correctness of csrc/peer_reduce.cuh inside the scaling run."""
        cases = [("C3", dict(n_sets=8 * world, n_points=640)),       # local + shared: Schur sums
                 ("C2", dict(n_sets=2 * world, n_points=20_000)),    # shared only
                 ("C4", dict(n_sets=2 * world, n_points=1024)),      # + geodesic acceleration
                 ("C2", dict(n_sets=1, n_points=40_000))]            # fewer sets than ranks: point split
        out = []
        for name, kw in cases:
            from gadfit_b200 import partition
            w = W.WORKLOADS[name](iteration_limit=6, **kw)
            first, end, split = partition([w.n_points] * w.n_sets, rank, world)
            s = attach(w.model)
            w.setup(s, owned=None if split else range(first, end))
            s.fit(w.lambda0)
            pars = np.array([[s.getParValue(k, j) for k in range(w.n_pars)] for j in range(w.n_sets)])
            tr = s.trace()
            comm = s.commStats()
            # every rank must end with the same parameters
            t = torch.tensor(pars, dtype=torch.float64, device=device)
            ref = t.clone()
            dist.broadcast(ref, 0)
            same = torch.tensor([1.0 if torch.equal(t, ref) else 0.0], dtype=torch.float64, device=device)
            dist.all_reduce(same, op=dist.ReduceOp.MIN)
            del s
            if rank == 0:
                a = LMsolver(w.model)
                a.setDevice(local_rank)
                w.setup(a)
                a.fit(w.lambda0)
                apars = np.array([[a.getParValue(k, j) for k in range(w.n_pars)] for j in range(w.n_sets)])
                atr = a.trace()
                floor = 1e-11 * np.max(np.abs(apars), axis=0, keepdims=True)
                err = float(np.max(np.maximum(np.abs(pars - apars) - floor, 0.0)
                                   / np.maximum(np.abs(apars), 1e-300)))
                chi2_err = max(abs(x["new_chi2"] - y["new_chi2"]) / abs(y["new_chi2"]) for x, y in zip(tr, atr))
                rec = dict(
                    workload=f"{name} {kw}", point_split=bool(split), trials=len(atr),
                    decisions_identical=[x["decision"] for x in tr] == [y["decision"] for y in atr],
                    lambdas_identical=[x["lambda_"] for x in tr] == [y["lambda_"] for y in atr],
                    chi2_max_rel_diff=chi2_err, parameter_max_rel_diff=err,
                    all_ranks_hold_identical_parameters=bool(same.item() == 1.0),
                    peer_reductions=comm["peer_reductions"], nccl_reductions=comm["nccl_reductions"])
                rec["ok"] = bool(rec["decisions_identical"] and rec["lambdas_identical"]
                                 and chi2_err <= 1e-10 and err <= 1e-10
                                 and rec["all_ranks_hold_identical_parameters"])
                out.append(rec)
                del a
            barrier()
        if rank != 0:
            return None
        return {"ok": all(r["ok"] for r in out), "tolerance": 1e-10,
                "what": f"{world} ranks together vs rank 0 alone on the same inputs", "cases": out}

    sets_per_gpu = args.sets or SETS_PER_GPU[args.workload]
    n_points = args.points or POINTS[args.workload]
    full_size = not args.sets and not args.points

    sampler = ClockSampler(local_rank)  # started early: its first sample takes a while
    sampler.start()
    m = measure(args.workload, sets_per_gpu, n_points, args.steps, args.warmup, args.repeats,
                with_e2e=not args.no_e2e, sampler=sampler)
    clocks = sampler.stop(*m["region"])
    roofline = roofline_of(args.workload, m, full_size)
    w = m["w"]

    # The other BASELINE shapes, quick (full size per GPU, one repeat): ms per
    # iteration, sweep times and roofline fraction of each, driver-run
    others = {}
    if args.workload == "C2" and full_size and not args.no_workloads:
        for name in ("C3", "C4", "C5"):
            steps = {"C3": 10, "C4": 8, "C5": 6}[name]
            q = measure(name, SETS_PER_GPU[name], POINTS[name], steps, 3, 3 if name == "C3" else 1,
                        with_e2e=False)
            r = roofline_of(name, q, True)
            tm = q["tm"]
            per = lambda k: tm[k][0] / max(tm[k][1], 1)  # noqa: E731
            others[name] = {
                "workload": DESCRIPTION[name], "steps": steps, "iterations_done": q["done"],
                "trial_steps": q["trials"], "ms_per_step": q["seconds"] / max(q["done"], 1) * 1e3,
                "value": q["value"], "unit": "jacobian_entries/s", "gpu_launches": q["launches"],
                "seconds_per_call": {"k1_jacobian_kernel": q["k1_sec"], "jacobian_phase": per("jacobian"),
                                     "k2_chi2_phase": per("chi2"), "k3_omega_phase": per("omega"),
                                     "linear_algebra_phase": per("linalg")},
                "phase_calls": {k: v[1] for k, v in tm.items()},
                "phase_timers": q["phase_timers"],
                "roofline": {k: r[k] for k in ("frac", "achieved", "peak", "unit", "flops_per_point",
                                               "flops_source", "frac_v1", "seconds_per_launch")},
                "final_chi2": q["chi2"],
            }

    parity = multi_rank_parity() if world > 1 else None

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            os.sched_setaffinity(0, all_cpus)  # the CPU baseline uses every host core
            b = run_cpu_baseline(args.workload, args.steps, args.warmup)
            cpu_baseline = {k: b[k] for k in ("value", "unit", "cores", "kind", "sample")}
            cpu_baseline["lm_iterations_per_s"] = b["lm_iterations_per_s"]
            cpu_baseline["jacobian_sweep_entries_per_s"] = b["jacobian_sweep_entries_per_s"]
        except Exception as exc:  # the baseline must never take the GPU line down
            cpu_baseline = {"value": None, "unit": "jacobian_entries/s", "cores": os.cpu_count(),
                            "kind": "unavailable", "sample": repr(exc)}

    if rank == 0:
        tm, done, seconds = m["tm"], m["done"], m["seconds"]
        line = {
            "metric": "jacobian_entries_per_s", "value": m["value"], "unit": "jacobian_entries/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": seconds / max(done, 1) * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": DESCRIPTION[args.workload],
                "data_sets_per_gpu": sets_per_gpu, "points_per_data_set": n_points,
                "parameters": {"shared": len(w.shared), "local_per_data_set": len(w.local)},
                "settings": w.settings | {"lambda0": w.lambda0, "iteration_limit": args.steps},
                "l2": "inputs (24 B/point x %d points = %.2f GB per GPU) are larger than the 126 MB L2"
                      % (m["points_per_gpu"], 24 * m["points_per_gpu"] / 1e9),
                "parallelism": "one process per GPU; data sets sharded; all-reduce of the shared "
                               "normal block + chi2 per trial step over NVLink peer memory" if world > 1
                               else "single GPU",
                "host_buffers": {"pinned": True, "numa": numa},
            },
            "lm_iterations_per_s": done / seconds,
            "repeats": {"n": len(m["all_seconds"]), "seconds": m["all_seconds"], "reported": "median"},
            "iterations_done": done, "trial_steps": m["trials"], "final_chi2": m["chi2"],
            "k1_sweep_entries_per_s": w.nnz_jacobian / m["k1_sec"] if m["k1_sec"] > 0 else None,
            "device_seconds_main_loop": m["device_main"],
            "phase_seconds": {k: v[0] for k, v in tm.items()},
            "phase_calls": {k: v[1] for k, v in tm.items()},
            "phase_timers": m["phase_timers"],
            "gpu_launches": m["launches"],
            "collectives": m["comm"],
            "clocks": clocks,
            "roofline": roofline,
            "e2e": m["e2e"],
            "cpu_baseline": cpu_baseline,
        }
        if others:
            line["workloads"] = others
        if parity is not None:
            line["parity"] = parity
        emit(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse_args()
    # Libraries (NCCL's version banner, OpenMP notices) may write to stdout;
    # stdout must carry exactly ONE JSON line, so everything else goes to
    # stderr until the line is printed.
    sys.stdout.flush()
    _stdout_fd = os.dup(1)
    os.dup2(2, 1)
    _emit = lambda text: os.write(_stdout_fd, (text + "\n").encode())  # noqa: E731
    if a.impl == "reference":
        main_reference(a, _emit)
    else:
        main_b200(a, _emit)
