#!/usr/bin/env python
"""bench.py -- particle-sim-updates/sec of the SABC population update (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c4|c1|c2|c3|c5] [--particles n_per_gpu]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...      (N > 1, one rank per GPU)
    python bench.py --impl reference ...      times the CPU oracle (C restatement of the reference, all host threads)

One "step" = one population update = n_particles particle updates (propose, prior, simulate + distance, ECDF, accept;
src/SimulatedAnnealingABC.jl:294-375).  Default workload = BASELINE.json configs[3], the config the metric's target is quoted
on: SIR tau-leap, 4 parameters, 50-step trajectories, 10^7 particles over 8 GPUs = 1.25e6 particles per GPU (weak scaling).
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

# synthetic observations (data, not spec): generated once from θ* with the oracle's model (tools/make_obs.py)
SIR_OBS = (45902.0, 2194.0, 22.0)          # θ* = (0.3, 0.1, 0.01, 0.5), pop 1e5, 50 steps
LOGISTIC_OBS = [15.806, 22.441, 30.036, 44.301, 57.071, 67.076, 87.987, 95.948, 106.452, 118.053, 143.42, 161.398, 169.342,
                192.71, 223.82, 240.195, 203.042, 228.929, 220.256, 246.703]   # θ* = (0.4, 200, 0.1)


def workload(name: str):
    """-> (model, prior, algorithm, default particles per GPU, description); SURVEY.md §8d definitions."""
    import sabc_b200 as sb
    m = sb.models
    U, Nm, prod = sb.Uniform, sb.Normal, sb.product_distribution
    if name == "c1":
        return m.gauss_mean(1.0), Nm(0, 1), "single_eps", 1000, "C1 1-D Gaussian mean, known variance, normal prior, type=:single"
    if name == "c5":
        return m.gauss_mean(1.0), Nm(0, 1), "single_eps", 10_000_000, "C5 1-D Gaussian particle sweep"
    if name == "c2":
        return (m.gauss_sample(10, 2.0, 42.5, n_para=2, second_is_sum=True), prod([Nm(0, 2), U(0, 2)]), "multi_eps", 100_000,
                "C2 Gaussian (mean, sd), 2 summary statistics, type=:multi")
    if name == "c3":
        return (m.logistic(LOGISTIC_OBS), prod([U(0, 1), U(50, 500), U(0, 0.5)]), "single_eps", 1_000_000,
                "C3 stochastic logistic growth, 3 parameters, 20-point series, type=:hybrid")
    if name == "c4":
        return (m.sir_tauleap(*SIR_OBS), prod([U(0.1, 1), U(0.05, 0.5), U(0.001, 0.05), U(0.2, 1)]), "single_eps", 1_250_000,
                "C4 SIR tau-leap, 4 parameters, 50-step trajectories, 3 statistics, 1e7 particles over 8 GPUs")
    raise SystemExit(f"unknown workload {name}")


def algorithmic_bytes_per_update(d: int, s: int, accept_frac: float) -> float:
    """FP64 SoA state.  read θ_i,u_i (8(d+s)) + cached log-prior (8) + two DE partner gathers (16d) + one
    32-byte ECDF leaf sector per statistic (32s) + accepted writes of θ,u,ρ,lp (a(8(d+2s)+8)) + ρ re-read of the rows that
    did not accept, for the ρ history sum ((1-a)8s)."""
    return 8 * (d + s) + 8 + 16 * d + 32 * s + accept_frac * (8 * (d + 2 * s) + 8) + (1 - accept_frac) * 8 * s


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md): one streaming nvidia-smi process
    sampling every 50 ms, started before and killed after the region."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.samples = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return
        time.sleep(0.06)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill(); out = ""
        self.samples = [[x.strip() for x in line.split(",")] for line in out.splitlines() if line.count(",") >= 6]

    def summary(self) -> dict:
        def num(x):
            try:
                return float(x)
            except ValueError:
                return None
        rows = [(num(s[0]), num(s[1]), num(s[2]), s) for s in self.samples]
        rows = [r for r in rows if r[0] is not None]
        # "under load": power above the idle floor (the first sample is taken before the region starts)
        pmax = max([r[2] for r in rows if r[2] is not None], default=0.0)
        loaded = [r for r in rows if r[2] is not None and r[2] >= 0.5 * pmax] or rows
        reasons = set()
        for r in loaded:
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3][3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median([r[0] for r in loaded]) if loaded else None,
                "sm_max_mhz": max([r[1] for r in rows if r[1] is not None], default=None),
                "power_w_max": pmax or None, "reasons": sorted(reasons), "samples": len(rows), "samples_under_load": len(loaded)}


def pin_to_gpu_numa_node(index: int) -> str:
    """Bind this process (and the pinned host buffers it allocates afterwards) to the NUMA node of GPU `index`: with one
    process per GPU the e2e copies otherwise cross the socket interconnect for half of the GPUs."""
    try:
        bus = subprocess.run(["nvidia-smi", f"--id={index}", "--query-gpu=pci.bus_id", "--format=csv,noheader"], capture_output=True,
                             text=True, timeout=10).stdout.strip().lower()
        bus = bus[-12:] if len(bus) > 12 else bus                      # 00000000:1B:00.0 -> 0000:1b:00.0
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        if node < 0:
            return "numa node unknown"
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return f"node {node} ({len(cpus)} cpus)"
        return "no allowed cpu on the node"
    except Exception as ex:                                              # plumbing only: never fail the run
        return f"not pinned ({type(ex).__name__})"


def measured_peak_hbm() -> tuple[float, str]:
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "MEASURED_PEAKS.json (of measured)"
    except Exception:
        return 6650.0, "B200_PROFILING.md fallback (of fallback)"


def cpu_oracle_throughput(model, prior, algorithm, *, target_seconds: float, steps: int, warmup: int, max_particles: int,
                          fixed_steps: bool = True):
    """Time the oracle's update loop (C restatement of the reference, OpenMP over particles exactly where the reference
    uses Threads.@threads) on a bounded sample of the workload.  Returns (updates/s, cores, description, ms_per_step)."""
    import oracle_binding as ob
    import sabc_b200 as sb
    ob.lib().orc_set_num_threads(len(os.sched_getaffinity(0)))      # torchrun exports OMP_NUM_THREADS=1: use every host thread we may run on
    cores = ob.lib().orc_num_threads()
    kw = dict(algorithm=algorithm, proposal=sb.DifferentialEvolution(n_para=model.n_para), v=1.0, delta=0.1)
    n_cal = min(max_particles, max(256 * cores, 20000))
    o = ob.OracleEngine(model, prior, n_particles=n_cal, resample=2 * n_cal, **kw)
    o.init(); o.update(n_cal); o.update(3 * n_cal)
    rate = 3 * n_cal / max(o.seconds, 1e-6)
    n = int(min(max_particles, max(n_cal, rate * target_seconds / max(steps + warmup, 1))))
    if n == max_particles and not fixed_steps:   # the full population already fits the time budget: spend the rest on more updates
        steps = max(steps, min(int(rate * target_seconds / n) - warmup, 200))
    o = ob.OracleEngine(model, prior, n_particles=n, resample=2 * n, **kw)
    o.init()
    if warmup:
        o.update(warmup * n)
    o.update(steps * n)
    sec = o.seconds
    return steps * n / sec, cores, f"{n} particles x {steps} population updates (same model, prior, proposal) in {sec:.2f} s", 1e3 * sec / steps


def main():
    # stdout carries exactly one JSON line: everything libraries print there (e.g. NCCL's version banner) goes to stderr
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(line: dict):
        os.write(real_stdout, (json.dumps(line) + "\n").encode())

    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c4")
    ap.add_argument("--particles", type=int, default=0, help="particles per GPU (default: the workload's)")
    ap.add_argument("--e2e-steps", type=int, default=0, help="steps of the host-buffer (e2e) leg; default min(steps, 50)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ecdf-knots", type=int, default=0, help="compressed ECDF mode: keep K quantiles (0 = the reference's full ECDF)")
    ap.add_argument("--flags", type=int, default=0, help="extra SABC_FLAG_* bits for the engine (tuning)")
    ap.add_argument("--graph", action="store_true", help="replay the CUDA graph in the timed region (kernel times then come from a separate pass)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    import sabc_b200 as sb
    model, prior, algorithm, n_default, desc = workload(args.workload)
    n_per_gpu = args.particles or n_default
    d, s = model.n_para, model.n_stats

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return
        val, cores, sample, ms = cpu_oracle_throughput(model, prior, algorithm, target_seconds=float(os.environ.get("SABC_BENCH_REF_SECONDS", "60")), steps=args.steps,
                                                       warmup=args.warmup, max_particles=n_per_gpu * max(world, args.gpus))
        line = {"impl": "reference", "metric": "particle-sim-updates/sec", "value": val, "unit": "particle-updates/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": desc, "particles_per_gpu": n_per_gpu, "proposal": "DifferentialEvolution"},
                "cpu_baseline": {"value": val, "unit": "particle-updates/s", "cores": cores, "kind": "port", "sample": sample,
                                 "note": "C/OpenMP restatement of the reference (oracle/), not the Julia package: julia is not installed"},
                "e2e": {"value": val, "unit": "particle-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        emit(line)
        return

    # ------------------------------------------------------------------ B200 arm
    dist = None
    numa = pin_to_gpu_numa_node(local_rank) if world > 1 else "single process, not pinned"
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()

    def max_over_ranks(x: float) -> float:
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x: float) -> float:
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    N = n_per_gpu * world
    time_kernels_live = not args.graph and world == 1
    flags = (sb.SABC_FLAG_TIME_KERNELS if time_kernels_live else 0) | args.flags
    kw = dict(n_particles=N, algorithm=algorithm, proposal=sb.DifferentialEvolution(n_para=d), resample=2 * N, v=1.0, delta=0.1,
              device=local_rank)
    comm = sb.api._distributed_setup("torch") if world > 1 else (0, 1, None)
    kw["ecdf_max_knots"] = args.ecdf_knots
    eng = sb.Engine(model, prior, rank=comm[0], world_size=comm[1], nccl_unique_id=comm[2], flags=flags, **kw)
    eng.init()

    # device-resident throughput ("value")
    eng.update(args.warmup * N)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    cnt0 = eng.get_state()[1].copy()
    eng.update(args.steps * N)             # blocking; timed inside with CUDA events on the engine's stream
    barrier()
    t = eng.timing()
    if rank == 0:
        sampler.stop()
    cnt1 = eng.get_state()[1]
    ms_total = max_over_ranks(t["update_ms"])
    value = args.steps * N / (ms_total * 1e-3)
    accept_frac = float(cnt1[1] - cnt0[1]) / float(args.steps * N)

    # kernel time of the dominant kernel (update_half): live in the timed region, or from a separate pass when the graph is replayed
    if time_kernels_live:
        kernel_ms, kernel_launches, how = t["kernel_ms"], t["kernel_launches"], "CUDA events around every update_half launch inside the timed region"
    else:
        e2 = sb.Engine(model, prior, rank=0, world_size=1, flags=sb.SABC_FLAG_TIME_KERNELS, **{**kw, "n_particles": n_per_gpu, "resample": 2 * n_per_gpu})
        e2.init(); e2.update(args.warmup * n_per_gpu); e2.update(args.steps * n_per_gpu)
        t2 = e2.timing()
        kernel_ms, kernel_launches, how = t2["kernel_ms"], t2["kernel_launches"], "separate pass of the same steps with CUDA events around every update_half launch (1 GPU slice)"
        e2.close()
    kinfo = eng.kernel_info()
    bytes_per_update = algorithmic_bytes_per_update(d, s, accept_frac)
    avg_kernel_ms = kernel_ms / max(kernel_launches, 1)
    bytes_per_launch = bytes_per_update * (n_per_gpu / 2)
    achieved = bytes_per_launch / (avg_kernel_ms * 1e-3) / 1e9
    peak, peak_src = measured_peak_hbm()
    traffic = None                         # DRAM bytes per launch of the dominant kernel from the committed ncu capture
    pipes = None                           # and its pipe utilisation from the same capture (what actually bounds a non-HBM kernel)
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json")))
        key = None
        if model.name == "sir_tauleap" and n_per_gpu == 1_250_000:
            key = "simulate_accept_kernel<sir_tauleap>"
        elif model.name == "gauss_mean" and n_per_gpu == 10_000_000:
            key = "update_half_kernel<gauss_mean, DE>@5000000"
        if key:
            traffic = tr.get(key)
            pipes = tr.get("pipes", {}).get(key)
    except Exception:
        pass

    # end-to-end through the host-buffer call (what Julia's update_population!(::SABCresult) would ccall): every step uploads
    # the slice (θ,u,ρ,ε,counters) from pinned memory, runs one population update and downloads the result
    e2e_steps = args.e2e_steps or min(args.steps, 50)
    import ctypes as C
    nl = eng.n_local
    sizes = [nl * d, nl * s, nl * s]
    bufs = []
    for sz in sizes:
        p = C.c_void_p()
        sb._lib.check(sb._lib.lib().sabc_host_alloc(C.byref(p), sz * 8))
        bufs.append(np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_double)), shape=(sz,)))
    th, uu, rr = eng.get_population()
    bufs[0][:] = th.ravel(order="F"); bufs[1][:] = uu.ravel(order="F"); bufs[2][:] = rr.ravel(order="F")
    eps_h, cnt_h = eng.get_state()
    eps_h = eps_h.copy(); cnt_h = cnt_h.copy()
    for _ in range(2):
        eng.update_host(bufs[0], bufs[1], bufs[2], eps_h, cnt_h, N)
    barrier()
    e2e_ms = 0.0
    for _ in range(e2e_steps):
        eng.update_host(bufs[0], bufs[1], bufs[2], eps_h, cnt_h, N)
        tt = eng.timing()
        e2e_ms += tt["host_ms"]
    barrier()
    e2e_ms = max_over_ranks(e2e_ms)
    e2e_value = e2e_steps * N / (e2e_ms * 1e-3)
    io_bytes = 8 * nl * (d + 2 * s) + 8 * eng.n_eps + 32
    for b in bufs:
        sb._lib.lib().sabc_host_free(b.ctypes.data_as(C.c_void_p))

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    line = {
        "metric": "particle-sim-updates/sec", "value": value, "unit": "particle-updates/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": desc, "particles_per_gpu": n_per_gpu, "n_particles": N, "n_para": d, "n_stats": s, "algorithm": algorithm,
                   "proposal": "DifferentialEvolution", "resample": 2 * N, "checkpoint_history": 1, "ecdf_max_knots": args.ecdf_knots,
                   "rng": "Philox4x32-10 counter streams", "accept_fraction": accept_frac,
                   "l2": f"per-GPU working set {(8 * nl * (d + 2 * s + 1) + 8 * s * (N + 2)) / 1e6:.0f} MB (state + ECDF tables) vs 126 MB L2; no explicit flush",
                   "host_numa": numa,
                   "launch": "direct launches with event pairs" if time_kernels_live else ("host-driven + NCCL" if world > 1 else "CUDA graph replay")},
        "e2e": {"value": e2e_value, "unit": "particle-updates/s", "h2d_bytes_per_step": io_bytes, "d2h_bytes_per_step": io_bytes,
                "steps": e2e_steps, "ms_per_step": e2e_ms / e2e_steps,
                "how": "sabc_update_host per step: pinned host (theta,u,rho,eps,counters) -> device, one population update, device -> host; "
                       "CUDA events from the first uploaded byte to the last downloaded byte, transfers pipelined with the half-sweeps"},
        "gpu_launches": int(t["total_launches"]),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                     "kernel": (f"simulate_accept_kernel<{model.name}> (split path: propose -> compacted simulate+accept -> stats)"
                                if model.name in ("sir_tauleap", "logistic") else f"update_half_kernel<{model.name}, DE>"), "avg_kernel_ms": avg_kernel_ms, "kernel_launches": int(kernel_launches),
                     "kernel_share_of_step": kernel_ms / t["update_ms"] if time_kernels_live else None,
                     "algorithmic_bytes_per_update": bytes_per_update, "updates_per_launch": n_per_gpu / 2, "peak_source": peak_src,
                     "grid": kinfo, "timing": how, "ncu_pipes": pipes,
                     "note": ("simulation-heavy model: FP64/INT-issue and divergence bound, not HBM bound (DESIGN.md section 4); profiles/ holds the pipe utilisation"
                              if model.name in ("sir_tauleap", "logistic") else "state streaming + ECDF leaf sectors (DESIGN.md section 4)")},
        "clocks": sampler.summary(),
    }
    if world == 1 and not args.no_cpu_baseline:
        val, cores, sample, _ = cpu_oracle_throughput(model, prior, algorithm, target_seconds=15.0, steps=3, warmup=1, max_particles=n_per_gpu, fixed_steps=False)
        line["cpu_baseline"] = {"value": val, "unit": "particle-updates/s", "cores": cores, "kind": "port", "sample": sample,
                                "note": "C/OpenMP restatement of the reference (oracle/); julia is not installed"}
    emit(line)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
