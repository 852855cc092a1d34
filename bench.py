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
    sampling every 10 ms, started before the timed region and killed after the end-to-end leg."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.samples = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "10"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return
        time.sleep(0.03)
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



def csrc_digest() -> str:
    """sha256 over the CUDA sources with comments and white space removed: stamps the committed ncu counts, so that a number taken
    from an older kernel is never reported (and an edited comment does not invalidate it)."""
    import hashlib
    import re
    h = hashlib.sha256()
    d = os.path.join(ROOT, "simulatedannealingabc.jl_b200", "csrc")
    for f in sorted(os.listdir(d)):
        if f.endswith((".cu", ".cuh", ".inl", ".h")):
            text = open(os.path.join(d, f), errors="replace").read()
            text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
            text = re.sub(r"//[^\n]*", "", text)
            h.update(f.encode()); h.update(re.sub(r"\s+", "", text).encode())
    return h.hexdigest()[:16]


def ncu_counts(key: str):
    """per-launch counts of the dominant kernel from the committed `ncu --set full` capture (tools/make_kernel_counts.py):
    DRAM bytes, warp instructions, thread instructions, issue-slot utilisation under ncu.  None if the capture was taken from
    different kernel sources than the ones this run uses."""
    try:
        tab = json.load(open(os.path.join(ROOT, "profiles", "r2_kernel_counts.json")))
        rec = tab["kernels"].get(key)
        if rec is None:
            return None, "no capture for this kernel / size"
        if tab.get("csrc_sha256_16") != csrc_digest():
            return None, f"capture is stale (taken at csrc {tab.get('csrc_sha256_16')}, running {csrc_digest()})"
        return rec, tab.get("source")
    except Exception as ex:
        return None, f"profiles/r2_kernel_counts.json unreadable ({type(ex).__name__})"


def run_config(sb, name, n_particles, steps, warmup, device, flags=0):
    """one device-resident measurement of another configuration (the `extra` object): (value, ms_per_step, kernel ms, accept fraction)."""
    model, prior, algorithm, _, desc = workload(name)
    eng = sb.Engine(model, prior, n_particles=n_particles, algorithm=algorithm, proposal=sb.DifferentialEvolution(n_para=model.n_para),
                    resample=2 * n_particles, v=1.0, delta=0.1, device=device, flags=sb.SABC_FLAG_TIME_KERNELS | flags)
    eng.init(); eng.update(warmup * n_particles)
    c0 = eng.get_state()[1].copy()
    eng.update(steps * n_particles)
    t = eng.timing(); c1 = eng.get_state()[1]
    acc = float(c1[1] - c0[1]) / float(steps * n_particles)
    out = {"workload": desc, "n_particles": n_particles, "steps": steps, "value": steps * n_particles / (t["update_ms"] * 1e-3),
           "ms_per_step": t["update_ms"] / steps, "avg_kernel_ms": t["kernel_ms"] / max(t["kernel_launches"], 1), "accept_fraction": acc}
    d, s_ = model.n_para, model.n_stats
    peak, _ = measured_peak_hbm()
    out["hbm_frac_algorithmic"] = algorithmic_bytes_per_update(d, s_, acc) * (n_particles / 2) / (out["avg_kernel_ms"] * 1e-3) / 1e9 / peak
    eng.close()
    return out


def mg_parity(sb, dist, rank, world, local_rank):
    """multi-GPU correctness evidence in the driver's own run, OUTSIDE the timed region (the oracle is the checker here):
    replicated mode against the oracle bit for bit, the invariants that define the sharded mode, and the strict resampling
    variant against the single-GPU multiset."""
    import oracle_binding as ob
    from helpers import model_cases
    res = {}

    def gather(x):
        out = [None] * world
        dist.all_gather_object(out, x)
        return out

    model, prior = model_cases()["sir_tauleap"]
    N = 1250 * world
    kw = dict(n_particles=N, algorithm="single_eps", proposal=sb.DifferentialEvolution(n_para=4), resample=N // 8, v=1.0, delta=0.1)
    comm = sb.api._distributed_setup("torch")
    eng = sb.Engine(model, prior, device=local_rank, rank=comm[0], world_size=comm[1], nccl_unique_id=comm[2], flags=sb.SABC_FLAG_MG_REPLICATED, **kw)
    eng.init(); eng.update(8 * N)
    orc = ob.OracleEngine(model, prior, **kw); orc.init(); orc.update(8 * N)
    same = all(np.array_equal(a, b) for a, b in zip(eng.get_population(), orc.get_population())) and \
        np.array_equal(eng.get_state()[0], orc.get_state()[0]) and np.array_equal(eng.get_state()[1], orc.get_state()[1]) and \
        all(np.array_equal(a, b) for a, b in zip(eng.get_history(), orc.get_history()))
    res["replicated_vs_oracle"] = "bit-exact" if all(gather(bool(same))) else "MISMATCH"
    eng.close(); orc.close()

    model, prior = model_cases()["gauss_sample_d2s2"]
    N = 4096 * world
    kw = dict(n_particles=N, algorithm="multi_eps", proposal=sb.DifferentialEvolution(n_para=2), resample=N // 2, v=1.0, delta=0.1, seed=7)
    comm = sb.api._distributed_setup("torch")
    eng = sb.Engine(model, prior, device=local_rank, rank=comm[0], world_size=comm[1], nccl_unique_id=comm[2], **kw)
    eng.init(); eng.update(20 * N)
    th, u, r = eng.get_population()
    st = gather([eng.get_state()[0].tolist(), eng.get_state()[1].tolist(), [h.tolist() for h in eng.get_history()]])
    all_u = np.concatenate(gather(u)); all_r = np.concatenate(gather(r))
    eh, uh, rh = eng.get_history()
    ok = all(x == st[0] for x in st) and np.allclose(uh[-1], all_u.mean(axis=0), rtol=1e-12, atol=1e-15) and \
        np.allclose(rh[-1], all_r.mean(axis=0), rtol=1e-10) and eng.get_state()[1][2] >= 2 and eng.get_state()[1][3] == 20
    res["sharded_invariants"] = "ok" if all(gather(bool(ok))) else "VIOLATED"
    res["sharded_resamplings"] = int(eng.get_state()[1][2])
    eng.close()

    comm = sb.api._distributed_setup("torch")
    kw["resample"] = 2 * N
    eng = sb.Engine(model, prior, device=local_rank, rank=comm[0], world_size=comm[1], nccl_unique_id=comm[2], flags=sb.SABC_FLAG_MG_STRICT_RESAMPLE, **kw)
    eng.init()
    th, u, r = eng.get_population()
    all_th = np.concatenate(gather(th)); all_u = np.concatenate(gather(u))
    eq = True
    if rank == 0:
        ref = sb.Engine(model, prior, device=local_rank, **kw); ref.init()
        th1, u1, _ = ref.get_population()
        key = lambda a: np.sort(np.ascontiguousarray(a).view([("", a.dtype)] * a.shape[1]).ravel())      # noqa: E731
        eq = bool(np.array_equal(key(all_th), key(th1)) and np.array_equal(key(all_u), key(u1)) and np.array_equal(eng.get_state()[0], ref.get_state()[0]))
        ref.close()
    res["strict_resample_multiset_vs_1gpu"] = "equal" if all(gather(eq)) else "DIFFERENT"
    eng.close()
    return res

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
    ap.add_argument("--no-extra", action="store_true", help="skip the `extra` object (other configurations) and, on several GPUs, `mg_parity`")
    ap.add_argument("--ecdf-knots", type=int, default=0, help="compressed ECDF mode: keep K quantiles (0 = the reference's full ECDF)")
    ap.add_argument("--flags", type=int, default=0, help="extra SABC_FLAG_* bits for the engine (tuning)")
    ap.add_argument("--live-kernel-timing", action="store_true", help="take `value` from the pass with CUDA event pairs around every launch of the dominant kernel "
                    "(direct launches) instead of the product's default path (CUDA graph replay); about 10 %% slower on C4")
    ap.add_argument("--graph", action="store_true", help="(default since round 2, kept for old command lines)")
    ap.add_argument("--single-process", action="store_true", help="with --gpus N and no torchrun: ONE process drives the N GPUs through one handle (sabc_config.n_gpus)")
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
    n_group = args.gpus if (args.single_process and world == 1 and args.gpus > 1) else 0     # GPUs behind ONE handle
    n_gpus_total = max(world, n_group, 1)
    N = n_per_gpu * n_gpus_total

    def config(**more):
        c = {"workload": desc, "particles_per_gpu": n_per_gpu, "n_particles": N, "n_para": d, "n_stats": s, "algorithm": algorithm,
             "proposal": "DifferentialEvolution", "resample": 2 * N, "checkpoint_history": 1, "ecdf_max_knots": args.ecdf_knots,
             "rng": "Philox4x32-10 counter streams"}
        c.update(more)
        return c

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return
        N = n_per_gpu * max(world, args.gpus)
        val, cores, sample, ms = cpu_oracle_throughput(model, prior, algorithm, target_seconds=float(os.environ.get("SABC_BENCH_REF_SECONDS", "60")), steps=args.steps,
                                                       warmup=args.warmup, max_particles=N)
        line = {"impl": "reference", "metric": "particle-sim-updates/sec", "value": val, "unit": "particle-updates/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": config(),
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

    time_kernels_live = args.live_kernel_timing and world == 1 and n_group == 0
    flags = (sb.SABC_FLAG_TIME_KERNELS if time_kernels_live else 0) | args.flags
    kw = dict(n_particles=N, algorithm=algorithm, proposal=sb.DifferentialEvolution(n_para=d), resample=2 * N, v=1.0, delta=0.1,
              device=local_rank)
    comm = sb.api._distributed_setup("torch") if world > 1 else (0, 1, None)
    kw["ecdf_max_knots"] = args.ecdf_knots
    eng = sb.Engine(model, prior, rank=comm[0], world_size=comm[1], nccl_unique_id=comm[2], flags=flags, n_gpus=n_group, **kw)
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
    cnt1 = eng.get_state()[1]
    ms_total = max_over_ranks(t["update_ms"])
    value = args.steps * N / (ms_total * 1e-3)
    accept_frac = float(cnt1[1] - cnt0[1]) / float(args.steps * N)

    # kernel time of the dominant kernel: live in the timed region, or from a separate pass when the graph is replayed
    if time_kernels_live:
        kernel_ms, kernel_launches, how = t["kernel_ms"], t["kernel_launches"], "CUDA events around every launch of the dominant kernel inside the timed region"
        share = kernel_ms / t["update_ms"]
    else:
        e2 = sb.Engine(model, prior, rank=0, world_size=1, flags=sb.SABC_FLAG_TIME_KERNELS, **{**kw, "n_particles": n_per_gpu, "resample": 2 * n_per_gpu})
        e2.init(); e2.update(args.warmup * n_per_gpu); e2.update(args.steps * n_per_gpu)
        t2 = e2.timing()
        kernel_ms, kernel_launches, how = t2["kernel_ms"], t2["kernel_launches"], ("second timed pass of the same K steps inside this run, direct launches with CUDA event pairs around every launch "
                                                                                   "of the dominant kernel (one GPU's slice); `value` comes from the first pass, the product's default path")
        share = t2["kernel_ms"] / t2["update_ms"]
        value_event_pass = args.steps * n_per_gpu / (t2["update_ms"] * 1e-3)
        e2.close()
    else_pass = None if time_kernels_live else value_event_pass
    kinfo = eng.kernel_info()
    bytes_per_update = algorithmic_bytes_per_update(d, s, accept_frac)
    avg_kernel_ms = kernel_ms / max(kernel_launches, 1)
    bytes_per_launch = bytes_per_update * (n_per_gpu / 2)
    hbm_achieved = bytes_per_launch / (avg_kernel_ms * 1e-3) / 1e9
    peak, peak_src = measured_peak_hbm()
    heavy = model.name in ("sir_tauleap", "logistic") or model.name.startswith("sir_gillespie")
    kname = (f"simulate_accept_kernel<{model.name}>" if heavy else f"update_half_kernel<{model.name}, DE>")
    counts, counts_src = ncu_counts(f"{kname}@{n_per_gpu // 2}")

    # end-to-end through the host-buffer call (what Julia's update_population!(::SABCresult) would ccall): every step uploads
    # the population (theta,u,rho,eps,counters) from pinned memory, runs one population update and downloads the result
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
    if rank == 0:
        sampler.stop()
    e2e_ms = max_over_ranks(e2e_ms)
    e2e_value = e2e_steps * N / (e2e_ms * 1e-3)
    io_bytes = 8 * nl * (d + 2 * s) + 8 * eng.n_eps + 32
    for b in bufs:
        sb._lib.lib().sabc_host_free(b.ctypes.data_as(C.c_void_p))
    eng.close()

    parity = None
    if world > 1 and not args.no_extra:
        parity = mg_parity(sb, dist, rank, world, local_rank)

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    clocks = sampler.summary()
    launch = ("direct launches with event pairs" if time_kernels_live else
              ("CUDA graph replay (one all-gather per update inside the graph), host looks " + "4 updates ahead" if world > 1 or n_group else "CUDA graph replay"))
    if (args.flags & sb.SABC_FLAG_NO_GRAPH) and not time_kernels_live:
        launch = "direct launches"
    roof = {"kernel": kname + (" (split path: propose -> compacted simulate+accept -> stats)" if heavy else ""),
            "avg_kernel_ms": avg_kernel_ms, "kernel_launches": int(kernel_launches), "kernel_share_of_step": share,
            "updates_per_launch": n_per_gpu / 2, "grid": kinfo, "timing": how, "value_per_gpu_in_the_event_pair_pass": else_pass,
            "hbm": {"achieved": hbm_achieved, "peak": peak, "unit": "GB/s", "frac": hbm_achieved / peak, "peak_source": peak_src,
                    "algorithmic_bytes_per_update": bytes_per_update},
            "traffic": counts["dram_bytes_per_launch"] if counts else None,
            "ncu": counts if counts else None, "ncu_source": counts_src}
    if heavy:
        # this kernel is bound by instruction issue and divergence, not by HBM: achieved = warp instructions
        # per second (count per launch from the committed ncu capture of this very kernel build, time measured live), peak = one
        # warp instruction per scheduler and cycle at the SM clock sampled under load
        sm_mhz = clocks.get("sm_mhz") or 1965.0
        peak_issue = 148 * 4 * sm_mhz * 1e6 / 1e9
        if counts:
            ach = counts["warp_inst_per_launch"] / (avg_kernel_ms * 1e-3) / 1e9
            roof.update({"bound": "issue", "achieved": ach, "peak": peak_issue, "unit": "Gwarp-inst/s", "frac": ach / peak_issue,
                         "lane_efficiency": counts["thread_inst_per_launch"] / counts["warp_inst_per_launch"] / 32.0,
                         "note": "instruction issue x divergence bound (frac = issue-slot utilisation; useful-lane share in lane_efficiency); the HBM figures are in `hbm`"})
        else:
            roof.update({"bound": "issue", "achieved": None, "peak": peak_issue, "unit": "Gwarp-inst/s", "frac": None,
                         "note": "instruction issue x divergence bound; no current ncu instruction count (" + str(counts_src) + "); the HBM figures are in `hbm`"})
    else:
        roof.update({"bound": "hbm", "achieved": hbm_achieved, "peak": peak, "unit": "GB/s", "frac": hbm_achieved / peak,
                     "note": "state streaming + scattered ECDF / partner sectors; the kernel itself is held by L1 data-pipe wavefronts (DESIGN.md section 4)"})
    line = {
        "metric": "particle-sim-updates/sec", "value": value, "unit": "particle-updates/s", "n_gpus": n_gpus_total, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": config(),
        "run": {"accept_fraction": accept_frac, "host_numa": numa, "launch": launch, "processes": ("one per GPU (torchrun)" if world > 1 else "one"),
                "gpus_per_handle": max(n_group, 1),
                "l2": f"per-GPU working set {(8 * n_per_gpu * (d + 2 * s + 1) + 2 * 8 * s * (N + 2)) / 1e6:.0f} MB (state + ECDF tables and index) vs 126 MB L2; no explicit flush"},
        "e2e": {"value": e2e_value, "unit": "particle-updates/s", "h2d_bytes_per_step": io_bytes, "d2h_bytes_per_step": io_bytes,
                "steps": e2e_steps, "ms_per_step": e2e_ms / e2e_steps,
                "how": "sabc_update_host per step: pinned host (theta,u,rho,eps,counters) -> device, one population update, device -> host; "
                       "CUDA events from the first uploaded byte to the last downloaded byte, transfers pipelined with the half-sweeps"},
        "gpu_launches": int(t["total_launches"]),
        "roofline": roof,
        "clocks": clocks,
    }
    if parity is not None:
        line["mg_parity"] = parity
    if world == 1 and n_group == 0 and not args.no_extra:
        # the other BASELINE.json configurations at their nominal sizes, device-resident, a second or two each
        extra = {}
        try:
            extra["c4_full_1e7_one_gpu"] = run_config(sb, "c4", 10_000_000, 10, 3, local_rank)
            extra["c5_1e7"] = run_config(sb, "c5", 10_000_000, 20, 3, local_rank)
            extra["c3_1e6"] = run_config(sb, "c3", 1_000_000, 20, 3, local_rank)
            extra["c2_1e5"] = run_config(sb, "c2", 100_000, 200, 3, local_rank)
            extra["c1_1e3"] = run_config(sb, "c1", 1000, 99, 3, local_rank)
        except Exception as ex:                                          # never lose the main line
            extra["error"] = f"{type(ex).__name__}: {ex}"
        line["extra"] = extra
    if world == 1 and not args.no_cpu_baseline:
        val, cores, sample, _ = cpu_oracle_throughput(model, prior, algorithm, target_seconds=15.0, steps=3, warmup=1, max_particles=n_per_gpu, fixed_steps=False)
        line["cpu_baseline"] = {"value": val, "unit": "particle-updates/s", "cores": cores, "kind": "port", "sample": sample,
                                "note": "C/OpenMP restatement of the reference (oracle/); julia is not installed"}
    emit(line)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
