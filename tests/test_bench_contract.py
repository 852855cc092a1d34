"""bench.py contract on CPU: the reference arm prints exactly one JSON line on stdout with the agreed keys (the B200 arm needs a
GPU and is exercised by the driver)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_reference(extra_env=None, args=()):
    env = dict(os.environ, SABC_BENCH_REF_SECONDS="2", **(extra_env or {}))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "3", *args],
                       capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout


def test_reference_arm_json_line():
    out = run_reference()
    lines = [l for l in out.splitlines() if l.strip()]
    assert len(lines) == 1, out
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "particle-sim-updates/sec" and d["unit"] == "particle-updates/s"
    assert d["steps"] == 2 and d["warmup"] == 3 and d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["dtype"] == "f64" and d["data"] == "synthetic" and "workload" in d["config"] and "SIR" in d["config"]["workload"]
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["gpu_launches"] == 0
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "population updates" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_quietly():
    out = run_reference({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}, ("--gpus", "2"))
    assert out.strip() == ""


def test_reference_arm_ignores_launcher_thread_cap():
    d = json.loads(run_reference({"OMP_NUM_THREADS": "1"}).strip())
    assert d["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))
