"""The C ABI from a plain C program (examples/c_host): compiled against include/sabc_b200.h with gcc, so the header -- not a
ctypes mirror of it -- defines the struct layout.  Without a GPU the program must report SABC_ERR_CUDA; with one it must
reproduce the committed C1 trajectory fixture."""
import json
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "simulatedannealingabc.jl_b200")


@pytest.fixture(scope="module")
def c_host(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("c_host") / "sabc_c_host")
    subprocess.run(["/usr/bin/gcc", "-O2", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "c_host", "sabc_c_host.c"),
                    "-o", exe, "-L", PKG, "-l:libsabc_b200.so", f"-Wl,-rpath,{PKG}"], check=True, capture_output=True)
    return exe


def _have_gpu():
    import ctypes as C
    import sabc_b200 as sb
    n = C.c_int(0)
    return sb._lib.lib().sabc_device_count(C.byref(n)) == 0 and n.value > 0


def test_c_host_fails_loudly_without_gpu(c_host):
    if _have_gpu():
        pytest.skip("a CUDA device is present")
    r = subprocess.run([c_host], capture_output=True, text=True)
    assert r.returncode == 3 and "sabc error -30" in r.stderr and "CUDA" in r.stderr


@pytest.mark.gpu
def test_c_host_reproduces_c1_fixture(gpu, c_host):
    r = subprocess.run([c_host], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    fx = [f for f in json.load(open(os.path.join(ROOT, "tests", "golden", "trajectories.json"))) if f["case"][0] == "gauss_mean"][0]
    out = dict(kv.split("=") for kv in r.stdout.split()[1:])
    got = [int(out[k]) for k in ("n_simulation", "n_accept", "n_resampling", "n_population_updates")]
    assert got == fx["counters"] and int(out["records"]) == len(fx["eps_history"])
    assert float.fromhex(out["eps"]) == float.fromhex(fx["eps"][0])              # bit-identical to the oracle-made fixture


@pytest.mark.gpu
def test_c_host_drives_two_gpus_through_one_handle(gpu, c_host):
    """cfg.n_gpus = 2 from plain C: same counters semantics, a finite posterior sample over the global array (the sharded run is not
    bit-comparable with the one-GPU fixture: partners come from the local halves)."""
    import ctypes as C
    import sabc_b200 as sb
    n = C.c_int(0)
    sb._lib.lib().sabc_device_count(C.byref(n))
    if n.value < 2:
        pytest.skip("needs 2 GPUs")
    r = subprocess.run([c_host, "2"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    line = [l for l in r.stdout.splitlines() if l.startswith("ok ")][0]     # NCCL prints its version banner on stdout as well
    out = dict(kv.split("=") for kv in line.split()[1:])
    assert int(out["n_simulation"]) == 100000 and int(out["n_population_updates"]) == 99 and int(out["n_resampling"]) >= 2
    assert 0.7 < float(out["mean"]) < 1.1 and 0 < float.fromhex(out["eps"]) < 0.05
