"""The product's sampler / model headers compiled for the HOST (g++) against the oracle, bit for bit, and the CPU check of
the PTRS acceptance filter (tools/check_ptrs_filter.cpp).  No GPU needed: these pin the header logic; the device build of
the same headers is pinned by the GPU parity tests."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CXX = ["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-mfma"]


def test_host_compiled_headers_match_the_oracle(tmp_path):
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, capture_output=True)
    exe = str(tmp_path / "host_header_check")
    ora = os.path.join(ROOT, "oracle")
    subprocess.run(CXX + [os.path.join(ROOT, "tests", "host_header_check.cpp"), os.path.join(ora, "liboracle.so"),
                          f"-Wl,-rpath,{ora}", "-o", exe], check=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "0 mismatches" in r.stdout


def test_ptrs_filter_never_changes_a_decision():
    r = subprocess.run([os.path.join(ROOT, "tools", "check_ptrs_filter.sh"), "150000"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert " 0 mismatches" in r.stdout.splitlines()[-1]
