"""Build libsabc_b200.so (CUDA, sm_100a) in-tree with nvcc.  No torch, no JIT cache: the .so sits next to this file
so that it travels to the GPU box with the repo snapshot."""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# SABC_LIB_OUT / SABC_EXTRA_NVCC_FLAGS: development knobs for building an experimental variant next to the product
# library (objects then go to a variant-specific directory); the product build uses neither.
LIB = os.environ.get("SABC_LIB_OUT") or os.path.join(HERE, "libsabc_b200.so")
EXTRA = os.environ.get("SABC_EXTRA_NVCC_FLAGS", "").split()
OBJDIR = CSRC if not os.environ.get("SABC_LIB_OUT") else LIB + ".obj"
SOURCES = ["engine.cu", "hooks.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false",                      # bit-exact spec: a*b+c is never contracted unless written as fma()
    "-Xcompiler", "-fPIC,-ffp-contract=off,-O2", "-ccbin", "/usr/bin/g++",
    "-I", "/usr/include",
]


def _newer(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(d) <= t for d in deps)


def _deps() -> list[str]:
    out = [os.path.join(HERE, "..", "include", "sabc_b200.h")]
    for f in os.listdir(CSRC):
        out.append(os.path.join(CSRC, f))
    return out


def build(force: bool = False, verbose: bool = False) -> str:
    deps = _deps()
    if not force and _newer(LIB, deps):
        return LIB
    objs = []

    def compile_one(src: str) -> str:
        os.makedirs(OBJDIR, exist_ok=True)
        obj = os.path.join(OBJDIR, src.replace(".cu", ".o"))
        if not force and _newer(obj, deps):
            return obj
        cmd = [NVCC, *FLAGS, *EXTRA, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    cmd = [NVCC, "-shared", "-o", LIB, *objs, "-ldl", "-ccbin", "/usr/bin/g++"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
