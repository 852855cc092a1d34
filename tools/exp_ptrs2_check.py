"""GPU check of the experimental PTRS attempt (csrc/ptrs2_experimental.cuh, built with -DSABC_EXPERIMENTAL_PTRS2):
SABC_B200_LIB=build/variants/libsabc_ptrs2.so python tools/exp_ptrs2_check.py [attempts]
prints the counts of sabc_ptrs2_check and exits 1 if any decision differs from the product's exact attempt."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sabc_b200 as sb  # noqa: E402

lib = C.CDLL(sb._lib.LIB_PATH)
lib.sabc_ptrs2_check.restype = C.c_int
lib.sabc_ptrs2_check.argtypes = [C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_void_p]
attempts = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000_000
rng = np.random.default_rng(1)
bad = 0
for name, lam in (("lambda 10..1e7", np.concatenate([10 ** rng.uniform(1, 7, 4000), rng.uniform(10, 40, 1000), [10.0, 16.0, 17.0, 25000.0]])),
                  ("C4 range 30..3e4", 10 ** rng.uniform(1.5, 4.5, 4000))):
    lam = np.ascontiguousarray(lam)
    counts = np.zeros(5, dtype=np.int64)
    rc = lib.sabc_ptrs2_check(lam.ctypes.data, lam.size, attempts, C.c_uint64(2024), counts.ctypes.data)
    assert rc == 0, rc
    n, und1, tests, und2, wrong = (int(c) for c in counts)
    print(f"{name}: {n:.3e} attempts, candidate undecided {100 * und1 / n:.3f} %, acceptance tests {tests:.3e}, "
          f"filter 2 undecided {100 * und2 / max(tests, 1):.2f} %, WRONG {wrong}")
    bad += wrong
sys.exit(1 if bad else 0)
