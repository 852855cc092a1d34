"""ctypes binding of libsabc_b200.so (include/sabc_b200.h).  There is no fallback: if the CUDA library is
missing or a call fails, an exception is raised."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# SABC_B200_LIB: development override to load an experimental build of the same CUDA library (tools/, A/B timing)
LIB_PATH = os.environ.get("SABC_B200_LIB") or os.path.join(HERE, "libsabc_b200.so")

c_double_p = C.POINTER(C.c_double)
c_int64_p = C.POINTER(C.c_int64)
c_int32_p = C.POINTER(C.c_int32)

SABC_FLAG_NO_GRAPH = 1
SABC_FLAG_TIME_KERNELS = 2
SABC_FLAG_FUSED = 4
SABC_FLAG_NO_PIPELINE = 8
SABC_FLAG_SORT_WORK = 16
SABC_FLAG_GENERIC_TAIL = 32
SABC_FLAG_MG_REPLICATED = 64
SABC_FLAG_MG_STRICT_RESAMPLE = 128

ERR_NAMES = {
    -1: "NSIM_TOO_SMALL", -2: "BAD_V", -3: "BAD_DELTA", -4: "NEG_DISTANCE", -5: "UBAR_ZERO", -6: "BAD_ALGORITHM",
    -7: "BAD_PROPOSAL", -8: "NO_POSITIVE", -20: "INVALID", -21: "STATE", -30: "CUDA", -31: "NCCL",
}


class SABCError(RuntimeError):
    """Mirror of Julia's ErrorException raised by error(...) in the reference."""

    def __init__(self, code: int, msg: str):
        super().__init__(f"[{ERR_NAMES.get(code, code)}] {msg}")
        self.code = code


class Config(C.Structure):
    _fields_ = [
        ("n_particles", C.c_int64), ("n_para", C.c_int32), ("n_stats", C.c_int32), ("algorithm", C.c_int32),
        ("proposal", C.c_int32), ("prop_par", C.c_double * 2), ("v", C.c_double), ("delta", C.c_double),
        ("resample", C.c_int64), ("seed", C.c_uint64), ("model_name", C.c_char_p), ("model_par", c_double_p),
        ("n_model_par", C.c_int32), ("device", C.c_int32), ("prior_kind", c_int32_p), ("prior_par", c_double_p),
        ("rank", C.c_int32), ("world_size", C.c_int32), ("nccl_unique_id", C.c_void_p), ("flags", C.c_uint32), ("ecdf_max_knots", C.c_int32),
        ("n_gpus", C.c_int32), ("gpu_ids", c_int32_p),
    ]


class Timing(C.Structure):
    _fields_ = [("update_ms", C.c_double), ("kernel_ms", C.c_double), ("kernel_launches", C.c_int64),
                ("total_launches", C.c_int64), ("h2d_ms", C.c_double), ("d2h_ms", C.c_double), ("host_ms", C.c_double),
                ("resample_ms", C.c_double), ("resample_events", C.c_int64)]


# every symbol include/sabc_b200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "sabc_abi_version": (C.c_int, []),
    "sabc_create": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(Config)]),
    "sabc_destroy": (C.c_int, [C.c_void_p]),
    "sabc_last_error": (C.c_char_p, []),
    "sabc_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "sabc_nccl_unique_id": (C.c_int, [C.c_void_p]),
    "sabc_init": (C.c_int, [C.c_void_p]),
    "sabc_update": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64]),
    "sabc_update_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64]),
    "sabc_set_tuning": (C.c_int, [C.c_void_p, C.c_double, C.c_double, C.c_int64, C.c_int32, c_double_p]),
    "sabc_local_particles": (C.c_int, [C.c_void_p, c_int64_p, c_int64_p]),
    "sabc_get_population": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "sabc_set_population": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "sabc_get_state": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "sabc_history_len": (C.c_int, [C.c_void_p, c_int64_p]),
    "sabc_get_history": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "sabc_get_ecdf": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, c_int64_p]),
    "sabc_set_ecdf": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int64]),
    "sabc_get_timing": (C.c_int, [C.c_void_p, C.POINTER(Timing)]),
    "sabc_update_kernel_info": (C.c_int, [C.c_void_p] + [C.POINTER(C.c_int)] * 4),
    "sabc_host_alloc": (C.c_int, [C.POINTER(C.c_void_p), C.c_int64]),
    "sabc_host_free": (C.c_int, [C.c_void_p]),
    "sabc_ecdf_build": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, c_int64_p]),
    "sabc_ecdf_transform": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p]),
    "sabc_accept_step": (C.c_int, [C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "sabc_update_epsilon_single": (C.c_int, [C.c_double, C.c_double, c_double_p]),
    "sabc_update_epsilon_multi": (C.c_int, [C.c_void_p, C.c_int32, C.c_double, C.c_void_p]),
    "sabc_resample_weights": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_double, C.c_void_p]),
    "sabc_resample_indices": (C.c_int, [C.c_void_p, C.c_int64, C.c_uint64, C.c_uint64, C.c_void_p]),
    "sabc_exact_mean_u": (C.c_int, [C.c_void_p, C.c_int64, c_double_p]),
    "sabc_treesum": (C.c_int, [C.c_void_p, C.c_int64, c_double_p]),
    "sabc_detmath": (C.c_int, [C.c_int32, C.c_void_p, C.c_int64, C.c_void_p]),
    "sabc_philox": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "sabc_poisson": (C.c_int, [C.c_void_p, C.c_int64, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p]),
    "sabc_ptrs_filter_check": (C.c_int, [C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_void_p, C.c_void_p]),
    "sabc_prior_logpdf": (C.c_int, [C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "sabc_model_simulate": (C.c_int, [C.c_char_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int64, C.c_uint64, C.c_uint32, C.c_uint64, C.c_void_p]),
    "sabc_model_info": (C.c_int, [C.c_char_p, c_int32_p, c_int32_p]),
    "sabc_propose": (C.c_int, [C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint64, C.c_void_p, C.c_void_p]),
    "sabc_mg_exchange_plan": (C.c_int, [C.c_void_p, C.c_int32, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "sabc_multinomial_split": (C.c_int, [C.c_int64, C.c_void_p, C.c_int32, C.c_uint64, C.c_uint32, C.c_void_p]),
    "sabc_register_model": (C.c_int, [C.c_void_p]),
    "sabc_model_count": (C.c_int, []),
    "sabc_model_name": (C.c_char_p, [C.c_int]),
}

_lib = None


def lib() -> C.CDLL:
    """Load the CUDA library (building it is __graft_entry__.build()'s job).  Raises if it is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: run `python simulatedannealingabc.jl_b200/build.py` "
                              "(there is no CPU fallback)")
        L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        raise SABCError(rc, lib().sabc_last_error().decode("utf-8", "replace"))


def ptr(a: np.ndarray | None):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def f64(a, order="F") -> np.ndarray:
    return np.require(np.asarray(a, dtype=np.float64), requirements=["F_CONTIGUOUS" if order == "F" else "C_CONTIGUOUS", "ALIGNED"])
