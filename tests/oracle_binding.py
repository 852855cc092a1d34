"""ctypes binding of oracle/liboracle.so -- the CPU checker (test infrastructure only)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB_PATH = os.path.join(ORACLE_DIR, "liboracle.so")

vp, i64, i32, u64, u32, dbl = C.c_void_p, C.c_int64, C.c_int32, C.c_uint64, C.c_uint32, C.c_double


class OrcConfig(C.Structure):
    _fields_ = [("n_particles", i64), ("n_para", i32), ("n_stats", i32), ("algorithm", i32), ("proposal", i32),
                ("prop_par", dbl * 2), ("v", dbl), ("delta", dbl), ("resample", i64), ("seed", u64), ("model_id", i32),
                ("n_model_par", i32), ("model_par", C.POINTER(dbl)), ("prior_kind", C.POINTER(i32)), ("prior_par", C.POINTER(dbl)),
                ("ecdf_max_knots", i32)]


MODEL_IDS = {"gauss_mean": 0, "gauss_sample": 1, "logistic": 2, "sir_tauleap": 3, "sir_gillespie": 4}


def model_id(name: str) -> int:
    for prefix in ("gauss_sample", "sir_gillespie"):
        if name.startswith(prefix):
            return MODEL_IDS[prefix]
    return MODEL_IDS[name]


def build_oracle() -> str:
    src = [os.path.join(ORACLE_DIR, f) for f in ("sabc_oracle.c", "sabc_oracle.h", "zig_tables.inc", "Makefile")]
    if not os.path.exists(LIB_PATH) or any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in src):
        subprocess.run(["make", "-C", ORACLE_DIR], check=True, capture_output=True)
    return LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        L = C.CDLL(build_oracle())
        sig = {
            "orc_log": (dbl, [dbl]), "orc_exp": (dbl, [dbl]), "orc_logfact": (dbl, [dbl]),
            "orc_sincos2pi": (None, [dbl, C.POINTER(dbl), C.POINTER(dbl)]),
            "orc_philox4x32_10": (None, [vp, vp, vp]),
            "orc_normal_pair": (None, [u64, u64, C.POINTER(dbl), C.POINTER(dbl)]),
            "orc_poisson": (i64, [dbl, u64, u32, u64, C.POINTER(u32)]),
            "orc_zig_normal": (dbl, [u64, u32, u64, C.POINTER(u32)]),
            "orc_normal_stream": (None, [u64, u32, u64, i32, vp]),
            "orc_treesum": (dbl, [vp, i64]),
            "orc_ecdf_build": (i64, [vp, i64, vp]),
            "orc_ecdf_eval": (None, [vp, i64, vp, i64, vp]),
            "orc_accept_step": (None, [i64, i32, vp, vp, vp, i32, vp, vp, vp, vp]),
            "orc_eps_single": (dbl, [dbl, dbl]), "orc_eps_single_bisect": (dbl, [dbl, dbl]),
            "orc_eps_multi": (C.c_int, [vp, i32, dbl, vp]),
            "orc_resample_weights": (None, [vp, i64, i32, vp, dbl, vp, vp]),
            "orc_resample_indices": (None, [vp, i64, u64, u64, vp]),
            "orc_exact_mean_u": (None, [vp, i64, C.POINTER(dbl)]),
            "orc_prior_logpdf": (dbl, [i32, vp, vp, vp]), "orc_lgamma": (dbl, [dbl]),
            "orc_prior_rand": (None, [i32, vp, vp, u64, u32, vp]),
            "orc_model_simulate": (C.c_int, [i32, i32, i32, vp, i32, vp, u64, u32, u64, vp]),
            "orc_propose": (C.c_int, [i32, vp, i32, vp, vp, i64, vp, u64, u32, u64, vp, C.POINTER(dbl)]),
            "orc_create": (C.c_int, [C.POINTER(vp), C.POINTER(OrcConfig)]), "orc_destroy": (C.c_int, [vp]),
            "orc_init": (C.c_int, [vp]), "orc_update": (C.c_int, [vp, i64, i64, C.POINTER(dbl)]),
            "orc_get_population": (C.c_int, [vp, vp, vp, vp]), "orc_set_population": (C.c_int, [vp, vp, vp, vp, vp, vp]),
            "orc_get_state": (C.c_int, [vp, vp, vp]), "orc_history_len": (i64, [vp]), "orc_get_history": (C.c_int, [vp, vp, vp, vp]),
            "orc_get_ecdf": (i64, [vp, i32, vp]), "orc_set_ecdf": (C.c_int, [vp, i32, vp, i64]),
            "orc_num_threads": (C.c_int, []), "orc_set_num_threads": (None, [C.c_int]), "orc_last_error": (C.c_char_p, []),
        }
        for n, (r, a) in sig.items():
            f = getattr(L, n); f.restype = r; f.argtypes = a
        _lib = L
    return _lib


def p(a):
    return None if a is None else a.ctypes.data_as(vp)


class OracleError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"[{code}] {msg}")
        self.code = code


class OracleEngine:
    """Drives the C oracle with the same arguments as sabc_b200.Engine."""

    def __init__(self, model, prior, *, n_particles, algorithm, proposal, resample, v, delta, seed=0x5ABC, ecdf_max_knots=0):
        comps = prior.components()
        self.N, self.d, self.s = int(n_particles), model.n_para, model.n_stats
        self.n_eps = self.s if algorithm == "multi_eps" else 1
        self._par = np.ascontiguousarray(model.par, dtype=np.float64)
        self._kind = np.array([c.kind for c in comps], dtype=np.int32)
        self._ppar = np.array([q for c in comps for q in c.params()], dtype=np.float64)
        cfg = OrcConfig()
        cfg.n_particles, cfg.n_para, cfg.n_stats = self.N, self.d, self.s
        cfg.algorithm = {"single_eps": 0, "multi_eps": 1}[algorithm]
        cfg.proposal = proposal.kind
        cfg.prop_par[0], cfg.prop_par[1] = proposal.params()
        cfg.v, cfg.delta, cfg.resample, cfg.seed = v, delta, int(resample), int(seed)
        cfg.model_id = model_id(model.name)
        cfg.n_model_par = self._par.size
        cfg.model_par = self._par.ctypes.data_as(C.POINTER(dbl))
        cfg.prior_kind = self._kind.ctypes.data_as(C.POINTER(i32))
        cfg.prior_par = self._ppar.ctypes.data_as(C.POINTER(dbl))
        cfg.ecdf_max_knots = int(ecdf_max_knots)
        self._h = vp()
        self._check(lib().orc_create(C.byref(self._h), C.byref(cfg)))
        self.seconds = 0.0

    def _check(self, rc):
        if rc != 0:
            raise OracleError(rc, lib().orc_last_error().decode("utf-8", "replace"))

    def close(self):
        if self._h:
            lib().orc_destroy(self._h)
            self._h = vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def init(self):
        self._check(lib().orc_init(self._h))

    def update(self, n_simulation, checkpoint_history=1):
        sec = dbl()
        self._check(lib().orc_update(self._h, int(n_simulation), int(checkpoint_history), C.byref(sec)))
        self.seconds = sec.value

    def get_population(self):
        th = np.empty((self.N, self.d), order="F"); u = np.empty((self.N, self.s), order="F"); r = np.empty((self.N, self.s), order="F")
        lib().orc_get_population(self._h, p(th), p(u), p(r))
        return th, u, r

    def set_population(self, theta, u, rho, eps, counters):
        th = np.asfortranarray(theta, dtype=np.float64); uu = np.asfortranarray(u, dtype=np.float64); rr = np.asfortranarray(rho, dtype=np.float64)
        ee = np.ascontiguousarray(eps, dtype=np.float64); cc = np.ascontiguousarray(counters, dtype=np.int64)
        lib().orc_set_population(self._h, p(th), p(uu), p(rr), p(ee), p(cc))

    def get_state(self):
        eps = np.zeros(self.n_eps); cnt = np.zeros(4, dtype=np.int64)
        lib().orc_get_state(self._h, p(eps), p(cnt))
        return eps, cnt

    def get_history(self):
        n = lib().orc_history_len(self._h)
        e = np.zeros((n, self.n_eps)); u = np.zeros((n, self.s)); r = np.zeros((n, self.s))
        lib().orc_get_history(self._h, p(e), p(u), p(r))
        return e, u, r

    def get_ecdf(self, stat):
        L = lib().orc_get_ecdf(self._h, stat, None)
        k = np.zeros(L)
        lib().orc_get_ecdf(self._h, stat, p(k))
        return k

    def set_ecdf(self, stat, knots):
        k = np.ascontiguousarray(knots, dtype=np.float64)
        self._check(lib().orc_set_ecdf(self._h, stat, p(k), k.size))
