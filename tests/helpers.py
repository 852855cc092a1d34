"""Shared test helpers: thin numpy wrappers of the product hooks (sabc_b200) and the oracle (oracle_binding)."""
from __future__ import annotations

import ctypes as C

import numpy as np
import oracle_binding as ob
import sabc_b200 as sb

L = sb._lib
ptr = sb._lib.ptr


def obs_logistic():
    # fixed synthetic observation for C3 (generated once from θ* = (0.4, 200, 0.1); data, not part of the spec)
    return np.array([13.9, 18.7, 27.0, 35.2, 47.9, 60.1, 82.5, 99.4, 121.0, 138.6, 158.3, 170.2, 181.9, 186.0, 196.4,
                     190.7, 203.8, 199.1, 205.6, 197.3])


def model_cases():
    m = sb.models
    return {
        "gauss_mean": (m.gauss_mean(1.0), sb.Normal(0.0, 1.0)),
        "gauss_sample_d1s1": (m.gauss_sample(100, 0.0), sb.Uniform(-10, 10)),
        "gauss_sample_d2s1": (m.gauss_sample(100, 0.0, n_para=2), sb.product_distribution([sb.Normal(0, 1), sb.Uniform(0, 1)])),
        "gauss_sample_d1s2": (m.gauss_sample(10, 0.0, 1.0), sb.Normal(0, 1)),
        "gauss_sample_d2s2": (m.gauss_sample(10, 2.0, 42.5, n_para=2, second_is_sum=True),
                              sb.product_distribution([sb.Normal(0, 2), sb.Uniform(0, 2)])),
        "gauss_sample_d2s2_lnexp": (m.gauss_sample(10, 2.0, 42.5, n_para=2, second_is_sum=True),
                                    sb.product_distribution([sb.LogNormal(0.5, 0.8), sb.Exponential(1.5)])),
        "gauss_sample_d2s2_gambeta": (m.gauss_sample(10, 2.0, 42.5, n_para=2, second_is_sum=True),
                                      sb.product_distribution([sb.Gamma(2.5, 0.8), sb.Beta(0.7, 3.0)])),
        "gauss_sample_d2s2_laplinvg": (m.gauss_sample(10, 2.0, 42.5, n_para=2, second_is_sum=True),
                                       sb.product_distribution([sb.Laplace(1.0, 1.5), sb.InverseGamma(3.0, 2.0)])),
        "gauss_sample_d2s2_cauweib": (m.gauss_sample(10, 2.0, 42.5, n_para=2, second_is_sum=True),
                                      sb.product_distribution([sb.Cauchy(1.0, 0.5), sb.Weibull(1.7, 1.2)])),
        "logistic": (m.logistic(obs_logistic()), sb.product_distribution([sb.Uniform(0, 1), sb.Uniform(50, 500), sb.Uniform(0, 0.5)])),
        "sir_gillespie_s3": (m.sir_gillespie(83.0, 24.0, 41.7), sb.product_distribution([sb.Uniform(0.1, 1), sb.Uniform(0.05, 0.5)])),
        "sir_gillespie_s1": (m.sir_gillespie(83.0, 24.0, 41.7, single_stat=True), sb.product_distribution([sb.Uniform(0.1, 1), sb.Uniform(0.05, 0.5)])),
        "sir_tauleap": (m.sir_tauleap(20000.0, 1500.0, 25.0),
                        sb.product_distribution([sb.Uniform(0.1, 1), sb.Uniform(0.05, 0.5), sb.Uniform(0.001, 0.05), sb.Uniform(0.2, 1)])),
    }


def make_pair(model, prior, **kw):
    """(product engine, oracle engine) with identical configuration."""
    flags = kw.pop("flags", 0)
    eng = sb.Engine(model, prior, flags=flags, **kw)
    orc = ob.OracleEngine(model, prior, **kw)
    return eng, orc


def assert_same_state(eng, orc, what=""):
    for name, a, b in zip(("theta", "u", "rho"), eng.get_population(), orc.get_population()):
        assert np.array_equal(a, b), f"{what}: {name} differs, max|diff|={np.nanmax(np.abs(a - b))}, n_diff={(a != b).sum()}"
    (e1, c1), (e2, c2) = eng.get_state(), orc.get_state()
    assert np.array_equal(e1, e2), f"{what}: eps {e1} vs {e2}"
    assert np.array_equal(c1, c2), f"{what}: counters {c1} vs {c2}"


# ---- product hooks ----
def g_detmath(op, x):
    x = np.ascontiguousarray(x, dtype=np.float64); out = np.empty_like(x)
    L.check(L.lib().sabc_detmath(op, ptr(x), x.size, ptr(out)))
    return out


def g_ecdf_build(x):
    x = np.ascontiguousarray(x, dtype=np.float64); k = np.zeros(x.size + 2); n = C.c_int64()
    L.check(L.lib().sabc_ecdf_build(ptr(x), x.size, ptr(k), C.byref(n)))
    return k[:n.value].copy()


def g_ecdf_transform(k, rho):
    k = np.ascontiguousarray(k, dtype=np.float64); rho = np.ascontiguousarray(rho, dtype=np.float64); u = np.empty_like(rho)
    L.check(L.lib().sabc_ecdf_transform(ptr(k), k.size, ptr(rho), rho.size, ptr(u)))
    return u


def g_accept(uo, un, eps, dlp, lf, U):
    uo = np.asfortranarray(uo, dtype=np.float64); un = np.asfortranarray(un, dtype=np.float64)
    m, s = uo.shape
    eps = np.ascontiguousarray(eps, dtype=np.float64); acc = np.zeros(m, dtype=np.uint8)
    L.check(L.lib().sabc_accept_step(m, s, ptr(uo), ptr(un), ptr(eps), eps.size, ptr(np.ascontiguousarray(dlp)),
                                     ptr(np.ascontiguousarray(lf)), ptr(np.ascontiguousarray(U)), ptr(acc)))
    return acc


# ---- oracle unit functions ----
def o_detmath(op, x):
    x = np.ascontiguousarray(x, dtype=np.float64); out = np.empty_like(x)
    s, c = C.c_double(), C.c_double()
    for i, v in enumerate(x):
        if op == 0: out[i] = ob.lib().orc_log(v)
        elif op == 1: out[i] = ob.lib().orc_exp(v)
        elif op in (2, 3):
            ob.lib().orc_sincos2pi(v, C.byref(s), C.byref(c)); out[i] = s.value if op == 2 else c.value
        else: out[i] = ob.lib().orc_logfact(v)
    return out


def o_ecdf_build(x):
    x = np.ascontiguousarray(x, dtype=np.float64); k = np.zeros(x.size + 2)
    n = ob.lib().orc_ecdf_build(ob.p(x), x.size, ob.p(k))
    if n < 0:
        raise ValueError("no positive distance")
    return k[:n].copy()


def o_ecdf_eval(k, rho):
    k = np.ascontiguousarray(k, dtype=np.float64); rho = np.ascontiguousarray(np.atleast_1d(rho), dtype=np.float64); u = np.empty_like(rho)
    ob.lib().orc_ecdf_eval(ob.p(k), k.size, ob.p(rho), rho.size, ob.p(u))
    return u


def o_accept(uo, un, eps, dlp, lf, U):
    uo = np.asfortranarray(uo, dtype=np.float64); un = np.asfortranarray(un, dtype=np.float64)
    m, s = uo.shape
    eps = np.ascontiguousarray(eps, dtype=np.float64); acc = np.zeros(m, dtype=np.uint8)
    ob.lib().orc_accept_step(m, s, ob.p(uo), ob.p(un), ob.p(eps), eps.size, ob.p(np.ascontiguousarray(dlp)),
                             ob.p(np.ascontiguousarray(lf)), ob.p(np.ascontiguousarray(U)), ob.p(acc))
    return acc
