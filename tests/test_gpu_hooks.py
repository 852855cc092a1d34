"""GPU parity, step by step: every hook of the C ABI runs the engine's device code on given arrays and must agree
BIT FOR BIT with the oracle (north_star check (1): ECDF transforms and accept decisions on identical inputs)."""
import ctypes as C

import numpy as np
import pytest

import oracle_binding as ob
import sabc_b200 as sb
from helpers import (L, g_accept, g_detmath, g_ecdf_build, g_ecdf_transform, model_cases, o_accept, o_detmath, o_ecdf_build,
                     o_ecdf_eval, ptr)

pytestmark = pytest.mark.gpu


def test_detmath_bit_exact(gpu):
    rng = np.random.default_rng(1)
    xs = {
        0: np.concatenate([rng.random(20000), 10 ** rng.uniform(-300, 300, 5000), [0.0, 1.0, 2.0 ** -53, 5e-324, 1e-310, np.inf]]),
        1: np.concatenate([rng.uniform(-745, 709, 20000), rng.uniform(-1, 1, 5000), [0.0, -1000.0, 1000.0, -708.5, -740.0]]),
        2: np.concatenate([rng.random(20000), [0.0, 0.125, 0.25, 0.5, 0.75, 1 - 2.0 ** -53]]),
        4: np.concatenate([np.arange(0, 40.0), [100.0, 1234.0, 1e5, 1e7]]),
    }
    xs[3] = xs[2]
    for op, x in xs.items():
        a, b = g_detmath(op, x), o_detmath(op, x)
        assert np.array_equal(a, b, equal_nan=True), f"op {op}: {np.sum(a != b)} mismatches"


def test_philox_known_answers(gpu):
    # Random123 kat_vectors for philox4x32-10
    kat = [([0] * 4, [0] * 2, [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
           ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
           ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0], [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1])]
    for ctr, key, want in kat:
        out = (C.c_uint32 * 4)()
        L.check(L.lib().sabc_philox((C.c_uint32 * 4)(*ctr), (C.c_uint32 * 2)(*key), out))
        assert list(out) == want


def test_poisson_bit_exact(gpu):
    rng = np.random.default_rng(2)
    lam = np.concatenate([rng.uniform(0, 10, 3000), rng.uniform(10, 50, 3000), 10 ** rng.uniform(1, 5, 3000), [0.0, -1.0, 9.999999, 10.0]])
    k = np.zeros(lam.size, dtype=np.int64); blocks = np.zeros(lam.size, dtype=np.uint32)
    L.check(L.lib().sabc_poisson(ptr(lam), lam.size, C.c_uint64(77), C.c_uint64(5), ptr(k), ptr(blocks)))
    for i, l in enumerate(lam):
        blk = C.c_uint32(0)
        ko = ob.lib().orc_poisson(l, 77, i, 5, C.byref(blk))
        assert ko == k[i] and blk.value == blocks[i], (i, l, ko, k[i])
    big = lam > 100
    assert abs((k[big] / lam[big]).mean() - 1) < 0.01


def test_ptrs_filters_agree_with_exact_test(gpu):
    """The two cheap filters in front of the exact PTRS acceptance test (csrc/philox.cuh) may only shortcut it: on 4e9
    candidates over lambda in [10, 1e7] neither takes a decision the exact test does not take, their error stays well
    inside the bound they use, and they leave only a small share of the candidates to the exact test."""
    rng = np.random.default_rng(11)
    lam = np.concatenate([10 ** rng.uniform(1, 7, 4000), rng.uniform(10, 40, 1000), [10.0, 10.5, 16.0, 17.0, 1e5, 25000.0]])
    counts = np.zeros(5, dtype=np.int64); ratio = np.zeros(2)
    L.check(L.lib().sabc_ptrs_filter_check(ptr(lam), lam.size, 4_000_000_000, C.c_uint64(99), ptr(counts), ptr(ratio)))
    slow, und1, und12, bad1, bad2 = (int(c) for c in counts)
    print("ptrs filters:", dict(reached=slow, undecided_mufu=und1, undecided_both=und12, bad_mufu=bad1, bad_fp64=bad2), ratio)
    assert slow > 5e8 and bad1 == 0 and bad2 == 0
    assert ratio[0] < 0.7 and ratio[1] < 0.7, ratio          # measured error / bound used
    # typical C4 rates (lambda 30 .. 3e4): nearly every candidate is decided by the first filter
    lam = 10 ** rng.uniform(1.5, 4.5, 4000)
    L.check(L.lib().sabc_ptrs_filter_check(ptr(lam), lam.size, 400_000_000, C.c_uint64(100), ptr(counts), ptr(ratio)))
    print("ptrs filters, C4 range:", counts, ratio)
    assert counts[3] == 0 and counts[4] == 0
    assert counts[1] < 0.05 * counts[0] and counts[2] < 0.002 * counts[0]


@pytest.mark.parametrize("n", [5, 100, 2048, 2049, 40000, 700000])
def test_ecdf_build_and_transform(gpu, n):
    rng = np.random.default_rng(n)
    x = rng.lognormal(0, 2, n)
    x[rng.random(n) < 0.05] = 0.0                       # zeros are dropped (cdf_estimators.jl:29)
    if n >= 100:
        x[:20] = x[20:40]                                # duplicates stay
    kg, ko = g_ecdf_build(x), o_ecdf_build(x)
    assert np.array_equal(kg, ko)
    rho = np.concatenate([rng.lognormal(0, 2.5, 5000), x[:200], [0.0, -1.0, np.inf, kg[-1], kg[-2], kg[1], kg[1] / 2, 1e300]])
    ug, uo = g_ecdf_transform(kg, rho), o_ecdf_eval(ko, rho)
    assert np.array_equal(ug, uo), f"{np.sum(ug != uo)} mismatches of {rho.size}"
    assert ug.min() >= 0 and ug.max() <= 1 + 1e-15


def test_ecdf_reference_testset(gpu):
    """test/runtests.jl:9-29 against the device path, plus the SURVEY App. F values."""
    rng = np.random.default_rng(0)
    for data, scale in ((rng.random(100) * 4, 1.0), (np.array([1, 2, 2, 3, 3, 3.0]), 3.0), (np.array([1, 0, 2, 0, 3.0]), 3.0)):
        k = g_ecdf_build(data)
        assert g_ecdf_transform(k, [0.0])[0] <= np.sqrt(np.finfo(float).eps)
        assert abs(g_ecdf_transform(k, [np.inf])[0] - 1) < 1e-8
        assert np.all(np.diff(g_ecdf_transform(k, np.sort(rng.random(100) * scale))) >= 0)
    k = g_ecdf_build(np.array([1, 2, 2, 3, 3, 3.0]))
    assert np.allclose(g_ecdf_transform(k, [2.0, 2.5, 3.0, np.inf]), [2 / 7, 0.5, 4 / 7, 1.0], rtol=0, atol=2e-16)
    k = g_ecdf_build(np.array([1, 0, 2, 0, 3.0]))
    assert np.allclose(g_ecdf_transform(k, [2.0, 2.5, 3.0]), [0.5, 0.625, 0.75], rtol=0, atol=2e-16)


def test_ecdf_no_positive_is_an_error(gpu):
    with pytest.raises(sb.SABCError) as ei:
        g_ecdf_build(np.zeros(10))
    assert ei.value.code == -8


@pytest.mark.parametrize("s,n_eps", [(1, 1), (3, 1), (3, 3), (20, 1)])
def test_accept_step_bit_exact(gpu, s, n_eps):
    rng = np.random.default_rng(s * 10 + n_eps)
    m = 50000
    uo, un = rng.random((m, s)), rng.random((m, s))
    eps = rng.uniform(0.01, 0.5, n_eps)
    dlp = rng.normal(0, 1, m); dlp[::17] = -np.inf
    lf = rng.normal(0, 0.3, m); U = rng.random(m); U[::1001] = 0.0
    a, b = g_accept(uo, un, eps, dlp, lf, U), o_accept(uo, un, eps, dlp, lf, U)
    assert np.array_equal(a, b)
    assert 0.05 < a.mean() < 0.95
    # decisions sitting exactly on the threshold: L = log U to the last bit
    uo2 = np.zeros((m, s)); un2 = np.zeros((m, s)); z = np.zeros(m)
    thr = np.array([ob.lib().orc_log(x) for x in U[:2000]])
    for shift in (-1, 0, 1):
        d = np.nextafter(thr, np.inf if shift > 0 else -np.inf) if shift else thr
        a = g_accept(uo2[:2000], un2[:2000], eps, d, z[:2000], U[:2000]); b = o_accept(uo2[:2000], un2[:2000], eps, d, z[:2000], U[:2000])
        assert np.array_equal(a, b)


def test_epsilon_updates(gpu):
    for ub, v, want in [(0.5, 1, 0.2969523957538746), (0.3, 1, 0.16043438531730103), (0.1, 1, 0.04104440219268037),
                        (0.01, 1, 0.0020911595859812284), (0.3, 0.5, 0.20708121264802898), (1e-4, 2, 2.9223527320399704e-06)]:
        out = C.c_double()
        L.check(L.lib().sabc_update_epsilon_single(ub, v, C.byref(out)))
        assert out.value == ob.lib().orc_eps_single(ub, v)
        assert abs(out.value - want) <= 1e-9 * want
    out = C.c_double()
    L.check(L.lib().sabc_update_epsilon_single(1e-17, 1.0, C.byref(out)))
    assert out.value == 0.0
    rng = np.random.default_rng(3)
    for s in (1, 2, 3, 5, 20):
        for _ in range(20):
            ub = rng.uniform(1e-4, 0.49, s); v = rng.uniform(0.1, 10)
            a = np.zeros(s); b = np.zeros(s)
            L.check(L.lib().sabc_update_epsilon_multi(ptr(ub), s, v, ptr(a)))
            assert ob.lib().orc_eps_multi(ob.p(ub), s, v, ob.p(b)) == 0
            assert np.array_equal(a, b)
    a = np.zeros(2)
    assert L.lib().sabc_update_epsilon_multi(ptr(np.array([0.3, 0.0])), 2, 1.0, ptr(a)) == -5


@pytest.mark.parametrize("n,s", [(1000, 1), (4097, 3), (300000, 2)])
def test_resampling_bit_exact(gpu, n, s):
    rng = np.random.default_rng(n)
    u = np.asfortranarray(rng.random((n, s)) ** 3)
    ubar = u.mean(axis=0)
    qg = np.zeros(n, dtype=np.uint64); qo = np.zeros(n, dtype=np.uint64)
    L.check(L.lib().sabc_resample_weights(ptr(u), n, s, ptr(ubar), 0.1, ptr(qg)))
    ob.lib().orc_resample_weights(ob.p(u), n, s, ob.p(ubar), 0.1, None, ob.p(qo))
    assert np.array_equal(qg, qo)
    ig = np.zeros(n, dtype=np.int64); io = np.zeros(n, dtype=np.int64)
    L.check(L.lib().sabc_resample_indices(ptr(qg), n, C.c_uint64(9), C.c_uint64(3), ptr(ig)))
    ob.lib().orc_resample_indices(ob.p(qo), n, 9, 3, ob.p(io))
    assert np.array_equal(ig, io)
    # multinomial sanity: selection frequency follows the weights
    w = qg.astype(np.float64); cnt = np.bincount(ig, minlength=n)
    assert abs(np.corrcoef(w, cnt)[0, 1]) > 0.05 or n < 2000
    assert ig.min() >= 0 and ig.max() < n


@pytest.mark.parametrize("n", [1, 31, 256, 257, 65536, 65537, 1000003])
def test_treesum_and_exact_mean(gpu, n):
    rng = np.random.default_rng(n)
    x = rng.lognormal(0, 3, n)
    out = C.c_double()
    L.check(L.lib().sabc_treesum(ptr(x), n, C.byref(out)))
    assert out.value == ob.lib().orc_treesum(ob.p(x), n)
    assert abs(out.value - np.sum(x)) <= 1e-12 * np.sum(x)
    u = rng.random(n); u[::7] = 0.0; u[::11] = 1.0
    mg, mo = C.c_double(), C.c_double()
    L.check(L.lib().sabc_exact_mean_u(ptr(u), n, C.byref(mg)))
    ob.lib().orc_exact_mean_u(ob.p(u), n, C.byref(mo))
    assert mg.value == mo.value
    assert abs(mg.value - u.mean()) < 1e-14


def test_prior_logpdf(gpu):
    rng = np.random.default_rng(5)
    kind = np.array([1, 0, 0], dtype=np.int32); par = np.array([0.5, 2.0, -1.0, 3.0, 0.0, 0.5])
    th = np.asfortranarray(np.column_stack([rng.normal(0, 3, 5000), rng.uniform(-2, 4, 5000), rng.uniform(-0.1, 0.6, 5000)]))
    lp = np.zeros(5000)
    L.check(L.lib().sabc_prior_logpdf(3, ptr(kind), ptr(par), ptr(th), 5000, ptr(lp)))
    want = np.array([ob.lib().orc_prior_logpdf(3, ob.p(kind), ob.p(par), ob.p(np.ascontiguousarray(th[i]))) for i in range(5000)])
    assert np.array_equal(lp, want)
    inside = (th[:, 1] >= -1) & (th[:, 1] <= 3) & (th[:, 2] >= 0) & (th[:, 2] <= 0.5)
    assert np.all(np.isneginf(lp[~inside])) and np.all(np.isfinite(lp[inside]))
    ref = -0.5 * ((th[:, 0] - 0.5) / 2) ** 2 - 0.5 * np.log(2 * np.pi) - np.log(2.0) - np.log(4.0) - np.log(0.5)
    assert np.allclose(lp[inside], ref[inside], rtol=1e-13)
    # Exponential(theta) and LogNormal(mu, sigma) against scipy
    from scipy import stats
    kind = np.array([2, 3], dtype=np.int32); par = np.array([1.5, 0.0, 0.5, 0.8])
    th = np.asfortranarray(np.column_stack([rng.uniform(-1, 6, 4000), rng.uniform(-0.5, 8, 4000)]))
    lp = np.zeros(4000)
    L.check(L.lib().sabc_prior_logpdf(2, ptr(kind), ptr(par), ptr(th), 4000, ptr(lp)))
    want = np.array([ob.lib().orc_prior_logpdf(2, ob.p(kind), ob.p(par), ob.p(np.ascontiguousarray(th[i]))) for i in range(4000)])
    assert np.array_equal(lp, want)
    ok = (th[:, 0] >= 0) & (th[:, 1] > 0)
    ref = stats.expon(scale=1.5).logpdf(th[ok, 0]) + stats.lognorm(s=0.8, scale=np.exp(0.5)).logpdf(th[ok, 1])
    assert np.allclose(lp[ok], ref, rtol=1e-12, atol=1e-12) and np.all(np.isneginf(lp[~ok]))


def test_prior_logpdf_gamma_beta(gpu):
    """Gamma(alpha, theta) and Beta(alpha, beta): device == oracle bit for bit, both == scipy, Distributions.jl edge values."""
    from scipy import stats
    rng = np.random.default_rng(6)
    for (a, th_), (al, be) in [((2.5, 0.8), (0.7, 3.0)), ((1.0, 2.0), (1.0, 1.0)), ((0.4, 1.5), (2.0, 0.5))]:
        kind = np.array([4, 5], dtype=np.int32); par = np.array([a, th_, al, be])
        th = np.asfortranarray(np.column_stack([rng.uniform(-1, 8, 4000), rng.uniform(-0.2, 1.2, 4000)]))
        th[:4] = [[0.0, 0.5], [1.0, 0.0], [1.0, 1.0], [0.0, 0.0]]
        lp = np.zeros(4000)
        L.check(L.lib().sabc_prior_logpdf(2, ptr(kind), ptr(par), ptr(th), 4000, ptr(lp)))
        want = np.array([ob.lib().orc_prior_logpdf(2, ob.p(kind), ob.p(par), ob.p(np.ascontiguousarray(th[i]))) for i in range(4000)])
        assert np.array_equal(lp, want, equal_nan=True)      # [0, 0] gives -Inf + Inf = NaN for some parameter sets, as in Julia
        ok = (th[:, 0] > 0) & (th[:, 1] > 0) & (th[:, 1] < 1)
        ref = stats.gamma(a, scale=th_).logpdf(th[ok, 0]) + stats.beta(al, be).logpdf(th[ok, 1])
        assert np.allclose(lp[ok], ref, rtol=1e-11, atol=1e-11)
        out = (th[:, 0] < 0) | (th[:, 1] < 0) | (th[:, 1] > 1)
        assert np.all(np.isneginf(lp[out]))


def test_prior_logpdf_more_families(gpu):
    """Cauchy, Laplace, Weibull, InverseGamma: device == oracle bit for bit (also off the support and at the edges)."""
    rng = np.random.default_rng(8)
    kind = np.array([6, 7, 8, 9], dtype=np.int32)
    for par in ([0.5, 2.0, -1.0, 0.7, 1.7, 2.5, 3.0, 2.0], [0.0, 1.0, 0.0, 1.0, 1.0, 3.0, 0.5, 0.1], [-3.0, 0.01, 2.0, 5.0, 0.6, 1.0, 20.0, 5.0]):
        par = np.array(par)
        th = np.asfortranarray(np.column_stack([rng.normal(0, 5, 4000), rng.normal(0, 3, 4000), rng.uniform(-1, 8, 4000), rng.uniform(-1, 8, 4000)]))
        th[:3] = [[0.0, 0.0, 0.0, 0.0], [1.0, -1.0, 0.0, 1.0], [0.5, 0.0, 1.0, 0.0]]
        lp = np.zeros(4000)
        L.check(L.lib().sabc_prior_logpdf(4, ptr(kind), ptr(par), ptr(th), 4000, ptr(lp)))
        want = np.array([ob.lib().orc_prior_logpdf(4, ob.p(kind), ob.p(par), ob.p(np.ascontiguousarray(th[i]))) for i in range(4000)])
        assert np.array_equal(lp, want, equal_nan=True)
        off = lp[(th[:, 2] < 0) | (th[:, 3] <= 0)]          # off the support: -Inf (NaN where Weibull(alpha < 1) adds +Inf at x = 0)
        assert np.all(np.isneginf(off) | np.isnan(off)) and np.isneginf(off).sum() >= off.size - 1


@pytest.mark.parametrize("name", list(model_cases().keys()))
def test_model_simulate_bit_exact(gpu, name):
    model, prior = model_cases()[name]
    rng = np.random.default_rng(11)
    n = 3000
    comps = prior.components()
    def draw(c):
        if isinstance(c, sb.Normal): return rng.normal(c.mu, c.sigma, n)
        if isinstance(c, sb.Uniform): return rng.uniform(c.a, c.b, n)
        if isinstance(c, sb.Exponential): return rng.exponential(c.theta, n)
        if isinstance(c, sb.Gamma): return rng.gamma(c.alpha, c.theta, n)
        if isinstance(c, sb.Beta): return rng.beta(c.alpha, c.beta, n)
        if isinstance(c, sb.Cauchy): return c.mu + c.sigma * rng.standard_cauchy(n)
        if isinstance(c, sb.Laplace): return rng.laplace(c.mu, c.theta, n)
        if isinstance(c, sb.Weibull): return c.theta * rng.weibull(c.alpha, n)
        if isinstance(c, sb.InverseGamma): return c.theta / rng.gamma(c.alpha, 1.0, n)
        return rng.lognormal(c.mu, c.sigma, n)
    th = np.column_stack([draw(c) for c in comps])
    rho = model.simulate(th, seed=123, particle_base=1000, sweep=7)
    par = np.ascontiguousarray(model.par)
    want = np.zeros((n, model.n_stats)); tmp = np.zeros(model.n_stats)
    for i in range(n):
        ob.lib().orc_model_simulate(ob.model_id(model.name), model.n_para, model.n_stats, ob.p(par), par.size,
                                    ob.p(np.ascontiguousarray(th[i])), 123, 1000 + i, 7, ob.p(tmp))
        want[i] = tmp
    assert np.array_equal(rho, want), f"{np.sum(rho != want)} mismatches"
    assert np.all(rho >= 0)


@pytest.mark.parametrize("d", [1, 2, 4])
@pytest.mark.parametrize("prop", [0, 1, 2])
def test_propose_bit_exact(gpu, d, prop):
    rng = np.random.default_rng(d * 3 + prop)
    n, M = 2000, 777
    act = np.asfortranarray(rng.normal(0, 1, (n, d))); ina = np.asfortranarray(rng.normal(0, 1, (M, d)))
    pp = np.array([[2.38 / np.sqrt(2 * d), 1e-5], [2.0, 0.0], [0.8, 0.0]][prop])
    A = rng.normal(0, 1, (d, d)); chol = np.linalg.cholesky(A @ A.T + np.eye(d)) if d > 1 else np.array([[0.7]])
    chol = np.ascontiguousarray(chol)
    out = np.zeros((n, d), order="F"); lf = np.zeros(n)
    L.check(L.lib().sabc_propose(prop, ptr(pp), d, ptr(act), n, ptr(ina), M, ptr(chol), C.c_uint64(42), 500, C.c_uint64(9), ptr(out), ptr(lf)))
    ina_rm = np.ascontiguousarray(ina)    # oracle wants M x d row-major
    to = np.zeros(d); lfo = C.c_double()
    for i in range(n):
        ob.lib().orc_propose(prop, ob.p(pp), d, ob.p(np.ascontiguousarray(act[i])), ob.p(ina_rm), M, ob.p(chol), 42, 500 + i, 9, ob.p(to), C.byref(lfo))
        assert np.array_equal(out[i], to) and lf[i] == lfo.value, (i, out[i], to, lf[i], lfo.value)
