"""CPU tests of the host side: the C-ABI library loads and exports every symbol include/sabc_b200.h declares, argument
validation mirrors the reference's errors, proposals' constructors behave like src/proposals.jl, and compute entry points
fail loudly (no CPU fallback) when no CUDA device is present."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import sabc_b200 as sb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "sabc_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sabc_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    names = declared_symbols()
    assert len(names) >= 40
    lib = sb._lib.lib()
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/sabc_b200.h but not exported"
        assert n in sb._lib.SYMBOLS, f"{n} has no ctypes signature"
    out = subprocess.run(["nm", "-D", "--defined-only", sb._lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (sabc_[a-z0-9_]+)", out))
    assert set(names) <= exported
    assert lib.sabc_abi_version() == 2


def test_cuda_code_is_sm_100a():
    out = subprocess.run(["cuobjdump", "-lelf", sb._lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out


def test_model_registry():
    names = sb.models.registered()
    for n in ("gauss_mean", "gauss_sample_d1s1", "gauss_sample_d2s2", "logistic", "sir_tauleap"):
        assert n in names
    d, s = C.c_int32(), C.c_int32()
    assert sb._lib.lib().sabc_model_info(b"sir_tauleap", C.byref(d), C.byref(s)) == 0 and (d.value, s.value) == (4, 3)
    assert sb._lib.lib().sabc_model_info(b"nope", C.byref(d), C.byref(s)) == -20


def test_proposal_constructors():
    """test/runtests.jl:202-209 and src/proposals.jl:29-36,85-99."""
    with pytest.raises(TypeError):
        sb.DifferentialEvolution(1, 0.1)            # positional arguments: MethodError in Julia
    with pytest.raises(TypeError):
        sb.DifferentialEvolution(1)
    with pytest.raises(ValueError):
        sb.DifferentialEvolution(gamma0=1, n_para=5)
    with pytest.raises(ValueError):
        sb.DifferentialEvolution(gamma0=1, n_para=5, sigma_gamma=1.4)
    with pytest.raises(ValueError):
        sb.DifferentialEvolution()
    assert sb.DifferentialEvolution(n_para=2).params() == (2.38 / 2.0, 1e-5)
    assert sb.DifferentialEvolution(**{"γ0": 0.7}).params() == (0.7, 1e-5)
    assert sb.StretchMove().params() == (2.0, 0.0)
    assert sb.RandomWalk(n_para=2).params() == (0.8, 0.0)
    for bad in (0.0, -0.1, 1.1):
        with pytest.raises(RuntimeError):
            sb.RandomWalk(n_para=1, beta=bad)


def test_argument_validation_before_any_device_work():
    model, prior = sb.models.gauss_mean(1.0), sb.Normal(0, 1)
    with pytest.raises(RuntimeError, match="too small"):                  # :155-156; test/runtests.jl:39-54
        sb.sabc(model, prior, n_particles=100, n_simulation=10)
    with pytest.raises(RuntimeError, match="too small"):
        sb.sabc(model, prior, n_particles=100, n_simulation=10, v=-0.1)
    with pytest.raises(RuntimeError, match="algorithm"):                  # :462-464
        sb.sabc(model, prior, n_particles=100, n_simulation=1000, algorithm="both_eps")
    with pytest.raises(TypeError, match="closures"):
        sb.sabc(lambda th: abs(th), prior, n_particles=100, n_simulation=1000)
    with pytest.raises(RuntimeError, match="type"):
        sb.sabc(model, prior, n_particles=100, n_simulation=1000, type="triple")


def test_extra_f_dist_arguments_bind_into_the_model():
    """f_dist(θ, args...; kwargs...) (:163,174,315,421): with a device-model factory as f_dist the extra arguments of sabc() /
    update_population!() become the model's data, exactly what a closure over them would have captured."""
    from sabc_b200.api import _resolve_model
    m = _resolve_model(sb.models.gauss_mean, (1.0,), {"sigma": 2.0, "n_obs": 4})
    assert isinstance(m, sb.DeviceModel) and m.name == "gauss_mean" and np.array_equal(m.par, [1.0, 1.0])
    m2 = _resolve_model(sb.models.gauss_sample, (10, 2.0, 42.5), {"n_para": 2, "second_is_sum": True})
    assert m2.name == "gauss_sample_d2s2" and (m2.n_para, m2.n_stats) == (2, 2)
    assert _resolve_model(m, (), {}) is m
    with pytest.raises(TypeError, match="already holds its data"):
        _resolve_model(m, (3.0,), {})
    with pytest.raises(TypeError, match="closures"):
        _resolve_model(lambda th, y: abs(th - y), (1.0,), {})
    with pytest.raises(TypeError):                                          # a factory called with arguments it does not take
        _resolve_model(sb.models.gauss_mean, (), {"not_a_parameter": 1})


def test_config_validation_in_the_library():
    L = sb._lib
    model, prior = sb.models.gauss_mean(1.0), sb.Normal(0, 1)
    ok = dict(n_particles=100, algorithm="single_eps", proposal=sb.DifferentialEvolution(n_para=1), resample=200, v=1.0, delta=0.1)
    bad_model = sb.DeviceModel("not_registered", 1, 1, np.zeros(2))
    with pytest.raises(sb.SABCError) as ei:
        sb.Engine(bad_model, prior, **ok)
    assert ei.value.code == -20
    with pytest.raises(sb.SABCError) as ei:                               # prior length must match the model
        sb.Engine(model, sb.product_distribution([sb.Normal(0, 1), sb.Uniform(0, 1)]), **ok)
    assert ei.value.code == -20
    with pytest.raises(sb.SABCError) as ei:
        sb.Engine(model, prior, **{**ok, "n_particles": 2})
    assert ei.value.code == -20
    assert L.lib().sabc_destroy(None) == 0                                # idempotent on NULL


def test_no_cpu_fallback():
    """Without a CUDA device every compute entry point must fail with SABC_ERR_CUDA instead of computing on the host."""
    n = C.c_int(0)
    rc = sb._lib.lib().sabc_device_count(C.byref(n))
    if rc == 0 and n.value > 0:
        pytest.skip("a CUDA device is present")
    x = np.ones(4); out = np.zeros(4)
    assert sb._lib.lib().sabc_detmath(0, sb._lib.ptr(x), 4, sb._lib.ptr(out)) == -30
    with pytest.raises(sb.SABCError) as ei:
        sb.Engine(sb.models.gauss_mean(1.0), sb.Normal(0, 1), n_particles=100, algorithm="single_eps",
                  proposal=sb.DifferentialEvolution(n_para=1), resample=200, v=1.0, delta=0.1)
    assert ei.value.code == -30
    assert b"CUDA" in sb._lib.lib().sabc_last_error()


def test_product_never_touches_the_oracle():
    """the product tree must not reference oracle/ (parity would be void)."""
    pkg = os.path.join(ROOT, "simulatedannealingabc.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".inl", ".jl")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle" not in text.lower() or f in ("hooks.cu",) and "liboracle" not in text, f"{f} mentions the oracle"
    out = subprocess.run(["ldd", sb._lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out


def test_exchange_plan_is_consistent():
    """multi-GPU resampling: the per-rank send/recv plans must tile every slice exactly once and match pairwise."""
    rng = np.random.default_rng(0)
    for world in (1, 2, 4, 8):
        for _ in range(50):
            n_local = int(rng.integers(1, 5000))
            counts = rng.multinomial(n_local * world, rng.dirichlet(np.ones(world) * rng.uniform(0.2, 50))).astype(np.int64)
            plans = []
            for me in range(world):
                arrs = [np.zeros(world, dtype=np.int64) for _ in range(4)]
                rc = sb._lib.lib().sabc_mg_exchange_plan(sb._lib.ptr(counts), world, n_local, me, *[sb._lib.ptr(a) for a in arrs])
                assert rc == 0
                plans.append(arrs)
            for me in range(world):
                s_off, s_cnt, r_off, r_cnt = plans[me]
                assert s_cnt.sum() == counts[me] and r_cnt.sum() == n_local
                filled = np.zeros(n_local, dtype=int)
                for g in range(world):
                    assert r_cnt[g] == plans[g][1][me]                     # what g sends to me is what I expect from g
                    filled[r_off[g]:r_off[g] + r_cnt[g]] += 1
                assert np.all(filled == 1)
                sent = np.zeros(int(counts[me]), dtype=int)
                for d in range(world):
                    sent[s_off[d]:s_off[d] + s_cnt[d]] += 1
                assert np.all(sent == 1)
    bad = np.array([3, 3], dtype=np.int64); arrs = [np.zeros(2, dtype=np.int64) for _ in range(4)]
    assert sb._lib.lib().sabc_mg_exchange_plan(sb._lib.ptr(bad), 2, 4, 0, *[sb._lib.ptr(a) for a in arrs]) == -20


def test_prior_constructors_validate_like_distributions_jl():
    """Distributions.jl throws DomainError for invalid parameters at construction; the mirror raises ValueError."""
    for bad in (lambda: sb.Uniform(1.0, 1.0), lambda: sb.Normal(0.0, 0.0), lambda: sb.Exponential(-1.0), lambda: sb.LogNormal(0.0, -1.0),
                lambda: sb.Gamma(0.0, 1.0), lambda: sb.Gamma(1.0, -2.0), lambda: sb.Beta(-1.0, 1.0), lambda: sb.Beta(1.0, 0.0),
                lambda: sb.Cauchy(0.0, 0.0), lambda: sb.Laplace(0.0, -1.0), lambda: sb.Weibull(0.0, 1.0), lambda: sb.InverseGamma(1.0, 0.0)):
        with pytest.raises(ValueError):
            bad()
    p = sb.product_distribution([sb.Gamma(2.0, 0.5), sb.Beta(2.0, 3.0), sb.Uniform(0, 1)])
    assert len(p) == 3 and [c.kind for c in p.components()] == [4, 5, 0] and p.components()[0].params() == (2.0, 0.5)
    with pytest.raises(TypeError):
        sb.product_distribution([sb.Gamma(2.0, 0.5), "Normal(0,1)"])


def test_multinomial_split_counts_are_multinomial():
    """the per-rank counts of the sharded resampling: sum to N, deterministic in (seed, resampling count), and each marginal is the
    Binomial(N, w_g / W) it must be (chi-square against scipy over many resampling counts), for weights from balanced to extreme."""
    from scipy import stats
    lib = sb._lib.lib()
    for N, w in ((100_000, [3, 1, 1, 1]), (1000, [1, 1]), (37, [5, 1, 0, 2]), (10_000_000, [2**40, 2**40 + 12345, 2**39, 2**41, 1, 2**40, 2**40, 2**40]),
                 (5000, [10**6, 1])):
        w = np.array(w, dtype=np.uint64); G = w.size
        reps = 4000 if N <= 100_000 else 600
        out = np.zeros((reps, G), dtype=np.int64)
        for r in range(reps):
            assert lib.sabc_multinomial_split(N, sb._lib.ptr(w), G, 0x5ABC, r, sb._lib.ptr(out[r])) == 0
        assert np.all(out.sum(axis=1) == N) and np.all(out >= 0)
        again = np.zeros(G, dtype=np.int64)
        lib.sabc_multinomial_split(N, sb._lib.ptr(w), G, 0x5ABC, 7, sb._lib.ptr(again))
        assert np.array_equal(again, out[7])
        p = w.astype(np.float64) / w.astype(np.float64).sum()
        for g in range(G):
            if p[g] == 0:
                assert np.all(out[:, g] == 0)
                continue
            m, v = N * p[g], N * p[g] * (1 - p[g])
            assert abs(out[:, g].mean() - m) < 5 * np.sqrt(v / reps) + 1e-9, (N, g, out[:, g].mean(), m)
            if v > 1:
                assert 0.85 < out[:, g].var() / v < 1.15, (N, g, out[:, g].var(), v)
            if N <= 1000:                                   # full chi-square of the marginal
                lo, hi = out[:, g].min(), out[:, g].max()
                obs = np.bincount(out[:, g] - lo, minlength=hi - lo + 1).astype(float)
                exp = stats.binom(N, p[g]).pmf(np.arange(lo, hi + 1)) * reps
                keep = exp >= 5
                o = np.append(obs[keep], obs[~keep].sum()); e = np.append(exp[keep], reps - exp[keep].sum())
                chi = ((o - e) ** 2 / np.maximum(e, 1e-9)).sum()
                assert stats.chi2.sf(chi, o.size - 1) > 1e-4, (N, g, chi)
