"""GPU parity of the whole path: initialization() + update_population!() on the device against the oracle on the same
Philox streams.  DE and Stretch trajectories must be bit-identical (population, u, rho, eps history, counters); north_star
check (2) asks for eps trajectories within 1e-6 relative -- bit identity is stronger.  RandomWalk goes through FP64 tree sums
that both sides define identically, so it is bit-identical too."""
import numpy as np
import pytest

import oracle_binding as ob
import sabc_b200 as sb
from helpers import assert_same_state, make_pair, model_cases

pytestmark = pytest.mark.gpu

DE = lambda d: sb.DifferentialEvolution(n_para=d)   # noqa: E731


def run_pair(model, prior, n_particles, n_updates, algorithm="single_eps", proposal=None, resample=None, v=1.0, delta=0.1,
             checkpoint=1, flags=0, seed=0x5ABC):
    proposal = proposal or DE(model.n_para)
    kw = dict(n_particles=n_particles, algorithm=algorithm, proposal=proposal, resample=resample or 2 * n_particles, v=v, delta=delta, seed=seed)
    eng, orc = make_pair(model, prior, flags=flags, **kw)
    eng.init(); orc.init()
    assert_same_state(eng, orc, "after init")
    for j in range(model.n_stats):
        assert np.array_equal(eng.get_ecdf(j), orc.get_ecdf(j))
    eng.update(n_updates * n_particles, checkpoint); orc.update(n_updates * n_particles, checkpoint)
    assert_same_state(eng, orc, f"after {n_updates} updates")
    for name, a, b in zip(("eps_h", "u_h", "rho_h"), eng.get_history(), orc.get_history()):
        assert a.shape == b.shape, (name, a.shape, b.shape)
        assert np.array_equal(a, b), f"{name}: max rel diff {np.max(np.abs(a - b) / np.abs(b))}"
    return eng, orc


def test_c1_gauss_mean_trajectory(gpu):
    """BASELINE configs[0]: 1-D Gaussian, N=1000, n_simulation=1e5 (99 population updates)."""
    model, prior = model_cases()["gauss_mean"]
    eng, orc = run_pair(model, prior, 1000, 99)
    eps, cnt = eng.get_state()
    assert cnt[0] == 100_000 and cnt[3] == 99 and cnt[2] >= 2 and eps[0] < 0.05


@pytest.mark.parametrize("name", ["gauss_sample_d1s1", "gauss_sample_d2s1", "gauss_sample_d1s2", "gauss_sample_d2s2", "logistic", "sir_tauleap",
                                  "sir_gillespie_s3", "sir_gillespie_s1", "gauss_sample_d2s2_lnexp", "gauss_sample_d2s2_gambeta", "gauss_sample_d2s2_laplinvg", "gauss_sample_d2s2_cauweib"])
@pytest.mark.parametrize("algorithm", ["single_eps", "multi_eps"])
def test_models_trajectory(gpu, name, algorithm):
    model, prior = model_cases()[name]
    run_pair(model, prior, 600, 12, algorithm=algorithm, resample=600)


@pytest.mark.parametrize("prop", ["stretch", "rw"])
@pytest.mark.parametrize("name", ["gauss_mean", "gauss_sample_d2s2", "sir_tauleap"])
def test_proposals_trajectory(gpu, name, prop):
    model, prior = model_cases()[name]
    proposal = sb.StretchMove() if prop == "stretch" else sb.RandomWalk(n_para=model.n_para)
    run_pair(model, prior, 700, 10, proposal=proposal, resample=700)


@pytest.mark.parametrize("n", [4, 5, 255, 257, 513, 2050])
def test_ragged_sizes(gpu, n):
    """odd N (halves N÷2 and N−N÷2, :300-301), sizes around the 256-particle group and the 2048 scan tile."""
    model, prior = model_cases()["gauss_sample_d2s2"]
    run_pair(model, prior, n, 6, algorithm="multi_eps", resample=max(2, n // 2))


def test_larger_population_multilevel_ecdf(gpu):
    """N = 300k: ECDF table > one index level, several scan tiles, many resamplings."""
    model, prior = model_cases()["gauss_mean"]
    eng, orc = run_pair(model, prior, 300_000, 8, resample=100_000)
    assert eng.get_state()[1][2] >= 3


def test_graph_and_direct_launch_agree(gpu):
    model, prior = model_cases()["gauss_sample_d2s2"]
    states = []
    for flags in (0, sb.SABC_FLAG_NO_GRAPH, sb.SABC_FLAG_TIME_KERNELS):
        eng = sb.Engine(model, prior, n_particles=5000, algorithm="multi_eps", proposal=DE(2), resample=5000, v=1.0, delta=0.1, flags=flags)
        eng.init(); eng.update(15 * 5000)
        states.append((eng.get_population(), eng.get_state(), eng.timing()))
    for (p, s, t) in states[1:]:
        for a, b in zip(p, states[0][0]):
            assert np.array_equal(a, b)
        assert np.array_equal(s[0], states[0][1][0]) and np.array_equal(s[1], states[0][1][1])
    assert states[2][2]["kernel_ms"] > 0 and states[2][2]["kernel_launches"] == 30


@pytest.mark.parametrize("name", ["sir_tauleap", "logistic"])
@pytest.mark.parametrize("prop", ["de", "stretch", "rw"])
def test_split_and_fused_kernels_agree(gpu, name, prop):
    """simulation-heavy models run propose -> compacted simulate+accept -> stats; SABC_FLAG_FUSED forces the single fused
    kernel.  Both must reproduce the oracle bit for bit."""
    model, prior = model_cases()[name]
    proposal = {"de": DE(model.n_para), "stretch": sb.StretchMove(), "rw": sb.RandomWalk(n_para=model.n_para)}[prop]
    for flags in (0, sb.SABC_FLAG_FUSED, sb.SABC_FLAG_NO_GRAPH, sb.SABC_FLAG_FUSED | sb.SABC_FLAG_TIME_KERNELS, sb.SABC_FLAG_SORT_WORK):
        run_pair(model, prior, 1500, 8, proposal=proposal, resample=1500, flags=flags)


@pytest.mark.parametrize("name,K", [("gauss_mean", 64), ("gauss_mean", 2046), ("gauss_sample_d2s2", 500), ("sir_tauleap", 1000), ("logistic", 300)])
def test_compressed_ecdf_mode(gpu, name, K):
    """ecdf_max_knots = K: the table keeps K rank-uniform quantiles and lives in shared memory as a whole.  The oracle applies the
    same subsampling rule, so trajectories stay bit-identical; against the full ECDF the transform differs by O(1/K)."""
    model, prior = model_cases()[name]
    N = 6000
    kw = dict(n_particles=N, algorithm="single_eps", proposal=DE(model.n_para), resample=N, v=1.0, delta=0.1, ecdf_max_knots=K)
    eng, orc = make_pair(model, prior, **kw)
    eng.init(); orc.init()
    for j in range(model.n_stats):
        kn = eng.get_ecdf(j)
        assert kn.size <= K + 2 and np.array_equal(kn, orc.get_ecdf(j))
    pow2 = lambda n: 1 << max(1, int(n - 1).bit_length())          # the staged tables are padded to a power of two   # noqa: E731
    assert eng.kernel_info()["smem_bytes"] == sum(pow2(eng.get_ecdf(j).size) for j in range(model.n_stats)) * 8
    eng.update(10 * N); orc.update(10 * N)
    assert_same_state(eng, orc, "compressed ECDF")
    full = sb.Engine(model, prior, **{**kw, "ecdf_max_knots": 0}); full.init()
    kf = full.get_ecdf(0); kc = eng.get_ecdf(0)
    x = np.quantile(kf[1:-1], np.linspace(0.01, 0.99, 200))
    import helpers
    assert np.abs(helpers.g_ecdf_transform(kf, x) - helpers.g_ecdf_transform(kc, x)).max() < 3.0 / min(K, kf.size)


@pytest.mark.parametrize("name", ["gauss_mean", "gauss_sample_d2s2", "sir_tauleap"])
@pytest.mark.parametrize("n", [100, 2049, 16384])
def test_small_population_tail_kernel(gpu, name, n):
    """n <= 16384: one single-CTA kernel replaces the generic tail (rho sums, trigger, resampling, eps, history); both forms and
    the oracle must agree bit for bit, including iterations that resample."""
    model, prior = model_cases()[name]
    alg = "multi_eps" if name == "gauss_sample_d2s2" else "single_eps"
    for flags in (0, sb.SABC_FLAG_GENERIC_TAIL, sb.SABC_FLAG_NO_GRAPH):
        eng, orc = run_pair(model, prior, n, 12, algorithm=alg, resample=max(4, n // 3), flags=flags)
        assert eng.get_state()[1][2] >= 3


def test_checkpoint_history_striding(gpu):
    """history every k-th update plus a final record (:367-382)."""
    model, prior = model_cases()["gauss_mean"]
    eng, orc = run_pair(model, prior, 500, 10, checkpoint=4)
    e, u, r = eng.get_history()
    assert e.shape[0] == 1 + 2 + 1      # init, ix=4, ix=8, final


def test_resume_matches_single_run(gpu):
    """update_population! continues an existing result (:251-271,387-397; test/runtests.jl:67-71)."""
    model, prior = model_cases()["gauss_sample_d1s2"]
    kw = dict(n_particles=800, algorithm="single_eps", proposal=DE(1), resample=800, v=1.0, delta=0.1)
    a = sb.Engine(model, prior, **kw); a.init(); a.update(20 * 800)
    b = sb.Engine(model, prior, **kw); b.init(); b.update(9 * 800); b.update(11 * 800)
    for x, y in zip(a.get_population(), b.get_population()):
        assert np.array_equal(x, y)
    assert np.array_equal(a.get_state()[0], b.get_state()[0]) and np.array_equal(a.get_state()[1], b.get_state()[1])
    # n_simulation < n_particles: no update at all (test/runtests.jl:75-78)
    before = b.get_state()[1].copy(); nh = b.get_history()[0].shape[0]
    b.update(50)
    assert np.array_equal(b.get_state()[1], before) and b.get_history()[0].shape[0] == nh


def test_growing_update_calls_keep_history(gpu):
    """update calls of growing length re-allocate the device history buffer; the replayed CUDA graph must not keep the old
    address (regression: found by bench.py --graph --steps 200 after a 3-step warm-up)."""
    model, prior = model_cases()["gauss_sample_d2s2"]
    kw = dict(n_particles=20_000, algorithm="multi_eps", proposal=DE(2), resample=20_000, v=1.0, delta=0.1)
    eng, orc = make_pair(model, prior, **kw)
    eng.init(); orc.init()
    for n_upd in (3, 40, 1500):
        eng.update(n_upd * 20_000); orc.update(n_upd * 20_000)
    assert_same_state(eng, orc, "growing calls")
    for a, b in zip(eng.get_history(), orc.get_history()):
        assert a.shape == b.shape and np.array_equal(a, b)


def test_host_round_trip(gpu):
    """SABCresult held on the host: upload, update, download (sabc_update_host) equals the device-resident run, and a
    second engine can be resumed from (theta,u,rho,eps,counters,ECDF knots) alone."""
    model, prior = model_cases()["gauss_sample_d2s2"]
    kw = dict(n_particles=1000, algorithm="multi_eps", proposal=DE(2), resample=1000, v=1.0, delta=0.1)
    a = sb.Engine(model, prior, **kw); a.init(); a.update(5 * 1000)
    th, u, rho = a.get_population(); eps, cnt = a.get_state()
    b = sb.Engine(model, prior, **kw)
    for j in range(2):
        b.set_ecdf(j, a.get_ecdf(j))
    th2, u2, rho2, eps2, cnt2 = th.copy(order="F"), u.copy(order="F"), rho.copy(order="F"), eps.copy(), cnt.copy()
    b.update_host(th2, u2, rho2, eps2, cnt2, 7 * 1000)
    a.update(7 * 1000)
    for x, y in zip(a.get_population(), (th2, u2, rho2)):
        assert np.array_equal(x, y)
    assert np.array_equal(a.get_state()[0], eps2) and np.array_equal(a.get_state()[1], cnt2)
    t = b.timing()
    assert t["h2d_ms"] > 0 and t["d2h_ms"] > 0


@pytest.mark.parametrize("name,n_upd,resample", [("gauss_sample_d2s2", 1, 10**9), ("gauss_sample_d2s2", 3, 20000), ("sir_tauleap", 2, 15000),
                                                  ("gauss_mean", 4, 30000)])
def test_host_round_trip_pipelined(gpu, name, n_upd, resample):
    """large slices overlap the transfers with the half-sweeps (sabc_update_host); the result must equal both the strictly
    sequential call (SABC_FLAG_NO_PIPELINE) and the oracle, also when a resampling falls into the last update."""
    model, prior = model_cases()[name]
    N = 70_001
    alg = "multi_eps" if name == "gauss_sample_d2s2" else "single_eps"
    kw = dict(n_particles=N, algorithm=alg, proposal=DE(model.n_para), resample=resample, v=1.0, delta=0.1)
    src = sb.Engine(model, prior, **kw); src.init(); src.update(2 * N)
    orc = ob.OracleEngine(model, prior, **kw); orc.init(); orc.update(2 * N); orc.update(n_upd * N)
    th, u, rho = src.get_population(); eps, cnt = src.get_state()
    outs = []
    for flags in (0, sb.SABC_FLAG_NO_PIPELINE):
        b = sb.Engine(model, prior, flags=flags, **kw)
        for j in range(model.n_stats):
            b.set_ecdf(j, src.get_ecdf(j))
        bufs = [th.copy(order="F"), u.copy(order="F"), rho.copy(order="F"), eps.copy(), cnt.copy()]
        b.update_host(*bufs, n_upd * N)
        assert b.timing()["host_ms"] > 0
        outs.append(bufs)
        dev = b.get_population()
        for x, y in zip(dev, bufs[:3]):
            assert np.array_equal(x, y)                     # what came home is what the device holds
    for x, y in zip(outs[0], outs[1]):
        assert np.array_equal(x, y)
    for x, y in zip(orc.get_population(), outs[0][:3]):
        assert np.array_equal(x, y)
    assert np.array_equal(orc.get_state()[0], outs[0][3]) and np.array_equal(orc.get_state()[1], outs[0][4])


def test_negative_distance_is_an_error(gpu):
    """:185 -- a model returning negative prior distances aborts initialization (obs far below gives |.| >= 0, so use the
    second statistic of gauss_sample with a NaN-free negative trick: obs2 = -inf makes |obs2 - x| = inf, not negative; the
    device check is exercised through the flag instead)."""
    model = sb.models.gauss_mean(float("nan"))       # rho = |y - NaN| = NaN: not negative, not positive -> no knots
    eng = sb.Engine(model, sb.Normal(0, 1), n_particles=100, algorithm="single_eps", proposal=DE(1), resample=200, v=1.0, delta=0.1)
    with pytest.raises(sb.SABCError) as ei:
        eng.init()
    assert ei.value.code == -8                       # build_cdf: maximum() of an empty collection


def test_public_api_counters(gpu):
    """test/runtests.jl:56-78 through the mirrored sabc()/update_population!() surface."""
    f_dist, prior = model_cases()["gauss_sample_d1s1"]
    for algorithm in ("multi_eps", "single_eps"):
        res = sb.sabc(f_dist, prior, n_particles=100, n_simulation=1000, algorithm=algorithm)
        assert res.state.n_simulation <= 1000 and res.state.n_population_updates == 9 and len(res.population) == 100
        sb.update_population(res, f_dist, prior, n_simulation=1000)
        assert res.state.n_simulation <= 2000 and res.state.n_population_updates == 19
        n_sim = res.state.n_simulation
        sb.update_population(res, f_dist, prior, n_simulation=50)
        assert res.state.n_simulation == n_sim
        assert "Approximate posterior sample with 100 particles" in repr(res)
    f2, p2 = model_cases()["gauss_sample_d2s2"]
    for algorithm in ("multi_eps", "single_eps"):
        res = sb.sabc(f2, p2, n_particles=100, n_simulation=1000, algorithm=algorithm)
        assert np.all(res.state.eps < 1) and res.population.shape == (100, 2) and res.u.shape == (100, 2)      # :140,179
    for p in (sb.DifferentialEvolution(n_para=2), sb.StretchMove(), sb.RandomWalk(n_para=2)):                  # :238-266
        res = sb.sabc(f2, p2, proposal=p, n_particles=100, n_simulation=1000)
        sb.update_population(res, f2, p2, proposal=p, n_simulation=1000)
        assert res.state.n_simulation <= 2000
    # f_dist(θ, args...; kwargs...): the extra arguments are bound into the device model (factory as f_dist)
    ra = sb.sabc(sb.models.gauss_sample, p2, 10, 2.0, 42.5, n_para=2, second_is_sum=True, n_particles=100, n_simulation=1000, algorithm="multi_eps")
    rb = sb.sabc(f2, p2, n_particles=100, n_simulation=1000, algorithm="multi_eps")
    assert np.array_equal(ra.population, rb.population) and np.array_equal(ra.state.eps, rb.state.eps)
    with pytest.raises(RuntimeError):
        sb.update_population(res, f2, p2, n_simulation=1000, v=-0.1)
    with pytest.raises(RuntimeError):
        sb.update_population(res, f2, p2, n_simulation=1000, delta=-0.1)


def test_posterior_matches_conjugate_and_oracle(gpu):
    """north_star check (3) on C1: slow annealing (v = 0.02) so the ensemble stays near equilibrium; mean and variance within
    2 MC standard errors of N(10/11, 1/11) with the MC error estimated from independent runs, and a KS test against the
    oracle's particles from other seeds."""
    from scipy import stats
    model, prior = model_cases()["gauss_mean"]
    N, n_upd = 4000, 400
    means, vars_, pops = [], [], []
    for seed in range(6):
        eng = sb.Engine(model, prior, n_particles=N, algorithm="single_eps", proposal=DE(1), resample=2 * N, v=0.02, delta=0.1, seed=100 + seed)
        eng.init(); eng.update(n_upd * N)
        th = eng.get_population()[0][:, 0]
        means.append(th.mean()); vars_.append(th.var()); pops.append(th)
    se_m = np.std(means, ddof=1) / np.sqrt(len(means)); se_v = np.std(vars_, ddof=1) / np.sqrt(len(vars_))
    assert abs(np.mean(means) - 10 / 11) < 2 * se_m, (np.mean(means), se_m)
    assert abs(np.mean(vars_) - 1 / 11) < 2 * se_v, (np.mean(vars_), se_v)
    orc = ob.OracleEngine(model, prior, n_particles=N, algorithm="single_eps", proposal=DE(1), resample=2 * N, v=0.02, delta=0.1, seed=999)
    orc.init(); orc.update(n_upd * N)
    tho = orc.get_population()[0][:, 0]
    # thin to reduce the within-run correlation the KS test ignores
    p = stats.ks_2samp(pops[0][::8], tho[::8]).pvalue
    assert p > 0.01, p


@pytest.mark.parametrize("name,N,alg", [("sir_tauleap", 1_250_000, "single_eps"), ("gauss_mean", 10_000_000, "single_eps"),
                                        ("logistic", 1_000_000, "single_eps"), ("gauss_sample_d2s2", 100_000, "multi_eps")])
def test_full_size_properties(gpu, name, N, alg):
    """BASELINE.json sizes (per GPU), where the oracle would take minutes: size-independent properties instead -- determinism
    (graph replay == direct launches, run twice), the exact integer means equal the means of the downloaded population, every
    particle inside the prior support, u in [0,1], sorted ECDF knots with L = n_positive + 2, counters."""
    model, prior = model_cases()[name]
    kw = dict(n_particles=N, algorithm=alg, proposal=DE(model.n_para), resample=N // 16, v=1.0, delta=0.1)
    outs = []
    for flags in (0, sb.SABC_FLAG_NO_GRAPH):
        eng = sb.Engine(model, prior, flags=flags, **kw)
        eng.init()
        th0, u0, rho0 = eng.get_population()
        for j in range(model.n_stats):
            k = eng.get_ecdf(j)
            assert k[0] == 0.0 and np.all(np.diff(k) >= 0) and k[-1] == 1.5 * k[-2] and k.size == np.count_nonzero(rho0[:, j] > 0) + 2
        eng.update(4 * N)
        outs.append((eng.get_population(), eng.get_state(), eng.get_history()))
        eng.close()
    (th, u, rho), (eps, cnt), (eh, uh, rh) = outs[0]
    for a, b in zip(outs[0][0], outs[1][0]):
        assert np.array_equal(a, b)
    assert np.array_equal(outs[0][1][0], outs[1][1][0]) and np.array_equal(outs[0][1][1], outs[1][1][1])
    assert cnt[0] == 5 * N and cnt[3] == 4 and cnt[2] >= 2 and 0 < cnt[1] < 4 * N
    assert np.all((u >= 0) & (u <= 1 + 1e-15)) and np.all(rho >= 0) and np.all(np.isfinite(th))
    for c, comp in enumerate(prior.components()):
        if isinstance(comp, sb.Uniform):
            assert th[:, c].min() >= comp.a and th[:, c].max() <= comp.b
    assert np.allclose(uh[-1], u.mean(axis=0), rtol=1e-12, atol=0)
    assert np.allclose(rh[-1], rho.mean(axis=0), rtol=1e-9, atol=0)
    assert np.all(np.diff(eh[:, 0]) < 0) and np.all(eps > 0)            # eps anneals monotonically in these runs


@pytest.mark.parametrize("name,N,alg", [("gauss_sample_d2s2", 100_000, "multi_eps"),        # C2
                                        ("logistic", 1_000_000, "single_eps"),              # C3 (split path, s = 20, "hybrid")
                                        ("sir_tauleap", 1_250_000, "single_eps"),           # C4 per-GPU slice (split path, 625 000-item work list)
                                        ("gauss_mean", 10_000_000, "single_eps")])          # C5 (deepest ECDF index, many scan tiles)
def test_full_size_parity_vs_oracle(gpu, name, N, alg):
    """BASELINE.json's configurations at their OWN sizes against the oracle, bit for bit: initialization (prior sample, global sort,
    ECDF knots, transform, first resampling, eps_0) + 3 population updates with resamplings in between.  The oracle runs on all host
    threads (a few seconds per case)."""
    import os
    ob.lib().orc_set_num_threads(len(os.sched_getaffinity(0)))
    model, prior = model_cases()[name]
    eng, orc = run_pair(model, prior, N, 3, algorithm=alg, resample=N // 8)
    cnt = eng.get_state()[1]
    assert cnt[0] == 4 * N and cnt[3] == 3 and cnt[2] >= 2, cnt           # at least one resampling after the initial one
    eng.close(); orc.close()


def test_zero_weights_fail_loudly_and_the_error_does_not_stick(gpu):
    """delta so large that every 32.32 fixed-point resampling weight flushes to zero: the single-GPU path reports it (like the sharded
    one) instead of collapsing the population onto one particle, and the next call -- with a sane delta -- runs (the error flag is
    cleared at the start of every call)."""
    model, prior = model_cases()["gauss_sample_d2s2"]
    for n in (3000, 40_000):                                    # single-CTA tail and generic tail
        eng = sb.Engine(model, prior, n_particles=n, algorithm="single_eps", proposal=DE(2), resample=n // 4, v=1.0, delta=1e6)
        with pytest.raises(sb.SABCError) as ei:
            eng.init()                                           # the initial resampling already has all-zero weights
        assert ei.value.code == -20 and "weights are zero" in str(ei.value)
        eng2 = sb.Engine(model, prior, n_particles=n, algorithm="single_eps", proposal=DE(2), resample=n // 4, v=1.0, delta=0.1)
        eng2.init(); eng2.update(3 * n)
        before = [a.copy() for a in eng2.get_population()]
        eng2.set_tuning(1.0, 1e6, n // 4, DE(2))
        with pytest.raises(sb.SABCError):
            eng2.update(30 * n)                                  # a resampling falls into this call
        eng2.set_tuning(1.0, 0.1, n // 4, DE(2))
        eng2.update(3 * n)                                       # not sticky
        assert eng2.get_state()[1][3] >= 6 and not all(np.array_equal(a, b) for a, b in zip(before, eng2.get_population()))


def test_rho_mean_of_a_very_large_population(gpu):
    """1.25e8 particles with two statistics: the radix-256 tree sums of the columns run concurrently, one CTA per column, each with
    its own scratch for ALL tree levels (a scratch sized for the first level only let column j's third level overwrite column j+1's
    second one beyond 1.2e8 particles).  rho_history[0] must equal the mean of the downloaded prior distances."""
    import ctypes as C
    free, total = C.c_size_t(), C.c_size_t()
    cudart = C.CDLL("libcudart.so")
    if cudart.cudaMemGetInfo(C.byref(free), C.byref(total)) != 0 or free.value < 60e9:
        pytest.skip("needs 60 GB of free device memory")
    model, prior = model_cases()["gauss_sample_d2s2"]
    N = 125_000_000
    eng = sb.Engine(model, prior, n_particles=N, algorithm="multi_eps", proposal=DE(2), resample=2 * N, v=1.0, delta=0.1)
    eng.init()
    rho = eng.get_population(theta=False, u=False)[2]
    rh = eng.get_history()[2]
    assert np.allclose(rh[0], rho.mean(axis=0), rtol=1e-11, atol=0), (rh[0], rho.mean(axis=0))
    eng.close()
