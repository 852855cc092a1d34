"""Device model plug-ins: the GPU form of the closure `f_dist(θ, args...; kwargs...)` (contract
src/SimulatedAnnealingABC.jl:421).  A DeviceModel names a kernel family registered in libsabc_b200.so and carries the
parameter blob (observations etc.) that the closure would have captured."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import _lib


@dataclass
class DeviceModel:
    name: str
    n_para: int
    n_stats: int
    par: np.ndarray = field(default_factory=lambda: np.zeros(0))

    def simulate(self, theta, seed=0, particle_base=0, sweep=0) -> np.ndarray:
        """Run the device model for each row of theta (n x d); returns distances n x s."""
        th = _lib.f64(np.atleast_2d(np.asarray(theta, dtype=np.float64).reshape(-1, self.n_para)))
        n = th.shape[0]
        rho = np.zeros((n, self.n_stats), order="F")
        par = np.ascontiguousarray(self.par, dtype=np.float64)
        _lib.check(_lib.lib().sabc_model_simulate(self.name.encode(), _lib.ptr(par), par.size, _lib.ptr(th), n,
                                                  C.c_uint64(seed), particle_base, C.c_uint64(sweep), _lib.ptr(rho)))
        return rho


def gauss_mean(y_obs_mean: float, sigma: float = 1.0, n_obs: int = 10) -> DeviceModel:
    """1-D Gaussian mean with known variance, sufficient-statistic form: ρ = |ȳ_sim − ȳ_obs|,
    ȳ_sim ~ N(θ, σ²/n)  (config C1/C5)."""
    return DeviceModel("gauss_mean", 1, 1, np.array([y_obs_mean, sigma / np.sqrt(n_obs)]))


def gauss_sample(n_obs: int, obs_mean: float, obs_second: float | None = None, *, n_para: int = 1, sigma: float = 1.0,
                 second_is_sum: bool = False) -> DeviceModel:
    """n_obs iid draws N(θ1, θ2 or σ); statistics |obs_mean − mean(y)| and optionally |obs_second − mean(y²)| (or Σy²
    with second_is_sum) -- the f_dist shapes of test/runtests.jl:35,86,128-131,167-170 and docs/src/usage.md:30-35."""
    n_stats = 1 if obs_second is None else 2
    par = np.array([n_obs, sigma, obs_mean, 0.0 if obs_second is None else obs_second, 1.0 if second_is_sum else 0.0])
    return DeviceModel(f"gauss_sample_d{n_para}s{n_stats}", n_para, n_stats, par)


LOGISTIC_T = 20


def logistic(x_obs, x0: float = 10.0) -> DeviceModel:
    """Stochastic logistic growth x' = max(0, x + r x (1 − x/K) + σ x z), θ = (r, K, σ), 20 points, ρ_t = |x_t − x_t^obs| (C3)."""
    x_obs = np.asarray(x_obs, dtype=np.float64)
    if x_obs.size != LOGISTIC_T:
        raise ValueError(f"logistic model needs {LOGISTIC_T} observations")
    return DeviceModel("logistic", 3, LOGISTIC_T, np.concatenate([[x0, LOGISTIC_T], x_obs]))


def sir_tauleap(obs_total: float, obs_peak: float, obs_tpeak: float, pop: float = 1e5, n_steps: int = 50, tau: float = 1.0) -> DeviceModel:
    """SIR tau-leaping, θ = (β, γ, ι, φ): infections ~ Poisson(β S I/pop τ), recoveries ~ Poisson(γ I τ), reported cases ~
    Poisson(φ·infections); statistics (Δtotal)², (Δpeak)², (Δt_peak)² as docs/src/example.md:143-147 (C4)."""
    return DeviceModel("sir_tauleap", 4, 3, np.array([pop, n_steps, tau, obs_total, obs_peak, obs_tpeak]))


def sir_gillespie(obs_total: float, obs_peak: float, obs_tpeak: float, *, S0: int = 99, I0: int = 1, R0: int = 0,
                  t_max: float = 160.0, single_stat: bool = False) -> DeviceModel:
    """The reference's documented example (docs/src/example.md:75-152): event-driven SIR, θ = (β, γ); distances abs2 of total
    infected, peak infected and time of the peak (`f_dist_multi_stats`), or their sum (`f_dist_single_stat`)."""
    s = 1 if single_stat else 3
    return DeviceModel(f"sir_gillespie_s{s}", 2, s, np.array([S0, I0, R0, t_max, obs_total, obs_peak, obs_tpeak], dtype=np.float64))


def registered() -> list[str]:
    L = _lib.lib()
    return [L.sabc_model_name(i).decode() for i in range(L.sabc_model_count())]
