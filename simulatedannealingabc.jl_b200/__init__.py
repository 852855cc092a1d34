"""sabc_b200 -- B200-native SABC population engine (hot path of SimulatedAnnealingABC.jl behind a C ABI).

    import sabc_b200 as sb
    res = sb.sabc(sb.models.gauss_mean(1.0), sb.Normal(0, 1), n_particles=1000, n_simulation=100_000)
    sb.update_population(res, model, prior, n_simulation=50_000)
"""
from . import _lib, models  # noqa: F401
from ._lib import SABCError, SABC_FLAG_FUSED, SABC_FLAG_GENERIC_TAIL, SABC_FLAG_MG_REPLICATED, SABC_FLAG_MG_STRICT_RESAMPLE, SABC_FLAG_NO_GRAPH, SABC_FLAG_NO_PIPELINE, SABC_FLAG_SORT_WORK, SABC_FLAG_TIME_KERNELS  # noqa: F401
from .api import Engine, SABCresult, SABCstate, sabc, update_population  # noqa: F401
from .distributions import (Beta, Cauchy, Exponential, Gamma, InverseGamma, Laplace, LogNormal, Normal, Product, Uniform,  # noqa: F401
                            Weibull, product_distribution)
from .models import DeviceModel  # noqa: F401
from .proposals import DifferentialEvolution, RandomWalk, StretchMove  # noqa: F401

__all__ = ["sabc", "update_population", "SABCresult", "SABCstate", "Engine", "DeviceModel", "models", "Normal", "Uniform", "Exponential", "LogNormal", "Gamma", "Beta", "Cauchy", "Laplace", "Weibull", "InverseGamma",
           "product_distribution", "DifferentialEvolution", "StretchMove", "RandomWalk", "SABCError"]
