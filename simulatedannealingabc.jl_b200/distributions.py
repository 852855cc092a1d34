"""Priors: the subset of Distributions.jl the reference's tests and docs use (Uniform, Normal and
product_distribution of them; test/runtests.jl:36,87-88,125,163-164, docs/src/usage.md:20-21) plus Exponential,
LogNormal, Gamma, Beta, Cauchy, Laplace, Weibull and InverseGamma with Distributions.jl's parametrisations.  The objects only carry parameters; sampling and
log-density run on the device (csrc/plugin.cuh)."""
from __future__ import annotations

from dataclasses import dataclass

PRIOR_UNIFORM, PRIOR_NORMAL, PRIOR_EXPONENTIAL, PRIOR_LOGNORMAL, PRIOR_GAMMA, PRIOR_BETA = 0, 1, 2, 3, 4, 5
PRIOR_CAUCHY, PRIOR_LAPLACE, PRIOR_WEIBULL, PRIOR_INVERSEGAMMA = 6, 7, 8, 9


class Distribution:
    def components(self):
        return [self]

    def __len__(self):
        return len(self.components())


@dataclass(frozen=True)
class Uniform(Distribution):
    a: float = 0.0
    b: float = 1.0

    def __post_init__(self):
        if not self.a < self.b:
            raise ValueError("Uniform: the condition a < b is not satisfied")  # Distributions.jl DomainError analogue

    kind = PRIOR_UNIFORM

    def params(self):
        return (float(self.a), float(self.b))


@dataclass(frozen=True)
class Normal(Distribution):
    mu: float = 0.0
    sigma: float = 1.0

    def __post_init__(self):
        if not self.sigma > 0:
            raise ValueError("Normal: the condition σ > 0 is not satisfied")

    kind = PRIOR_NORMAL

    def params(self):
        return (float(self.mu), float(self.sigma))


@dataclass(frozen=True)
class Exponential(Distribution):
    theta: float = 1.0            # scale, as in Distributions.Exponential(θ)

    def __post_init__(self):
        if not self.theta > 0:
            raise ValueError("Exponential: the condition θ > 0 is not satisfied")

    kind = PRIOR_EXPONENTIAL

    def params(self):
        return (float(self.theta), 0.0)


@dataclass(frozen=True)
class LogNormal(Distribution):
    mu: float = 0.0
    sigma: float = 1.0

    def __post_init__(self):
        if not self.sigma > 0:
            raise ValueError("LogNormal: the condition σ > 0 is not satisfied")

    kind = PRIOR_LOGNORMAL

    def params(self):
        return (float(self.mu), float(self.sigma))


@dataclass(frozen=True)
class Gamma(Distribution):
    alpha: float = 1.0            # shape
    theta: float = 1.0            # scale, as in Distributions.Gamma(α, θ)

    def __post_init__(self):
        if not (self.alpha > 0 and self.theta > 0):
            raise ValueError("Gamma: the condition α > zero(α) && θ > zero(θ) is not satisfied")

    kind = PRIOR_GAMMA

    def params(self):
        return (float(self.alpha), float(self.theta))


@dataclass(frozen=True)
class Beta(Distribution):
    alpha: float = 1.0
    beta: float = 1.0

    def __post_init__(self):
        if not (self.alpha > 0 and self.beta > 0):
            raise ValueError("Beta: the condition α > zero(α) && β > zero(β) is not satisfied")

    kind = PRIOR_BETA

    def params(self):
        return (float(self.alpha), float(self.beta))


@dataclass(frozen=True)
class Cauchy(Distribution):
    mu: float = 0.0
    sigma: float = 1.0

    def __post_init__(self):
        if not self.sigma > 0:
            raise ValueError("Cauchy: the condition σ > zero(σ) is not satisfied")

    kind = PRIOR_CAUCHY

    def params(self):
        return (float(self.mu), float(self.sigma))


@dataclass(frozen=True)
class Laplace(Distribution):
    mu: float = 0.0
    theta: float = 1.0

    def __post_init__(self):
        if not self.theta > 0:
            raise ValueError("Laplace: the condition θ > zero(θ) is not satisfied")

    kind = PRIOR_LAPLACE

    def params(self):
        return (float(self.mu), float(self.theta))


@dataclass(frozen=True)
class Weibull(Distribution):
    alpha: float = 1.0            # shape
    theta: float = 1.0            # scale

    def __post_init__(self):
        if not (self.alpha > 0 and self.theta > 0):
            raise ValueError("Weibull: the condition α > zero(α) && θ > zero(θ) is not satisfied")

    kind = PRIOR_WEIBULL

    def params(self):
        return (float(self.alpha), float(self.theta))


@dataclass(frozen=True)
class InverseGamma(Distribution):
    alpha: float = 1.0            # shape
    theta: float = 1.0            # scale

    def __post_init__(self):
        if not (self.alpha > 0 and self.theta > 0):
            raise ValueError("InverseGamma: the condition α > zero(α) && θ > zero(θ) is not satisfied")

    kind = PRIOR_INVERSEGAMMA

    def params(self):
        return (float(self.alpha), float(self.theta))


UNIVARIATE = (Uniform, Normal, Exponential, LogNormal, Gamma, Beta, Cauchy, Laplace, Weibull, InverseGamma)


class Product(Distribution):
    def __init__(self, dists):
        self.dists = list(dists)
        if not self.dists or not all(isinstance(d, UNIVARIATE) for d in self.dists):
            raise TypeError("product_distribution supports Uniform, Normal, Exponential, LogNormal, Gamma, Beta, Cauchy, Laplace, Weibull and InverseGamma components on the device path")

    def components(self):
        return self.dists


def product_distribution(dists) -> Product:
    return Product(dists)
