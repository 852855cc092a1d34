"""Priors: the subset of Distributions.jl the reference's tests and docs use (Uniform, Normal and
product_distribution of them; test/runtests.jl:36,87-88,125,163-164, docs/src/usage.md:20-21).  The objects only
carry parameters; sampling and log-density run on the device (csrc/plugin.cuh)."""
from __future__ import annotations

from dataclasses import dataclass

PRIOR_UNIFORM, PRIOR_NORMAL, PRIOR_EXPONENTIAL, PRIOR_LOGNORMAL = 0, 1, 2, 3


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


UNIVARIATE = (Uniform, Normal, Exponential, LogNormal)


class Product(Distribution):
    def __init__(self, dists):
        self.dists = list(dists)
        if not self.dists or not all(isinstance(d, UNIVARIATE) for d in self.dists):
            raise TypeError("product_distribution supports Uniform, Normal, Exponential and LogNormal components on the device path")

    def components(self):
        return self.dists


def product_distribution(dists) -> Product:
    return Product(dists)
