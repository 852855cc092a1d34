"""Proposal constructors with the reference's surface and error behaviour (src/proposals.jl:24-36,85-99,132-135)."""
from __future__ import annotations

import math

PROP_DE, PROP_STRETCH, PROP_RW = 0, 1, 2


class Proposal:
    kind: int

    def params(self) -> tuple[float, float]:
        raise NotImplementedError


class DifferentialEvolution(Proposal):
    """DifferentialEvolution(; n_para, σ_gamma=1e-5) | DifferentialEvolution(; γ0, σ_gamma=1e-5); keyword-only like the
    reference (positional arguments are a MethodError there, a TypeError here; test/runtests.jl:204-205)."""
    kind = PROP_DE

    def __init__(self, *, gamma0=None, n_para=None, sigma_gamma=1e-5, **kw):
        if "γ0" in kw:
            gamma0 = kw.pop("γ0")
        if "σ_gamma" in kw:
            sigma_gamma = kw.pop("σ_gamma")
        if kw:
            raise TypeError(f"unexpected keyword arguments {sorted(kw)}")
        if gamma0 is not None and n_para is None:
            self.gamma0 = float(gamma0)
        elif n_para is not None and gamma0 is None:
            self.gamma0 = 2.38 / math.sqrt(2 * n_para)                      # src/proposals.jl:93
        else:
            raise ValueError("Provide either `γ0` or `n_para`, not both.")  # ArgumentError, src/proposals.jl:96
        self.sigma_gamma = float(sigma_gamma)

    def params(self):
        return (self.gamma0, self.sigma_gamma)


class StretchMove(Proposal):
    kind = PROP_STRETCH

    def __init__(self, *, a=2.0):
        self.a = float(a)

    def params(self):
        return (self.a, 0.0)


class RandomWalk(Proposal):
    kind = PROP_RW

    def __init__(self, *, n_para, beta=0.8, **kw):
        if "β" in kw:
            beta = kw.pop("β")
        if kw:
            raise TypeError(f"unexpected keyword arguments {sorted(kw)}")
        if not (0 < beta <= 1):
            raise RuntimeError("Mixing parameter `β` must be between zero and one.")   # src/proposals.jl:30
        self.beta = float(beta)
        self.n_para = int(n_para)

    def params(self):
        return (self.beta, 0.0)
