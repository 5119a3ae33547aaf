"""Value types returned by `posterior` (reference src/distributions.jl). Host-side, tiny;
never used by the samplers."""
from __future__ import annotations

import math

import numpy as np


class Gaussian1D:
    """src/distributions.jl:15-20"""

    def __init__(self, mu: float, vv: float):
        if not vv > 0:
            raise ValueError("Gaussian1D: the condition vv > zero(vv) is not satisfied.")
        self.mu, self.vv = float(mu), float(vv)

    def pdf(self, x: float) -> float:                       # :36
        return math.exp(-(x - self.mu) ** 2 / (2 * self.vv)) / math.sqrt(2 * math.pi * self.vv)

    def logpdf(self, x: float) -> float:                    # :37
        return math.log(self.pdf(x))

    def mean(self) -> float:
        return self.mu

    def sample(self, n=None, rng=None):                     # :32-33
        rng = rng or np.random.default_rng()
        return self.mu + math.sqrt(self.vv) * rng.standard_normal(n)

    def __repr__(self):
        return f"Gaussian1D distribution\nmean: mu={self.mu}, variance: vv={self.vv}"


class Dirichlet:
    """src/distributions.jl:44-65"""

    def __init__(self, alpha, dim: int | None = None):
        if dim is not None:                                 # Dirichlet(alpha::Real, dim) = fill(alpha/dim, dim), :51
            self.alpha = np.full(dim, float(alpha) / dim)
        else:
            self.alpha = np.asarray(alpha, dtype=np.float64).copy()

    def mean(self) -> np.ndarray:                           # :61
        return self.alpha / self.alpha.sum()

    def pdf(self, x: int) -> float:                         # :64 ("in fact expectations")
        return float(self.mean()[x])

    def logpdf(self, x: int) -> float:
        return math.log(self.pdf(x))

    def __repr__(self):
        return f"Dirichlet distribution\ncardinality = {len(self.alpha)}\nalpha       = {self.alpha}"
