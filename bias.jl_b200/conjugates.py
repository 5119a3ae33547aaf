"""Conjugate component types of the reference API (src/conjugates.jl) and the Sent datum
(src/common.jl:12-18).  These are the host-side prototypes a user passes to LDA / BMM / HDP;
the samplers read only their hyper-parameters and run on the GPU.  The scalar methods are
kept for API completeness (`posterior`, interactive use) and are not on the sampling path.

Python indices are 0-based (word ids in [0, dd)); the Julia shim converts from 1-based.
"""
from __future__ import annotations

import math

import numpy as np

from . import _lib
from .distributions import Dirichlet, Gaussian1D


class Sent:
    """Sparse count vector: words w with counts c (src/common.jl:12-18)."""

    def __init__(self, w, c=None):
        if c is None:                                        # Sent(n) -> zeros
            n = int(w)
            self.w, self.c = np.zeros(n, dtype=np.int64), np.zeros(n, dtype=np.int64)
        else:
            self.w, self.c = np.asarray(w, dtype=np.int64), np.asarray(c, dtype=np.int64)
            if self.w.shape != self.c.shape:
                raise ValueError("Sent: w and c must have the same length")

    def __len__(self):
        return len(self.w)

    def __repr__(self):
        return f"Sent({self.w.tolist()},{self.c.tolist()})"


def sparsify_sentence(x) -> Sent:
    """src/common.jl:94-103 — unique words in first-appearance order with their counts."""
    x = np.asarray(x)
    _, first = np.unique(x, return_index=True)
    words = x[np.sort(first)]
    return Sent(words, np.array([(x == w).sum() for w in words]))


class Gaussian1DGaussian1D:
    """src/conjugates.jl:30-105"""
    kind = _lib.G1G1

    def __init__(self, m0: float, v0: float, vv: float):
        self.m0, self.v0 = float(m0), float(v0)
        self.mu, self.vv = float(m0), float(vv)
        self.nn = 0

    def deepcopy(self):
        c = Gaussian1DGaussian1D(self.m0, self.v0, self.vv)
        c.mu, c.nn = self.mu, self.nn
        return c

    def additem(self, x: float):                             # :60-63
        self.nn += 1
        self.mu = ((self.nn - 1) / self.nn) * self.mu + x / self.nn

    def delitem(self, x: float):                             # :64-74
        if self.nn < 1:
            raise RuntimeError("there is no more datapoints in the conjugate component")
        if self.nn == 1:
            self.mu, self.nn = self.m0, 0
        else:
            self.mu = self.mu * (self.nn / (self.nn - 1)) - x / (self.nn - 1)
            self.nn -= 1

    def posterior(self) -> Gaussian1D:                       # :77-89
        if self.nn > 0:
            denom = self.vv + self.nn * self.v0
            return Gaussian1D((self.nn * self.v0 * self.mu + self.vv * self.m0) / denom, self.v0 * self.vv / denom)
        return Gaussian1D(self.m0, self.v0)

    def logpredictive(self, x: float) -> float:              # :92-103
        post = self.posterior()
        dummy = post.vv + self.vv
        return -(x - post.mu) ** 2 / (2 * dummy) - 0.5 * math.log(2 * math.pi * dummy)

    def loglikelihood(self, x: float) -> float:              # :105
        return self.posterior().logpdf(x)

    def datum_kind(self, sample_item) -> int:
        return _lib.F64

    def spec(self, datum: int) -> _lib.Conjugate:
        return _lib.Conjugate(self.kind, datum, 0, 0.0, self.m0, self.v0, self.vv)

    def __repr__(self):
        return ("Gaussian1DGaussian1D conjugate component\n"
                f"likelihood parameters: mu={self.mu}, vv={self.vv}\n"
                f"prior parameters     : m0={self.m0}, v0={self.v0}\n"
                f"number of data points: nn={self.nn}")


class MultinomialDirichlet:
    """src/conjugates.jl:131-213.  MultinomialDirichlet(dd, aa) stores aa/dd (:142)."""
    kind = _lib.MD

    def __init__(self, dd: int, aa: float):
        self.dd = int(dd)
        self.aa = float(aa) / self.dd
        self.pp = np.zeros(self.dd)
        self.mm = 0
        self.mi = np.zeros(self.dd, dtype=np.int64)
        self.nn = 0

    def deepcopy(self):
        c = MultinomialDirichlet(self.dd, self.aa * self.dd)
        c.aa = self.aa
        c.mm, c.mi, c.nn = self.mm, self.mi.copy(), self.nn
        return c

    def additem(self, x):                                    # :153-157, :163-170
        if isinstance(x, Sent):
            np.add.at(self.mi, x.w, x.c)
            self.mm += int(x.c.sum())
        else:
            self.mi[x] += 1
            self.mm += 1
        self.nn += 1

    def delitem(self, x):                                    # :158-162, :171-178
        if isinstance(x, Sent):
            np.subtract.at(self.mi, x.w, x.c)
            self.mm -= int(x.c.sum())
        else:
            self.mi[x] -= 1
            self.mm -= 1
        self.nn -= 1

    def logpredictive(self, x) -> float:                     # :180, :181-198
        if not isinstance(x, Sent):
            return math.log((self.aa + self.mi[x]) / (self.aa * self.dd + self.mm))
        lll = self.aa + self.mi.astype(np.float64)
        np.add.at(lll, x.w, x.c.astype(np.float64))
        mm = int(x.c.sum())
        lg = np.vectorize(math.lgamma)
        return (math.lgamma(mm + 1) - lg(x.c + 1.0).sum() + math.lgamma(self.aa * self.dd + self.mm)
                - math.lgamma(self.aa * self.dd + self.mm + mm) + (lg(lll) - lg(self.aa + self.mi)).sum())

    def posterior(self) -> Dirichlet:                        # :200-203
        return Dirichlet(self.mi + self.aa)

    def loglikelihood(self, x) -> float:                     # :205-213 (posterior mean, not a log)
        mean = self.posterior().mean()
        if isinstance(x, Sent):
            return float((x.c * mean[x.w]).sum())
        return float(mean[x])

    def datum_kind(self, sample_item) -> int:
        return _lib.SENT if isinstance(sample_item, Sent) else _lib.INT

    def spec(self, datum: int) -> _lib.Conjugate:
        return _lib.Conjugate(self.kind, datum, self.dd, self.aa, 0.0, 1.0, 1.0)

    def __repr__(self):
        return ("MultinomialDirichlet component\n"
                f"cardinality: dd={self.dd}, Dirichlet prior parameter: aa={self.aa}\n"
                f"data: mm={self.mm}, nn={self.nn}")
