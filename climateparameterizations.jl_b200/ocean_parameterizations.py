"""Host-side mirror of the OceanParameterizations module surface the NDE path uses
(src/OceanParameterizations.jl:7-21): Dᶜ, Dᶠ, ZeroMeanUnitVarianceScaling, MinMaxScaling, scale, unscale.

These are set-up helpers (a few dozen scalars); the device path applies the operators as stencils and takes
mu/sigma as plain numbers in the model description.
"""
from __future__ import annotations

import numpy as np


def Dc(N: int, delta: float) -> np.ndarray:
    """Dᶜ(N, Δ): N×(N+1) face→centre difference operator (src/differentiation_operators.jl:6-14)."""
    D = np.zeros((N, N + 1))
    i = np.arange(N)
    D[i, i] = -1.0
    D[i, i + 1] = 1.0
    return D / delta


def Df(N: int, delta: float) -> np.ndarray:
    """Dᶠ(N, Δ): (N+1)×N centre→face difference operator with zero first/last rows (:21-29)."""
    D = np.zeros((N + 1, N))
    i = np.arange(1, N)
    D[i, i - 1] = -1.0
    D[i, i] = 1.0
    return D / delta


class ZeroMeanUnitVarianceScaling:
    """μ = mean(data), σ = std(data) (n-1 estimator, Julia's default) over the whole array (feature_scaling.jl:7-23)."""

    def __init__(self, data=None, *, mu=None, sigma=None):
        if data is not None:
            a = np.asarray(data, dtype=np.float64)
            mu, sigma = float(a.mean()), float(a.std(ddof=1))
        self.mu, self.sigma = mu, sigma

    def __call__(self, x):
        return (np.asarray(x) - self.mu) / self.sigma

    scale = __call__

    def unscale(self, y):
        return self.sigma * np.asarray(y) + self.mu

    def inv(self):
        return self.unscale


class MinMaxScaling:
    """feature_scaling.jl:29-47 (not used by any NDE; kept for API completeness)."""

    def __init__(self, data, a=0.0, b=1.0):
        arr = np.asarray(data, dtype=np.float64)
        self.a, self.b, self.data_min, self.data_max = a, b, float(arr.min()), float(arr.max())

    def __call__(self, x):
        return self.a + (np.asarray(x) - self.data_min) * (self.b - self.a) / (self.data_max - self.data_min)

    scale = __call__

    def unscale(self, y):
        return self.data_min + (np.asarray(y) - self.a) * (self.data_max - self.data_min) / (self.b - self.a)

    def inv(self):
        return self.unscale


def scale(x, s):
    return s(x)


def unscale(y, s):
    return s.unscale(y)
