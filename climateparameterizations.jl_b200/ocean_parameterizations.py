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


# ---- LES -> coarse grid (host-side data path; src/DataWrangling/coarse_graining.jl) -------------------
Center, Face = "Center", "Face"


def coarse_grain(Phi, n: int, location: str = Center) -> np.ndarray:
    """coarse_grain(Φ, n, Center|Face) (src/DataWrangling/coarse_graining.jl:8-40).

    Center: block means of an evenly spaced cell-centred profile (n must divide len(Φ)).
    Face: end points preserved; the interior is block-averaged (exactly when (N-2)/(n-2) is an integer, otherwise over
    the index windows round(2 + (i-2)Δ) .. round(2 + (i-1)Δ), Julia's round-half-to-even, 1-based inclusive)."""
    Phi = np.asarray(Phi)
    N = Phi.shape[0]
    if location == Center:
        if N % n != 0:
            raise ValueError(f"n = {n} must evenly divide length(Φ) = {N}")
        return Phi.reshape(n, N // n, *Phi.shape[1:]).mean(axis=1)
    if location != Face:
        raise ValueError("location must be Center or Face")
    out = np.empty((n,) + Phi.shape[1:], dtype=Phi.dtype)
    out[0], out[n - 1] = Phi[0], Phi[N - 1]
    delta = (N - 2) / (n - 2)
    if float(delta).is_integer():
        out[1:n - 1] = coarse_grain(Phi[1:N - 1], n - 2, Center)
    else:
        for i in range(2, n):  # Julia index i = 2..n-1
            i1 = int(np.rint(2 + (i - 2) * delta))
            i2 = int(np.rint(2 + (i - 1) * delta))
            out[i - 1] = Phi[i1 - 1:i2].mean(axis=0)
    return out


def coarse_grain_linear_interpolation(Phi, n: int, location: str = Face) -> np.ndarray:
    """coarse_grain_linear_interpolation(Φ, n, Face) (coarse_graining.jl:47-62): end points preserved, interior points
    linearly interpolated at positions 1 + (i-1)(N-1)/(n-1) (1-based) — the 129 -> 33 face regridding of the LES data."""
    if location != Face:
        raise ValueError("only Face-located fields are interpolated in the reference")
    Phi = np.asarray(Phi)
    N = Phi.shape[0]
    out = np.empty((n,) + Phi.shape[1:], dtype=np.result_type(Phi.dtype, np.float32))
    out[0], out[n - 1] = Phi[0], Phi[N - 1]
    gap = (N - 1) / (n - 1)
    for i in range(2, n):
        pos = 1 + (i - 1) * gap
        f = int(np.floor(pos))
        out[i - 1] = (f + 1 - pos) * Phi[f - 1] + (pos - f) * Phi[f]
    return out
